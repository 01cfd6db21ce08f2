"""Host-side dispatch queries of the C ABI that need no GPU (pure arithmetic on the problem descriptor, no kernel launch, no device memory):
which problems get an activation tape, how large it is, and how the environment switches of the hidden-128 path change that.  Pointers in the
descriptors are dummies (validated for non-NULL only, never dereferenced by these entry points)."""
import ctypes as C

import pytest

from py_psnode_b200 import _native as N

DUMMY = 0x1000          # non-NULL, never dereferenced


def _mlp(dims):
    m = N.Mlp()
    m.n_layers = len(dims) - 1
    for k in range(len(dims) - 1):
        m.in_dim[k], m.out_dim[k] = dims[k], dims[k + 1]
        m.W[k], m.b[k] = DUMMY, DUMMY
    return m


def _ode_problem(X, Z, hidden, depth, B, T, impl, method=2):
    p = N.Problem()
    p.kind, p.method, p.impl = N.ODE, method, N.IMPL_BY_NAME[impl]
    p.B, p.T, p.X, p.Z = B, T, X, Z
    for s in (p.t, p.x, p.z, p.x_sol):
        s.p, s.st, s.sb = DUMMY, B * 16, 16
    p.a0, p.a0_sb = DUMMY, X + Z
    p.de = _mlp([3 * (X + Z)] + [hidden] * (depth - 1) + [X])
    return p


@pytest.fixture
def lib():
    return N.lib()


def test_tape_sizes_per_kernel_family(lib, monkeypatch):
    monkeypatch.delenv("PSNODE_WIDE4", raising=False)
    monkeypatch.delenv("PSNODE_WIDE4_BWD", raising=False)
    B, T = 100, 11                              # 7 groups of 16 trajectories, 10 steps
    groups, steps, stages = 7, 10, 4
    # 4-layer ODE_01 at hidden 128 (the training script's default): a1 | a2 | a3 blocks (3 x 2048 floats) + the 16 x 16 stage input per group-stage
    p = _ode_problem(16, 2, 128, 4, B, T, "auto")
    assert lib.psnode_tape_floats(C.byref(p)) == groups * steps * stages * (3 * 2048 + 256)
    assert lib.psnode_tape_covers_input_grads(C.byref(p)) == 0          # its sweep does not produce input-series gradients
    assert lib.psnode_forward_workspace(C.byref(p)) >= 256
    p.method = 0                                                        # Euler: one stage per step
    assert lib.psnode_tape_floats(C.byref(p)) == groups * steps * 1 * (3 * 2048 + 256)
    p.method = 2
    # narrower nets run there only on request (impl = wide): zero-padded to 128 neurons, same record size
    assert lib.psnode_tape_floats(C.byref(_ode_problem(16, 2, 96, 4, B, T, "auto"))) == groups * steps * stages * (3 * 2048 + 256)
    assert lib.psnode_tape_floats(C.byref(_ode_problem(16, 2, 40, 4, B, T, "auto"))) == 0
    assert lib.psnode_tape_floats(C.byref(_ode_problem(16, 2, 40, 4, B, T, "wide"))) == groups * steps * stages * (3 * 2048 + 256)
    # outside its limits (state wider than 16, more than 8 held inputs, other depths): no tape, generic kernels
    assert lib.psnode_tape_floats(C.byref(_ode_problem(17, 2, 128, 4, B, T, "auto"))) == 0
    assert lib.psnode_tape_floats(C.byref(_ode_problem(16, 9, 128, 4, B, T, "auto"))) == 0
    assert lib.psnode_tape_floats(C.byref(_ode_problem(16, 2, 128, 3, B, T, "auto"))) == 0
    assert lib.psnode_tape_floats(C.byref(_ode_problem(16, 2, 128, 4, B, T, "generic"))) == 0
    # the latent ODE_02 net (X = Z = H = 128, 2 layers): a1 + y blocks per group-stage, and its sweep covers the input-series gradients
    q = _ode_problem(128, 128, 128, 2, B, T, "auto")
    assert lib.psnode_tape_floats(C.byref(q)) == groups * steps * stages * 2 * 2048
    assert lib.psnode_tape_covers_input_grads(C.byref(q)) == 1
    # the H = 64 net keeps its own (tc8) tape
    assert lib.psnode_tape_floats(C.byref(_ode_problem(16, 2, 64, 4, B, T, "auto"))) > 0


def test_environment_switches_of_the_hidden_128_path(lib, monkeypatch):
    p = _ode_problem(16, 2, 128, 4, 64, 6, "auto")
    monkeypatch.setenv("PSNODE_WIDE4_BWD", "0")                         # generic recomputing sweep: no tape is recorded
    assert lib.psnode_tape_floats(C.byref(p)) == 0
    monkeypatch.setenv("PSNODE_WIDE4_BWD", "1")
    assert lib.psnode_tape_floats(C.byref(p)) > 0


def test_invalid_descriptors_are_rejected_not_sized(lib):
    p = _ode_problem(16, 2, 128, 4, 64, 6, "auto")
    p.de.in_dim[0] = 3 * 18 + 1                                         # first layer does not match cat(a0, s - a0, s)
    assert lib.psnode_tape_floats(C.byref(p)) == 0 and lib.psnode_forward_workspace(C.byref(p)) == 0
    p = _ode_problem(16, 2, 128, 4, 64, 6, "auto")
    p.a0 = None
    assert lib.psnode_tape_floats(C.byref(p)) == 0
