"""GPU parity: the CUDA path (through the Python call surface -> C ABI) against the reference's own outputs
(tests/golden/*.npz, produced by the unmodified reference) at rtol=1e-5 / atol=1e-6."""
import pytest
import torch

from helpers import AE, DE, ATOL, RTOL, golden_names, load_golden, params_of, tm, tol_report

pytestmark = pytest.mark.gpu

SOLVERS = {}


def _solver(name, impl):
    from py_psnode_b200 import Euler, Midpoint, RK4
    return {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[name](impl=impl)


def run_ode_case(d, impl, dev="cuda:0", requires_grad=False):
    from py_psnode_b200 import ODE_Event
    de = DE(params_of(d, "de")).to(dev)
    t, x, z = tm(d["t"], dev), tm(d["x"], dev), tm(d["z"], dev)
    name = str(d["_name"])
    ev = ODE_Event()
    event_fn = jump_fn = None
    if "noevent" not in name:
        ev.set_event(t=torch.from_numpy(d["event_t"]).to(dev), z=torch.from_numpy(d["z_jump"]).to(dev))
        event_fn, jump_fn = ev.event_fn, ev.jump_change_fn
    a0 = torch.cat((x[0], z[0]), dim=-1)
    sol = _solver(str(d["solver"]), impl).integrate_ODE(x_func=de, t=t, x=x, z=z, all_initial=a0, event_fn=event_fn,
                                                        jump_change_fn=jump_fn, input_true_x=bool(d["teacher_x"]))
    return sol


def run_dae_case(d, impl, dev="cuda:0"):
    from py_psnode_b200 import DAE_Event
    de, ae = DE(params_of(d, "de")).to(dev), AE(params_of(d, "ae")).to(dev)
    t, x, z, v, i = (tm(d[k], dev) for k in ("t", "x", "z", "v", "i"))
    name = str(d["_name"])
    ev = DAE_Event()
    event_fn = jump_fn = None
    if "noevent" not in name:
        ev.set_event(t=torch.from_numpy(d["event_t"]).to(dev), z=torch.from_numpy(d["z_jump"]).to(dev),
                     v=torch.from_numpy(d["v_jump"]).to(dev))
        event_fn, jump_fn = ev.event_fn, ev.jump_change_fn
    x_init = torch.from_numpy(d["x_init"]).to(dev)
    a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
    return _solver(str(d["solver"]), impl).integrate_DAE(
        x_init=x_init, x_func=de, i_func=ae, t=t, x=x, z=z, v=v, i=i, all_initial=a0, event_fn=event_fn,
        jump_change_fn=jump_fn, input_true_x=bool(d["teacher_x"]), input_true_i=bool(d["teacher_i"]))


@pytest.mark.parametrize("impl", ["generic", "auto"])
@pytest.mark.parametrize("name", [n for n in golden_names("ode0") if "model" not in n])
def test_ode_forward_matches_reference(native_lib, name, impl):
    d = load_golden(name)
    d["_name"] = name
    with torch.no_grad():
        got = run_ode_case(d, impl).cpu()
    want = torch.from_numpy(d["x_sol"])
    want64 = torch.from_numpy(d["x_sol64"])
    assert got.shape == want.shape
    assert torch.equal(got[0], want[0]), "x_sol[0] must equal the given initial state exactly"
    assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), tol_report(got, want, want64)


@pytest.mark.parametrize("impl", ["generic", "auto"])
@pytest.mark.parametrize("name", golden_names("dae0"))
def test_dae_forward_matches_reference(native_lib, name, impl):
    d = load_golden(name)
    d["_name"] = name
    with torch.no_grad():
        gx, gi = run_dae_case(d, impl)
    gx, gi = gx.cpu(), gi.cpu()
    wx, wi = torch.from_numpy(d["x_sol"]), torch.from_numpy(d["i_sol"])
    assert gx.shape == wx.shape and gi.shape == wi.shape
    assert torch.allclose(gx, wx, rtol=RTOL, atol=ATOL), "x: " + tol_report(gx, wx, torch.from_numpy(d["x_sol64"]))
    assert torch.allclose(gi, wi, rtol=RTOL, atol=ATOL), "i: " + tol_report(gi, wi, torch.from_numpy(d["i_sol64"]))


TC_CASES = ["ode01_rk4_small", "ode01_midpoint_small", "ode01_rk4_noevent", "ode01_rk4_nomatch", "ode01_rk4_2events_pad",
            "ode01_rk4_long"]


@pytest.mark.parametrize("impl", ["tc", "tc8"])
@pytest.mark.parametrize("name", TC_CASES)
def test_ode_forward_tensor_core_matches_reference(native_lib, name, impl):
    """tcgen05 3xTF32 kernels (impl="tc": 4 warps per group, "tc8": 8 warps per group) against the reference's fp32 output
    at the same rtol=1e-5 / atol=1e-6."""
    from py_psnode_b200 import _native
    d = load_golden(name)
    d["_name"] = name
    with torch.no_grad():
        got = run_ode_case(d, impl).cpu()
    assert _native.last_kernel().startswith(f"psn_{impl}_ode_kernel")
    want = torch.from_numpy(d["x_sol"])
    want64 = torch.from_numpy(d["x_sol64"])
    assert torch.equal(got[0], want[0])
    assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), tol_report(got, want, want64)


TC_DAE_CASES = ["dae01_euler_small", "dae01_midpoint_small", "dae01_rk4_small", "dae01_rk4_noevent", "dae01_rk4_inputgrads"]


@pytest.mark.parametrize("impl", ["tc", "tc8"])
@pytest.mark.parametrize("name", TC_DAE_CASES)
def test_dae_forward_tensor_core_matches_reference(native_lib, name, impl):
    """tcgen05 DAE kernels (DE net from TMEM, AE net from shared memory) against the reference's fp32 x_sol / i_sol."""
    from py_psnode_b200 import _native
    d = load_golden(name)
    d["_name"] = name
    with torch.no_grad():
        gx, gi = run_dae_case(d, impl)
    assert _native.last_kernel().startswith(f"psn_{impl}_dae_kernel"), _native.last_kernel()
    gx, gi = gx.cpu(), gi.cpu()
    wx, wi = torch.from_numpy(d["x_sol"]), torch.from_numpy(d["i_sol"])
    assert torch.equal(gx[0], torch.from_numpy(d["x_init"]))
    assert torch.allclose(gx, wx, rtol=RTOL, atol=ATOL), "x: " + tol_report(gx, wx, torch.from_numpy(d["x_sol64"]))
    assert torch.allclose(gi, wi, rtol=RTOL, atol=ATOL), "i: " + tol_report(gi, wi, torch.from_numpy(d["i_sol64"]))


@pytest.mark.parametrize("teacher", [False, True])
def test_dae_teacher_forcing_falls_back_to_generic(native_lib, teacher):
    """impl="tc" with teacher forcing is refused (EUNSUPPORTED); "auto" silently takes the generic kernel."""
    from py_psnode_b200 import _native
    d = load_golden("dae01_rk4_tx" if teacher else "dae01_rk4_small")
    d["_name"] = "dae01"
    with torch.no_grad():
        run_dae_case(d, "auto")
    assert _native.last_kernel().startswith("psn_generic_fwd_kernel" if teacher else "psn_tc8_dae_kernel")
    if teacher:
        with pytest.raises(RuntimeError):
            run_dae_case(d, "tc")


def test_dae_tensor_core_matches_generic_at_scale(native_lib):
    """B = 5000 trajectories (ragged last group, two groups per CTA), 150 steps, one event: tensor-core DAE vs generic kernel."""
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, RK4
    torch.manual_seed(5)
    dev = "cuda:0"
    B, N, X, Z, V, I, H = 5000, 150, 16, 1, 2, 4, 64
    T = N + 1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I).to(dev)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z).to(dev)
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda w: torch.randn(T, B, w, device=dev) * 0.1
    x, z, v, i = mk(X), mk(Z), mk(V), mk(I)
    x_init = torch.randn(B, X, device=dev) * 0.1
    ev = DAE_Event()
    ev.set_event(t=t[N // 3].view(B, 1, 1).clone(), z=torch.randn(B, 1, Z, device=dev) * 0.1, v=torch.randn(B, 1, V, device=dev) * 0.1)
    a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
    outs = {}
    with torch.no_grad():
        for impl in ("generic", "tc", "tc8"):
            outs[impl] = RK4(impl=impl).integrate_DAE(x_init=x_init, x_func=de, i_func=ae, t=t, x=x, z=z, v=v, i=i, all_initial=a0,
                                                      event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
    for impl in ("tc", "tc8"):
        for k, nm in ((0, "x"), (1, "i")):
            assert torch.allclose(outs[impl][k], outs["generic"][k], rtol=RTOL, atol=ATOL), \
                f"{impl} {nm}: " + tol_report(outs[impl][k].cpu(), outs["generic"][k].cpu())


def test_tensor_core_matches_fused_at_scale(native_lib):
    """B = 5000 trajectories (ragged last group, two groups per CTA), 200 steps: tensor-core vs CUDA-core fused kernel."""
    from py_psnode_b200 import DE_Func, ODE_Event, RK4
    torch.manual_seed(3)
    dev = "cuda:0"
    B, N, X, Z, H = 5000, 200, 16, 2, 64
    T = N + 1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H).to(dev)
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    x = torch.randn(T, B, X, device=dev) * 0.1
    z = torch.randn(T, B, Z, device=dev) * 0.1
    ev = ODE_Event()
    ev.set_event(t=t[N // 3].view(B, 1, 1).clone(), z=torch.randn(B, 1, Z, device=dev) * 0.1)
    a0 = torch.cat((x[0], z[0]), dim=-1)
    outs = {}
    with torch.no_grad():
        for impl in ("fused", "tc", "tc8"):
            outs[impl] = RK4(impl=impl).integrate_ODE(x_func=de, t=t, x=x, z=z, all_initial=a0, event_fn=ev.event_fn,
                                                      jump_change_fn=ev.jump_change_fn)
    assert torch.allclose(outs["tc"], outs["fused"], rtol=RTOL, atol=ATOL), tol_report(outs["tc"].cpu(), outs["fused"].cpu())
    assert torch.allclose(outs["tc8"], outs["fused"], rtol=RTOL, atol=ATOL), tol_report(outs["tc8"].cpu(), outs["fused"].cpu())
    assert torch.equal(outs["tc8"], outs["tc"]), "tc and tc8 run the same arithmetic in the same order"


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("name", ["ode01_rk4_small", "ode01_rk4_2events_pad", "dae01_rk4_small", "ode01_euler_cfg1"])
def test_host_buffer_entry_matches_reference(native_lib, name, pinned):
    """C ABI `psnode_forward_host` (HOST pointers; staged copies for pageable memory, zero-copy for pinned >= 1 MB)."""
    from py_psnode_b200 import _native as N, engine
    d = load_golden(name)
    dae = str(d["kind"]) == "dae"
    fix = (lambda q: q.pin_memory()) if pinned else (lambda q: q)
    t, x, z = (fix(torch.from_numpy(d[k]).permute(1, 0, 2).contiguous()) for k in ("t", "x", "z"))
    de = [q for wb in params_of(d, "de") for q in wb]
    method = {"euler": N.EULER, "midpoint": N.MIDPOINT, "rk4": N.RK4}[str(d["solver"])]
    ev_t, zj = torch.from_numpy(d["event_t"]), torch.from_numpy(d["z_jump"])
    if dae:
        v, i = (fix(torch.from_numpy(d[k]).permute(1, 0, 2).contiguous()) for k in ("v", "i"))
        ae = [q for wb in params_of(d, "ae") for q in wb]
        x_init = torch.from_numpy(d["x_init"])
        a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
        cfg = engine.Config(kind=N.DAE, method=method, impl=N.IMPL_AUTO, X=x_init.shape[1], Z=z.shape[2], V=v.shape[2], I=i.shape[2],
                            teacher_x=False, teacher_i=False, n_de=len(de) // 2, n_ae=len(ae) // 2, has_event=True)
        tens = [t, None, z, v, i, x_init, a0, ev_t, zj, torch.from_numpy(d["v_jump"]), *de, *ae]
    else:
        a0 = torch.cat((x[0], z[0]), dim=-1)
        cfg = engine.Config(kind=N.ODE, method=method, impl=N.IMPL_AUTO, X=x.shape[2], Z=z.shape[2], V=0, I=0, teacher_x=False,
                            teacher_i=False, n_de=len(de) // 2, n_ae=0, has_event=True)
        tens = [t, x, z, None, None, None, a0, ev_t, zj, None, *de]
    xs, is_, up, down = engine.forward_host(cfg, tens)
    assert up > 0 and down == xs.numel() * 4 + (is_.numel() * 4 if is_ is not None else 0)
    assert torch.allclose(xs, torch.from_numpy(d["x_sol"]), rtol=RTOL, atol=ATOL), tol_report(xs, torch.from_numpy(d["x_sol"]))
    if dae:
        assert torch.allclose(is_, torch.from_numpy(d["i_sol"]), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("kind", ["ode", "dae"])
def test_host_buffer_dma_path_equals_zero_copy_path(native_lib, kind, monkeypatch):
    """PSNODE_HOST_PATH=dma: the grid is integrated in 8 time chunks whose inputs / trajectory rows move on the copy engines
    while the neighbouring chunk integrates.  Same kernels, same per-step arithmetic -> bit-identical to the in-place path,
    including an event that falls on a chunk boundary and the re-evaluated i_0 of every DAE chunk."""
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, ODE_Event, RK4
    torch.manual_seed(91)
    B, N, X, Z, V, I, H = 272, 200, 16, 2, 2, 4, 64
    T = N + 1
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1).contiguous().pin_memory()
    mk = lambda w: (torch.randn(T, B, w) * 0.1).pin_memory()
    x, z, v, i = mk(X), mk(Z), mk(V), mk(I)
    event_t = torch.stack((t[25, :, 0], t[113, :, 0]), dim=1).view(B, 2, 1).clone()      # step 25 is the first chunk boundary
    zj, vj = torch.randn(B, 2, Z) * 0.1, torch.randn(B, 2, V) * 0.1
    outs = {}
    for mode in ("inplace", "dma"):
        if mode == "dma":
            monkeypatch.setenv("PSNODE_HOST_PATH", "dma")
        else:
            monkeypatch.delenv("PSNODE_HOST_PATH", raising=False)
        torch.manual_seed(92)
        if kind == "ode":
            de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H)
            ev = ODE_Event()
            ev.set_event(t=event_t, z=zj)
            a0 = torch.cat((x[0], z[0]), dim=-1)
            xs = RK4().integrate_ODE_host(x_func=de, t=t, x=x, z=z, all_initial=a0, event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
            outs[mode] = (xs.clone(),)
        else:
            de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I)
            ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z)
            ev = DAE_Event()
            ev.set_event(t=event_t, z=zj, v=vj)
            x_init = x[0].clone()
            a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
            xs, is_ = RK4().integrate_DAE_host(x_init=x_init, x_func=de, i_func=ae, t=t, x=x, z=z, v=v, i=i, all_initial=a0,
                                               event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
            outs[mode] = (xs.clone(), is_.clone())
    for a, b in zip(outs["dma"], outs["inplace"]):
        assert torch.isfinite(a).all() and torch.equal(a, b), float((a - b).abs().max())
