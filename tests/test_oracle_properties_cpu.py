"""Size-independent properties of the oracle (CPU, float64), on top of the golden fixtures that pin it to the reference: what the three
schemes of neural_dae/my_fixed_grid.py:15-59 must do on right-hand sides with known solutions, the zero-order-hold / event semantics of
my_solvers.py:52-80 and neural_base.py:52-65, and the algebraic-variable bookkeeping of integrate_DAE (my_solvers.py:95,108-121).  The same
properties hold for the CUDA path through the parity tests against this oracle."""
import math

import pytest
import torch

from oracle import psnode_oracle as O

torch.set_default_dtype(torch.float32)


def _affine_net(X, Z, hidden, A, d, seed=0):
    """A 4-layer ELU net that equals f(s) = A s_x + d on the region visited: every pre-activation is kept far above 0 (ELU = identity there),
    so the net is affine in its input; built so that the composed map picks out the `s` block of cat(a0, s - a0, s)."""
    g = torch.Generator().manual_seed(seed)
    S = X + Z
    W1 = torch.zeros(hidden, 3 * S, dtype=torch.float64)
    P = torch.randn(hidden, X, generator=g, dtype=torch.float64) * 0.3          # hidden = P x + 50
    W1[:, 2 * S:2 * S + X] = P
    b1 = torch.full((hidden,), 50.0, dtype=torch.float64)
    W2 = torch.eye(hidden, dtype=torch.float64); b2 = torch.zeros(hidden, dtype=torch.float64)
    W3 = torch.eye(hidden, dtype=torch.float64); b3 = torch.zeros(hidden, dtype=torch.float64)
    Pinv = torch.linalg.pinv(P)                                                  # X x hidden: Pinv (P x + 50) = x + 50 Pinv 1
    W4 = A @ Pinv
    b4 = d - 50.0 * (W4 @ torch.ones(hidden, dtype=torch.float64))
    return [(W1, b1), (W2, b2), (W3, b3), (W4, b4)]


def _grid(T, B, t_end):
    return (torch.arange(T, dtype=torch.float64) * (t_end / (T - 1))).view(T, 1, 1).repeat(1, B, 1)


@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_constant_slope_is_integrated_exactly(method):
    X, Z, B, T = 3, 2, 4, 9
    c = torch.tensor([0.5, -1.0, 2.0], dtype=torch.float64)
    net = _affine_net(X, Z, 8, torch.zeros(X, X, dtype=torch.float64), c)
    t = _grid(T, B, 1.0)
    x = torch.randn(T, B, X, dtype=torch.float64)
    z = torch.randn(T, B, Z, dtype=torch.float64)
    a0 = torch.cat((x[0], z[0]), dim=-1)
    sol = O.integrate_ode(method, net, t, x, z, a0)
    want = x[0].unsqueeze(0) + c * (t - t[0])
    assert torch.allclose(sol, want, rtol=0, atol=1e-11)
    assert torch.equal(sol[0], x[0])


@pytest.mark.parametrize("method,order", [("euler", 1), ("midpoint", 2), ("rk4", 4)])
def test_convergence_order_on_a_linear_system(method, order):
    """dx/dt = A x + d with a rotation-plus-decay A: halving the step divides the end-point error by 2^order."""
    X, Z, B = 2, 1, 3
    A = torch.tensor([[-0.3, 2.0], [-2.0, -0.3]], dtype=torch.float64)
    d = torch.tensor([0.2, -0.1], dtype=torch.float64)
    net = _affine_net(X, Z, 6, A, d)
    x0 = torch.randn(B, X, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    xs = -torch.linalg.solve(A, d)                                              # steady state; exact: xs + expm(A t)(x0 - xs)
    exact = xs + (torch.matrix_exp(A * 1.0) @ (x0 - xs).T).T
    errs = []
    for n in (20, 40, 80):
        T = n + 1
        t = _grid(T, B, 1.0)
        x = x0.unsqueeze(0).expand(T, B, X)
        z = torch.zeros(T, B, Z, dtype=torch.float64)
        a0 = torch.cat((x0, z[0]), dim=-1)
        sol = O.integrate_ode(method, net, t, x, z, a0)
        errs.append((sol[-1] - exact).abs().max().item())
    for e0, e1 in zip(errs, errs[1:]):
        assert math.log2(e0 / e1) == pytest.approx(order, abs=0.15), (method, errs)


def _random_net(dims, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(o, i, generator=g, dtype=torch.float64) / math.sqrt(i), torch.randn(o, generator=g, dtype=torch.float64) * 0.1)
            for i, o in zip(dims[:-1], dims[1:])]


def test_event_replaces_the_held_input_of_exactly_one_step():
    """An event at grid time t[k] makes step k+1 hold z_jump[:, e] instead of z[k] (all stages of that step, every sample; sample 0's clock
    decides): identical to integrating with that one row of the series replaced; an event time off the grid changes nothing."""
    X, Z, B, T = 4, 2, 5, 12
    net = _random_net([3 * (X + Z), 16, 16, 16, X], 7)
    g = torch.Generator().manual_seed(8)
    t = _grid(T, B, 0.5)
    x = torch.randn(T, B, X, generator=g, dtype=torch.float64) * 0.2
    z = torch.randn(T, B, Z, generator=g, dtype=torch.float64) * 0.2
    a0 = torch.cat((x[0], z[0]), dim=-1)
    k = 5
    event_t = torch.stack((t[k, :, 0], t[k, :, 0] + 1e-3), dim=1).view(B, 2, 1).clone()     # second event never matches a grid time
    z_jump = torch.randn(B, 2, Z, generator=g, dtype=torch.float64)
    got = O.integrate_ode("rk4", net, t, x, z, a0, event_t, z_jump)
    z_rep = z.clone()
    z_rep[k] = z_jump[:, 0]
    want = O.integrate_ode("rk4", net, t, x, z_rep, a0)
    plain = O.integrate_ode("rk4", net, t, x, z, a0)
    assert torch.equal(got, want)
    assert torch.equal(got[:k + 1], plain[:k + 1]) and not torch.allclose(got[k + 1:], plain[k + 1:])
    both = torch.stack((t[k, :, 0], t[k, :, 0]), dim=1).view(B, 2, 1)
    with pytest.raises(ValueError):                                             # the reference's `.view` throws on two matches
        O.integrate_ode("rk4", net, t, x, z, a0, both, z_jump)


def test_dae_algebraic_trajectory_is_the_ae_net_on_the_returned_state():
    """i_sol[j] = i_func(x_sol[j], z[j], v[j]) with the UN-jumped inputs (my_solvers.py:121), also on an event step; the state only sees
    the algebraic variable of the previous grid point (held through the step)."""
    X, Z, V, I, B, T = 3, 1, 2, 2, 4, 8
    S = X + Z + V + I
    de = _random_net([3 * S, 12, 12, 12, X], 11)
    ae = _random_net([S + X + Z + V, 12, 12, 12, I], 12)
    g = torch.Generator().manual_seed(13)
    t = _grid(T, B, 0.4)
    mk = lambda w: torch.randn(T, B, w, generator=g, dtype=torch.float64) * 0.2
    x, z, v, i = mk(X), mk(Z), mk(V), mk(I)
    x_init = torch.randn(B, X, generator=g, dtype=torch.float64) * 0.2
    a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
    event_t = t[3, :, 0].view(B, 1, 1).clone()
    zj = torch.randn(B, 1, Z, generator=g, dtype=torch.float64)
    vj = torch.randn(B, 1, V, generator=g, dtype=torch.float64)
    xs, is_ = O.integrate_dae("midpoint", de, ae, x_init, t, x, z, v, i, a0, event_t, zj, vj)
    assert torch.equal(xs[0], x_init)
    for j in range(T):
        assert torch.allclose(is_[j], O.ae_eval(ae, a0, xs[j], z[j], v[j]), rtol=0, atol=1e-13)
    # changing the algebraic net's LAST-row inputs z[T-1], v[T-1] moves only i_sol[T-1]
    z2, v2 = z.clone(), v.clone()
    z2[-1] += 1.0
    v2[-1] -= 1.0
    xs2, is2 = O.integrate_dae("midpoint", de, ae, x_init, t, x, z2, v2, i, a0, event_t, zj, vj)
    assert torch.equal(xs2, xs) and torch.equal(is2[:-1], is_[:-1]) and not torch.allclose(is2[-1], is_[-1])
