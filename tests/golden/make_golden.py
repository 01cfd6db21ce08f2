#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run from anywhere INSIDE THE BUILD CONTAINER (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (xxh0523/Py_PSNODE @ d366e75) is imported from /root/reference with stub modules for
`ray`, `ray.worker` and `matplotlib*`; its own `neural_dae.Euler/Midpoint/RK4`
(`neural_dae/my_fixed_grid.py`, `neural_dae/my_solvers.py`) and the script-local `DE_Func` / `AE_Func`
classes (`neural_00_ODE_01_no_encode.py:58-68`, `neural_00_ODE_02_direct_encode.py:49-57`,
`neural_01_DAE_01_no_encode.py:61-83`, `neural_01_DAE_02_direct_encode.py:70-100`) are called
unchanged on seeded synthetic inputs.  Inputs, weights, fp32 outputs, fp64 outputs (same code on
`.double()` copies) and autograd gradients are written to one .npz per case.

This script must never import the repo's own `neural_dae` shim: it puts /root/reference first on
sys.path and refuses to run if the imported package is not the reference's.
"""
import copy
import os
import sys
import types

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _install_stubs():
    ray = types.ModuleType("ray")
    ray_worker = types.ModuleType("ray.worker")
    ray_worker.init = lambda *a, **k: None
    ray.worker = ray_worker
    sys.modules["ray"] = ray
    sys.modules["ray.worker"] = ray_worker
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    mpl.rcParams = {}
    plt = types.ModuleType("matplotlib.pyplot")
    markers = types.ModuleType("matplotlib.markers")
    mpl.pyplot = plt
    mpl.markers = markers
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules["matplotlib.markers"] = markers


def _load_reference():
    # drop the repo root / script dir from the path so the repo's own shim can never be picked up
    repo_root = os.path.dirname(os.path.dirname(HERE))
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (repo_root, HERE)]
    sys.path.insert(0, REF)
    _install_stubs()
    import importlib
    nd = importlib.import_module("neural_dae")
    assert os.path.abspath(nd.__file__).startswith(REF), nd.__file__
    mods = {
        "ode01": importlib.import_module("neural_00_ODE_01_no_encode"),
        "ode02": importlib.import_module("neural_00_ODE_02_direct_encode"),
        "dae01": importlib.import_module("neural_01_DAE_01_no_encode"),
        "dae02": importlib.import_module("neural_01_DAE_02_direct_encode"),
    }
    return nd, mods


def main():
    import numpy as np
    import torch

    torch.set_num_threads(1)   # single-threaded => deterministic summation order in the CPU kernels
    nd, mods = _load_reference()
    solvers = {"euler": nd.Euler, "midpoint": nd.Midpoint, "rk4": nd.RK4}

    def seq_params(seq):
        out = []
        for m in seq:
            if isinstance(m, torch.nn.Linear):
                out.append((m.weight, m.bias))
        return out

    def make_inputs(g, B, N, widths, dt=0.01, scale=0.1, pad_tail=0):
        """t[b,j] = dt*j (identical across b, as in the reference's simulator data); series ~ N(0, scale^2).
        pad_tail > 0 marks the last steps as padding with t = -1 (reference: neural_base.py mask/padding)."""
        T = N + 1
        t = (torch.arange(T, dtype=torch.float32) * dt).view(1, T, 1).repeat(B, 1, 1).contiguous()
        if pad_tail:
            t[:, T - pad_tail:] = -1.0
        series = {k: (torch.randn(B, T, w, generator=g) * scale) for k, w in widths.items()}
        return t, series

    def run_ode(name, script, solver, X, Z, H, B, N, event=True, teacher=False, grads=True, latent=False,
                input_grads=False, n_events=1, pad_tail=0, seed=0):
        g = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        M = mods[script]
        de = M.DE_Func(x_dim=X, z_dim=Z, hidden_dim=H)
        t_bt, s = make_inputs(g, B, N, {"x": X, "z": Z}, pad_tail=pad_tail)
        x_bt, z_bt = s["x"], s["z"]
        ev = nd.ODE_Event()
        if event:
            # E events per sample; event k of every sample sits at grid index idx[k] (exact float equality)
            idx = [((k + 1) * N) // (n_events + 1) for k in range(n_events)]
            event_t = torch.stack([t_bt[:, j] for j in idx], dim=1).contiguous()          # (B,E,1)
            z_jump = torch.randn(B, n_events, Z, generator=g) * 0.1
        else:
            event_t = torch.full((B, 1, 1), -5.0)
            z_jump = torch.zeros(B, 1, Z)

        def forward(de_mod, dtype, want_grads):
            tt, xx, zz = t_bt.to(dtype), x_bt.to(dtype), z_bt.to(dtype)
            zj = z_jump.to(dtype)
            leaves = {}
            if want_grads and input_grads:
                xx = xx.clone().requires_grad_(True); zz = zz.clone().requires_grad_(True); zj = zj.clone().requires_grad_(True)
                leaves.update(x=xx, z=zz, z_jump=zj)
            ev.set_event(t=event_t.to(dtype), z=zj)
            a0 = torch.cat((xx.permute(1, 0, 2)[0], zz.permute(1, 0, 2)[0]), dim=-1)
            if want_grads and not input_grads:
                a0 = a0.detach().clone().requires_grad_(True)
                leaves["all_initial"] = a0
            sol = solvers[solver]().integrate_ODE(
                x_func=de_mod, t=tt.permute(1, 0, 2), x=xx.permute(1, 0, 2), z=zz.permute(1, 0, 2), all_initial=a0,
                event_fn=ev.event_fn if event is not None else None,
                jump_change_fn=ev.jump_change_fn if event is not None else None, input_true_x=teacher)
            return sol, leaves

        with torch.no_grad():
            x_sol, _ = forward(de, torch.float32, False)
            x_sol64, _ = forward(copy.deepcopy(de).double(), torch.float64, False)
        out = dict(kind="ode", script=script, solver=solver, X=X, Z=Z, H=H, B=B, N=N, teacher_x=teacher,
                   t=t_bt.numpy(), x=x_bt.numpy(), z=z_bt.numpy(), event_t=event_t.numpy(), z_jump=z_jump.numpy(),
                   x_sol=x_sol.numpy(), x_sol64=x_sol64.numpy(), n_layers=len(seq_params(de.x_dot)))
        for li, (W, b) in enumerate(seq_params(de.x_dot)):
            out[f"de_W{li}"] = W.detach().numpy(); out[f"de_b{li}"] = b.detach().numpy()
        if grads:
            gx = torch.randn(x_sol.shape, generator=g)
            out["gx"] = gx.numpy()
            for dtype, tag, mod in ((torch.float32, "", de), (torch.float64, "64", copy.deepcopy(de).double())):
                mod.zero_grad()
                sol, leaves = forward(mod, dtype, True)
                (sol * gx.to(dtype)).sum().backward()
                for li, (W, b) in enumerate(seq_params(mod.x_dot)):
                    out[f"g{tag}_de_W{li}"] = W.grad.numpy(); out[f"g{tag}_de_b{li}"] = b.grad.numpy()
                for k, v in leaves.items():
                    out[f"g{tag}_{k}"] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: x_sol {tuple(x_sol.shape)} max|x|={x_sol.abs().max():.3f} fp32-vs-fp64 {(x_sol.double()-x_sol64).abs().max():.2e}")

    def run_dae(name, script, solver, X, Z, V, I, H, B, N, event=True, teacher_x=False, teacher_i=False, grads=True,
                latent=False, input_grads=False, seed=0):
        g = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        M = mods[script]
        if latent:   # DAE_02: every solver-side width equals hidden (z may be absent: z_dim == 0)
            de = M.DE_Func(x_dim=X, z_dim=Z, v_dim=V, i_dim=I, hidden_dim=H)
            ae = M.AE_Func(x_dim=X, z_dim=Z, v_dim=V, i_dim=I, hidden_dim=H)
            wX, wZ, wV, wI = H, (H if Z != 0 else 0), H, H
        else:
            de = M.DE_Func(x_dim=X, z_dim=Z, v_dim=V, i_dim=I, hidden_dim=H)
            ae = M.AE_Func(x_dim=X, z_dim=Z, v_dim=V, i_dim=I, hidden_dim=H)
            wX, wZ, wV, wI = X, Z, V, I
        t_bt, s = make_inputs(g, B, N, {"x": wX, "z": wZ, "v": wV, "i": wI})
        x_bt, z_bt, v_bt, i_bt = s["x"], s["z"], s["v"], s["i"]
        x_init = torch.randn(B, wX, generator=g) * 0.1
        ev = nd.DAE_Event()
        if event:
            event_t = t_bt[:, N // 2].view(B, 1, 1).contiguous()
            z_jump = torch.randn(B, 1, wZ, generator=g) * 0.1
            v_jump = torch.randn(B, 1, wV, generator=g) * 0.1
        else:
            event_t = torch.full((B, 1, 1), -5.0)
            z_jump = torch.zeros(B, 1, wZ); v_jump = torch.zeros(B, 1, wV)

        def forward(de_mod, ae_mod, dtype, want_grads):
            tt, xx, zz, vv, ii = (a.to(dtype) for a in (t_bt, x_bt, z_bt, v_bt, i_bt))
            zj, vj, xi = z_jump.to(dtype), v_jump.to(dtype), x_init.to(dtype)
            leaves = {}
            if want_grads:
                xi = xi.clone().requires_grad_(True); leaves["x_init"] = xi
            if want_grads and input_grads:
                zz = zz.clone().requires_grad_(True); vv = vv.clone().requires_grad_(True)
                zj = zj.clone().requires_grad_(True); vj = vj.clone().requires_grad_(True)
                xx = xx.clone().requires_grad_(True); ii = ii.clone().requires_grad_(True)
                leaves.update(z=zz, v=vv, z_jump=zj, v_jump=vj, x=xx, i=ii)
            ev.set_event(t=event_t.to(dtype), z=zj, v=vj)
            a0 = torch.cat((xi, zz.permute(1, 0, 2)[0], vv.permute(1, 0, 2)[0], ii.permute(1, 0, 2)[0]), dim=-1)
            if want_grads and not input_grads:
                a0 = a0.detach().clone().requires_grad_(True); leaves["all_initial"] = a0
            xs, is_ = solvers[solver]().integrate_DAE(
                x_init=xi, x_func=de_mod, i_func=ae_mod, t=tt.permute(1, 0, 2), x=xx.permute(1, 0, 2), z=zz.permute(1, 0, 2),
                v=vv.permute(1, 0, 2), i=ii.permute(1, 0, 2), all_initial=a0,
                event_fn=ev.event_fn if event is not None else None,
                jump_change_fn=ev.jump_change_fn if event is not None else None,
                input_true_x=teacher_x, input_true_i=teacher_i)
            return xs, is_, leaves

        with torch.no_grad():
            x_sol, i_sol, _ = forward(de, ae, torch.float32, False)
            x_sol64, i_sol64, _ = forward(copy.deepcopy(de).double(), copy.deepcopy(ae).double(), torch.float64, False)
        out = dict(kind="dae", script=script, solver=solver, X=wX, Z=wZ, V=wV, I=wI, H=H, B=B, N=N, teacher_x=teacher_x,
                   teacher_i=teacher_i, t=t_bt.numpy(), x=x_bt.numpy(), z=z_bt.numpy(), v=v_bt.numpy(), i=i_bt.numpy(),
                   x_init=x_init.numpy(), event_t=event_t.numpy(), z_jump=z_jump.numpy(), v_jump=v_jump.numpy(),
                   x_sol=x_sol.numpy(), i_sol=i_sol.numpy(), x_sol64=x_sol64.numpy(), i_sol64=i_sol64.numpy(),
                   n_layers=len(seq_params(de.x_dot)))
        for li, (W, b) in enumerate(seq_params(de.x_dot)):
            out[f"de_W{li}"] = W.detach().numpy(); out[f"de_b{li}"] = b.detach().numpy()
        for li, (W, b) in enumerate(seq_params(ae.i_calculator)):
            out[f"ae_W{li}"] = W.detach().numpy(); out[f"ae_b{li}"] = b.detach().numpy()
        if grads:
            gx = torch.randn(x_sol.shape, generator=g); gi = torch.randn(i_sol.shape, generator=g)
            out["gx"] = gx.numpy(); out["gi"] = gi.numpy()
            for dtype, tag, dm, am in ((torch.float32, "", de, ae),
                                       (torch.float64, "64", copy.deepcopy(de).double(), copy.deepcopy(ae).double())):
                dm.zero_grad(); am.zero_grad()
                xs, is_, leaves = forward(dm, am, dtype, True)
                ((xs * gx.to(dtype)).sum() + (is_ * gi.to(dtype)).sum()).backward()
                for li, (W, b) in enumerate(seq_params(dm.x_dot)):
                    out[f"g{tag}_de_W{li}"] = W.grad.numpy(); out[f"g{tag}_de_b{li}"] = b.grad.numpy()
                for li, (W, b) in enumerate(seq_params(am.i_calculator)):
                    out[f"g{tag}_ae_W{li}"] = W.grad.numpy(); out[f"g{tag}_ae_b{li}"] = b.grad.numpy()
                for k, v in leaves.items():
                    out[f"g{tag}_{k}"] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: x_sol {tuple(x_sol.shape)} i_sol {tuple(i_sol.shape)} max|x|={x_sol.abs().max():.3f} "
              f"fp32-vs-fp64 x {(x_sol.double()-x_sol64).abs().max():.2e} i {(i_sol.double()-i_sol64).abs().max():.2e}")

    def run_ode_model(name, seed=0):
        """The whole reference ODE_Model.forward (neural_00_ODE_01_no_encode.py:78-91), i.e. the caller of the
        boundary, with its permuted (non-contiguous) views, plus the masked-MSE training loss gradient (:353-355)."""
        g = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        M = mods["ode01"]
        X, Z, H, B, N = 16, 2, 64, 6, 40
        model = M.ODE_Model(x_dim=X, z_dim=Z, hidden_dim=H)
        model.solver = nd.RK4()
        t_bt, s = make_inputs(g, B, N, {"x": X, "z": Z})
        event_t = t_bt[:, N // 2].view(B, 1, 1).contiguous()
        z_jump = torch.randn(B, 1, Z, generator=g) * 0.1
        mask = torch.ones(B, N + 1, X)
        x_pred = model(t_bt, s["x"], s["z"], event_t, z_jump)
        loss = torch.sum(torch.nn.functional.mse_loss(x_pred, s["x"], reduction="none") * mask) / torch.sum(mask)
        loss.backward()
        out = dict(kind="ode_model", X=X, Z=Z, H=H, B=B, N=N, t=t_bt.numpy(), x=s["x"].numpy(), z=s["z"].numpy(),
                   event_t=event_t.numpy(), z_jump=z_jump.numpy(), mask=mask.numpy(), x_pred=x_pred.detach().numpy(),
                   loss=loss.detach().numpy(), solver="rk4", n_layers=4)
        for li, (W, b) in enumerate(seq_params(model.de_func.x_dot)):
            out[f"de_W{li}"] = W.detach().numpy(); out[f"de_b{li}"] = b.detach().numpy()
            out[f"g_de_W{li}"] = W.grad.numpy(); out[f"g_de_b{li}"] = b.grad.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: loss {loss.item():.6f}")

    # ---- ODE (integrate_ODE, my_solvers.py:52-80) ------------------------------------------------
    run_ode("ode01_euler_cfg1", "ode01", "euler", X=4, Z=1, H=32, B=8, N=100)                       # BASELINE configs[0]
    run_ode("ode01_rk4_small", "ode01", "rk4", X=16, Z=2, H=64, B=8, N=60)                          # configs[1] dims
    run_ode("ode01_midpoint_small", "ode01", "midpoint", X=16, Z=2, H=64, B=8, N=40)
    run_ode("ode01_rk4_teacher", "ode01", "rk4", X=16, Z=2, H=64, B=6, N=30, teacher=True, input_grads=True)
    run_ode("ode01_rk4_noevent", "ode01", "rk4", X=16, Z=2, H=64, B=5, N=30, event=None)
    run_ode("ode01_rk4_nomatch", "ode01", "rk4", X=16, Z=2, H=64, B=5, N=30, event=False)          # event_t = -5 never matches
    run_ode("ode01_rk4_2events_pad", "ode01", "rk4", X=16, Z=2, H=64, B=7, N=48, n_events=2, pad_tail=6, input_grads=True)
    run_ode("ode01_rk4_long", "ode01", "rk4", X=16, Z=2, H=64, B=4, N=1000, grads=False)            # full-length drift check
    run_ode("ode01_rk4_odd", "ode01", "rk4", X=5, Z=3, H=24, B=3, N=20, input_grads=True)           # ragged widths
    run_ode("ode02_euler_latent", "ode02", "euler", X=32, Z=32, H=32, B=6, N=30, input_grads=True)  # configs[3] structure
    run_ode("ode02_rk4_latent", "ode02", "rk4", X=32, Z=32, H=32, B=6, N=30, input_grads=True)
    run_ode_model("ode01_model_rk4")
    # ---- DAE (integrate_DAE, my_solvers.py:82-131) -----------------------------------------------
    run_dae("dae01_euler_small", "dae01", "euler", X=16, Z=1, V=2, I=4, H=64, B=8, N=60)
    run_dae("dae01_rk4_small", "dae01", "rk4", X=16, Z=1, V=2, I=4, H=64, B=8, N=60)                # configs[2] dims
    run_dae("dae01_midpoint_small", "dae01", "midpoint", X=16, Z=1, V=2, I=4, H=64, B=6, N=30)
    run_dae("dae01_rk4_noevent", "dae01", "rk4", X=16, Z=1, V=2, I=4, H=64, B=5, N=30, event=None)
    run_dae("dae01_rk4_tx", "dae01", "rk4", X=16, Z=1, V=2, I=4, H=64, B=5, N=30, teacher_x=True, input_grads=True)
    run_dae("dae01_rk4_ti", "dae01", "rk4", X=16, Z=1, V=2, I=4, H=64, B=5, N=30, teacher_i=True, input_grads=True)
    run_dae("dae01_rk4_txi", "dae01", "rk4", X=16, Z=1, V=2, I=4, H=64, B=5, N=30, teacher_x=True, teacher_i=True, input_grads=True)
    run_dae("dae01_rk4_inputgrads", "dae01", "rk4", X=16, Z=1, V=2, I=4, H=64, B=5, N=30, input_grads=True)
    run_dae("dae02_euler_latent", "dae02", "euler", X=3, Z=1, V=2, I=2, H=32, B=6, N=30, latent=True, input_grads=True)  # configs[4] structure
    run_dae("dae02_rk4_latent", "dae02", "rk4", X=3, Z=1, V=2, I=2, H=32, B=6, N=30, latent=True, input_grads=True)
    run_dae("dae02_rk4_latent_noz", "dae02", "rk4", X=3, Z=0, V=2, I=2, H=32, B=4, N=20, latent=True, input_grads=True)  # z_dim == 0 branch


if __name__ == "__main__":
    main()
