"""The four scripts at their argparse defaults (--batch 64, --hidden 128; solver Euler as hard-coded in the models, and RK4): time of one
integrate call and of one forward + backward on the GPU, next to the oracle port on the host cores (same sizes, all threads).
    gpurun -- python tests/probe_script_default.py [steps]"""
import sys, time, torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, AE_Func, Euler, RK4, _native
from oracle import psnode_oracle as O
dev = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
B, H, T = 64, 128, N + 1
def params(mod):
    return [(m.weight.detach().cpu(), m.bias.detach().cpu()) for m in mod if isinstance(m, torch.nn.Linear)]
def bench(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for name in ("ode01", "ode02", "dae01", "dae02"):
    torch.manual_seed(0)
    dae, latent = name.startswith("dae"), name.endswith("02")
    X, Z, V, I = (H, H, H, H) if latent else (16, 2, 2, 4)
    if not dae: V = I = 0
    depth = 2 if latent else 4
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I, depth=depth) if dae else DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, depth=depth)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z, depth=depth) if dae else None
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda w: torch.randn(T, B, w) * 0.05
    x, z = mk(X), mk(Z)
    v, i = (mk(V), mk(I)) if dae else (None, None)
    a0 = torch.cat((x[0], z[0]) + ((v[0], i[0]) if dae else ()), dim=-1)
    for S, sname in ((Euler, "euler"), (RK4, "rk4")):
        ded, aed = de.to(dev), (ae.to(dev) if dae else None)
        td, xd, a0d = t.to(dev), x.to(dev), a0.to(dev)
        zd = z.to(dev).requires_grad_(latent)
        vd, idv = (v.to(dev).requires_grad_(latent), i.to(dev)) if dae else (None, None)
        def call():
            if dae:
                return S().integrate_DAE(x_init=xd[0], x_func=ded, i_func=aed, t=td, x=xd, z=zd, v=vd, i=idv, all_initial=a0d)
            return (S().integrate_ODE(x_func=ded, t=td, x=xd, z=zd, all_initial=a0d),)
        def fwd():
            with torch.no_grad(): call()
        def train():
            for p in list(ded.parameters()) + (list(aed.parameters()) if dae else []): p.grad = None
            sum(o.square().mean() for o in call()).backward()
        f_ms = bench(fwd); kf = _native.last_kernel()
        tr_ms = bench(train); kb = _native.last_kernel()
        # CPU: oracle port, all threads, forward only
        pd = params(de.cpu().x_dot)
        t0 = time.perf_counter()
        with torch.no_grad():
            if dae: O.integrate_dae(sname, pd, params(ae.cpu().i_calculator), x[0], t, x, z, v, i, a0)
            else: O.integrate_ode(sname, pd, t, x, z, a0)
        c_ms = (time.perf_counter() - t0) * 1e3
        print(f"{name} hidden=128 batch=64 {N} {sname} steps: GPU forward {f_ms:.2f} ms [{kf}], forward+backward {tr_ms:.2f} ms [{kb}]; CPU port forward {c_ms:.0f} ms "
              f"({torch.get_num_threads()} threads) -> x{c_ms / f_ms:.0f}", flush=True)
