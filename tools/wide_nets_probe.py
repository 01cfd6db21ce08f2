"""Timing of the generic kernels on the latent (`*_02_direct_encode`) widths at BASELINE configs[3]/[4]-like per-GPU sizes.

    gpurun -- python tools/wide_nets_probe.py
"""
import sys, time, torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, AE_Func, RK4, Euler, _native
dev = "cuda:0"
torch.manual_seed(0)
def run(name, B, N, H, dae, solver):
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H if dae else 0, i_dim=H if dae else 0, depth=2).to(dev)
    ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2).to(dev) if dae else None
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda: (torch.randn(T, B, H, device=dev) * 0.05)
    x, z = mk(), mk().requires_grad_(True)
    if dae:
        v, i = mk().requires_grad_(True), mk()
        x_init = (torch.randn(B, H, device=dev) * 0.05).requires_grad_(True)
    def fwd():
        if dae:
            a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
            return solver.integrate_DAE(x_init=x_init, x_func=de, i_func=ae, t=t, x=x, z=z, v=v, i=i, all_initial=a0)[0]
        a0 = torch.cat((x[0], z[0]), dim=-1)
        return solver.integrate_ODE(x_func=de, t=t, x=x, z=z, all_initial=a0)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.no_grad(): out = fwd()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        out = fwd(); k = _native.last_kernel()
        out.sum().backward()
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name}: B={B} N={N} H={H} fwd {1e3*(t1-t0):.1f} ms ({B*N/(t1-t0)/1e6:.2f} M traj-steps/s)  fwd+bwd {1e3*(t2-t1):.1f} ms  kernel {k}", flush=True)
run("cfg4/GPU rk4", 4096, 500, 128, False, RK4())
run("cfg4/GPU euler", 4096, 500, 128, False, Euler())
run("cfg5-ish rk4", 1024, 100, 256, True, RK4())
