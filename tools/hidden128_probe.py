"""The `*_01` nets at the scripts' argparse default --hidden 128 (neural_00_ODE_01_no_encode.py:245-247): 4-layer DE_Func 54-128-128-128-16
at the cfg2 batch (B = 4096 x 1000 RK4 steps).  These widths have no tensor-core kernel (H = 64 only): CUDA-core generic kernels.
    gpurun -- python tools/hidden128_probe.py"""
import sys, time, torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, AE_Func, RK4, _native
dev = "cuda:0"
torch.manual_seed(0)
B, N = 4096, 1000
T = N + 1
for H in (64, 128):
    X, Z = 16, 2
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H).to(dev)
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    z = torch.randn(T, B, Z, device=dev) * 0.1
    x0 = torch.randn(B, X, device=dev) * 0.1
    a0 = torch.cat((x0, z[0]), dim=-1)
    xv = x0.unsqueeze(0).expand(T, B, X)
    def fwd():
        return RK4().integrate_ODE(x_func=de, t=t, x=xv, z=z, all_initial=a0)
    with torch.no_grad():
        fwd(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): fwd()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    k = _native.last_kernel()
    plist = list(de.parameters())
    def step():
        for p in plist: p.grad = None
        fwd().square().mean().backward()
    step(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2): step()
    torch.cuda.synchronize(); dtr = (time.perf_counter() - t0) / 2
    print(f"ODE_01 hidden={H}: fwd {dt*1e3:.1f} ms ({B*N/dt/1e6:.0f} M traj-steps/s, {k}), training step {dtr*1e3:.1f} ms ({_native.last_kernel()})", flush=True)
