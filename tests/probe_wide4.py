"""The ODE_01 net at the scripts' argparse default --hidden 128 (neural_00_ODE_01_no_encode.py:245-247; DE_Func 54-128-128-128-16) at the
cfg2 batch (B = 4096 x 1000 RK4 steps) and at the scripts' default batch (64): tensor-core kernel (impl = wide -> psn_wide4_fwd_kernel)
against the CUDA-core generic kernel, device-timed, with the error of both against the oracle's float64 run on the first 32 trajectories.
    gpurun -- python tests/probe_wide4.py            # timing + accuracy table
    gpurun -- python tests/probe_wide4.py one 200    # one forward call of N steps (for ncu captures)"""
import sys
import torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, RK4, _native

dev = "cuda:0"
torch.manual_seed(0)
X, Z, H = 16, 2, 128


def problem(B, N):
    T = N + 1
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    z = torch.randn(T, B, Z, device=dev) * 0.1
    x0 = torch.randn(B, X, device=dev) * 0.1
    a0 = torch.cat((x0, z[0]), dim=-1)
    return t, x0.unsqueeze(0).expand(T, B, X), z, a0


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H).to(dev)
if len(sys.argv) > 1 and sys.argv[1] == "one":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    t, xv, z, a0 = problem(4096, N)
    with torch.no_grad():
        for _ in range(3):
            RK4(impl="wide").integrate_ODE(x_func=de, t=t, x=xv, z=z, all_initial=a0)
    torch.cuda.synchronize()
    sys.exit(0)

from oracle import psnode_oracle as O
for B, N in ((4096, 1000), (64, 1000)):
    t, xv, z, a0 = problem(B, N)
    res = {}
    with torch.no_grad():
        for impl in ("generic", "wide", "auto"):
            ms, out = timed(lambda: RK4(impl=impl).integrate_ODE(x_func=de, t=t, x=xv, z=z, all_initial=a0), 3)
            res[impl] = out
            print(f"B={B} N={N} impl={impl:8s} {ms:8.2f} ms = {B * N / ms / 1e3:7.1f} M traj-steps/s   {_native.last_kernel()}", flush=True)
    print(f"   max|wide - generic| = {(res['wide'] - res['generic']).abs().max().item():.3e}   "
          f"bitwise repeat: {torch.equal(res['wide'], RK4(impl='wide').integrate_ODE(x_func=de, t=t, x=xv, z=z, all_initial=a0).detach())}", flush=True)
    nb = 32
    pc = [(m.weight.detach().cpu(), m.bias.detach().cpu()) for m in de.x_dot if isinstance(m, torch.nn.Linear)]
    tc, xc, zc, ac = t[:, :nb].cpu(), xv[:, :nb].cpu().contiguous(), z[:, :nb].cpu(), a0[:nb].cpu()
    w32 = O.integrate_ode("rk4", pc, tc, xc, zc, ac)
    w64 = O.integrate_ode("rk4", [(W.double(), b.double()) for W, b in pc], tc.double(), xc.double(), zc.double(), ac.double())
    for name, got in (("oracle fp32", w32), ("generic", res["generic"][:, :nb].cpu()), ("wide4", res["wide"][:, :nb].cpu())):
        print(f"   {name:12s} max|. - float64| = {(got.double() - w64).abs().max().item():.3e}   "
              f"allclose(fp32 oracle, 1e-5/1e-6): {torch.allclose(got, w32, rtol=1e-5, atol=1e-6)}", flush=True)
    plist = list(de.parameters())

    def step():
        for p in plist:
            p.grad = None
        RK4().integrate_ODE(x_func=de, t=t, x=xv, z=z, all_initial=a0).square().mean().backward()
    ms, _ = timed(step, 2)
    print(f"   training step (impl=auto forward + reverse sweep): {ms:.1f} ms   last kernel {_native.last_kernel()}", flush=True)
