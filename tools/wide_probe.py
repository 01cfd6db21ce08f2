"""Timing of the latent-width (impl = wide) kernels at BASELINE configs[3] per-GPU and whole-batch sizes (CUDA events).

    gpurun -- python tools/wide_probe.py [fwd|train]
"""
import sys, torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, RK4, Euler, _native
dev = "cuda:0"
torch.manual_seed(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"

def run(name, B, N, solver, train):
    H = 128
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, depth=2).to(dev)
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    x = torch.randn(T, B, H, device=dev) * 0.1
    z = (torch.randn(T, B, H, device=dev) * 0.1).requires_grad_(train)
    def step():
        a0 = torch.cat((x[0], z[0]), dim=-1)
        if not train:
            with torch.no_grad():
                return solver.integrate_ODE(x_func=de, t=t, x=x, z=z, all_initial=a0)
        out = solver.integrate_ODE(x_func=de, t=t, x=x, z=z, all_initial=a0)
        out.sum().backward()
        return out
    for _ in range(2):
        step()
    k = _native.last_kernel()
    n0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 3
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: B={B} N={N} {'train' if train else 'fwd'} {ms:.2f} ms ({B*N/ms/1e3:.1f} M traj-steps/s) launches/step {(_native.launch_count()-n0)//reps} last kernel {k}", flush=True)

train = mode == "train"
run("cfg4/GPU rk4", 4096, 500, RK4(), train)
run("cfg4/GPU euler", 4096, 500, Euler(), train)
run("cfg4 whole batch rk4", 16384, 500, RK4(), train)
