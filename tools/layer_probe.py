"""Timing of the per-layer GEMM path (impl = layer) at the BASELINE configs[4] per-GPU shard: DAE_02, H = 256, B = 8192.
    gpurun -- python tools/layer_probe.py [steps]"""
import sys, torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, AE_Func, RK4, Euler, _native
dev = "cuda:0"
torch.manual_seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
def run(name, B, N, H, solver, impl):
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2).to(dev)
    ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2).to(dev)
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda: (torch.randn(T, B, H, device=dev) * 0.05)
    z, v = mk(), mk()
    x_init, i0 = torch.randn(B, H, device=dev) * 0.05, torch.randn(B, H, device=dev) * 0.05
    xv, iv = x_init.unsqueeze(0).expand(T, B, H), i0.unsqueeze(0).expand(T, B, H)
    a0 = torch.cat((x_init, z[0], v[0], i0), dim=-1)
    def step():
        with torch.no_grad():
            return solver(impl=impl).integrate_DAE(x_init=x_init, x_func=de, i_func=ae, t=t, x=xv, z=z, v=v, i=iv, all_initial=a0)
    step()
    k = _native.last_kernel(); n0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    reps = 2
    for _ in range(reps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: B={B} N={N} H={H} {ms:.2f} ms ({B*N/ms/1e3:.2f} M traj-steps/s) launches/call {(_native.launch_count()-n0)//reps} {ms*1e3/((_native.launch_count()-n0)//reps):.2f} us/launch kernel {k}", flush=True)
run("cfg5/GPU rk4 layer", 8192, N, 256, RK4, "layer")
run("cfg5/GPU euler layer", 8192, N, 256, Euler, "layer")
run("dae02 H=128 rk4 layer", 8192, N, 128, RK4, "layer")
run("cfg5/GPU rk4 generic", 8192, min(N, 20), 256, RK4, "generic")
