import sys, time, torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, AE_Func, Euler, RK4, _native
dev = "cuda:0"
N, H = 500, int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = N + 1
def bench(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for B in (64, 256, 1024):
    torch.manual_seed(0)
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2).to(dev)
    ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2).to(dev)
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda: torch.randn(T, B, H, device=dev) * 0.05
    x, z, v, i = mk(), mk().requires_grad_(True), mk().requires_grad_(True), mk()
    a0 = torch.cat((x[0], z[0].detach(), v[0].detach(), i[0]), dim=-1)
    for impl in ("layer", "generic"):
        for S in (Euler, RK4):
            def call():
                return S(impl=impl).integrate_DAE(x_init=x[0], x_func=de, i_func=ae, t=t, x=x, z=z, v=v, i=i, all_initial=a0)
            def fwd():
                with torch.no_grad(): call()
            def train():
                for p in list(de.parameters()) + list(ae.parameters()): p.grad = None
                z.grad = None; v.grad = None
                sum(o.square().mean() for o in call()).backward()
            print(f"dae02 H={H} B={B} {S.__name__} impl={impl}: fwd {bench(fwd):.2f} ms, fwd+bwd {bench(train):.2f} ms [{_native.last_kernel()}]", flush=True)
