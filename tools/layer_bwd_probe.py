"""Timing of the layer path's reverse sweep (psn_lg_backward) at the BASELINE configs[4] per-GPU shard: DAE_02, H = 256, B = 8192.
    gpurun -- python tools/layer_bwd_probe.py [steps] [H] [reps]"""
import sys, torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, AE_Func, RK4, _native
dev = "cuda:0"
torch.manual_seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
B, T = 8192, N + 1
de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2).to(dev)
ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2).to(dev)
t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
mk = lambda: (torch.randn(T, B, H, device=dev) * 0.05)
z, v = mk().requires_grad_(True), mk().requires_grad_(True)
x_init, i0 = torch.randn(B, H, device=dev) * 0.05, torch.randn(B, H, device=dev) * 0.05
xv, iv = x_init.unsqueeze(0).expand(T, B, H), i0.unsqueeze(0).expand(T, B, H)
a0 = torch.cat((x_init, z[0].detach(), v[0].detach(), i0), dim=-1)
wx, wi = mk(), mk()
def fwd():
    return RK4().integrate_DAE(x_init=x_init, x_func=de, i_func=ae, t=t, x=xv, z=z, v=v, i=iv, all_initial=a0)
def step():
    xs, is_ = fwd()
    return xs, is_
xs, is_ = step()
gx, gi = wx, wi
torch.autograd.backward([xs, is_], [gx, gi])
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
n0 = _native.launch_count()
for _ in range(reps):
    torch.cuda.synchronize(); e[0].record()
    xs, is_ = step()
    e[1].record()
    n1 = _native.launch_count()
    torch.autograd.backward([xs, is_], [gx, gi])
    e[2].record(); torch.cuda.synchronize()
    tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    nb = _native.launch_count() - n1
print(f"layer path H={H} B={B} N={N}: forward {tf/reps:.2f} ms ({tf/reps/N*1e3:.1f} us/step), reverse sweep {tb/reps:.2f} ms ({tb/reps/N*1e3:.1f} us/step, "
      f"{nb} launches, {nb/N:.1f} per step), kernel {_native.last_kernel()}", flush=True)
