"""Fused encoder / decoder entry against the unfused pipeline at the BASELINE configs[4] shard (DAE_02, latent 256, B = 8192):
time and peak HBM.   gpurun -- python tools/encoded_probe.py [steps] [B]"""
import sys, torch, torch.nn as nn
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, AE_Func, RK4, _native
dev = "cuda:0"
torch.manual_seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
H, XR, ZR, VR, IR = 256, 32, 1, 2, 2
T = N + 1
codec = lambda i, h, o: nn.Sequential(nn.Linear(i, h), nn.ELU(), nn.Linear(h, o)).to(dev)
de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2).to(dev)
ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2).to(dev)
z_enc, v_enc, x_dec, i_dec = codec(ZR, H, H), codec(VR, H, H), codec(H, H, XR), codec(H, H, IR)
t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
z, v = torch.randn(T, B, ZR, device=dev), torch.randn(T, B, VR, device=dev)
x_init, i0 = torch.randn(B, H, device=dev) * 0.05, torch.randn(B, H, device=dev) * 0.05
with torch.no_grad():
    a0 = torch.cat((x_init, z_enc(z[0]), v_enc(v[0]), i0), dim=-1)
def unfused():
    with torch.no_grad():
        Zh, Vh = z_enc(z), v_enc(v)
        xs, is_ = RK4().integrate_DAE(x_init=x_init, x_func=de, i_func=ae, t=t, x=x_init.unsqueeze(0).expand(T, B, H), z=Zh, v=Vh,
                                      i=i0.unsqueeze(0).expand(T, B, H), all_initial=a0)
        return x_dec(xs), i_dec(is_)
def fused():
    return RK4().integrate_DAE_encoded(x_init=x_init, x_func=de, i_func=ae, t=t, z=z, v=v, all_initial=a0, z_encoder=z_enc, v_encoder=v_enc,
                                       x_decoder=x_dec, i_decoder=i_dec)
def timeit(name, fn, reps=2):
    from py_psnode_b200 import engine
    engine._workspaces.clear(); torch.cuda.empty_cache()
    base = torch.cuda.memory_allocated()
    torch.cuda.reset_peak_memory_stats()
    out = fn(); torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - base
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: B={B} N={N}: {ms:.1f} ms ({B*N/ms/1e3:.2f} M traj-steps/s), peak extra HBM {peak/2**30:.2f} GiB, last kernel {_native.last_kernel()}", flush=True)
    return out
fx, fi = timeit("fused  (integrate_DAE_encoded)", fused)
ux, ui = timeit("unfused (torch encoders -> integrate_DAE -> torch decoders)", unfused)
print("max |fused - unfused| x:", float((fx - ux).abs().max()), " i:", float((fi - ui).abs().max()))
