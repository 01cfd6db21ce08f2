"""Training step of the ODE_01 net at the script default --hidden 128, cfg2 batch (B = 4096 x 1000 RK4 steps), masked-MSE loss fused into the
sweep: tensor-core reverse sweep (PSNODE_WIDE4_BWD=1: psn_wide4_fwd_kernel<rk4,tape> + psn_wide4_bwd_kernel + psn_wide_grad_kernel) against the
generic recomputing sweep (PSNODE_WIDE4_BWD=0), device-timed, with the gradient difference between the two.
    gpurun -- python tools/wide4_train_probe.py [N]"""
import os, sys
import torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, RK4, _native, engine

dev = "cuda:0"
torch.manual_seed(0)
X, Z, H = 16, 2, 128
B = 4096
N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1000
T = N + 1
de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H).to(dev)
t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
z = torch.randn(T, B, Z, device=dev) * 0.1
x0 = torch.randn(B, X, device=dev) * 0.1
a0 = torch.cat((x0, z[0]), dim=-1)
xv = x0.unsqueeze(0).expand(T, B, X)
target = torch.randn(T, B, X, device=dev) * 0.1
mask = torch.ones(T, B, 1, device=dev)
plist = list(de.parameters())


def step():
    for p in plist:
        p.grad = None
    num, _ = RK4().integrate_ODE_loss(x_func=de, t=t, x=xv, z=z, all_initial=a0, target=target, mask=mask)
    (num / mask.sum()).backward()


res = {}
modes = (("0", "0"), ("1", "0"), ("1", "1"))         # (PSNODE_WIDE4_BWD, PSNODE_WIDE4_PF)
if "fast" in sys.argv:
    modes = modes[1:]
for flag, pfl in modes:
    os.environ["PSNODE_WIDE4_BWD"] = flag
    os.environ["PSNODE_WIDE4_PF"] = pfl
    engine.release_tape_pool()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    res[flag + pfl] = [p.grad.clone() for p in plist]
    print(f"PSNODE_WIDE4_BWD={flag} PSNODE_WIDE4_PF={pfl}: training step {ms:.1f} ms = {B * N / ms / 1e3:.1f} M traj-steps/s, peak {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, "
          f"last kernel {_native.last_kernel()}", flush=True)
base = "00" if "00" in res else "10"
for p, a, b, c in zip(plist, res[base], res["10"], res["11"]):
    print(f"   {tuple(p.shape)}: max|tensor-core sweep - {'generic' if base == '00' else 'itself'}| = {(a - b).abs().max().item():.3e}, "
          f"prefetch variant bit-identical: {torch.equal(b, c)} (scale {a.abs().max().item():.3e})", flush=True)
