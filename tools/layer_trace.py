"""Per-launch %globaltimer trace of the layer path (PSNODE_LG_TRACE=1): gaps between consecutive GEMM launches at the cfg5 shard.
    gpurun -- env PSNODE_LG_TRACE=1 python tools/layer_trace.py [steps]"""
import sys, torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, AE_Func, RK4
dev = "cuda:0"
torch.manual_seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
B, H, T = 8192, 256, N + 1
de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2).to(dev)
ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2).to(dev)
t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
mk = lambda: (torch.randn(T, B, H, device=dev) * 0.05)
z, v = mk(), mk()
x_init, i0 = torch.randn(B, H, device=dev) * 0.05, torch.randn(B, H, device=dev) * 0.05
a0 = torch.cat((x_init, z[0], v[0], i0), dim=-1)
with torch.no_grad():
    for _ in range(2):
        RK4(impl="layer").integrate_DAE(x_init=x_init, x_func=de, i_func=ae, t=t, x=x_init.unsqueeze(0).expand(T, B, H), z=z, v=v,
                                        i=i0.unsqueeze(0).expand(T, B, H), all_initial=a0)
torch.cuda.synchronize()
