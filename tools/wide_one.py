"""One cfg4-shard training step (for ncu captures):  python tools/wide_one.py [euler|rk4] [B] [N]"""
import sys, torch
sys.path.insert(0, '.')
from py_psnode_b200 import DE_Func, RK4, Euler
dev = "cuda:0"
torch.manual_seed(0)
S = RK4 if (len(sys.argv) < 2 or sys.argv[1] == "rk4") else Euler
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
N = int(sys.argv[3]) if len(sys.argv) > 3 else 500
H, T = 128, N + 1
de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, depth=2).to(dev)
t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
x = torch.randn(T, B, H, device=dev) * 0.1
z = (torch.randn(T, B, H, device=dev) * 0.1).requires_grad_(True)
for _ in range(2):
    a0 = torch.cat((x[0], z[0]), dim=-1)
    out = S().integrate_ODE(x_func=de, t=t, x=x, z=z, all_initial=a0)
    out.sum().backward()
torch.cuda.synchronize()
