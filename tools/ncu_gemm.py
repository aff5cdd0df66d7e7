import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
M, N, K = 7168, 1280, int(os.environ.get("K", "5120"))
x = torch.randn(M, K, device="cuda").half(); w = torch.randn(N, K, device="cuda").half(); b = torch.randn(N, device="cuda").half()
for _ in range(3):
    y = _cabi.linear(x, w, b)
torch.cuda.synchronize()
