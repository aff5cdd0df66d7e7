"""One attention-core launch per mode at SDXL geometry, for ncu (uses PAID_LIB_PATH if set)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
N, S, h, d = 7, int(os.environ.get("S", "4096")), int(os.environ.get("H", "10")), 64
q = torch.randn(N, S, h * d, device="cuda").half(); k = torch.randn(N, S, h * d, device="cuda").half(); v = torch.randn(N, S, h * d, device="cuda").half()
coef = torch.linspace(0, 1, N, device="cuda")
for _ in range(2):
    _cabi.attn_core(q, k, v, coef, h, _cabi.PAID_OUTER, True)
    _cabi.attn_core(q, k, v, coef, h, _cabi.PAID_PLAIN, False)
torch.cuda.synchronize()
