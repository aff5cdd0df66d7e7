"""Runs the traced build of the dual-warpgroup kernel (tools/build_variant.sh trace -DPAID_DW_TRACE=1) on the SDXL shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
N = 7
coef = torch.linspace(0, 1, N, device="cuda")
for S, L, h in ((4096, 4096, 10), (1024, 1024, 20), (4096, 77, 10), (1024, 77, 20)):
    q, k, v = (torch.randn(N, T, h * 64, device="cuda").half() for T in (S, L, L))
    for rep in range(2):
        _cabi.attn_core(q, k, v, coef, h, _cabi.PAID_PLAIN, False)
        torch.cuda.synchronize()
        print(f"--- S={S} L={L} heads={h} plain rep {rep}", flush=True)
