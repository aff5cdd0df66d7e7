"""One attention-core launch per (SDXL attention-layer class, mode) for an `ncu --set full` capture:
    ncu --set full --clock-control none --import-source on -k regex:attn_tc -c 8 -o gpurun_out/attn_core python tools/ncu_core.py
Launch order (= row order of the report): for each class in CLASSES: fused-outer, then plain.  N = 7 frames."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
CLASSES = [(4096, 4096, 10), (4096, 77, 10), (1024, 1024, 20), (1024, 77, 20)]   # (S, L, heads), head_dim 64
if __name__ == "__main__":
    N = 7
    coef = torch.linspace(0, 1, N, device="cuda")
    for S, L, h in CLASSES:
        q = torch.randn(N, S, h * 64, device="cuda").half(); k = torch.randn(N, L, h * 64, device="cuda").half(); v = torch.randn_like(k)
        _cabi.attn_core(q, k, v, coef, h, _cabi.PAID_OUTER, True)
        _cabi.attn_core(q, k, v, coef, h, _cabi.PAID_PLAIN, False)
    torch.cuda.synchronize()
