import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
N = 7
coef = torch.tensor([0.0, 0.23, 0.36, 0.5, 0.64, 0.77, 1.0], device="cuda")
for (S, C, h, L, Cc) in ((4096, 640, 10, None, 640), (4096, 640, 10, 77, 2048), (1024, 1280, 20, None, 1280), (1024, 1280, 20, 77, 2048)):
    x = torch.randn(N, S, C, device="cuda").half()
    ctx = None if L is None else torch.randn(N, L, Cc, device="cuda").half()
    w = [torch.randn(C, C, device="cuda").half() / C ** 0.5, torch.randn(C, Cc, device="cuda").half() / Cc ** 0.5,
         torch.randn(C, Cc, device="cuda").half() / Cc ** 0.5, torch.randn(C, C, device="cuda").half() / C ** 0.5,
         torch.randn(C, device="cuda").half()]
    for mode, fused in ((_cabi.PAID_OUTER, True), (_cabi.PAID_PLAIN, False), (_cabi.PAID_INNER, True)):
        for it in range(20):
            y = _cabi.attn_forward(x, ctx, *w, coef, h, mode, fused)
        torch.cuda.synchronize()
        print("ok", S, C, h, L, mode, fused, float(y.float().abs().mean()), flush=True)
