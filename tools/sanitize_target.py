"""Small command for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of libpaid_attn.so once or
twice at the smallest geometries that still exercise its barrier protocol (several key tiles, ragged last tile, two Q
blocks, a second segment, the CTA-pair GEMM with a GEGLU epilogue), through the C ABI.
    compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi  # noqa: E402

torch.manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda").half()
N = 3
coef = torch.tensor([0.0, 0.4, 1.0], device="cuda")
for (S, C, h, L, Cc) in ((320, 128, 2, None, 128), (200, 128, 2, 77, 96), (130, 160, 2, None, 160)):   # head_dim 64, 64, 80
    x, ctx = r(N, S, C), (None if L is None else r(N, L, Cc))
    w = [r(C, C) / C ** 0.5, r(C, Cc) / Cc ** 0.5, r(C, Cc) / Cc ** 0.5, r(C, C) / C ** 0.5, r(C)]
    for mode, fused in ((_cabi.PAID_OUTER, True), (_cabi.PAID_INNER, True), (_cabi.PAID_PLAIN, False)):
        _cabi.attn_forward(x, ctx, *w, coef, h, mode, fused)
    _cabi.attn_forward(torch.cat([x, x]), None if ctx is None else torch.cat([ctx, ctx]), *w, coef, h, _cabi.PAID_OUTER, True,
                       plain_tail=N)
xs = r(600, 256)
_cabi.linear(xs, r(512, 256) / 16, r(512))                      # CTA-pair GEMM (256-wide tiles), ragged M
_cabi.linear(xs, r(320, 256) / 16, None)                        # 1-CTA kernel
_cabi.linear_geglu(xs, r(2 * 512, 256) / 16, r(2 * 512))        # GEGLU epilogue
_cabi.add_layer_norm(r(300, 640), r(300, 640), r(640), r(640))
fm = r(2, 64, 16, 16).contiguous(memory_format=torch.channels_last)
_cabi.group_norm_nhwc(fm, r(64), r(64), 32, silu=True, pre_bias=r(2, 64))
_cabi.residual_bias_add(fm, fm.clone(memory_format=torch.channels_last), r(64))
_cabi.geglu(r(100, 256))
torch.cuda.synchronize()
print("done", _cabi.launch_count(), "kernels; last attention kernel:", _cabi.last_kernel())
