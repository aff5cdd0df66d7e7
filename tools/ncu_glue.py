"""One launch of each HBM-bound glue kernel at SDXL N=7 shapes, for an `ncu --set full` capture:
    ncu --set full --clock-control none -k regex:"gn_|layer_norm|geglu" -c 12 -o gpurun_out/glue python tools/ncu_glue.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
N = 7
for C, side in ((320, 128), (640, 64), (1280, 32), (2560, 32)):
    x = torch.randn(N, C, side, side, device="cuda").half().contiguous(memory_format=torch.channels_last)
    g = torch.ones(C, device="cuda").half(); b = torch.zeros(C, device="cuda").half()
    _cabi.group_norm_nhwc(x, g, b, 32, 1e-5, True, None)                     # gn_stats + gn_apply
    if C == 320:
        _cabi.residual_bias_add(x, x.clone(memory_format=torch.channels_last), b)
for S, C in ((4096, 640), (1024, 1280)):
    x = torch.randn(N, S, C, device="cuda").half(); d = torch.randn_like(x)
    g = torch.ones(C, device="cuda").half(); b = torch.zeros(C, device="cuda").half()
    _cabi.add_layer_norm(x, d, g, b, 1e-5)
    _cabi.geglu(torch.randn(N * S, 8 * C, device="cuda").half())
torch.cuda.synchronize()
