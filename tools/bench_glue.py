"""CUDA-event timing of the HBM-bound glue kernels at SDXL N=7 shapes (L2 flushed between launches), both GroupNorm
register / occupancy variants (PAID_GN_VARIANT).  Prints one JSON line per kernel: time, algorithmic bytes, GB/s."""
import json, os, statistics, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=9, warm=3):
    ts = []
    for i in range(warm + iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= warm: ts.append(a.elapsed_time(b))
    return statistics.median(ts)
N = 7
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
def report(name, ms, nbytes, **kw):
    print(json.dumps(dict(kernel=name, ms=round(ms, 4), MB=round(nbytes / 1e6, 1), GBps=round(nbytes / ms / 1e6, 0),
                          frac_of_hbm_peak=round(nbytes / ms / 1e6 / peak, 3), **kw)), flush=True)
for C, side in ((320, 128), (960, 128), (640, 64), (1920, 64), (1280, 32), (2560, 32)):
    x = torch.randn(N, C, side, side, device="cuda").half().contiguous(memory_format=torch.channels_last)
    g = torch.ones(C, device="cuda").half(); b = torch.zeros(C, device="cuda").half(); pb = torch.randn(N, C, device="cuda").half()
    for variant in ("0", "1"):
        os.environ["PAID_GN_VARIANT"] = variant
        for pre in (None, pb):
            ms = timeit(lambda: _cabi.group_norm_nhwc(x, g, b, 32, 1e-5, True, pre))
            report("group_norm_nhwc+silu (stats + apply)", ms, 3 * x.numel() * 2, C=C, side=side, variant=int(variant), pre_bias=pre is not None)
    os.environ.pop("PAID_GN_VARIANT")
    if C in (320, 640, 1280):
        y = x.clone(memory_format=torch.channels_last)
        report("residual_bias_add", timeit(lambda: _cabi.residual_bias_add(x, y, b)), 3 * x.numel() * 2, C=C, side=side)
for S, C in ((4096, 640), (1024, 1280)):
    x = torch.randn(N, S, C, device="cuda").half(); d = torch.randn_like(x)
    g = torch.ones(C, device="cuda").half(); b = torch.zeros(C, device="cuda").half()
    report("add_layer_norm", timeit(lambda: _cabi.add_layer_norm(x, d, g, b, 1e-5)), 4 * x.numel() * 2, S=S, C=C)
    report("layer_norm (no delta)", timeit(lambda: _cabi.add_layer_norm(x, None, g, b, 1e-5)), 2 * x.numel() * 2, S=S, C=C)
    h = torch.randn(N * S, 8 * C, device="cuda").half()
    report("geglu", timeit(lambda: _cabi.geglu(h)), 3 * h.numel(), S=S, C=C)
