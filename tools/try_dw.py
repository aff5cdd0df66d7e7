"""Persistent dual-warpgroup kernel (attn_dw.cu, PLAIN / INNER) against the one-warpgroup kernel (FLAG_ONE_WARPGROUP) and
the generic SIMT kernel: agreement on ragged / SDXL / SD1.5 shapes, repeatability, then CUDA-event timing (L2 flushed)."""
import json, os, statistics, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
OLD, GEN = _cabi.FLAG_ONE_WARPGROUP, _cabi.FLAG_GENERIC_KERNELS
MODES = [("plain", _cabi.PAID_PLAIN, False), ("inner_fused", _cabi.PAID_INNER, True), ("inner_pure", _cabi.PAID_INNER, False)]
torch.manual_seed(0)
worst = 0.0
shapes = ((3, 130, 77, 2, 64), (5, 700, 333, 3, 64), (4, 257, 64, 1, 64), (3, 64, 5, 2, 40), (3, 1, 1, 1, 64), (3, 129, 65, 5, 16),
          (3, 200, 77, 2, 56), (7, 1024, 1024, 20, 64), (4, 300, 640, 2, 64), (7, 1024, 77, 20, 64), (2, 4096, 4096, 2, 64),
          (5, 300, 130, 3, 40), (3, 257, 128, 2, 64), (3, 257, 129, 2, 64), (3, 257, 192, 2, 64), (3, 257, 193, 2, 64))
if os.environ.get("TRY_DW_TIMING_ONLY") == "1": shapes = ()
for N, S, L, h, d in shapes:
    q, k, v = (torch.randn(N, T, h * d, device="cuda").half() for T in (S, L, L))
    if L == 640: k = k * torch.linspace(0.2, 6.0, L, device="cuda").view(1, L, 1).half()     # growing logits: rescale path
    coef = torch.linspace(0, 1, N, device="cuda")
    for name, mode, fused in MODES:
        ref = _cabi.attn_core(q, k, v, coef, h, mode, fused, flags=GEN).float()
        old = _cabi.attn_core(q, k, v, coef, h, mode, fused, flags=OLD).float()
        out = _cabi.attn_core(q, k, v, coef, h, mode, fused).float()
        torch.cuda.synchronize()
        rms = ref.pow(2).mean().sqrt()
        err, err_old = float((out - ref).pow(2).mean().sqrt() / rms), float((old - ref).pow(2).mean().sqrt() / rms)
        worst = max(worst, err)
        again = _cabi.attn_core(q, k, v, coef, h, mode, fused).float()
        ok = bool(torch.isfinite(out).all()) and err < 1e-3 and torch.equal(out, again)
        print(json.dumps(dict(N=N, S=S, L=L, h=h, d=d, mode=name, err_vs_generic=err, old_err_vs_generic=err_old, ok=ok)), flush=True)
        assert ok, (N, S, L, h, d, name, err)
# accumulate / out_scale / kv_broadcast (the IP-Adapter forms)
N, S, L, h, d = 5, 300, 16, 4, 64
q = torch.randn(N, S, h * d, device="cuda").half(); k1 = torch.randn(1, L, h * d, device="cuda").half(); v1 = torch.randn_like(k1)
base = torch.randn(N, S, h * d, device="cuda").half(); fs = torch.linspace(0.1, 1, N, device="cuda")
outs = []
for fl in (GEN, OLD, 0):
    o = base.clone()
    _cabi.attn_core(q, k1, v1, None, h, _cabi.PAID_PLAIN, False, flags=fl, out=o, accumulate=True, out_scale=0.7, out_frame_scale=fs, kv_broadcast=True)
    outs.append(o.float())
torch.cuda.synchronize()
e = float((outs[2] - outs[0]).pow(2).mean().sqrt() / outs[0].pow(2).mean().sqrt())
print(json.dumps(dict(accumulate_broadcast_err=e))); assert e < 1e-3
print(json.dumps({"dw_vs_generic_worst_rel_rms": worst}), flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=9, warm=3):
    ts = []
    for i in range(warm + iters):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= warm: ts.append(a.elapsed_time(b))
    return statistics.median(ts)
N = 7
coef = torch.linspace(0, 1, N, device="cuda")
for S, L, h in ((4096, 4096, 10), (1024, 1024, 20), (1024, 77, 20), (4096, 77, 10)):
    q, k, v = (torch.randn(N, T, h * 64, device="cuda").half() for T in (S, L, L))
    for name, mode, fused, mult in (("plain", _cabi.PAID_PLAIN, False, 2), ("inner_fused", _cabi.PAID_INNER, True, 4),
                                    ("outer_fused", _cabi.PAID_OUTER, True, 6)):   # OUTER runs attn_tc.cu under both flags
        t0 = timeit(lambda: _cabi.attn_core(q, k, v, coef, h, mode, fused, flags=OLD))
        t1 = timeit(lambda: _cabi.attn_core(q, k, v, coef, h, mode, fused))
        tf = mult * 2.0 * N * S * L * h * 64 / t1 / 1e9
        print(json.dumps(dict(S=S, L=L, heads=h, mode=name, one_wg_ms=round(t0, 4), dual_wg_ms=round(t1, 4), speedup=round(t0 / t1, 3), alg_tflops=round(tf, 1))), flush=True)
