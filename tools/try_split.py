"""Experimental split-row softmax (PAID_ATTN_SPLIT=1) against the default attention kernel: agreement on ragged and SDXL
shapes in every mode, then CUDA-event timing (L2 flushed) of both on the SDXL self-attention shapes."""
import json, os, statistics, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
MODES = [(_cabi.PAID_OUTER, True), (_cabi.PAID_OUTER, False), (_cabi.PAID_INNER, True), (_cabi.PAID_INNER, False), (_cabi.PAID_PLAIN, False)]
def run(split, *a, **k):
    if split: os.environ["PAID_ATTN_SPLIT"] = "1"
    else: os.environ.pop("PAID_ATTN_SPLIT", None)
    return _cabi.attn_core(*a, **k)
torch.manual_seed(0)
worst = 0.0
for N, S, L, h, d in ((5, 700, 333, 3, 64), (3, 130, 77, 2, 64), (4, 257, 64, 1, 64), (3, 64, 5, 2, 40), (7, 1024, 1024, 20, 64), (4, 300, 640, 2, 64)):
    q, k, v = (torch.randn(N, T, h * d, device="cuda").half() for T in (S, L, L))
    if L == 640: k = k * torch.linspace(0.2, 6.0, L, device="cuda").view(1, L, 1).half()     # growing logits: rescale path
    coef = torch.linspace(0, 1, N, device="cuda")
    for mode, fused in MODES:
        ref = run(False, q, k, v, coef, h, mode, fused).float()
        out = run(True, q, k, v, coef, h, mode, fused).float()
        torch.cuda.synchronize()
        err = float((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
        worst = max(worst, err)
        again = run(True, q, k, v, coef, h, mode, fused).float()
        assert torch.isfinite(out).all() and err < 1e-3 and torch.equal(out, again), (N, S, L, h, d, mode, fused, err)
print(json.dumps({"split_vs_default_worst_rel_rms": worst}), flush=True)
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
    for name, mode, fused in (("fused_outer", _cabi.PAID_OUTER, True), ("plain", _cabi.PAID_PLAIN, False)):
        t0 = timeit(lambda: run(False, q, k, v, coef, h, mode, fused))
        t1 = timeit(lambda: run(True, q, k, v, coef, h, mode, fused))
        print(json.dumps(dict(S=S, L=L, heads=h, mode=name, default_ms=round(t0, 4), split_ms=round(t1, 4), speedup=round(t0 / t1, 3))), flush=True)
