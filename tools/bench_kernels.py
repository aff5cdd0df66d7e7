"""Kernel-level timing on the GPU box (CUDA events, L2 flushed between iterations).
usage: python tools/bench_kernels.py [linear] [core] [layer]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi  # noqa: E402

FLUSH = None


def timeit(fn, iters=10, warm=3):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def bench_linear():
    for M, N, K in ((7 * 1024, 1280, 1280), (7 * 4096, 640, 640), (7 * 77, 1280, 2048), (7 * 1024, 1280, 5120)):
        x = torch.randn(M, K, device="cuda").half()
        w = torch.randn(N, K, device="cuda").half()
        b = torch.randn(N, device="cuda").half()
        fl = 2.0 * M * N * K
        for name, fn in (("tc", lambda: _cabi.linear(x, w, b)), ("generic", lambda: _cabi.linear(x, w, b, flags=1)),
                         ("torch", lambda: torch.nn.functional.linear(x, w, b))):
            ms = timeit(fn)
            print(json.dumps(dict(op="linear", impl=name, M=M, N=N, K=K, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))))


def bench_core():
    for (N, S, L, h, d) in ((7, 1024, 1024, 20, 64), (7, 4096, 4096, 10, 64), (7, 1024, 77, 20, 64)):
        q = torch.randn(N, S, h * d, device="cuda").half()
        k = torch.randn(N, L, h * d, device="cuda").half()
        v = torch.randn(N, L, h * d, device="cuda").half()
        coef = torch.linspace(0, 1, N, device="cuda")
        A = 2.0 * N * S * L * h * d
        for mode, fused, mult in ((_cabi.PAID_PLAIN, False, 2), (_cabi.PAID_OUTER, True, 6), (_cabi.PAID_INNER, True, 4)):
            for name, flags in (("default", 0), ("generic", 1)):
                ms = timeit(lambda: _cabi.attn_core(q, k, v, coef, h, mode, fused, flags=flags), iters=5, warm=2)
                print(json.dumps(dict(op="core", impl=name, kernel=_cabi.last_kernel(), N=N, S=S, L=L, h=h, d=d, mode=mode,
                                      fused=fused, ms=round(ms, 4), alg_tflops=round(mult * A / ms / 1e9, 1))))
        qq, kk, vv = (t.view(N, -1, h, d).transpose(1, 2) for t in (q, k, v))
        ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv), iters=5, warm=2)
        print(json.dumps(dict(op="core", impl="torch_sdpa_plain", N=N, S=S, L=L, ms=round(ms, 4), alg_tflops=round(2 * A / ms / 1e9, 1))))


if __name__ == "__main__":
    what = sys.argv[1:] or ["linear", "core"]
    if "linear" in what:
        bench_linear()
    if "core" in what:
        bench_core()
