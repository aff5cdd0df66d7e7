"""Feed-forward / projection GEMM shapes of the SDXL UNet (N = 7 frames) on libpaid_attn's tcgen05 GEMM vs cuBLAS (torch):
CUDA events, L2 flushed.  python tools/bench_gemm.py"""
import json, os, statistics, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=9, warm=3):
    ts = []
    for i in range(warm + iters):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= warm: ts.append(a.elapsed_time(b))
    return statistics.median(ts)
for name, M, K, D in (("ff.proj+geglu 32x32", 7168, 1280, 5120), ("ff.proj+geglu 64x64", 28672, 640, 2560)):
    x = torch.randn(M, K, device="cuda").half(); w = (torch.randn(2 * D, K, device="cuda") / K ** 0.5).half(); b = torch.randn(2 * D, device="cuda").half()
    fl = 2.0 * M * 2 * D * K
    t_own = timeit(lambda: _cabi.linear_geglu(x, w, b))
    t_cublas = timeit(lambda: F.linear(x, w, b))
    t_cublas_geglu = timeit(lambda: _cabi.geglu(F.linear(x, w, b)))
    print(json.dumps(dict(op=name, M=M, K=K, D=D, own_fused_ms=round(t_own, 4), own_tflops=round(fl / t_own / 1e9, 1), cublas_gemm_ms=round(t_cublas, 4),
                          cublas_tflops=round(fl / t_cublas / 1e9, 1), cublas_plus_geglu_kernel_ms=round(t_cublas_geglu, 4))), flush=True)
for name, M, K, N in (("ff.out 32x32", 7168, 5120, 1280), ("ff.out 64x64", 28672, 2560, 640), ("proj_in/out 32x32", 7168, 1280, 1280),
                      ("proj_in/out 64x64", 28672, 640, 640), ("to_q cross 32x32", 7168, 1280, 1280), ("to_k/v cross", 539, 2048, 1280)):
    x = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).half(); b = torch.randn(N, device="cuda").half()
    fl = 2.0 * M * N * K
    t_own = timeit(lambda: _cabi.linear(x, w, b)); t_cublas = timeit(lambda: F.linear(x, w, b))
    print(json.dumps(dict(op=name, M=M, K=K, N=N, own_ms=round(t_own, 4), own_tflops=round(fl / t_own / 1e9, 1), cublas_ms=round(t_cublas, 4),
                          cublas_tflops=round(fl / t_cublas / 1e9, 1))), flush=True)
