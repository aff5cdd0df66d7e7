"""One launch of libpaid_attn's tcgen05 GEMM per feed-forward / projection shape of the SDXL UNet (N = 7) for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:linear_tc -c 7 -o gpurun_out/gemm python tools/ncu_gemm.py
Launch order = SHAPES order."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200 import _cabi
SHAPES = [("ff.proj+geglu 32x32", 7168, 1280, 5120, True), ("ff.proj+geglu 64x64", 28672, 640, 2560, True), ("ff.out 32x32", 7168, 5120, 1280, False),
          ("ff.out 64x64", 28672, 2560, 640, False), ("out-proj / to_q 32x32", 7168, 1280, 1280, False), ("out-proj / to_q 64x64", 28672, 640, 640, False),
          ("qkv grouped 32x32", 7168, 1280, 1280, "qkv")]
if __name__ == "__main__":
    for name, M, K, N, kind in SHAPES:
        x = torch.randn(M, K, device="cuda").half()
        if kind == "qkv":      # the grouped q/k/v launch of a self-attention layer goes through paid_attn_forward; time it via a plain call
            w = [torch.randn(N, K, device="cuda").half() / K ** 0.5 for _ in range(4)]
            b = torch.randn(N, device="cuda").half()
            S = 1024
            _cabi.attn_forward(x.view(7, S, K), None, w[0], w[1], w[2], w[3], b, None, 20, _cabi.PAID_PLAIN, False)
        elif kind:
            w = (torch.randn(2 * N, K, device="cuda") / K ** 0.5).half(); b = torch.randn(2 * N, device="cuda").half()
            _cabi.linear_geglu(x, w, b)
        else:
            w = (torch.randn(N, K, device="cuda") / K ** 0.5).half(); b = torch.randn(N, device="cuda").half()
            _cabi.linear(x, w, b)
    torch.cuda.synchronize()
