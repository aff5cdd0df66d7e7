"""Hot-path time of one SDXL UNet forward (N=7): paid_attn_forward per attention-layer class x layer count."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from attention_interpolation_diffusion_b200 import _cabi
from bench_kernels import timeit
N = int(sys.argv[1]) if len(sys.argv) > 1 else 7
coef = torch.linspace(0, 1, N, device="cuda")
classes = [(4096, 640, 10, None, 640, 10), (4096, 640, 10, 77, 2048, 10), (1024, 1280, 20, None, 1280, 60), (1024, 1280, 20, 77, 2048, 60)]
tot = {"outer": 0.0, "plain": 0.0}
for (S, C, h, L, Cc, count) in classes:
    x = torch.randn(N, S, C, device="cuda").half()
    ctx = None if L is None else torch.randn(N, L, Cc, device="cuda").half()
    w = [torch.randn(C, C, device="cuda").half() / C ** 0.5, torch.randn(C, Cc, device="cuda").half() / Cc ** 0.5,
         torch.randn(C, Cc, device="cuda").half() / Cc ** 0.5, torch.randn(C, C, device="cuda").half() / C ** 0.5,
         torch.randn(C, device="cuda").half()]
    row = dict(S=S, L=L or S, C=C, count=count)
    for name, mode, fused in (("outer", _cabi.PAID_OUTER, True), ("plain", _cabi.PAID_PLAIN, False)):
        ms = timeit(lambda: _cabi.attn_forward(x, ctx, *w, coef, h, mode, fused), iters=7, warm=2)
        row[name + "_ms"] = round(ms, 4)
        tot[name] += ms * count
    q = torch.randn(N, S, C, device="cuda").half(); k = torch.randn(N, L or S, C, device="cuda").half(); v = torch.randn_like(k)
    row["core_outer_ms"] = round(timeit(lambda: _cabi.attn_core(q, k, v, coef, h, _cabi.PAID_OUTER, True), iters=7, warm=2), 4)
    row["core_plain_ms"] = round(timeit(lambda: _cabi.attn_core(q, k, v, coef, h, _cabi.PAID_PLAIN, False), iters=7, warm=2), 4)
    print(json.dumps(row))
print(json.dumps({"hot_path_ms_per_AID_forward": round(tot["outer"], 2), "hot_path_ms_per_plain_forward": round(tot["plain"], 2),
                  "per_sequence_s(25 AID + 75 plain)": round((25 * tot["outer"] + 75 * tot["plain"]) / 1000, 3)}))
