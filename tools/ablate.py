"""Time the attention core for each experimental library variant (gpurun_scratch/libpaid_*.so)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch, json
sys.path.insert(0, %r)
sys.path.insert(0, %r + "/tools")
from attention_interpolation_diffusion_b200 import _cabi
from bench_kernels import timeit
res = {}
for (N, S, L, h, d) in ((7, 4096, 4096, 10, 64), (7, 1024, 1024, 20, 64)):
    q = torch.randn(N, S, h * d, device="cuda").half(); k = torch.randn(N, L, h * d, device="cuda").half(); v = torch.randn(N, L, h * d, device="cuda").half()
    coef = torch.linspace(0, 1, N, device="cuda")
    for mode, fused, name in ((0, False, "plain"), (1, True, "outer")):
        res[f"{S}_{name}"] = round(timeit(lambda: _cabi.attn_core(q, k, v, coef, h, mode, fused), iters=7, warm=2), 4)
print(json.dumps(res))
''' % (ROOT, ROOT)
for lib in sorted(glob.glob(os.path.join(ROOT, "gpurun_scratch", "libpaid_*.so"))):
    env = dict(os.environ, PAID_LIB_PATH=lib)
    try:
        r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=120)
        print(os.path.basename(lib), (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1])
    except subprocess.TimeoutExpired:
        print(os.path.basename(lib), "timeout")
