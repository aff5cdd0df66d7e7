"""Bring-up ladder for the tcgen05 attention kernel: compares it with the generic kernel on the same
inputs, one case per subprocess (a device trap poisons the CUDA context)."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    # N, S, L, heads, mode, fused, note
    (1, 128, 64, 1, 0, 0, "plain one step"),
    (1, 256, 64, 1, 0, 0, "plain two q tiles"),
    (1, 256, 128, 1, 0, 0, "plain two steps"),
    (1, 256, 1024, 2, 0, 0, "plain 16 steps 2 heads"),
    (3, 256, 77, 2, 0, 0, "plain ragged L"),
    (3, 200, 300, 2, 0, 0, "plain ragged S and L"),
    (3, 256, 256, 2, 1, 0, "outer pure"),
    (3, 256, 256, 2, 1, 1, "outer fused"),
    (5, 512, 320, 3, 1, 1, "outer fused N=5"),
    (5, 512, 320, 3, 2, 0, "inner pure"),
    (5, 512, 320, 3, 2, 1, "inner fused"),
    (7, 1024, 1024, 20, 1, 1, "sdxl 32x32 outer fused"),
    (4, 1024, 77, 20, 1, 1, "sdxl cross outer fused"),
]


def run_case(idx):
    import torch
    from attention_interpolation_diffusion_b200 import _cabi
    N, S, L, h, mode, fused, note = CASES[idx]
    d = 64
    torch.manual_seed(idx)
    q = torch.randn(N, S, h * d, device="cuda").half()
    k = torch.randn(N, L, h * d, device="cuda").half()
    v = torch.randn(N, L, h * d, device="cuda").half()
    coef = torch.linspace(0, 1, N, device="cuda") if N > 1 else torch.zeros(1, device="cuda")
    ref = _cabi.attn_core(q, k, v, coef, h, mode, fused, flags=1).float()
    out = _cabi.attn_core(q, k, v, coef, h, mode, fused).float()
    kern = _cabi.last_kernel()
    torch.cuda.synchronize()
    err = (out - ref)
    rms = ref.pow(2).mean().sqrt().item()
    res = dict(case=idx, note=note, kernel=kern, rel_rms=err.pow(2).mean().sqrt().item() / rms,
               max_abs=err.abs().max().item(), ref_rms=rms, nan=int(torch.isnan(out).sum()))
    if res["rel_rms"] > 2e-3 or res["nan"]:
        # where is it wrong?  per-frame, per-head, per 32-row block, per 8-column block error
        e = err.abs().view(N, S, h, d)
        res["by_frame"] = [round(x, 4) for x in e.amax(dim=(1, 2, 3)).tolist()]
        res["by_head"] = [round(x, 4) for x in e.amax(dim=(0, 1, 3)).tolist()][:8]
        rows = e.amax(dim=(0, 2, 3))
        res["by_row32"] = [round(rows[i:i + 32].max().item(), 4) for i in range(0, min(S, 512), 32)]
        cols = e.amax(dim=(0, 1, 2))
        res["by_col8"] = [round(cols[i:i + 8].max().item(), 4) for i in range(0, d, 8)]
        res["sample_out"] = [round(x, 4) for x in out.view(N, S, h, d)[0, 0, 0, :8].tolist()]
        res["sample_ref"] = [round(x, 4) for x in ref.view(N, S, h, d)[0, 0, 0, :8].tolist()]
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(int(sys.argv[1]))
    else:
        for i in range(len(CASES)):
            try:
                r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=120)
                out = r.stdout.strip().splitlines()
                print(out[-1] if out else json.dumps(dict(case=i, rc=r.returncode, err=r.stderr[-400:])))
            except subprocess.TimeoutExpired:
                print(json.dumps(dict(case=i, error="timeout")))
