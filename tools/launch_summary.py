"""Group an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel family:
    python tools/launch_summary.py gpurun_out/x_launches.csv profiles/out.txt "header line"."""
import collections
import csv
import re
import sys

FAMILIES = [("paid::attn_tc_kernel (outer)", r"paid::.*attn_tc_kernel"), ("paid::attn_dw_kernel (plain / inner)", r"paid::.*attn_dw_kernel"),
            ("paid::linear_tc_pair_kernel, GEGLU epilogue (cta_group::2)", r"paid::.*linear_tc_pair_kernel<[^>,]*, *(\(bool\))?(1|true)>"),
            ("paid::linear_tc_pair_kernel (cta_group::2)", r"paid::.*linear_tc_pair"),
            ("paid::linear_tc_kernel", r"paid::.*linear_tc"), ("paid::geglu_kernel", r"paid::.*geglu"),
            ("paid::group_norm kernels", r"paid::.*gn_"), ("paid::add_layer_norm_kernel", r"paid::.*layer_norm"),
            ("paid:: other", r"paid::"), ("cuBLAS nvjet GEMM (FF, proj_in/out, embeddings)", r"nvjet"),
            ("cuDNN conv", r"cutlass3x|cudnn|implicit_gemm|conv"), ("at::direct_copy_kernel", r"direct_copy"),
            ("at::CUDAFunctor_add", r"CUDAFunctor_add"), ("at::RowwiseMomentsCUDAKernel", r"RowwiseMoments"),
            ("at::GroupNormKernelImplInternal", r"GroupNormKernelImpl"), ("at::ComputeFusedParams", r"ComputeFusedParams"),
            ("at::layer_norm", r"layer_norm|LayerNorm"), ("at::silu", r"silu"), ("at::Cat", r"CatArray"),
            ("at::upsample", r"upsample"), ("at::other elementwise", r"elementwise|vectorized")]


def main():
    path, out, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    first = next((i for i, r in enumerate(rows[1:], 1) if "paid::" in r[ik]), 1)   # drop the weight-initialisation kernels in front
    rows = [hdr] + rows[first:]
    t, n = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[iu], 1e-6)
        fam = next((f for f, pat in FAMILIES if re.search(pat, r[ik])), r[ik][:40])
        t[fam] += v
        n[fam] += 1
    tot = sum(t.values())
    lines = [f"# {header}", f"# {len(rows) - 1} launches, total kernel time {tot:.2f} ms; per-launch times are cold-cache and "
             "serialised: compare SHARES, not absolutes"]
    lines += [f"{v:9.3f} ms {100 * v / tot:5.1f}%  x{n[k]:5d}  {k}" for k, v in t.most_common()]
    text = "\n".join(lines) + "\n"
    print(text)
    if out:
        open(out, "w").write(text)


if __name__ == "__main__":
    main()
