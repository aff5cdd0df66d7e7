"""GPU kernel time of one SDXL UNet forward (N=7) inside the real harness, grouped by kernel (torch profiler, so caches
are as warm as in the step loop): plain forward and AID forward."""
import collections, os, re, sys, time, torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200.pipeline import InterpolationPipeline
from attention_interpolation_diffusion_b200.unet_harness import build_unet
torch.backends.cudnn.benchmark = True
N = 7
net = build_unet("sdxl", "cuda", torch.float16, seed=0)
pipe = InterpolationPipeline(net, use_cuda_graphs=False)
pipe.load_aid(t=None, is_fused=True, atype="fused_outer", size=N, alpha=4, beta=4)
lat = torch.randn(N, 4, 128, 128, device="cuda").half().contiguous(memory_format=torch.channels_last)
ctx = torch.randn(N, 77, 2048, device="cuda").half()
added = {"text_embeds": torch.randn(N, 1280, device="cuda").half(), "time_ids": torch.zeros(N, 6, device="cuda").half()}
PAID = r"(attn_dw_kernel|attn_tc_kernel|linear_geglu_generic_kernel|linear_tc_pair_kernel|linear_tc_kernel|add_layer_norm_kernel|gn_stats_kernel|gn_apply_kernel|residual_bias_add_kernel|geglu_kernel|lerp_endpoints_kernel|attn_generic_kernel|linear_generic_kernel)"
AT = r"(GeluCUDAKernelImpl|direct_copy_kernel|silu|MulFunctor|AddFunctor|CUDAFunctor_add|LayerNorm\w*|layer_norm\w*|GroupNorm\w*|RowwiseMoments\w*|CatArrayBatchedCopy\w*|upsample\w*)"
for mode in ("plain", "aid"):
    pipe.deactivate_aid() if mode == "plain" else pipe.set_coefs(torch.linspace(0, 1, N))
    with torch.no_grad():
        for _ in range(3): net(lat, 500, ctx, added)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): net(lat, 500, ctx, added)
        torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 5
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            net(lat, 500, ctx, added); torch.cuda.synchronize()
    agg, cnt = collections.Counter(), collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            nm = e.name
            m = re.search(PAID, nm)
            if m: name = "paid::" + m.group(1)
            elif "nvjet" in nm: name = "cuBLAS nvjet GEMM"
            elif re.search(r"cutlass|cudnn|ndhwc|nhwc", nm): name = "cuDNN conv"
            else:
                m = re.search(AT, nm)
                name = ("at::" + m.group(1)) if m else nm.split("<")[0].split("(")[0][-50:]
            agg[name] += e.device_time; cnt[name] += 1
    tot = sum(agg.values())
    print(f"=== {mode} forward, SDXL N={N}: wall {wall*1e3:.1f} ms (eager launches), kernel sum {tot/1e3:.1f} ms")
    for k, v in agg.most_common(20):
        print(f"  {v/1e3:7.2f} ms {100*v/tot:5.1f}%  x{cnt[k]:4d}  avg {v/cnt[k]:7.1f} us  {k}")
