"""GPU kernel time of one plain UNet forward (N=7), grouped by kernel, harness variants: channels_last on/off."""
import os, sys, time, collections, torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200.pipeline import InterpolationPipeline
from attention_interpolation_diffusion_b200.unet_harness import build_unet, UNetHarness, CONFIGS
torch.backends.cudnn.benchmark = True
N = 7
for cl in (True,):
    torch.manual_seed(0)
    with torch.device("cuda"):
        net = UNetHarness(CONFIGS["sdxl"])
    net = net.half().eval().requires_grad_(False)
    if cl:
        net = net.to(memory_format=torch.channels_last)
    pipe = InterpolationPipeline(net, use_cuda_graphs=False)
    pipe.deactivate_aid()
    lat = torch.randn(N, 4, 128, 128, device="cuda").half()
    if cl: lat = lat.contiguous(memory_format=torch.channels_last)
    ctx = torch.randn(N, 77, 2048, device="cuda").half()
    added = {"text_embeds": torch.randn(N, 1280, device="cuda").half(), "time_ids": torch.zeros(N, 6, device="cuda").half()}
    with torch.no_grad():
        for _ in range(3): net(lat, 500, ctx, added)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): net(lat, 500, ctx, added)
        torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 5
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            net(lat, 500, ctx, added); torch.cuda.synchronize()
    agg = collections.Counter(); cnt = collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            import re
            nm = e.name
            m = re.search(r"(GeluCUDAKernelImpl|direct_copy_kernel|silu|MulFunctor|AddFunctor|CUDAFunctor_add|binary_internal::\w+|LayerNorm\w*|GroupNorm\w*|RowwiseMoments\w*|ComputeFusedParams\w*|CatArrayBatchedCopy\w*|upsample\w*|cos|sin|exp)", nm)
            base = "vec" if "vectorized" in nm else ("elt" if "elementwise_kernel" in nm else "")
            name = (base + ":" + m.group(1)) if m else nm.split("<")[0].split("(")[0][-60:]
            agg[name] += e.device_time; cnt[name] += 1
    tot = sum(agg.values())
    print(f"=== channels_last={cl}: wall {wall*1e3:.1f} ms/forward, kernel sum {tot/1e3:.1f} ms")
    for k, v in agg.most_common(24):
        print(f"  {v/1e3:7.2f} ms x{cnt[k]:4d} {k}")
    del net, pipe
    torch.cuda.empty_cache()
