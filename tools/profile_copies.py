import os, sys, torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200.pipeline import InterpolationPipeline
from attention_interpolation_diffusion_b200.unet_harness import build_unet
torch.backends.cudnn.benchmark = True
N = 7
net = build_unet("sdxl", "cuda", torch.float16)
pipe = InterpolationPipeline(net, use_cuda_graphs=False); pipe.deactivate_aid()
lat = torch.randn(N, 4, 128, 128, device="cuda").half().contiguous(memory_format=torch.channels_last)
ctx = torch.randn(N, 77, 2048, device="cuda").half()
added = {"text_embeds": torch.randn(N, 1280, device="cuda").half(), "time_ids": torch.zeros(N, 6, device="cuda").half()}
with torch.no_grad():
    for _ in range(3): net(lat, 500, ctx, added)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
        net(lat, 500, ctx, added); torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=4) if e.key in ("aten::copy_", "aten::add", "aten::add_", "aten::native_group_norm", "aten::native_layer_norm", "aten::contiguous", "aten::clone")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:14]:
    st = [s for s in e.stack if "attention_interpolation" in s or "torch/nn/functional" in s][:2]
    print(f"{e.device_time_total/1e3:7.2f} ms x{e.count:4d} {e.key:26s} {str(e.input_shapes)[:70]:70s} {[s.split('/')[-1][:60] for s in st]}")
