"""Where does one UNet forward spend its time?  torch.profiler kernel table for one AID forward and one
plain forward of the SDXL harness (N frames), plus wall vs summed-kernel time (launch-bound check)."""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200.pipeline import InterpolationPipeline  # noqa: E402
from attention_interpolation_diffusion_b200.prior import generate_beta_tensor  # noqa: E402
from attention_interpolation_diffusion_b200.unet_harness import build_unet  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "sdxl"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 7
torch.backends.cudnn.benchmark = True
net = build_unet(model, "cuda", torch.float16)
pipe = InterpolationPipeline(net)
pipe.load_aid(t=None, is_fused=True, size=N, alpha=4, beta=4)
cfg = net.cfg
lat = torch.randn(N, 4, cfg.sample_size, cfg.sample_size, device="cuda").half().contiguous(memory_format=torch.channels_last)
ctx = torch.randn(N, 77, cfg.cross_attention_dim, device="cuda").half()
added = {"text_embeds": torch.randn(N, 1280, device="cuda").half(), "time_ids": torch.zeros(N, 6, device="cuda").half()} if cfg.text_time else None
coef = generate_beta_tensor(N, 4, 4)

def fwd(aid):
    if aid:
        pipe.set_coefs(coef)
    else:
        pipe.deactivate_aid()
    with torch.no_grad():
        return net(lat, 500, ctx, added)

for aid in (True, False):
    for _ in range(3):
        fwd(aid)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fwd(aid)
    t_issue = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    t_wall = (time.perf_counter() - t0) / 5
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fwd(aid)
        torch.cuda.synchronize()
    ev = prof.key_averages()
    tot = sum(e.device_time_total for e in ev) / 1e3
    print(f"=== {'AID' if aid else 'plain'} forward N={N}: wall {t_wall*1e3:.1f} ms, cpu issue {t_issue*1e3:.1f} ms, "
          f"sum of kernel time {tot:.1f} ms")
    rows = sorted(ev, key=lambda e: -e.device_time_total)[:22]
    for e in rows:
        if e.device_time_total > 0:
            print(f"{e.device_time_total/1e3:9.2f} ms  x{e.count:5d}  {e.key[:110]}")
