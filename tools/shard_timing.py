"""torchrun --nproc-per-node R tools/shard_timing.py : per-rank time of one plain and one AID UNet forward (captured graphs) of the
frame-sharded SDXL sequence (7 frames per rank), and of the same forwards with the collectives removed from the picture
(world-size-1 shard of 7 frames).  One JSON line per rank."""
import json, os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attention_interpolation_diffusion_b200.pipeline import InterpolationPipeline
from attention_interpolation_diffusion_b200.sharding import FrameShard
from attention_interpolation_diffusion_b200.unet_harness import build_unet
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True
net = build_unet("sdxl", dev, torch.float16, seed=1002)
frames = 7 * world
shard = FrameShard(rank, world, frames) if world > 1 else None
pipe = InterpolationPipeline(net, shard=shard)
pipe.load_aid(t=None, is_fused=True, atype="fused_outer", size=frames, alpha=4, beta=4)
g = torch.Generator("cpu").manual_seed(1002)
r = lambda *s: torch.randn(*s, generator=g).to(dev).half()
args = dict(latent_start=r(1, 4, 128, 128), latent_end=r(1, 4, 128, 128), embeds_start=r(1, 77, 2048), embeds_end=r(1, 77, 2048),
            negative_embeds=r(1, 77, 2048), guide_embeds=r(1, 77, 2048), pooled_start=r(1, 1280), pooled_end=r(1, 1280),
            pooled_negative=r(1, 1280), pooled_guide=r(1, 1280), size=frames, alpha=4.0, beta=4.0)
pipe.interpolate(**args, num_inference_steps=4)            # captures the three forwards
n = 7
lat = torch.randn(n, 4, 128, 128, device=dev).half().contiguous(memory_format=torch.channels_last)
def time_forward(aid, tag, reps=20):
    key = next(k for k in pipe._graphs if k[0] == aid and k[1] == tag)
    gf = pipe._graphs[key]
    pipe._kv_tag[0] = tag
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): gf.graph.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = dict(rank=rank, world=world, frames_local=n, frame_ids=shard.frame_ids if shard else list(range(7)),
           aid_forward_ms=round(time_forward(True, "cond"), 3), plain_cond_ms=round(time_forward(False, "cond"), 3),
           plain_uncond_ms=round(time_forward(False, "uncond"), 3), overlap=os.environ.get("PAID_SHARD_OVERLAP", "1"))
print(json.dumps(out), flush=True)
if world > 1:
    pipe._graphs.clear(); torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
os._exit(0)
