"""torchrun --nproc-per-node R tools/check_sharded.py : the frame-sharded denoise (NCCL broadcast of the endpoint
K/V) against the single-GPU run of the same sequence, computed on rank 0."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from attention_interpolation_diffusion_b200.pipeline import InterpolationPipeline
from attention_interpolation_diffusion_b200.sharding import FrameShard
from attention_interpolation_diffusion_b200.unet_harness import build_unet
import paid_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
model = sys.argv[1] if len(sys.argv) > 1 else "tiny"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 7
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
net = build_unet(model, f"cuda:{local}", torch.float16, seed=1002)
cfg = net.cfg
g = torch.Generator("cpu").manual_seed(1002)
r = lambda *s: torch.randn(*s, generator=g).cuda().half()
side, cc = cfg.sample_size, cfg.cross_attention_dim
args = dict(latent_start=r(1, 4, side, side), latent_end=r(1, 4, side, side), embeds_start=r(1, 77, cc), embeds_end=r(1, 77, cc),
            negative_embeds=r(1, 77, cc), guide_embeds=r(1, 77, cc), pooled_start=r(1, 1280), pooled_end=r(1, 1280),
            pooled_negative=r(1, 1280), pooled_guide=r(1, 1280), size=frames, num_inference_steps=steps)
shard = FrameShard(rank, world, frames)
local_out = InterpolationPipeline(net, shard=shard).interpolate(**args)
outs = [torch.empty(hi - lo, *local_out.shape[1:], dtype=local_out.dtype, device=local_out.device) for lo, hi in shard.shards]
# gather variable-size shards with point-to-point broadcasts
for rk, (lo, hi) in enumerate(shard.shards):
    if hi > lo:
        buf = local_out.contiguous() if rk == rank else outs[rk]
        dist.broadcast(buf, src=rk)
        outs[rk] = buf
if rank == 0:
    full = InterpolationPipeline(net, use_cuda_graphs=False).interpolate(**args)
    sharded = torch.cat(outs, dim=0)
    ok, m = O.within_tolerance(sharded.float().cpu(), full.float().cpu(), 2e-3, 5e-2)
    print("SHARDED_VS_SINGLE", model, "frames", frames, "world", world, "ok" if ok else "MISMATCH", m, flush=True)
dist.barrier()
dist.destroy_process_group()
