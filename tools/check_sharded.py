"""torchrun --nproc-per-node R tools/check_sharded.py [--frames N] [--model tiny|sdxl] : the frame-sharded denoise (one NCCL
broadcast of the endpoint K/V per self-attention layer, side stream, captured in CUDA graphs) against the single-GPU run of the
same sequence, for the text processors (outer / inner) and the three IP-Adapter processors.  One JSON line per case from rank 0.
Also the body of tests/test_multirank_gpu.py."""
import argparse, json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import paid_oracle as O
from attention_interpolation_diffusion_b200.pipeline import InterpolationPipeline
from attention_interpolation_diffusion_b200.sharding import FrameShard
from attention_interpolation_diffusion_b200.unet_harness import build_unet

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=9)
ap.add_argument("--model", default="tiny")
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
net = build_unet(a.model, dev, torch.float16, seed=7)
cfg = net.cfg
g = torch.Generator("cpu").manual_seed(1002)
r = lambda *s: torch.randn(*s, generator=g).to(dev).half()
side, cc, T = cfg.sample_size, cfg.cross_attention_dim, 4
base = dict(latent_start=r(1, 4, side, side), latent_end=r(1, 4, side, side), embeds_start=r(1, 77, cc), embeds_end=r(1, 77, cc),
            negative_embeds=r(1, 77, cc), size=a.frames, num_inference_steps=a.steps, alpha=3.0, beta=3.0)
if cfg.text_time:
    base.update(pooled_start=r(1, 1280), pooled_end=r(1, 1280), pooled_negative=r(1, 1280))
guide = dict(guide_embeds=r(1, 77, cc), **({"pooled_guide": r(1, 1280)} if cfg.text_time else {}))
ip = dict(ip_start=r(1, T, cc), ip_end=r(1, T, cc))
CASES = [("text fused_outer (PAID guide prompt)", "load_aid", dict(atype="fused_outer"), {**base, **guide}),
         ("text fused_inner", "load_aid", dict(atype="fused_inner"), base),
         ("ip fused_outer", "load_aid_ip_adapter", dict(early="fused_outer", num_tokens=T, scale=0.7), {**base, **ip}),
         ("ip fused_inner", "load_aid_ip_adapter", dict(early="fused_inner", num_tokens=T, scale=0.7), {**base, **ip}),
         ("ip scale_control", "load_aid_ip_adapter", dict(early="scale_control", num_tokens=T, scale=0.7), {**base, **ip})]
ok_all = True
for name, loader, lkw, args in CASES:
    for graphs in (False, True):
        outs = {}
        for sharded in (False, True):
            shard = FrameShard(rank, world, a.frames) if sharded else None
            pipe = InterpolationPipeline(net, shard=shard, use_cuda_graphs=graphs)
            torch.manual_seed(0)                       # same random-init IP-Adapter weights in every variant / rank
            getattr(pipe, loader)(t=None, is_fused=True, size=a.frames, alpha=3, beta=3, **lkw)
            out = pipe.interpolate(**args)
            if sharded:
                parts = [torch.empty(len(ids), *out.shape[1:], dtype=out.dtype, device=dev) for ids in shard.shards]
                for rk in range(world):                # variable-size shards: one broadcast per rank
                    if rk == rank:
                        parts[rk].copy_(out)
                    dist.broadcast(parts[rk], src=rk)
                outs["sharded"] = shard.unshard(parts).float().cpu()
                outs["broadcasts"] = shard.broadcasts
            else:
                outs["single"] = out.float().cpu()
            del pipe
        ok, m = O.within_tolerance(outs["sharded"], outs["single"], 2e-3, 5e-2)
        bit = bool(torch.equal(outs["sharded"], outs["single"]))
        ok_all = ok_all and ok and bool(torch.isfinite(outs["sharded"]).all())
        if rank == 0:
            print(json.dumps(dict(case=name, graphs=graphs, world=world, frames=a.frames, model=a.model, ok=bool(ok), bit_identical=bit,
                                  rel_rms=m["rel_rms"], max_abs_over_rms=m.get("max_abs_over_rms"), broadcasts=outs["broadcasts"])), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
