"""The AID surface of the reference pipelines, on the UNet harness.

Mirrors (names, argument meaning, control flow) the parts of
``InterpolationStableDiffusion(XL)Pipeline`` that drive the hot path:

* ``load_aid`` / ``activate_aid`` / ``deactivate_aid``  pipeline_interpolated_sdxl.py:1066-1136
* ``interpolate_single`` (3 frames ``[start, t, end]``)  pipeline_interpolated_sdxl.py:1693-2411
* ``interpolate`` (N frames in one batch)                gradio_src/pipeline_interpolated_stable_diffusion.py:163-304

Text encoders, VAE and checkpoints are out of scope and absent from
this image, so prompts are given as embeddings and the result is the final latents.  The
step loop keeps the reference's structure: per step one conditional UNet pass with AID
active for the first ``int(steps * warmup_ratio)`` steps (sdxl:2230-2248), one
unconditional pass with AID off (sdxl:2272-2293), classifier-free guidance
(sdxl:2296-2298) and a scheduler step (deterministic DDIM here; the reference defers to
the checkpoint's scheduler).
"""
from __future__ import annotations

import os

from typing import Optional

import torch

from .attention import PaidIPAdapterAttnProcessor
from .interpolation import (InnerInterpolatedAttnProcessor, InnerInterpolatedIPAttnProcessor,
                            OuterInterpolatedAttnProcessor, OuterInterpolatedIPAttnProcessor,
                            ScaleControlIPAttnProcessor)
from .prior import generate_beta_tensor
from .sharding import FrameShard
from .unet_harness import UNetHarness


class DDIMScheduler:
    """Deterministic DDIM (eta = 0), scaled-linear betas 0.00085..0.012, 1000 train steps, leading spacing."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n: int):
        ratio = self.num_train_timesteps // n
        self.timesteps = (torch.arange(n) * ratio).flip(0) + 1
        self.ratio = ratio

    def scale_model_input(self, x, t):
        return x

    def step(self, eps: torch.Tensor, t: int, x: torch.Tensor) -> torch.Tensor:
        a_t = float(self.alphas_cumprod[t])
        prev = t - self.ratio
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else 1.0
        x0 = (x.float() - (1 - a_t) ** 0.5 * eps.float()) / a_t ** 0.5
        return (a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps.float()).to(x.dtype)


def slerp(v0: torch.Tensor, v1: torch.Tensor, t: float, threshold: float = 0.9995) -> torch.Tensor:
    """Latent interpolation used to build the frames (semantics of reference interpolation.py:861-918:
    spherical over the last dim, linear where the directions are colinear or undefined)."""
    u0 = v0 / v0.norm(dim=-1, keepdim=True)
    u1 = v1 / v1.norm(dim=-1, keepdim=True)
    cosine = (u0 * u1).sum(-1, keepdim=True)
    linear = cosine.abs().isnan() | (cosine.abs() > threshold)
    omega = cosine.arccos()
    sph = (torch.sin(omega * (1 - t)) * v0 + torch.sin(omega * t) * v1) / torch.sin(omega)
    return torch.where(linear, torch.lerp(v0, v1, t), sph)


class _GraphedForward:
    """One UNet forward (fixed processor state and shapes) captured into a CUDA graph.  The ~2400 kernel launches
    of a forward otherwise cost ~100 ms of host time, more than the GPU needs to execute them."""

    def __init__(self, unet, sample, t, ctx, added):
        from . import _cabi
        dev = sample.device
        self.sample, self.ctx = sample.clone(), ctx.clone()
        self.t = torch.zeros(1, device=dev, dtype=torch.float32)
        self.added = None if added is None else {k: v.clone() for k, v in added.items()}
        self.t.fill_(float(t))
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):      # warm-up outside capture: cuDNN autotune, lazy init, coef upload, workspace
            for _ in range(2):
                unet(self.sample, self.t, self.ctx, self.added)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.workspace = _cabi._workspaces.get(dev)      # keep the scratch buffer the graph points into alive
        n0 = _cabi.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = unet(self.sample, self.t, self.ctx, self.added)
        self.launches = _cabi.launch_count() - n0       # libpaid_attn kernels inside one replay

    def __call__(self, sample, t, ctx, added):
        self.sample.copy_(sample)
        self.ctx.copy_(ctx)
        self.t.fill_(float(t))
        if added is not None:
            for k, v in added.items():
                self.added[k].copy_(v)
        self.graph.replay()
        return self.out


class InterpolationPipeline:
    def __init__(self, unet: UNetHarness, shard: Optional[FrameShard] = None, use_cuda_graphs: bool = True):
        self.unet = unet
        self.scheduler = DDIMScheduler()
        self.shard = shard
        self.use_cuda_graphs = use_cuda_graphs
        self._graphs: dict = {}
        self.graph_kernel_launches = 0     # libpaid_attn kernels executed through graph replays
        self.load_aid()

    # ---- processor install / toggle (sdxl:1066-1136) ---------------------------------------------
    def load_aid(self, t: Optional[float] = 0.5, is_fused: bool = True, atype: str = "fused_outer", size: int = 7,
                 alpha: float = 1, beta: float = 1):
        cls = {"fused_outer": OuterInterpolatedAttnProcessor, "fused_inner": InnerInterpolatedAttnProcessor}[atype]
        attn_procs = {}
        for name, old in self.unet.attn_processors.items():
            if not name.startswith("encoder"):
                original = getattr(old, "original_attn", None) or old
                if isinstance(original, (OuterInterpolatedAttnProcessor, InnerInterpolatedAttnProcessor)):
                    original = None
                proc = cls(t=t, size=size, is_fused=is_fused, alpha=alpha, beta=beta, original_attn=original)
                proc.shard = self.shard
                attn_procs[name] = proc
            else:
                attn_procs[name] = old
        self.unet.set_attn_processor(attn_procs)
        self._graphs.clear()               # captured forwards bake the processor objects in

    def load_aid_ip_adapter(self, num_tokens: int = 16, scale: float = 1.0, t: Optional[float] = 0.5,
                            is_fused: bool = True, early: str = "fused_outer", size: int = 7, alpha: float = 1,
                            beta: float = 1):
        """Mirror of load_aid_ip_adapter (sdxl:1089-1126).  The reference first calls diffusers' load_ip_adapter, which
        gives every cross-attention layer an IPAdapterAttnProcessor2_0 with to_k_ip / to_v_ip; without checkpoints
        those are random-init ``PaidIPAdapterAttnProcessor`` objects here.  Then every processor is wrapped."""
        cls = {"fused_outer": OuterInterpolatedIPAttnProcessor, "fused_inner": InnerInterpolatedIPAttnProcessor,
               "scale_control": ScaleControlIPAttnProcessor}[early]
        mods = self.unet._attention_modules()
        attn_procs = {}
        for name, m in mods.items():
            old = m.processor
            if name.endswith("attn2.processor"):
                old = PaidIPAdapterAttnProcessor(m.to_q.in_features, m.to_k.in_features, (num_tokens,), scale)
                old = old.to(device=m.to_q.weight.device, dtype=m.to_q.weight.dtype)
            elif isinstance(old, (OuterInterpolatedAttnProcessor, InnerInterpolatedAttnProcessor)):
                old = old.original_attn
            attn_procs[name] = cls(t=t, size=size, is_fused=is_fused, alpha=alpha, beta=beta, ip_attn=old)
            attn_procs[name].shard = self.shard
        self.unet.set_attn_processor(attn_procs)
        self._graphs.clear()

    def activate_aid(self, it: float):
        for name, proc in self.unet.attn_processors.items():
            if not name.startswith("encoder"):
                proc.activate(it)

    def deactivate_aid(self):
        for name, proc in self.unet.attn_processors.items():
            if not name.startswith("encoder"):
                proc.deactivate()

    def set_coefs(self, coef: torch.Tensor):
        """N-frame extension of activate_aid: one coefficient per frame."""
        for name, proc in self.unet.attn_processors.items():
            if not name.startswith("encoder"):
                proc.set_coefs(coef)

    # ---- the step loop ---------------------------------------------------------------------------
    @torch.no_grad()
    def _denoise(self, latents, cond, uncond, added_cond, added_uncond, coef, num_inference_steps, guidance_scale,
                 warmup_ratio):
        """latents (n,4,H,W); cond / uncond (n,77,Cc): the local frames.  Returns final latents."""
        self.scheduler.set_timesteps(num_inference_steps)
        warmup_steps = int(num_inference_steps * warmup_ratio)
        latents = latents * self.scheduler.init_noise_sigma
        # multi-rank shards launch eagerly: the per-layer NCCL broadcast inside a captured forward is untested on this
        # stack (PAID_SHARD_GRAPHS=1 opts in, every rank captures the same sequence of collectives)
        graphs = (self.use_cuda_graphs and latents.is_cuda and
                  (self.shard is None or self.shard.world_size == 1 or os.environ.get("PAID_SHARD_GRAPHS") == "1"))
        for i, t in enumerate(self.scheduler.timesteps.tolist()):
            model_in = self.scheduler.scale_model_input(latents, t)
            noise_text = self._forward(i < warmup_steps, coef, model_in, t, cond, added_cond, graphs)
            if graphs:
                noise_text = noise_text.clone()      # the next replay may reuse the same static output
            noise_uncond = self._forward(False, coef, model_in, t, uncond, added_uncond, graphs)
            noise = noise_uncond + guidance_scale * (noise_text - noise_uncond)
            latents = self.scheduler.step(noise, t, latents)
        return latents

    def _set_mode(self, aid: bool, coef):
        if aid:
            self.set_coefs(coef)      # AID on (conditional pass of the first warmup_steps steps, sdxl:2245-2246)
        else:
            self.deactivate_aid()     # stock attention (sdxl:2248, 2272)

    def _forward(self, aid: bool, coef, sample, t, ctx, added, graphs: bool):
        if not graphs:
            self._set_mode(aid, coef)
            return self.unet(sample, t, ctx, added)
        key = (aid, tuple(float(c) for c in coef) if aid else None, tuple(sample.shape), tuple(ctx.shape), sample.dtype)
        g = self._graphs.get(key)
        if g is None:
            self._set_mode(aid, coef)
            g = self._graphs[key] = _GraphedForward(self.unet, sample, t, ctx, added)
        self.graph_kernel_launches += g.launches
        return g(sample, t, ctx, added)

    def _added(self, n, text_embeds, device, dtype):
        if not self.unet.cfg.text_time:
            return None
        s = self.unet.cfg.sample_size * 8
        ids = torch.tensor([[s, s, 0, 0, s, s]], device=device, dtype=dtype).expand(n, -1).contiguous()
        return {"text_embeds": text_embeds, "time_ids": ids}

    @torch.no_grad()
    def interpolate(self, latent_start, latent_end, embeds_start, embeds_end, negative_embeds, size: int = 7,
                    alpha: float = 4.0, beta: float = 4.0, guide_embeds=None, pooled_start=None, pooled_end=None,
                    pooled_negative=None, pooled_guide=None, num_inference_steps: int = 50,
                    guidance_scale: Optional[float] = None, warmup_ratio: float = 0.5, coef: Optional[torch.Tensor] = None,
                    ip_start=None, ip_end=None, ip_negative=None):
        """N-frame AID / PAID in one batch.  latent_* (1,4,H,W); embeds_* (1,77,Cc); pooled_* (1,1280) for SDXL.
        Interior frames take lerped embeddings, or ``guide_embeds`` when given (PAID).  Returns this rank's frames.

        Image-conditioned morphing (after ``load_aid_ip_adapter``; reference sdxl:2144-2197): ``ip_start`` / ``ip_end``
        (1,T,Cc) are the projected IP-Adapter image tokens of the two endpoint images; frame i gets their lerp by c_i
        (the reference's ``init == "linear"``), the unconditional pass ``ip_negative`` (default zeros).  The tokens
        travel appended to the text embeddings (one of the two forms the reference's processors accept,
        interpolation.py:259-266)."""
        g = self.unet.cfg.guidance_scale if guidance_scale is None else guidance_scale
        if coef is None:
            coef = generate_beta_tensor(size, alpha, beta)
        coef = coef.clone().float()
        coef[0], coef[-1] = 0, 1
        ts = [float(c) for c in coef]
        lat = torch.cat([slerp(latent_start, latent_end, t) for t in ts], dim=0)
        lat[0], lat[-1] = latent_start[0], latent_end[0]

        def frames(a, b, guide):
            if guide is not None:
                return torch.cat([a] + [guide] * (size - 2) + [b], dim=0)
            return torch.cat([torch.lerp(a, b, t) for t in ts], dim=0)

        cond = frames(embeds_start, embeds_end, guide_embeds)
        uncond = negative_embeds.expand(size, -1, -1).contiguous()
        if ip_start is not None:
            ip_neg = torch.zeros_like(ip_start) if ip_negative is None else ip_negative
            cond = torch.cat([cond, frames(ip_start, ip_end, None)], dim=1)
            uncond = torch.cat([uncond, ip_neg.expand(size, -1, -1)], dim=1)
        pooled_c = pooled_u = None
        if self.unet.cfg.text_time:
            pooled_c = frames(pooled_start, pooled_end, pooled_guide)
            pooled_u = pooled_negative.expand(size, -1).contiguous()
        if self.shard is not None:
            sl = self.shard.local
            lat, cond, uncond = sl(lat).contiguous(), sl(cond).contiguous(), sl(uncond).contiguous()
            if pooled_c is not None:
                pooled_c, pooled_u = sl(pooled_c).contiguous(), sl(pooled_u).contiguous()
        n = lat.shape[0]
        return self._denoise(lat, cond, uncond, self._added(n, pooled_c, lat.device, lat.dtype),
                             self._added(n, pooled_u, lat.device, lat.dtype), coef, num_inference_steps, g, warmup_ratio)

    @torch.no_grad()
    def interpolate_candidates(self, ts, latent_start, latent_end, embeds_start, embeds_end, negative_embeds, **kw):
        """Frames for several interpolation parameters ``ts`` in ONE batch ``[start, t_1, ..., t_K, end]``.  Interior
        frames interact with the rest of the batch only through the endpoint K/V, so frame i equals the middle frame
        of the reference's 3-frame ``interpolate_single(t_i)`` (SURVEY.md section 4 property 2): the K sequential
        denoises of the reference's exploration loop (prior.py:119-199 runs one per candidate) become one sharded
        batch.  The candidate SELECTION of that loop (CLIP distances, Beta fit) stays with the caller."""
        ts = [float(t) for t in ts]
        assert all(0 < t < 1 for t in ts), "t must be between 0 and 1"
        return self.interpolate(latent_start, latent_end, embeds_start, embeds_end, negative_embeds, size=len(ts) + 2,
                                coef=torch.tensor([0.0, *ts, 1.0]), **kw)

    @torch.no_grad()
    def interpolate_single(self, it: float, latent_start, latent_end, embeds_start, embeds_end, negative_embeds,
                           guide_embeds=None, warmup_ratio: float = 0.5, num_inference_steps: int = 50,
                           guidance_scale: Optional[float] = None, **pooled):
        """The reference's 3-frame ``[start, it, end]`` call (sdxl:1693)."""
        assert 0 < it < 1, "t must be between 0 and 1"
        return self.interpolate(latent_start, latent_end, embeds_start, embeds_end, negative_embeds, size=3,
                                guide_embeds=guide_embeds, num_inference_steps=num_inference_steps,
                                guidance_scale=guidance_scale, warmup_ratio=warmup_ratio,
                                coef=torch.tensor([0.0, it, 1.0]), **pooled)
