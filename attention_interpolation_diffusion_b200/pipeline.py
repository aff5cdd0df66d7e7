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

from .attention import PaidAttnProcessor, PaidIPAdapterAttnProcessor
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
    of a forward otherwise cost ~100 ms of host time, more than the GPU needs to execute them.  A frame-sharded
    forward captures its NCCL broadcasts and the side stream they run on as well (every rank captures and replays
    the same sequence of collectives)."""

    def __init__(self, unet, sample, t, ctx, added, shard=None):
        from . import _cabi
        dev = sample.device
        self.sample, self.ctx = sample.clone(), ctx.clone()
        self.t = torch.zeros(1, device=dev, dtype=torch.float32)
        self.added = None if added is None else {k: v.clone() for k, v in added.items()}
        self.t.fill_(float(t))
        self.shard = shard

        def run():
            if shard is not None:
                shard.begin_forward(dev)
            out = unet(self.sample, self.t, self.ctx, self.added)
            if shard is not None:
                shard.end_forward(dev)
            return out

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):      # warm-up outside capture: cuDNN autotune, lazy init, NCCL channels, workspace
            for _ in range(2):
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.workspace = _cabi._workspaces.get(dev)      # keep the scratch buffer the graph points into alive
        n0 = _cabi.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = run()
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
    MAX_GRAPHS = 6       # captured forwards kept alive (each owns a private memory pool): 3 per (shape, processor set)

    def __init__(self, unet: UNetHarness, shard: Optional[FrameShard] = None, use_cuda_graphs: bool = True,
                 cache_static_kv: bool = True, merge_plain_passes: bool = True, merge_aid_passes: bool = True):
        self.unet = unet
        # After the warm-up steps the conditional and the unconditional pass of a step both run stock attention
        # (sdxl:2245-2248): the reference still calls the UNet twice with n frames; here they run as ONE call with 2 n frames
        # (every op of the UNet is per sample, so the results are the same rows), which halves the launches of those steps
        # and doubles the rows of every GEMM.  Text-only processors on CUDA only (the IP-Adapter variants read per-frame
        # state while deactivated).
        self.merge_plain_passes = merge_plain_passes
        # The warm-up steps batch the same way: the unconditional frames ride as the tail of the conditional call, the
        # processors interpolate the first n frames and run stock attention on the last n (PaidAttnParams.plain_tail).
        self.merge_aid_passes = merge_aid_passes and merge_plain_passes
        self.scheduler = DDIMScheduler()
        self.shard = shard
        self.use_cuda_graphs = use_cuda_graphs
        self.cache_static_kv = cache_static_kv
        self._graphs: dict = {}            # insertion-ordered: oldest first (LRU eviction)
        self._coef_buf: Optional[torch.Tensor] = None     # fp32 coefficients of the local frames, shared by all processors
        self._kv_tag = [None]              # the pass whose cached cross-attention K/V the processors read ("cond" / "uncond" / "both")
        self.graph_kernel_launches = 0     # libpaid_attn kernels executed through graph replays
        self.load_aid()

    # ---- processor install / toggle (sdxl:1066-1136) ---------------------------------------------
    def _installed(self):
        for name, proc in self.unet.attn_processors.items():
            if not name.startswith("encoder"):
                yield name, proc

    @staticmethod
    def _stock(proc):
        """The stock attention processor underneath whatever AID wrapper is installed (the reference captures
        ``AttnProcessor2_0`` as ``original_attn``, sdxl:1076)."""
        for _ in range(4):
            inner = getattr(proc, "original_attn", None) or getattr(proc, "ip_attn", None)
            if inner is None:
                break
            proc = inner
        wrappers = (OuterInterpolatedAttnProcessor, InnerInterpolatedAttnProcessor, OuterInterpolatedIPAttnProcessor,
                    InnerInterpolatedIPAttnProcessor, ScaleControlIPAttnProcessor, PaidIPAdapterAttnProcessor)
        return PaidAttnProcessor() if proc is None or isinstance(proc, wrappers) else proc

    def _after_install(self):
        self._graphs.clear()               # captured forwards bake the processor objects in
        self._coef_buf = None
        for _, m in self.unet._attention_modules().items():
            m.paid_kv, m.paid_kv_tag = None, self._kv_tag

    def load_aid(self, t: Optional[float] = 0.5, is_fused: bool = True, atype: str = "fused_outer", size: int = 7,
                 alpha: float = 1, beta: float = 1):
        cls = {"fused_outer": OuterInterpolatedAttnProcessor, "fused_inner": InnerInterpolatedAttnProcessor}[atype]
        attn_procs = {}
        for name, old in self.unet.attn_processors.items():
            if not name.startswith("encoder"):
                proc = cls(t=t, size=size, is_fused=is_fused, alpha=alpha, beta=beta, original_attn=self._stock(old))
                proc.shard = self.shard
                attn_procs[name] = proc
            else:
                attn_procs[name] = old
        self.unet.set_attn_processor(attn_procs)
        self._after_install()

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
            else:
                old = self._stock(old)
            attn_procs[name] = cls(t=t, size=size, is_fused=is_fused, alpha=alpha, beta=beta, ip_attn=old)
            attn_procs[name].shard = self.shard
        self.unet.set_attn_processor(attn_procs)
        self._after_install()

    def activate_aid(self, it: float):
        for _, proc in self._installed():
            proc.activate(it)

    def deactivate_aid(self):
        for _, proc in self._installed():
            proc.deactivate()

    def set_coefs(self, coef: torch.Tensor):
        """N-frame extension of activate_aid: one coefficient per frame.  On CUDA the values also go into ONE fp32 device
        buffer that every processor reads (its address is what captured graphs hold; a new schedule rewrites it in place)."""
        for _, proc in self._installed():
            proc.set_coefs(coef)

    def _bind_coefs(self, coef: torch.Tensor, device: torch.device):
        """Install the schedule on every processor and refresh the shared device buffer (local frames only)."""
        self.set_coefs(coef)
        if device.type != "cuda":
            return
        local = coef.detach().float().clone()
        local[0], local[-1] = 0, 1
        if self.shard is not None:
            local = local[self.shard.frame_ids]
        if self._coef_buf is None or self._coef_buf.numel() != local.numel() or self._coef_buf.device != device:
            self._coef_buf = torch.empty(local.numel(), dtype=torch.float32, device=device)
            self._graphs.clear()
        self._coef_buf.copy_(local)
        for _, proc in self._installed():
            proc.bind_coef_buffer(self._coef_buf)

    def _refresh_static_kv(self, cond, uncond, uncond_uniform: bool, cond_endpoints, both=None):
        """K / V of the prompt embeddings of every cross-attention layer, once per sequence (SURVEY.md section 8f rank 1;
        the reference re-projects them in all 100 UNet calls, interpolation.py:623-624).  The unconditional pass carries
        the same negative prompt in every frame: one (L, C) K/V pair serves all frames (kv_broadcast)."""
        if not (self.cache_static_kv and cond.is_cuda):
            for _, m in self.unet._attention_modules().items():
                m.paid_kv = None
            return
        passes = [("cond", cond, False, cond_endpoints), ("uncond", uncond, uncond_uniform, None)]
        if both is not None:             # conditional + unconditional frames in one call (the interpolated ones need the endpoints)
            passes.append(("both", both, False, cond_endpoints))
        for name, m in self.unet._attention_modules().items():
            if not name.endswith("attn2.processor"):
                continue
            proc = m.processor
            if m.paid_kv is None:
                m.paid_kv = {"cond": {}, "uncond": {}, "both": {}}
            for tag, ctx, uniform, ends in passes:
                entry = m.paid_kv[tag]
                entry.pop("reallocated", None)
                proc.project_static(m, ctx, uniform, entry, ends)
                if entry.pop("reallocated", False):
                    self._graphs.clear()       # captured forwards point into the old buffers

    # ---- the step loop ---------------------------------------------------------------------------
    @torch.no_grad()
    def _denoise(self, latents, cond, uncond, added_cond, added_uncond, coef, num_inference_steps, guidance_scale,
                 warmup_ratio, uncond_uniform: bool = False, cond_endpoints=None):
        """latents (n,4,H,W); cond / uncond (n,77,Cc): the local frames.  Returns final latents."""
        self.scheduler.set_timesteps(num_inference_steps)
        warmup_steps = int(num_inference_steps * warmup_ratio)
        latents = latents * self.scheduler.init_noise_sigma
        multi_rank = self.shard is not None and self.shard.world_size > 1
        # PAID_SHARD_GRAPHS=0: launch a multi-rank forward eagerly (A/B timing of the captured NCCL broadcasts)
        graphs = self.use_cuda_graphs and latents.is_cuda and not (multi_rank and os.environ.get("PAID_SHARD_GRAPHS") == "0")
        self._bind_coefs(coef, latents.device)
        if self.shard is not None:
            self.shard.static_endpoints = cond_endpoints
        n = latents.shape[0]
        text_only = all(type(p) in (OuterInterpolatedAttnProcessor, InnerInterpolatedAttnProcessor) for _, p in self._installed())
        merge_plain = self.merge_plain_passes and latents.is_cuda and text_only and warmup_steps < num_inference_steps
        # frame-sharded over more than two ranks the warm-up steps keep the reference's two calls per step: the guidance rows
        # inside the interpolated call are validated on GPUs at world sizes 1 and 2 only
        few_ranks = self.shard is None or self.shard.world_size <= 2 or os.environ.get("PAID_MERGE_AID_ANY_WORLD") == "1"
        merge_aid = self.merge_aid_passes and latents.is_cuda and text_only and warmup_steps > 0 and few_ranks
        merge = merge_plain or merge_aid
        both = torch.cat([cond, uncond]) if merge else None
        added_both = None if (not merge or added_cond is None) else {k: torch.cat([added_cond[k], added_uncond[k]]) for k in added_cond}
        self._refresh_static_kv(cond, uncond, uncond_uniform, cond_endpoints if self.shard is not None else None, both)
        for i, t in enumerate(self.scheduler.timesteps.tolist()):
            model_in = self.scheduler.scale_model_input(latents, t)
            if (merge_aid and i < warmup_steps) or (merge_plain and i >= warmup_steps):
                aid = i < warmup_steps
                out = self._forward(aid, "both", torch.cat([model_in, model_in]), t, both, added_both, graphs, tail=n if aid else 0)
                noise_text, noise_uncond = out[:n], out[n:]
            else:
                noise_text = self._forward(i < warmup_steps, "cond", model_in, t, cond, added_cond, graphs)
                if graphs:
                    noise_text = noise_text.clone()      # the next replay may reuse the same static output
                noise_uncond = self._forward(False, "uncond", model_in, t, uncond, added_uncond, graphs)
            noise = noise_uncond + guidance_scale * (noise_text - noise_uncond)
            latents = self.scheduler.step(noise, t, latents)
        for _, proc in self._installed():
            proc.cfg_tail = 0               # processors called outside the step loop see plain interpolation batches again
        return latents

    def _set_mode(self, aid: bool, tail: int = 0):
        for _, proc in self._installed():
            proc.activated = bool(aid)      # AID on: conditional pass of the first warmup_steps steps (sdxl:2245-2248, 2272)
            proc.cfg_tail = tail if aid else 0   # unconditional frames appended to an interpolated call

    def _forward(self, aid: bool, tag: str, sample, t, ctx, added, graphs: bool, tail: int = 0):
        self._kv_tag[0] = tag
        if not graphs:
            self._set_mode(aid, tail)
            if self.shard is not None:
                self.shard.begin_forward(sample.device)
            out = self.unet(sample, t, ctx, added)
            if self.shard is not None:
                self.shard.end_forward(sample.device)
            return out
        # the coefficient VALUES are not part of the key: they live in the shared device buffer (_bind_coefs)
        key = (aid, tag, tail, tuple(sample.shape), tuple(ctx.shape), sample.dtype)
        g = self._graphs.pop(key, None)
        if g is None:
            self._set_mode(aid, tail)
            while len(self._graphs) >= self.MAX_GRAPHS:
                self._graphs.pop(next(iter(self._graphs)))     # least recently used
            g = _GraphedForward(self.unet, sample, t, ctx, added, self.shard)
        self._graphs[key] = g                                  # most recently used last
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
        Interior frames take lerped embeddings, or ``guide_embeds`` when given (PAID).  Returns this rank's frames
        (frame-sharded: in the shard's local order, ``FrameShard.frame_ids``).

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
        cond_endpoints = torch.stack([cond[0], cond[-1]]).contiguous()     # the endpoint prompts (every rank holds them)
        if self.shard is not None:
            sl = self.shard.local
            lat, cond, uncond = sl(lat).contiguous(), sl(cond).contiguous(), sl(uncond).contiguous()
            if pooled_c is not None:
                pooled_c, pooled_u = sl(pooled_c).contiguous(), sl(pooled_u).contiguous()
        n = lat.shape[0]
        return self._denoise(lat, cond, uncond, self._added(n, pooled_c, lat.device, lat.dtype),
                             self._added(n, pooled_u, lat.device, lat.dtype), coef, num_inference_steps, g, warmup_ratio,
                             uncond_uniform=True, cond_endpoints=cond_endpoints)

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
