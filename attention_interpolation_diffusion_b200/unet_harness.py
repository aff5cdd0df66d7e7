"""UNet-shaped harness: the caller of the hot path, so that interpolation-frames/sec of a
whole denoising loop is measurable without diffusers or checkpoints (neither exists in this
image).

Architecture follows the public SD1.5 / SDXL-base ``unet/config.json`` geometry: same block
structure, channel widths, transformer depths, head counts and cross-attention widths, hence
the same 32 / 140 attention-processor call sites with the same (S, L, C, heads) as the
reference sees inside diffusers' ``UNet2DConditionModel`` (pipeline_interpolated_sdxl.py:2261).
Weights are random-init.  Everything outside the attention processors is plain PyTorch
(cuDNN / cuBLAS) plumbing; attention goes through ``Attention.processor`` exactly as in
diffusers, so the processors of ``interpolation.py`` install with ``set_attn_processor``.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .attention import Attention


@dataclass
class UNetConfig:
    name: str
    sample_size: int
    block_out_channels: Tuple[int, ...]
    down_has_attn: Tuple[bool, ...]
    up_has_attn: Tuple[bool, ...]
    transformer_layers: Tuple[int, ...]      # per down level
    heads: Tuple[int, ...]                   # per down level
    cross_attention_dim: int
    use_linear_projection: bool
    layers_per_block: int = 2
    in_channels: int = 4
    text_time: bool = False                  # SDXL addition_embed_type == "text_time"
    guidance_scale: float = 7.5


SD15 = UNetConfig("sd15", 64, (320, 640, 1280, 1280), (True, True, True, False), (False, True, True, True),
                  (1, 1, 1, 1), (8, 8, 8, 8), 768, False, guidance_scale=7.5)
SDXL = UNetConfig("sdxl", 128, (320, 640, 1280), (False, True, True), (True, True, False),
                  (1, 2, 10), (5, 10, 20), 2048, True, text_time=True, guidance_scale=5.0)
# reduced geometry for tests / smoke: same structure, two levels, tiny widths (d = 64 and d = 40)
TINY = UNetConfig("tiny", 16, (128, 256), (True, True), (True, True), (1, 2), (2, 4), 96, True,
                  layers_per_block=1, text_time=True, guidance_scale=5.0)
CONFIGS = {"sd15": SD15, "sdxl": SDXL, "tiny": TINY}


def sinusoidal(t: torch.Tensor, dim: int) -> torch.Tensor:
    """flip_sin_to_cos=True, freq_shift=0."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# tests (and PAID_NATIVE_GLUE=0 for A/B timing) flip this to compare the fused glue kernels with the PyTorch composition
NATIVE_GLUE = os.environ.get("PAID_NATIVE_GLUE", "1") != "0"


def _native(x: torch.Tensor) -> bool:
    """Half-precision CUDA tensors take libpaid_attn's fused glue kernels; anything else (the CPU tests that host the
    oracle processors in this harness) takes the plain PyTorch composition of the same ops."""
    if x.is_cuda and NATIVE_GLUE and x.dtype not in (torch.float16, torch.bfloat16):
        raise TypeError(f"the CUDA path computes in fp16 / bf16 (got {x.dtype}); there is no fp32 GPU path")
    return NATIVE_GLUE and x.is_cuda


def group_norm(gn: nn.GroupNorm, x: torch.Tensor, silu: bool = False, pre_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """act(GroupNorm(x + pre_bias[:, :, None, None])): one statistics pass + one normalise/SiLU pass over the
    channels-last tensor (``paid_group_norm_nhwc``) instead of PyTorch's NHWC->NCHW copy, moments, normalise, SiLU and
    NCHW->NHWC copy kernels."""
    if _native(x):
        from . import _cabi
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        return _cabi.group_norm_nhwc(x, gn.weight, gn.bias, gn.num_groups, gn.eps, silu, pre_bias)
    if pre_bias is not None:
        x = x + pre_bias[:, :, None, None]
    h = gn(x)
    return F.silu(h) if silu else h


def add_layer_norm(ln: nn.LayerNorm, x: torch.Tensor, delta: Optional[torch.Tensor]):
    """(x + delta, LayerNorm(x + delta)): the residual add of the previous sub-layer fused with the norm in front of
    the next one (``paid_add_layer_norm``, one pass over the row)."""
    if _native(x):
        from . import _cabi
        return _cabi.add_layer_norm(x, delta, ln.weight, ln.bias, ln.eps)
    if delta is not None:
        x = x + delta
    return x, ln(x)


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_ch):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        if _native(x):
            # the convs run without their biases (PyTorch adds a conv bias as a separate broadcast kernel): conv1's
            # bias and the time embedding go into the load of the second norm, conv2's (and the shortcut's) bias into
            # the residual add
            from . import _cabi
            c1, c2, sc = self.conv1, self.conv2, self.conv_shortcut
            h = F.conv2d(group_norm(self.norm1, x, silu=True), c1.weight, None, padding=1)
            tb = F.linear(F.silu(temb), self.time_emb_proj.weight, self.time_emb_proj.bias + c1.bias)
            h = F.conv2d(group_norm(self.norm2, h, silu=True, pre_bias=tb), c2.weight, None, padding=1)
            if not x.is_contiguous(memory_format=torch.channels_last):
                x = x.contiguous(memory_format=torch.channels_last)
            skip = x if sc is None else F.conv2d(x, sc.weight, None)
            return _cabi.residual_bias_add(skip, h, c2.bias if sc is None else c2.bias + sc.bias)
        h = self.conv1(group_norm(self.norm1, x, silu=True))
        # h + time_emb[:, :, None, None] is folded into the load of the second norm
        h = self.conv2(group_norm(self.norm2, h, silu=True, pre_bias=self.time_emb_proj(F.silu(temb))))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Linear(dim, dim * 8)
        self.out = nn.Linear(dim * 4, dim)

    def forward(self, x):
        if _native(x):
            # both Linears on libpaid_attn's tcgen05 GEMM, the GEGLU in the first one's epilogue (the (rows, 8C)
            # intermediate never reaches HBM)
            from . import _cabi
            return _cabi.linear(_cabi.linear_geglu(x, self.proj.weight, self.proj.bias), self.out.weight, self.out.bias)
        a, g = self.proj(x).chunk(2, dim=-1)
        return self.out(a * F.gelu(g))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim // heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads, dim // heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx, pending=None):
        """x = x + attn1(norm1(x)); x = x + attn2(norm2(x), ctx); x = x + ff(norm3(x)), with every residual add fused
        into the LayerNorm that follows it.  ``pending`` is the previous block's feed-forward output, not yet added to
        x; returns (x, pending) in the same form."""
        x, h = add_layer_norm(self.norm1, x, pending)
        x, h = add_layer_norm(self.norm2, x, self.attn1(h))
        x, h = add_layer_norm(self.norm3, x, self.attn2(h, encoder_hidden_states=ctx))
        return x, self.ff(h)


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, cross_dim, depth, linear_proj):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(32, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, cross_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        b, c, hh, ww = x.shape
        res = x
        h = group_norm(self.norm, x)
        native = _native(x)
        if native:
            # channels-last feature map = (tokens, C) matrix: proj_in / proj_out (Linear, or 1x1 conv for SD1.5) are plain
            # GEMMs on libpaid_attn's tcgen05 kernel
            from . import _cabi
            h = _cabi.linear(h.permute(0, 2, 3, 1).reshape(b, hh * ww, c), self.proj_in.weight.reshape(c, c), self.proj_in.bias)
        elif self.linear_proj:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(b, hh * ww, c))
        else:
            h = self.proj_in(h).permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        h = h.contiguous()
        pending = None
        for blk in self.transformer_blocks:
            h, pending = blk(h, ctx, pending)
        h = h + pending
        if native:
            h = _cabi.linear(h, self.proj_out.weight.reshape(c, c), self.proj_out.bias).reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        elif self.linear_proj:
            h = self.proj_out(h).reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        else:
            h = self.proj_out(h.reshape(b, hh, ww, c).permute(0, 3, 1, 2))
        return h + res


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb_ch, n_layers, attn, heads, cross_dim, depth, linear_proj, downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_ch) for i in range(n_layers)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(cout, heads, cross_dim, depth, linear_proj) for _ in range(n_layers)]) if attn else None
        self.downsamplers = nn.ModuleList([nn.Conv2d(cout, cout, 3, stride=2, padding=1)]) if downsample else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, ch, temb_ch, heads, cross_dim, depth, linear_proj):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch), ResnetBlock2D(ch, ch, temb_ch)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, cross_dim, depth, linear_proj)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin, cout, prev, temb_ch, n_layers, attn, heads, cross_dim, depth, linear_proj, upsample):
        super().__init__()
        res = []
        for i in range(n_layers):
            skip = cin if i == n_layers - 1 else cout
            rin = prev if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, temb_ch))
        self.resnets = nn.ModuleList(res)
        self.attentions = nn.ModuleList(
            [Transformer2DModel(cout, heads, cross_dim, depth, linear_proj) for _ in range(n_layers)]) if attn else None
        self.upsamplers = nn.ModuleList([nn.Conv2d(cout, cout, 3, padding=1)]) if upsample else None

    def forward(self, x, skips, temb, ctx):
        for i, r in enumerate(self.resnets):
            x = r(torch.cat([x, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class UNetHarness(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        boc = cfg.block_out_channels
        temb_ch = boc[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = nn.Sequential(nn.Linear(boc[0], temb_ch), nn.SiLU(), nn.Linear(temb_ch, temb_ch))
        if cfg.text_time:
            self.add_embedding = nn.Sequential(nn.Linear(1280 + 6 * 256, temb_ch), nn.SiLU(), nn.Linear(temb_ch, temb_ch))
        nl = len(boc)
        self.down_blocks = nn.ModuleList()
        ch = boc[0]
        for i in range(nl):
            self.down_blocks.append(DownBlock(ch, boc[i], temb_ch, cfg.layers_per_block, cfg.down_has_attn[i],
                                              cfg.heads[i], cfg.cross_attention_dim, cfg.transformer_layers[i],
                                              cfg.use_linear_projection, downsample=i < nl - 1))
            ch = boc[i]
        self.mid_block = MidBlock(boc[-1], temb_ch, cfg.heads[-1], cfg.cross_attention_dim, cfg.transformer_layers[-1],
                                  cfg.use_linear_projection)
        rev = boc[::-1]
        rheads, rdepth = cfg.heads[::-1], cfg.transformer_layers[::-1]
        self.up_blocks = nn.ModuleList()
        prev = rev[0]
        for i in range(nl):
            cout, cin = rev[i], rev[min(i + 1, nl - 1)]
            self.up_blocks.append(UpBlock(cin, cout, prev, temb_ch, cfg.layers_per_block + 1, cfg.up_has_attn[i],
                                          rheads[i], cfg.cross_attention_dim, rdepth[i], cfg.use_linear_projection,
                                          upsample=i < nl - 1))
            prev = cout
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], cfg.in_channels, 3, padding=1)
        # resolution level of every attention layer (side = sample_size >> level)
        tagged = [(b, i) for i, b in enumerate(self.down_blocks)] + [(self.mid_block, nl - 1)] + \
                 [(b, nl - 1 - i) for i, b in enumerate(self.up_blocks)]
        for blk, level in tagged:
            for m in blk.modules():
                if isinstance(m, Attention):
                    m.level = level

    # ---- the diffusers plugin surface used by load_aid (pipeline_interpolated_sdxl.py:1069-1086) ----
    def _attention_modules(self) -> Dict[str, Attention]:
        return {f"{name}.processor": m for name, m in self.named_modules() if isinstance(m, Attention)}

    @property
    def attn_processors(self) -> Dict[str, object]:
        return {name: m.processor for name, m in self._attention_modules().items()}

    def set_attn_processor(self, processor):
        mods = self._attention_modules()
        if isinstance(processor, dict):
            if len(processor) != len(mods):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does "
                                 f"not match the number of attention layers: {len(mods)}.")
            for name, m in mods.items():
                m.set_processor(processor[name])
        else:
            for m in mods.values():
                m.set_processor(processor)

    def attention_geometry(self) -> List[dict]:
        """(S, L, C, heads) of every attention call of one forward, in call order."""
        cfg, out = self.cfg, []
        for name, m in self._attention_modules().items():
            C, level = m.to_q.in_features, m.level
            side = cfg.sample_size >> level
            is_self = name.endswith("attn1.processor")
            out.append(dict(name=name, S=side * side, L=side * side if is_self else 77, C=C,
                            Cc=C if is_self else cfg.cross_attention_dim, heads=m.heads, self_attn=is_self))
        return out

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs: Optional[dict] = None):
        cfg = self.cfg
        b = sample.shape[0]
        t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=sample.device)
        t = t.reshape(-1).expand(b)
        temb = self.time_embedding(sinusoidal(t, cfg.block_out_channels[0]).to(sample.dtype))
        if cfg.text_time:
            ids = added_cond_kwargs["time_ids"]
            tid = sinusoidal(ids.reshape(-1), 256).reshape(b, -1).to(sample.dtype)
            temb = temb + self.add_embedding(torch.cat([added_cond_kwargs["text_embeds"], tid], dim=-1))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states)
            skips.extend(outs)
        x = self.mid_block(x, temb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states)
        return self.conv_out(group_norm(self.conv_norm_out, x, silu=True))


def build_unet(name: str, device="cuda", dtype=torch.float16, seed: int = 1002) -> UNetHarness:
    """Random-init UNet of the named geometry (default nn init under the reference's seed, gradio_src/app.py:131)."""
    torch.manual_seed(seed)
    with torch.device(device):
        net = UNetHarness(CONFIGS[name])
    net = net.to(dtype=dtype).eval().requires_grad_(False)
    if torch.device(device).type == "cuda":
        net = net.to(memory_format=torch.channels_last)   # NHWC convs; the token view of a feature map is then free
    return net
