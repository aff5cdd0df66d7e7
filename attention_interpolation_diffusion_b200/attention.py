"""Host-side stand-in for ``diffusers.models.attention_processor.Attention``.

diffusers is not installed in this image, so the UNet harness and the tests use
this module.  It exposes exactly the members the reference processors read from
``attn`` (interpolation.py:588-677), so the same processor
objects also plug into a real diffusers ``Attention``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _cabi


def static_kv(attn, encoder_hidden_states):
    """Per-sequence K/V cache of a cross-attention layer.  The prompt embeddings do not change over the denoising steps
    (pipeline_interpolated_sdxl.py:2232-2345 feeds the same ``prompt_embeds`` to every UNet call), so the step loop
    projects them once per sequence (``InterpolationPipeline._refresh_static_kv`` -> ``project_static`` of the layer's
    processor) into ``attn.paid_kv[tag]`` for tag = "cond" / "uncond", and names the running pass in
    ``attn.paid_kv_tag[0]``.  Returns that entry, or None (no cache attached, self-attention, other pass): then the
    call projects K / V itself, as the reference does on every call."""
    store = getattr(attn, "paid_kv", None)
    if store is None or encoder_hidden_states is None:
        return None
    return store.get(attn.paid_kv_tag[0])


def _static_buffer(entry: dict, name: str, shape, like: torch.Tensor) -> torch.Tensor:
    """Persistent buffer of a cache entry (re-used across sequences: captured CUDA graphs hold its address)."""
    t = entry.get(name)
    if t is None or tuple(t.shape) != tuple(shape) or t.dtype != like.dtype or t.device != like.device:
        t = entry[name] = torch.empty(*shape, dtype=like.dtype, device=like.device)
        entry["reallocated"] = True
    return t


def project_text_static(attn, ctx: torch.Tensor, uniform: bool, entry: dict, endpoints: Optional[torch.Tensor] = None,
                        flags: int = 0, prefix: str = "", wk=None, wv=None):
    """K / V of a step-invariant context ``ctx`` (n, L, Cc) into ``entry`` (keys prefix + "k" / "v"): one (1, L, C) pair
    when every frame carries the same context (``uniform``, the unconditional pass), else per frame.  ``endpoints``
    (2, L, Cc): the two endpoint prompts of a frame-sharded sequence -> prefix + "kv_ext" (4, L, C), projected locally."""
    wk = attn.to_k.weight if wk is None else wk
    wv = attn.to_v.weight if wv is None else wv
    C = wk.shape[0]
    src = (ctx[:1] if uniform else ctx).contiguous()
    n, L = src.shape[0], src.shape[1]
    probe = src.new_empty(n, 1, C)      # only its shape / dtype are read
    k = _static_buffer(entry, prefix + "k", (n, L, C), src)
    v = _static_buffer(entry, prefix + "v", (n, L, C), src)
    _cabi.project_kv(probe, src, wk, wv, attn.heads, k, v, flags)
    entry[prefix + "broadcast"] = bool(uniform)
    if endpoints is not None:
        kv = _static_buffer(entry, prefix + "kv_ext", (4, L, C), src)
        ends = endpoints.contiguous()
        for f in range(2):
            _cabi.project_endpoints(src.new_empty(2, 1, C), ends, wk, wv, attn.heads, f, kv[2 * f], kv[2 * f + 1], flags)


class PaidAttnProcessor:
    """Stock (non-interpolated) attention through the PLAIN mode of libpaid_attn --
    the role ``AttnProcessor2_0`` plays in the reference (``original_attn``,
    pipeline_interpolated_sdxl.py:1076)."""

    def project_static(self, attn, ctx, uniform, entry, endpoints=None):
        project_text_static(attn, ctx, uniform, entry, None)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        check_unet_preconditions(attn, hidden_states, attention_mask)
        st = static_kv(attn, encoder_hidden_states) or {}
        return _cabi.attn_forward(
            hidden_states, encoder_hidden_states, attn.to_q.weight, attn.to_k.weight, attn.to_v.weight,
            attn.to_out[0].weight, attn.to_out[0].bias, None, attn.heads, _cabi.PAID_PLAIN, False, attn.scale,
            k_pre=st.get("k"), v_pre=st.get("v"), kv_pre_broadcast=st.get("broadcast", False))


def split_ip_states(encoder_hidden_states, num_tokens, batch):
    """(text, image tokens) from the forms the reference accepts (interpolation.py:254-266): a tuple
    ``(text, [ip])`` or one tensor with the image tokens appended.  Image tokens may come with the reference's 3x
    row repetition (rows [s,s,s,t,t,t,e,e,e] for a batch of 3, sdxl:2146-2185): un-repeated like its ``[::3]``."""
    if isinstance(encoder_hidden_states, tuple):
        text, ip = encoder_hidden_states
        ip = ip[0] if isinstance(ip, (list, tuple)) else ip
    else:
        end = encoder_hidden_states.shape[1] - num_tokens[0]
        text, ip = encoder_hidden_states[:, :end, :], encoder_hidden_states[:, end:, :]
    if ip.ndim == 4:                      # (B, num_images=1, T, Cc) of newer diffusers
        ip = ip.reshape(ip.shape[0], -1, ip.shape[-1])
    if ip.shape[0] == 3 * batch:
        ip = ip[::3]
    if ip.shape[0] != batch:
        raise ValueError(f"image tokens for {ip.shape[0]} frames, batch of {batch}")
    return text.contiguous(), ip.contiguous()


class PaidIPAdapterAttnProcessor(nn.Module):
    """Stock IP-Adapter attention (role of diffusers ``IPAdapterAttnProcessor2_0``, the ``ip_attn`` the reference
    wraps at pipeline_interpolated_sdxl.py:1110-1126): plain text attention + scale * plain image-token attention."""

    def __init__(self, hidden_size: int, cross_attention_dim: int, num_tokens=(4,), scale=1.0):
        super().__init__()
        self.num_tokens = tuple(num_tokens)
        self.scale = [scale] if not isinstance(scale, (list, tuple)) else list(scale)
        self.to_k_ip = nn.ModuleList([nn.Linear(cross_attention_dim, hidden_size, bias=False)])
        self.to_v_ip = nn.ModuleList([nn.Linear(cross_attention_dim, hidden_size, bias=False)])

    def project_static(self, attn, ctx, uniform, entry, endpoints=None):
        text, ip = split_ip_states(ctx, self.num_tokens, ctx.shape[0])
        project_text_static(attn, text, uniform, entry)
        project_text_static(attn, ip, uniform, entry, prefix="ip_", wk=self.to_k_ip[0].weight, wv=self.to_v_ip[0].weight)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        check_unet_preconditions(attn, hidden_states, attention_mask)
        x = hidden_states
        st = static_kv(attn, encoder_hidden_states)
        q = _cabi.linear(x, attn.to_q.weight)
        if st is not None:
            k, v, kip, vip, bc = st["k"], st["v"], st["ip_k"], st["ip_v"], st["broadcast"]
        else:
            text, ip = split_ip_states(encoder_hidden_states, self.num_tokens, x.shape[0])
            k, v = _cabi.linear(text, attn.to_k.weight), _cabi.linear(text, attn.to_v.weight)
            kip, vip, bc = _cabi.linear(ip, self.to_k_ip[0].weight), _cabi.linear(ip, self.to_v_ip[0].weight), False
        hid = _cabi.attn_core(q, k, v, None, attn.heads, _cabi.PAID_PLAIN, False, attn.scale, kv_broadcast=bc)
        _cabi.attn_core(q, kip, vip, None, attn.heads, _cabi.PAID_PLAIN, False, attn.scale, out=hid, accumulate=True,
                        out_scale=float(self.scale[0]), kv_broadcast=bc)
        return _cabi.linear(hid, attn.to_out[0].weight, attn.to_out[0].bias)


def check_unet_preconditions(attn, hidden_states, attention_mask):
    """The kernels implement the UNet transformer-block case of the processors; the
    branches that are dead there (interpolation.py:586-611, 618-621, 669-677) are
    rejected instead of silently falling back to PyTorch."""
    if hidden_states.ndim != 3:
        raise NotImplementedError("libpaid_attn handles 3-D (batch, tokens, channels) hidden_states only")
    if attention_mask is not None:
        raise NotImplementedError("attention_mask must be None (the reference's fused modes are inconsistent with a mask)")
    if getattr(attn, "spatial_norm", None) is not None or getattr(attn, "group_norm", None) is not None:
        raise NotImplementedError("spatial_norm / group_norm attention blocks are not on the UNet transformer path")
    if getattr(attn, "norm_cross", None):
        raise NotImplementedError("norm_cross is not used by the SD / SDXL UNets")
    if getattr(attn, "residual_connection", False) or getattr(attn, "rescale_output_factor", 1.0) != 1.0:
        raise NotImplementedError("residual_connection / rescale_output_factor are not used by the SD / SDXL UNets")
    if attn.to_q.bias is not None or attn.to_k.bias is not None or attn.to_v.bias is not None:
        raise NotImplementedError("q/k/v projections with bias are not used by the SD / SDXL UNets")


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, bias: bool = False, processor=None):
        super().__init__()
        inner = heads * dim_head
        if inner != query_dim:
            raise NotImplementedError("SD / SDXL UNet attention has inner_dim == query_dim")
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.upcast_attention = False
        self.upcast_softmax = False
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        cdim = query_dim if cross_attention_dim is None else cross_attention_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cdim, inner, bias=bias)
        self.to_v = nn.Linear(cdim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(dropout)])
        self.processor = processor if processor is not None else PaidAttnProcessor()

    def set_processor(self, processor):
        self.processor = processor

    # ---- the helper methods of diffusers' Attention that the REFERENCE processors call (interpolation.py:604, 614,
    # 637-659).  The processors of this package never use them (their kernels read heads in place
    # and never materialise the probabilities); they make this class a faithful host for the unmodified reference
    # processors, which is how oracle/gen_e2e_golden.py produces the end-to-end reference latents.
    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        if attention_mask is not None:
            raise NotImplementedError("attention_mask must be None")
        return None

    def head_to_batch_dim(self, tensor, out_dim=3):
        b, t, c = tensor.shape
        return tensor.reshape(b, t, self.heads, c // self.heads).permute(0, 2, 1, 3).reshape(b * self.heads, t, c // self.heads)

    def batch_to_head_dim(self, tensor):
        bh, t, d = tensor.shape
        return tensor.reshape(bh // self.heads, self.heads, t, d).permute(0, 2, 1, 3).reshape(bh // self.heads, t, d * self.heads)

    def get_attention_scores(self, query, key, attention_mask=None):
        if attention_mask is not None:
            raise NotImplementedError("attention_mask must be None")
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        scores = torch.baddbmm(torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device),
                               query, key.transpose(-1, -2), beta=0, alpha=self.scale)
        if self.upcast_softmax:
            scores = scores.float()
        return scores.softmax(dim=-1).to(dtype)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)
