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


class PaidAttnProcessor:
    """Stock (non-interpolated) attention through the PLAIN mode of libpaid_attn --
    the role ``AttnProcessor2_0`` plays in the reference (``original_attn``,
    pipeline_interpolated_sdxl.py:1076)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        check_unet_preconditions(attn, hidden_states, attention_mask)
        return _cabi.attn_forward(
            hidden_states, encoder_hidden_states, attn.to_q.weight, attn.to_k.weight, attn.to_v.weight,
            attn.to_out[0].weight, attn.to_out[0].bias, None, attn.heads, _cabi.PAID_PLAIN, False, attn.scale)


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

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        check_unet_preconditions(attn, hidden_states, attention_mask)
        x = hidden_states
        text, ip = split_ip_states(encoder_hidden_states, self.num_tokens, x.shape[0])
        q = _cabi.linear(x, attn.to_q.weight)
        hid = _cabi.attn_core(q, _cabi.linear(text, attn.to_k.weight), _cabi.linear(text, attn.to_v.weight), None,
                              attn.heads, _cabi.PAID_PLAIN, False, attn.scale)
        _cabi.attn_core(q, _cabi.linear(ip, self.to_k_ip[0].weight), _cabi.linear(ip, self.to_v_ip[0].weight), None,
                        attn.heads, _cabi.PAID_PLAIN, False, attn.scale, out=hid, accumulate=True,
                        out_scale=float(self.scale[0]))
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

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)
