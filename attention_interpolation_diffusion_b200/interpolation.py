"""Drop-in attention processors for PAID / AID, backed by libpaid_attn (sm_100a CUDA).

Same class names, constructor arguments, state API and ``__call__`` protocol as
the reference's ``interpolation.py``:

* ``InterpolatedAttnProcessor``        reference interpolation.py:10-48
* ``OuterInterpolatedAttnProcessor``   reference interpolation.py:548-679
* ``InnerInterpolatedAttnProcessor``   reference interpolation.py:682-804

so they install with ``unet.set_attn_processor({...})`` exactly like the reference's
(pipeline_interpolated_sdxl.py:1066-1086).  The body of ``__call__`` is ONE call into
the C ABI (``paid_attn_forward``): projections, endpoint attentions, alpha-lerp,
self-attention fusion and output projection all run in the library's kernels.  There
is no PyTorch / CPU fallback: unsupported situations raise.

Differences from the reference, all deliberate:
* ``coef`` stays fp32 on the device (the reference rounds it to the model dtype and
  re-uploads it from the host on every call, interpolation.py:662-663).
* ``set_coefs`` lets an N-frame processor change its coefficients (the reference's
  ``activate(t)`` can only install the 3-entry ``[0, t, 1]``).
* ``shard`` (a ``FrameShard``) runs the call on a slice of the frames, with the two
  endpoint K/V broadcast over NCCL (sharding.py); the reference has no multi-GPU path.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _cabi
from .attention import PaidAttnProcessor, check_unet_preconditions
from .prior import generate_beta_tensor

_coef_cache: dict = {}


def _device_coef(coef: torch.Tensor, device: torch.device) -> torch.Tensor:
    """fp32 device copy of a coefficient vector, shared by every processor holding the
    same values (one H2D copy per distinct schedule instead of one per layer per call)."""
    key = (tuple(float(v) for v in coef.tolist()), str(device))
    t = _coef_cache.get(key)
    if t is None:
        if len(_coef_cache) > 4096:
            _coef_cache.clear()
        t = coef.detach().to(device=device, dtype=torch.float32).contiguous()
        _coef_cache[key] = t
    return t


class InterpolatedAttnProcessor(nn.Module):
    mode = _cabi.PAID_PLAIN

    def __init__(self, t: Optional[float] = None, size: int = 7, is_fused: bool = False, alpha: float = 1,
                 beta: float = 1):
        super().__init__()
        if t is None:
            ts = generate_beta_tensor(size, alpha=alpha, beta=beta)
            ts[0], ts[-1] = 0, 1
        else:
            assert t > 0 and t < 1, "t must be between 0 and 1"
            ts = torch.tensor([0, t, 1])
            size = 3
        self.size = size
        self.coef = ts
        self.is_fused = is_fused
        self.activated = True
        self.shard = None            # optional sharding.FrameShard
        self.kernel_flags = 0        # _cabi.FLAG_* (tests use FLAG_GENERIC_KERNELS as a cross-check)

    def deactivate(self):
        self.activated = False

    def activate(self, t):
        self.activated = True
        assert t > 0 and t < 1, "t must be between 0 and 1"
        self.coef = torch.tensor([0, t, 1])

    def set_coefs(self, coef: torch.Tensor):
        """N-frame extension: install a full coefficient vector (ends are forced to 0 / 1)."""
        coef = coef.detach().to("cpu", torch.float32).clone()
        coef[0], coef[-1] = 0, 1
        self.size = coef.numel()
        self.coef = coef
        self.activated = True

    def load_end_point(self, key_begin, value_begin, key_end, value_end):
        # kept for API parity; like in the reference, nothing reads these
        self.key_begin, self.value_begin, self.key_end, self.value_end = key_begin, value_begin, key_end, value_end

    # ------------------------------------------------------------------------------------------
    def _plain(self, attn, hidden_states, encoder_hidden_states, attention_mask, temb):
        original = getattr(self, "original_attn", None)
        if original is None:
            original = PaidAttnProcessor()
        return original(attn, hidden_states, encoder_hidden_states, attention_mask, temb)

    def _interpolated(self, attn, hidden_states, encoder_hidden_states, attention_mask):
        check_unet_preconditions(attn, hidden_states, attention_mask)
        x = hidden_states
        w = (attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_out[0].weight, attn.to_out[0].bias)
        if self.shard is not None:
            return self.shard.run(self, attn, x, encoder_hidden_states, w)
        if x.shape[0] != self.size or self.coef.numel() != self.size:
            raise ValueError(f"batch size {x.shape[0]} / {self.coef.numel()} coefficients != processor size {self.size} "
                             "(the frames of one interpolation sequence must form the batch)")
        coef = _device_coef(self.coef, x.device)
        return _cabi.attn_forward(x, encoder_hidden_states, *w, coef, attn.heads, self.mode, self.is_fused, attn.scale,
                                  flags=self.kernel_flags)

    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, temb: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not self.activated:
            return self._plain(attn, hidden_states, encoder_hidden_states, attention_mask, temb)
        return self._interpolated(attn, hidden_states, encoder_hidden_states, attention_mask)


class OuterInterpolatedAttnProcessor(InterpolatedAttnProcessor):
    r"""Outer attention interpolation: for frame t with coefficient c_t

        (1 - c_t) * Attn(Q_t, K_1, V_1) + c_t * Attn(Q_t, K_m, V_m)

    and, fused with self-attention (``is_fused``),

        (1 - c_t) * Attn(Q_t, [K_t; K_1], [V_t; V_1]) + c_t * Attn(Q_t, [K_t; K_m], [V_t; V_m]).
    """
    mode = _cabi.PAID_OUTER

    def __init__(self, t: Optional[float] = None, size: int = 7, is_fused: bool = False, alpha: float = 1,
                 beta: float = 1, original_attn=None):
        super().__init__(t=t, size=size, is_fused=is_fused, alpha=alpha, beta=beta)
        self.original_attn = original_attn


class InnerInterpolatedAttnProcessor(InterpolatedAttnProcessor):
    r"""Inner attention interpolation: the endpoint keys / values are interpolated first,

        Attn(Q_t, (1 - c_t) K_1 + c_t K_m, (1 - c_t) V_1 + c_t V_m)

    and, fused with self-attention, the frame's own K_t / V_t are appended to them.
    """
    mode = _cabi.PAID_INNER

    def __init__(self, t: Optional[float] = None, size: int = 7, is_fused: bool = False, alpha: float = 1,
                 beta: float = 1, original_attn=None):
        super().__init__(t=t, size=size, is_fused=is_fused, alpha=alpha, beta=beta)
        self.original_attn = original_attn
