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
from .attention import PaidAttnProcessor, check_unet_preconditions, split_ip_states
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


# ------------------------------------------------------------------------------------------------------
# IP-Adapter variants (reference interpolation.py:51-545; installed by load_aid_ip_adapter, sdxl:1089-1126).
# The text part is the processor call above; the image-token part is a second attention of the SAME queries,
# accumulated into the text result by the attention kernel itself (PaidCoreParams.accumulate / out_scale).
# The reference hard-codes a batch of 3 (expand(3, ...), [::3], [6:9]); here any number of frames works and
# the image tokens are passed per frame as (N, T, Cc) (the reference's 3x-repeated rows are accepted too).
# ------------------------------------------------------------------------------------------------------
class _InterpolatedIPAttnProcessor(InterpolatedAttnProcessor):
    text_mode = _cabi.PAID_OUTER

    def __init__(self, t: Optional[float] = None, size: int = 7, is_fused: bool = False, alpha: float = 1,
                 beta: float = 1, ip_attn=None):
        super().__init__(t=t, size=size, is_fused=is_fused, alpha=alpha, beta=beta)
        self.num_tokens = ip_attn.num_tokens if hasattr(ip_attn, "num_tokens") else (16,)
        self.scale = ip_attn.scale if hasattr(ip_attn, "scale") else None
        self.ip_attn = ip_attn

    def _parts(self, attn, hidden_states, encoder_hidden_states, attention_mask):
        check_unet_preconditions(attn, hidden_states, attention_mask)
        x = hidden_states
        sh = self.shard
        frames = x.shape[0]
        if sh is not None:
            if frames != sh.local_frames or self.size != sh.num_frames:
                raise ValueError(f"local batch {frames} / processor size {self.size} do not match the shard "
                                 f"({sh.local_frames} of {sh.num_frames} frames)")
        elif frames != self.size:
            raise ValueError(f"batch size {frames} != processor size {self.size}")
        if self.coef.numel() != self.size:
            raise ValueError(f"{self.coef.numel()} coefficients != processor size {self.size}")
        text, ip = (None, None) if encoder_hidden_states is None else split_ip_states(
            encoder_hidden_states, self.num_tokens, frames)
        coef = _device_coef(self.coef if sh is None else self.coef[sh.lo:sh.hi], x.device)
        src = x if text is None else text
        q = _cabi.linear(x, attn.to_q.weight, flags=self.kernel_flags)
        k = _cabi.linear(src, attn.to_k.weight, flags=self.kernel_flags)
        v = _cabi.linear(src, attn.to_v.weight, flags=self.kernel_flags)
        return x, ip, coef, q, k, v

    def _endpoints(self, k, v, need_begin: bool = True):
        """Frame-sharded call: keyword arguments for ``attn_core`` that carry the endpoint K/V of the whole sequence
        (rows of the local k / v on the ranks that own frame 0 / N-1, broadcast to the others -- the one collective of
        the path, sharding.py).  Unsharded: the endpoints are rows 0 and -1 of the batch (no arguments)."""
        sh = self.shard
        if sh is None:
            return {}
        from .sharding import broadcast_endpoints
        kv = torch.empty(4, k.shape[1], k.shape[2], dtype=k.dtype, device=k.device)
        own_b, own_e = sh.rank == sh.begin_owner, sh.rank == sh.end_owner
        if own_b and need_begin:
            kv[0].copy_(k[0]), kv[1].copy_(v[0])
        if own_e:
            kv[2].copy_(k[-1]), kv[3].copy_(v[-1])
        if not need_begin:
            kv[0:2].zero_()            # never read by the kernel's consumers; keep the buffer defined
        if sh.world_size > 1:
            if need_begin:
                broadcast_endpoints(kv, sh.begin_owner, sh.end_owner, sh.group)
            else:
                torch.distributed.broadcast(kv[2:4], src=sh.end_owner, group=sh.group)
        return dict(kv_ext=kv, begin_frame=0 if own_b else -1, end_frame=k.shape[0] - 1 if own_e else -1)

    def _ip_kv(self, ip):
        return (_cabi.linear(ip, self.ip_attn.to_k_ip[0].weight, flags=self.kernel_flags),
                _cabi.linear(ip, self.ip_attn.to_v_ip[0].weight, flags=self.kernel_flags))

    def _out(self, attn, hid):
        return _cabi.linear(hid, attn.to_out[0].weight, attn.to_out[0].bias, flags=self.kernel_flags)


class OuterInterpolatedIPAttnProcessor(_InterpolatedIPAttnProcessor):
    r"""Outer interpolation of the text attention plus ``scale[0]`` times the outer interpolation of the image-token
    attention (reference interpolation.py:214-387)."""
    mode = _cabi.PAID_OUTER

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        if not self.activated:
            return self.ip_attn(attn, hidden_states, encoder_hidden_states, attention_mask, temb)
        x, ip, coef, q, k, v = self._parts(attn, hidden_states, encoder_hidden_states, attention_mask)
        hid = _cabi.attn_core(q, k, v, coef, attn.heads, _cabi.PAID_OUTER, self.is_fused, attn.scale, flags=self.kernel_flags,
                              **self._endpoints(k, v))
        if ip is not None:
            kip, vip = self._ip_kv(ip)
            _cabi.attn_core(q, kip, vip, coef, attn.heads, _cabi.PAID_OUTER, self.is_fused, attn.scale, flags=self.kernel_flags,
                            out=hid, accumulate=True, out_scale=float(self.scale[0]), **self._endpoints(kip, vip))
        return self._out(attn, hid)


class InnerInterpolatedIPAttnProcessor(_InterpolatedIPAttnProcessor):
    r"""Inner interpolation of the text attention plus ``scale[0]`` times the attention over each frame's OWN image
    tokens -- the reference computes lerped image K/V but then attends with ``key`` / ``value`` of the frame itself
    (interpolation.py:512-527); that behaviour is kept (reference interpolation.py:390-545)."""
    mode = _cabi.PAID_INNER

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        if not self.activated:
            return self.ip_attn(attn, hidden_states, encoder_hidden_states, attention_mask, temb)
        x, ip, coef, q, k, v = self._parts(attn, hidden_states, encoder_hidden_states, attention_mask)
        hid = _cabi.attn_core(q, k, v, coef, attn.heads, _cabi.PAID_INNER, self.is_fused, attn.scale, flags=self.kernel_flags,
                              **self._endpoints(k, v))
        if ip is not None:
            kip, vip = self._ip_kv(ip)
            _cabi.attn_core(q, kip, vip, None, attn.heads, _cabi.PAID_PLAIN, False, attn.scale, flags=self.kernel_flags,
                            out=hid, accumulate=True, out_scale=float(self.scale[0]))
        return self._out(attn, hid)


class ScaleControlIPAttnProcessor(_InterpolatedIPAttnProcessor):
    r"""Image-prompt strength control: text attention (outer-interpolated while activated, plain otherwise) plus
    ``coef[n]`` times the attention over the END frame's image tokens (reference interpolation.py:51-211)."""
    mode = _cabi.PAID_OUTER

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        x, ip, coef, q, k, v = self._parts(attn, hidden_states, encoder_hidden_states, attention_mask)
        if self.activated:
            hid = _cabi.attn_core(q, k, v, coef, attn.heads, _cabi.PAID_OUTER, self.is_fused, attn.scale, flags=self.kernel_flags,
                                  **self._endpoints(k, v))
        else:
            hid = _cabi.attn_core(q, k, v, None, attn.heads, _cabi.PAID_PLAIN, False, attn.scale, flags=self.kernel_flags)
        if ip is not None:
            kip, vip = self._ip_kv(ip[-1:].contiguous())          # the end image for every frame (ip[0][6:9])
            if self.shard is not None:                            # ... which lives on the rank that owns frame N-1
                kv = self._endpoints(kip, vip, need_begin=False)["kv_ext"]
                kip, vip = kv[2:3], kv[3:4]
            _cabi.attn_core(q, kip, vip, None, attn.heads, _cabi.PAID_PLAIN, False, attn.scale, flags=self.kernel_flags,
                            out=hid, accumulate=True, out_frame_scale=coef, kv_broadcast=True)
        return self._out(attn, hid)
