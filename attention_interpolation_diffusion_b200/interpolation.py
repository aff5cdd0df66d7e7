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
from .attention import (PaidAttnProcessor, check_unet_preconditions, project_text_static, split_ip_states, static_kv)
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
        self.cfg_tail = 0            # frames appended to the batch that get stock attention in the same call: the step loop
                                     # runs the unconditional pass of a warm-up step as the tail of the conditional one
        self.coef_device = None      # optional fp32 device buffer holding the coefficients of the LOCAL frames; the step
                                     # loop binds one buffer to every processor and rewrites it in place, so captured
                                     # CUDA graphs serve any schedule (bind_coef_buffer)

    def bind_coef_buffer(self, buf: Optional[torch.Tensor]):
        self.coef_device = buf

    def _coef_on(self, device, local_ids=None) -> torch.Tensor:
        """fp32 device coefficients of the local frames: the bound buffer, else a cached upload of ``self.coef``."""
        if self.coef_device is not None:
            return self.coef_device
        return _device_coef(self.coef if local_ids is None else self.coef[local_ids], device)

    def project_static(self, attn, ctx, uniform, entry, endpoints=None):
        """Per-sequence K / V of a step-invariant cross-attention context (attention.static_kv)."""
        project_text_static(attn, ctx, uniform, entry, endpoints, self.kernel_flags)

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
        st = static_kv(attn, encoder_hidden_states)
        if self.shard is not None:
            return self.shard.run(self, attn, x, encoder_hidden_states, w, static=st)
        if x.shape[0] - self.cfg_tail != self.size or self.coef.numel() != self.size:
            raise ValueError(f"batch size {x.shape[0]} (of which {self.cfg_tail} guidance rows) / {self.coef.numel()} "
                             f"coefficients != processor size {self.size} (the frames of one interpolation sequence must "
                             "form the batch)")
        st = st or {}
        return _cabi.attn_forward(x, encoder_hidden_states, *w, self._coef_on(x.device), attn.heads, self.mode, self.is_fused,
                                  attn.scale, flags=self.kernel_flags, k_pre=st.get("k"), v_pre=st.get("v"),
                                  plain_tail=self.cfg_tail)

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

    def project_static(self, attn, ctx, uniform, entry, endpoints=None):
        """Text and image-token K / V of the sequence (both step-invariant), and for a frame-sharded sequence the endpoint
        frames' K / V of both, projected locally from the endpoint contexts every rank holds."""
        text, ip = split_ip_states(ctx, self.num_tokens, ctx.shape[0])
        et = ei = None
        if endpoints is not None:
            et, ei = split_ip_states(endpoints, self.num_tokens, 2)
        project_text_static(attn, text, uniform, entry, et, self.kernel_flags)
        project_text_static(attn, ip, uniform, entry, ei, self.kernel_flags, prefix="ip_",
                            wk=self.ip_attn.to_k_ip[0].weight, wv=self.ip_attn.to_v_ip[0].weight)

    def _parts(self, attn, hidden_states, encoder_hidden_states, attention_mask):
        """(x, st, coef, q, k, v, kip, vip): projections of the call; K / V from the per-sequence cache when attached."""
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
        coef = self._coef_on(x.device, None if sh is None else sh.frame_ids)
        q = _cabi.linear(x, attn.to_q.weight, flags=self.kernel_flags)
        st = static_kv(attn, encoder_hidden_states)
        if st is not None:
            return x, st, coef, q, st["k"], st["v"], st.get("ip_k"), st.get("ip_v")
        text, ip = (None, None) if encoder_hidden_states is None else split_ip_states(
            encoder_hidden_states, self.num_tokens, frames)
        src = x if text is None else text
        k = _cabi.linear(src, attn.to_k.weight, flags=self.kernel_flags)
        v = _cabi.linear(src, attn.to_v.weight, flags=self.kernel_flags)
        kip = vip = None
        if ip is not None:
            kip, vip = self._ip_kv(ip)
        return x, {}, coef, q, k, v, kip, vip

    def _endpoints(self, attn, k, v, st: dict, key: str = "kv_ext", cross: bool = False):
        """Frame-sharded call: keyword arguments for ``attn_core`` that carry the endpoint K/V of the whole sequence.
        Rank 0 holds both endpoint frames (its local frames 0 and 1).  Cross-attention endpoints come from the
        per-sequence cache (projected locally on every rank, no collective); self-attention endpoints are rows of rank
        0's K / V, broadcast to the others -- the one collective of the path (sharding.py).  Unsharded: the endpoints
        are rows 0 and -1 of the batch (no arguments)."""
        sh = self.shard
        if sh is None:
            return {}
        if sh.owns_endpoints and (cross or sh.world_size == 1):
            return dict(begin_frame=0, end_frame=1)
        if cross:
            if key not in st:
                raise RuntimeError("a frame-sharded IP-Adapter cross-attention call needs the per-sequence K/V cache with "
                                   "the endpoint contexts (InterpolationPipeline attaches it)")
            return dict(kv_ext=st[key], begin_frame=-1, end_frame=-1)
        kv = sh.kv_buffer(id(attn), k.shape[1], k.shape[2], k)
        if sh.owns_endpoints:
            kv[0].copy_(k[0]), kv[1].copy_(v[0]), kv[2].copy_(k[1]), kv[3].copy_(v[1])
        ev = sh.exchange(kv, ready_on_main=sh.owns_endpoints)
        if sh.owns_endpoints:
            return dict(begin_frame=0, end_frame=1)
        if ev is not None:
            torch.cuda.current_stream(k.device).wait_event(ev)
        return dict(kv_ext=kv, begin_frame=-1, end_frame=-1)

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
        cross = encoder_hidden_states is not None
        x, st, coef, q, k, v, kip, vip = self._parts(attn, hidden_states, encoder_hidden_states, attention_mask)
        hid = _cabi.attn_core(q, k, v, coef, attn.heads, _cabi.PAID_OUTER, self.is_fused, attn.scale, flags=self.kernel_flags,
                              **self._endpoints(attn, k, v, st, "kv_ext", cross))
        if kip is not None:
            _cabi.attn_core(q, kip, vip, coef, attn.heads, _cabi.PAID_OUTER, self.is_fused, attn.scale, flags=self.kernel_flags,
                            out=hid, accumulate=True, out_scale=float(self.scale[0]),
                            **self._endpoints(attn, kip, vip, st, "ip_kv_ext", True))
        return self._out(attn, hid)


class InnerInterpolatedIPAttnProcessor(_InterpolatedIPAttnProcessor):
    r"""Inner interpolation of the text attention plus ``scale[0]`` times the attention over each frame's OWN image
    tokens -- the reference computes lerped image K/V but then attends with ``key`` / ``value`` of the frame itself
    (interpolation.py:512-527); that behaviour is kept (reference interpolation.py:390-545)."""
    mode = _cabi.PAID_INNER

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        if not self.activated:
            return self.ip_attn(attn, hidden_states, encoder_hidden_states, attention_mask, temb)
        cross = encoder_hidden_states is not None
        x, st, coef, q, k, v, kip, vip = self._parts(attn, hidden_states, encoder_hidden_states, attention_mask)
        hid = _cabi.attn_core(q, k, v, coef, attn.heads, _cabi.PAID_INNER, self.is_fused, attn.scale, flags=self.kernel_flags,
                              **self._endpoints(attn, k, v, st, "kv_ext", cross))
        if kip is not None:
            _cabi.attn_core(q, kip, vip, None, attn.heads, _cabi.PAID_PLAIN, False, attn.scale, flags=self.kernel_flags,
                            out=hid, accumulate=True, out_scale=float(self.scale[0]))
        return self._out(attn, hid)


class ScaleControlIPAttnProcessor(_InterpolatedIPAttnProcessor):
    r"""Image-prompt strength control: text attention (outer-interpolated while activated, plain otherwise) plus
    ``coef[n]`` times the attention over the END frame's image tokens (reference interpolation.py:51-211)."""
    mode = _cabi.PAID_OUTER

    def project_static(self, attn, ctx, uniform, entry, endpoints=None):
        super().project_static(attn, ctx, uniform, entry, endpoints)
        # the end image for every frame (reference: ip[0][6:9]): the last frame of the batch, or, frame-sharded, the
        # end-frame context every rank holds
        text, ip = split_ip_states(ctx if endpoints is None else endpoints, self.num_tokens, (ctx if endpoints is None else endpoints).shape[0])
        project_text_static(attn, ip[-1:], True, entry, None, self.kernel_flags, prefix="ip_end_",
                            wk=self.ip_attn.to_k_ip[0].weight, wv=self.ip_attn.to_v_ip[0].weight)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        cross = encoder_hidden_states is not None
        x, st, coef, q, k, v, kip, vip = self._parts(attn, hidden_states, encoder_hidden_states, attention_mask)
        if self.activated:
            hid = _cabi.attn_core(q, k, v, coef, attn.heads, _cabi.PAID_OUTER, self.is_fused, attn.scale, flags=self.kernel_flags,
                                  **self._endpoints(attn, k, v, st, "kv_ext", cross))
        else:
            hid = _cabi.attn_core(q, k, v, None, attn.heads, _cabi.PAID_PLAIN, False, attn.scale, flags=self.kernel_flags,
                                  kv_broadcast=st.get("broadcast", False))
        if kip is not None:
            if "ip_end_k" in st:
                kend, vend = st["ip_end_k"], st["ip_end_v"]
            elif self.shard is not None and self.shard.world_size > 1:
                raise RuntimeError("a frame-sharded scale-control call needs the per-sequence K/V cache (the end frame's "
                                   "image tokens live on rank 0)")
            else:                      # the end image for every frame (ip[0][6:9]): the last frame of the batch
                end = 1 if self.shard is not None else kip.shape[0] - 1
                kend, vend = kip[end:end + 1].contiguous(), vip[end:end + 1].contiguous()
            _cabi.attn_core(q, kend, vend, None, attn.heads, _cabi.PAID_PLAIN, False, attn.scale, flags=self.kernel_flags,
                            out=hid, accumulate=True, out_frame_scale=coef, kv_broadcast=True)
        return self._out(attn, hid)
