"""Frame sharding of one interpolation sequence over the GPUs of a box.

The frames of a sequence interact only through the K/V of the two endpoint frames
(reference interpolation.py:627-630: rows 0 and -1 of the batch), and the endpoint
frames depend on nothing else (property 3).  So the N frames are
dealt out to the ranks, BOTH endpoint frames to rank 0, and per interpolated
self-attention call

* rank 0 projects (K_0, V_0, K_{N-1}, V_{N-1}) of its two endpoint frames
  (``paid_attn_project_endpoints``) into the layer's ``(4, S, C)`` buffer and broadcasts
  it -- ONE NCCL broadcast per self-attention layer, the only collective of the path;
* the broadcast runs on a side stream: every rank queues the q/k/v projection of its
  own frames meanwhile, and ``paid_attn_forward`` waits for the transfer
  (``kv_ext_ready_event``) only in front of the attention core.

Cross-attention needs no collective at all: the endpoint prompts are known to every
rank, so each rank projects their K/V locally, once per sequence (the prompts do not
change over the denoising steps).  Deactivated (plain) calls need no communication.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _cabi


def plan_frame_shards(num_frames: int, world_size: int) -> List[List[int]]:
    """Frame ids per rank.  Deal order ``[0, N-1, 1, 2, ..., N-2]`` cut into balanced contiguous pieces (earlier ranks
    take the remainder), so rank 0 holds both endpoint frames as its local frames 0 and 1 and every rank holds at
    least one frame."""
    if num_frames < 2:
        raise ValueError("an interpolation sequence has at least the two endpoint frames")
    if world_size < 1:
        raise ValueError("world_size must be positive")
    if world_size > 1 and world_size > num_frames - 1:
        raise ValueError(f"{num_frames} frames cannot be sharded over {world_size} ranks: rank 0 holds both endpoint "
                         f"frames and every other rank needs at least one (world_size <= {num_frames - 1})")
    order = [0, num_frames - 1] + list(range(1, num_frames - 1))
    base, rem = divmod(num_frames, world_size)
    out, lo = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append(order[lo:lo + n])
        lo += n
    return out


@dataclass
class FrameShard:
    """The frames of an N-frame sequence held by this rank (``frame_ids``, global frame numbers in local order)."""
    rank: int
    world_size: int
    num_frames: int
    group: Optional[object] = None
    overlap: bool = field(default_factory=lambda: os.environ.get("PAID_SHARD_OVERLAP", "1") != "0")
    # ^ broadcast on a side stream (False / PAID_SHARD_OVERLAP=0: on the compute stream, for A/B timing)
    shards: List[List[int]] = field(init=False)
    frame_ids: List[int] = field(init=False)

    def __post_init__(self):
        self.shards = plan_frame_shards(self.num_frames, self.world_size)
        self.frame_ids = self.shards[self.rank]
        self.owner = 0                         # group rank that holds both endpoint frames
        self._index: Dict[torch.device, torch.Tensor] = {}
        self._kv: Dict[tuple, torch.Tensor] = {}           # (layer id, L, C, dtype) -> (4, L, C) endpoint K/V buffer
        self._side: Dict[torch.device, torch.cuda.Stream] = {}
        self.static_endpoints = None           # (2, L, Cc) endpoint prompts of the running conditional pass (pipeline)
        self.broadcasts = 0                    # collectives issued (tests / profiles)

    # ---- frame bookkeeping ------------------------------------------------------------------------
    @property
    def local_frames(self) -> int:
        return len(self.frame_ids)

    @property
    def owns_endpoints(self) -> bool:
        return self.rank == self.owner

    def local(self, t: torch.Tensor) -> torch.Tensor:
        """Rows of a per-frame tensor (N, ...) that belong to this rank, in local order."""
        idx = self._index.get(t.device)
        if idx is None:
            idx = self._index[t.device] = torch.tensor(self.frame_ids, device=t.device, dtype=torch.long)
        return t.index_select(0, idx)

    def unshard(self, parts: List[torch.Tensor]) -> torch.Tensor:
        """Inverse of ``local`` over all ranks: parts[r] are rank r's frames; returns them in global frame order."""
        out = torch.empty(self.num_frames, *parts[0].shape[1:], dtype=parts[0].dtype, device=parts[0].device)
        for ids, p in zip(self.shards, parts):
            out[torch.tensor(ids, device=p.device)] = p
        return out

    # ---- the collective ---------------------------------------------------------------------------
    def _src(self) -> int:
        return self.owner if self.group is None else dist.get_global_rank(self.group, self.owner)

    def side_stream(self, device) -> torch.cuda.Stream:
        s = self._side.get(device)
        if s is None:
            s = self._side[device] = torch.cuda.Stream(device)
        return s

    def kv_buffer(self, key, L: int, C: int, like: torch.Tensor) -> torch.Tensor:
        """Persistent (4, L, C) endpoint buffer of one layer: K_begin, V_begin, K_end, V_end.  One per layer, so a
        broadcast may land while earlier layers still compute, and captured CUDA graphs keep valid pointers."""
        k = (key, L, C, like.dtype, like.device)
        buf = self._kv.get(k)
        if buf is None:
            buf = self._kv[k] = torch.empty(4, L, C, dtype=like.dtype, device=like.device)
        return buf

    def exchange(self, kv: torch.Tensor, ready_on_main: bool):
        """Broadcast ``kv`` from the endpoint owner.  Returns the event the consumer has to wait for (None: the data is
        already ordered on the compute stream).  On CUDA the collective is issued on the side stream."""
        if self.world_size == 1:
            return None
        self.broadcasts += 1
        if not kv.is_cuda or not self.overlap:
            dist.broadcast(kv, src=self._src(), group=self.group)
            return None
        main = torch.cuda.current_stream(kv.device)
        side = self.side_stream(kv.device)
        # Every rank forks the side stream HERE, at the layer that consumes the data.  The owner has to (its projection
        # kernels were queued on the compute stream); a receiver must not post its receive earlier either: an NCCL kernel
        # spins on the GPU until its peer arrives, and 70 receives posted at the start of the forward kept ~8 % of the SMs
        # busy-waiting through the whole forward (profiles/r2_shard_timing.jsonl: 85.9 ms instead of 79.7 per AID forward).
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dist.broadcast(kv, src=self._src(), group=self.group)
            ev = torch.cuda.Event()
            ev.record(side)
        return ev

    def begin_forward(self, device):
        """Called once per UNet forward, before the first layer: the side stream must not run ahead into buffers the
        previous forward may still be reading."""
        if self.world_size > 1 and self.overlap and torch.device(device).type == "cuda":
            self.side_stream(device).wait_stream(torch.cuda.current_stream(device))

    def end_forward(self, device):
        """... and the compute stream joins the side stream at the end (required for CUDA-graph capture: every forked
        stream has to be joined before the capture ends)."""
        if self.world_size > 1 and self.overlap and torch.device(device).type == "cuda":
            torch.cuda.current_stream(device).wait_stream(self.side_stream(device))

    # ---- one interpolated attention call on the local frames --------------------------------------
    def endpoint_kv(self, key, x, ctx, wk, wv, heads: int, flags: int = 0, static_ctx: Optional[torch.Tensor] = None):
        """(kv_ext, event, begin_frame, end_frame) for an interpolated call on this rank's frames.

        Self-attention (ctx None): rank 0 projects its two endpoint frames and broadcasts.  Cross-attention with
        ``static_ctx`` (2, L, Cc) = the endpoint prompts: projected locally, no collective (callers cache the result)."""
        L = x.shape[1] if ctx is None else ctx.shape[1]
        kv = self.kv_buffer(key, L, x.shape[2], x)
        if ctx is not None and static_ctx is not None:
            probe = x[:1]          # only shapes / dtype of x matter for a cross-attention projection
            for f in range(2):
                _cabi.project_endpoints(probe, static_ctx, wk, wv, heads, f, kv[2 * f], kv[2 * f + 1], flags)
            return kv, None, -1, -1
        if self.owns_endpoints:
            for f in range(2):     # local frames 0 and 1 are the sequence's first and last frame
                _cabi.project_endpoints(x, ctx, wk, wv, heads, f, kv[2 * f], kv[2 * f + 1], flags)
        ev = self.exchange(kv, ready_on_main=self.owns_endpoints)
        if self.owns_endpoints:
            return None, None, 0, 1            # its own rows serve: nothing to wait for
        return kv, ev, -1, -1

    def run(self, proc, attn, x, encoder_hidden_states, w, static=None):
        """Interpolated attention of the local frames (called by the text processors)."""
        from .interpolation import _device_coef

        tail = getattr(proc, "cfg_tail", 0)    # guidance rows appended to the local frames (stock attention, same call)
        if x.shape[0] - tail != self.local_frames:
            raise ValueError(f"local batch {x.shape[0]} (of which {tail} guidance rows) != shard size {self.local_frames}")
        if proc.size != self.num_frames:
            raise ValueError(f"processor size {proc.size} != sequence length {self.num_frames}")
        wq, wk, wv, wo, bo = w
        st = static or {}
        if "kv_ext" in st:                     # cross-attention endpoints of this sequence, projected once
            kv, ev, bf, ef = st["kv_ext"], None, -1, -1
        else:
            kv, ev, bf, ef = self.endpoint_kv(id(attn), x, encoder_hidden_states, wk, wv, attn.heads, proc.kernel_flags,
                                              static_ctx=self.static_endpoints if encoder_hidden_states is not None else None)
        if self.owns_endpoints:
            bf, ef = 0, 1                      # rank 0 holds the endpoint frames themselves
        coef = proc.coef_device if getattr(proc, "coef_device", None) is not None else \
            _device_coef(proc.coef[self.frame_ids], x.device)
        return _cabi.attn_forward(
            x, encoder_hidden_states, wq, wk, wv, wo, bo, coef, attn.heads, proc.mode, proc.is_fused, attn.scale,
            begin_frame=bf, end_frame=ef, kv_ext=None if self.owns_endpoints else kv, flags=proc.kernel_flags,
            k_pre=st.get("k"), v_pre=st.get("v"), kv_ext_ready=ev, plain_tail=tail)
