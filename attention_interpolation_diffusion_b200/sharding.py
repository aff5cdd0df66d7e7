"""Frame sharding of one interpolation sequence over the GPUs of a box.

The frames of a sequence interact only through the K/V of the two endpoint frames
(reference interpolation.py:627-630: rows 0 and -1 of the batch), and the endpoint
frames depend on nothing else (property 3).  So rank r owns a
contiguous slice of the N frames; per interpolated attention call the owner of frame 0
projects (K_0, V_0), the owner of frame N-1 projects (K_{N-1}, V_{N-1})
(``paid_attn_project_endpoints``), the two pairs are broadcast over NCCL and every rank
runs ``paid_attn_forward`` on its slice with ``kv_ext``.  That broadcast is the only
collective of the path; deactivated (plain) calls need none.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _cabi


def plan_frame_shards(num_frames: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [lo, hi) slices; earlier ranks take the remainder.  Ranks beyond the
    frame count get an empty slice."""
    if num_frames < 2:
        raise ValueError("an interpolation sequence has at least the two endpoint frames")
    base, rem = divmod(num_frames, world_size)
    out, lo = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append((lo, lo + n))
        lo += n
    return out


def endpoint_owners(shards: List[Tuple[int, int]], num_frames: int) -> Tuple[int, int]:
    begin = next(r for r, (lo, hi) in enumerate(shards) if lo <= 0 < hi)
    end = next(r for r, (lo, hi) in enumerate(shards) if lo <= num_frames - 1 < hi)
    return begin, end


def broadcast_endpoints(kv: torch.Tensor, begin_owner: int, end_owner: int, group=None):
    """kv (4, L, C) = K_begin, V_begin, K_end, V_end.  Rows 0:2 are valid on begin_owner, rows 2:4 on
    end_owner; afterwards all four are valid everywhere.  One broadcast when one rank owns both."""
    if begin_owner == end_owner:
        dist.broadcast(kv, src=begin_owner, group=group)
    else:
        dist.broadcast(kv[0:2], src=begin_owner, group=group)
        dist.broadcast(kv[2:4], src=end_owner, group=group)


@dataclass
class FrameShard:
    """Slice [lo, hi) of an N-frame sequence held by this rank."""
    rank: int
    world_size: int
    num_frames: int
    group: Optional[object] = None

    def __post_init__(self):
        self.shards = plan_frame_shards(self.num_frames, self.world_size)
        self.lo, self.hi = self.shards[self.rank]
        self.begin_owner, self.end_owner = endpoint_owners(self.shards, self.num_frames)

    @property
    def local_frames(self) -> int:
        return self.hi - self.lo

    def local(self, t: torch.Tensor) -> torch.Tensor:
        """Slice a per-frame tensor (N, ...) to this rank's frames."""
        return t[self.lo:self.hi]

    def run(self, proc, attn, x, encoder_hidden_states, w):
        """Interpolated attention of the local frames (called by the processors)."""
        from .interpolation import _device_coef

        if x.shape[0] != self.local_frames:
            raise ValueError(f"local batch {x.shape[0]} != shard size {self.local_frames}")
        if proc.size != self.num_frames:
            raise ValueError(f"processor size {proc.size} != sequence length {self.num_frames}")
        wq, wk, wv, wo, bo = w
        L = x.shape[1] if encoder_hidden_states is None else encoder_hidden_states.shape[1]
        kv = torch.empty(4, L, x.shape[2], dtype=x.dtype, device=x.device)
        own_b, own_e = self.rank == self.begin_owner, self.rank == self.end_owner
        if own_b:
            _cabi.project_endpoints(x, encoder_hidden_states, wk, wv, attn.heads, 0, kv[0], kv[1], proc.kernel_flags)
        if own_e:
            _cabi.project_endpoints(x, encoder_hidden_states, wk, wv, attn.heads, self.local_frames - 1, kv[2], kv[3],
                                    proc.kernel_flags)
        if self.world_size > 1:
            broadcast_endpoints(kv, self.begin_owner, self.end_owner, self.group)
        coef = _device_coef(proc.coef[self.lo:self.hi], x.device)
        return _cabi.attn_forward(
            x, encoder_hidden_states, wq, wk, wv, wo, bo, coef, attn.heads, proc.mode, proc.is_fused, attn.scale,
            begin_frame=0 if own_b else -1, end_frame=self.local_frames - 1 if own_e else -1, kv_ext=kv,
            flags=proc.kernel_flags)
