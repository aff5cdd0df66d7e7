"""paid-b200: the PAID / AID interpolated-attention hot path as hand-written sm_100a CUDA
behind the reference's diffusers-AttnProcessor plugin surface."""
from . import _cabi
from .attention import Attention, PaidAttnProcessor, PaidIPAdapterAttnProcessor
from .interpolation import (InnerInterpolatedAttnProcessor, InnerInterpolatedIPAttnProcessor,
                            InterpolatedAttnProcessor, OuterInterpolatedAttnProcessor,
                            OuterInterpolatedIPAttnProcessor, ScaleControlIPAttnProcessor)
from .exploration import BetaPriorExplorer
from .prior import generate_beta_tensor
from .sharding import FrameShard, plan_frame_shards

__all__ = [
    "Attention", "PaidAttnProcessor", "InterpolatedAttnProcessor", "OuterInterpolatedAttnProcessor",
    "InnerInterpolatedAttnProcessor", "OuterInterpolatedIPAttnProcessor", "InnerInterpolatedIPAttnProcessor",
    "ScaleControlIPAttnProcessor", "PaidIPAdapterAttnProcessor", "generate_beta_tensor", "BetaPriorExplorer", "FrameShard", "plan_frame_shards", "_cabi",
]
