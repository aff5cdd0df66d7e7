"""CPU: the UNet harness exposes the reference's processor-install surface with the SD1.5 / SDXL call-site
counts and geometries, and the pipeline toggles AID like the reference."""
from collections import Counter

import pytest
import torch

from attention_interpolation_diffusion_b200.interpolation import (InnerInterpolatedAttnProcessor,
                                                                  OuterInterpolatedAttnProcessor)
from attention_interpolation_diffusion_b200.pipeline import DDIMScheduler, InterpolationPipeline, slerp
from attention_interpolation_diffusion_b200.unet_harness import CONFIGS, UNetHarness


def meta_unet(name):
    with torch.device("meta"):
        return UNetHarness(CONFIGS[name])


def test_call_site_counts_and_geometry():
    g = Counter((a["S"], a["L"], a["C"], a["heads"]) for a in meta_unet("sdxl").attention_geometry())
    assert g == {(4096, 4096, 640, 10): 10, (4096, 77, 640, 10): 10, (1024, 1024, 1280, 20): 60, (1024, 77, 1280, 20): 60}
    g = Counter((a["S"], a["L"], a["C"], a["heads"]) for a in meta_unet("sd15").attention_geometry())
    assert sum(g.values()) == 32 and g[(4096, 4096, 320, 8)] == 5 and g[(64, 77, 1280, 8)] == 1
    assert sum(p.numel() for p in meta_unet("sdxl").parameters()) == 2567463684  # SDXL-base UNet size


def test_load_aid_wraps_every_processor():
    net = meta_unet("tiny")
    pipe = InterpolationPipeline(net)
    procs = net.attn_processors
    assert len(procs) == 22 and all(isinstance(p, OuterInterpolatedAttnProcessor) for p in procs.values())
    assert all(p.size == 3 and p.is_fused and p.original_attn is not None for p in procs.values())
    assert all(n.endswith(".processor") for n in procs)
    pipe.deactivate_aid()
    assert not any(p.activated for p in procs.values())
    pipe.activate_aid(0.25)
    assert all(p.activated and float(p.coef[1]) == 0.25 for p in procs.values())
    pipe.load_aid(atype="fused_inner")
    assert all(isinstance(p, InnerInterpolatedAttnProcessor) for p in net.attn_processors.values())
    with pytest.raises(ValueError):
        net.set_attn_processor({"x": None})


def test_scheduler_and_slerp():
    s = DDIMScheduler()
    s.set_timesteps(50)
    assert len(s.timesteps) == 50 and s.timesteps[0] == 981 and s.timesteps[-1] == 1
    import paid_oracle as O
    a, b = torch.randn(1, 4, 8, 8), torch.randn(1, 4, 8, 8)
    assert torch.allclose(slerp(a, b, 0.3), O.slerp(a, b, 0.3), atol=1e-6)
