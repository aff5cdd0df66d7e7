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


def test_oracle_processor_hosts_the_reference_semantics_in_the_harness():
    """OracleAttnProcessor (test infrastructure for the end-to-end comparison, oracle/gen_e2e_golden.py): one layer equals
    the oracle's forward_direct, and a whole CPU denoise of the tiny UNet through the pipeline runs with it (activate /
    deactivate / set_coefs driven by the pipeline exactly as for the product processors)."""
    import paid_oracle as O
    from oracle_processor import OracleAttnProcessor
    from attention_interpolation_diffusion_b200.attention import Attention
    from attention_interpolation_diffusion_b200.unet_harness import build_unet
    torch.manual_seed(3)
    attn = Attention(64, 48, 2, 32)
    x, ctx = torch.randn(4, 20, 64), torch.randn(4, 9, 48)
    w = O.LayerWeights(attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_out[0].weight, attn.to_out[0].bias, 2)
    for mode in (O.MODE_OUTER, O.MODE_INNER):
        proc = OracleAttnProcessor(mode, True, size=4)
        proc.set_coefs(torch.tensor([0.0, 0.3, 0.8, 1.0]))
        attn.set_processor(proc)
        with torch.no_grad():
            y = attn(x, encoder_hidden_states=ctx)
            assert torch.allclose(y, O.forward_direct(x, ctx, w, proc.coef, mode, True), atol=1e-5)
            proc.deactivate()
            assert torch.allclose(attn(x, encoder_hidden_states=ctx), O.forward_direct(x, ctx, w, None, O.MODE_PLAIN, False), atol=1e-5)
    net = build_unet("tiny", "cpu", torch.float32, seed=2)
    pipe = InterpolationPipeline(net, use_cuda_graphs=False)
    net.set_attn_processor({n: OracleAttnProcessor(O.MODE_OUTER, True, 3, 0.5) for n in net.attn_processors})
    g = torch.Generator("cpu").manual_seed(4)
    r = lambda *s: torch.randn(*s, generator=g)
    out = pipe.interpolate(latent_start=r(1, 4, 16, 16), latent_end=r(1, 4, 16, 16), embeds_start=r(1, 77, 96),
                           embeds_end=r(1, 77, 96), negative_embeds=r(1, 77, 96), pooled_start=r(1, 1280),
                           pooled_end=r(1, 1280), pooled_negative=r(1, 1280), size=3, coef=torch.tensor([0.0, 0.5, 1.0]),
                           num_inference_steps=4)
    assert out.shape == (3, 4, 16, 16) and torch.isfinite(out).all()


def test_candidates_in_one_batch_equal_sequential_three_frame_runs():
    """Pipeline-level form of property 2, with the reference's processor semantics on the CPU:
    denoising K candidate t's in one batch [start, t_1..t_K, end] gives, frame by frame, the middle frame of the
    reference's sequential 3-frame interpolate_single(t_i) runs (prior.py:119-199 runs them one after the other)."""
    import paid_oracle as O
    from oracle_processor import OracleAttnProcessor
    from attention_interpolation_diffusion_b200.unet_harness import build_unet
    net = build_unet("tiny", "cpu", torch.float64, seed=6)
    pipe = InterpolationPipeline(net, use_cuda_graphs=False)
    net.set_attn_processor({n: OracleAttnProcessor(O.MODE_OUTER, True, 3, 0.5) for n in net.attn_processors})
    g = torch.Generator("cpu").manual_seed(8)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    args = dict(latent_start=r(1, 4, 16, 16), latent_end=r(1, 4, 16, 16), embeds_start=r(1, 77, 96), embeds_end=r(1, 77, 96),
                negative_embeds=r(1, 77, 96), pooled_start=r(1, 1280), pooled_end=r(1, 1280), pooled_negative=r(1, 1280),
                num_inference_steps=4)
    ts = [0.25, 0.6, 0.8]
    batch = pipe.interpolate_candidates(ts, **args)
    assert batch.shape == (5, 4, 16, 16)
    for i, t in enumerate(ts):
        single = pipe.interpolate_single(t, **args)
        assert torch.allclose(batch[i + 1], single[1], atol=1e-9), (t, float((batch[i + 1] - single[1]).abs().max()))
        assert torch.allclose(batch[0], single[0], atol=1e-9) and torch.allclose(batch[-1], single[2], atol=1e-9)
