"""Writes tests/golden/e2e_sd15_c1.npz: final latents of BASELINE configs[0] (SD1.5 64x64 latent, 2 endpoint prompts,
3-frame AID, 10 steps) computed on the CPU in fp32 by the UNMODIFIED reference processors
(/root/reference/interpolation.py OuterInterpolatedAttnProcessor, imported by gen_golden.import_reference) installed in
the UNet harness; the deactivated passes go to a restatement of diffusers' AttnProcessor2_0 (``StockProcessor`` below,
the ``original_attn`` the reference captures at pipeline_interpolated_sdxl.py:1076 -- diffusers is not installed here).
TEST INFRASTRUCTURE: the GPU test test_e2e_sd15_c1_drift compares the CUDA pipeline (fp16) with it.

    python oracle/gen_e2e_golden.py            (about 10 minutes on 8 cores; needs /root/reference)
    python oracle/gen_e2e_golden.py --port     (the oracle port instead of the reference: cross-check, prints the gap)

Weights: default nn init on the CPU under torch.manual_seed(1002) (the reference's seed, gradio_src/app.py:131); the GPU
test rebuilds the same weights on the CPU and moves them to the device.  Inputs: c1_inputs() below."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

C1 = dict(model="sd15", frames=3, t=0.5, steps=10, warmup_ratio=0.5, seed=1002)


def c1_unet_cpu():
    from attention_interpolation_diffusion_b200.unet_harness import build_unet
    return build_unet(C1["model"], "cpu", torch.float32, seed=C1["seed"])


def c1_inputs(dtype=torch.float32):
    g = torch.Generator("cpu").manual_seed(C1["seed"])
    r = lambda *s: torch.randn(*s, generator=g).to(dtype)
    return dict(latent_start=r(1, 4, 64, 64), latent_end=r(1, 4, 64, 64), embeds_start=r(1, 77, 768), embeds_end=r(1, 77, 768),
                negative_embeds=r(1, 77, 768))


def c1_call_kwargs():
    return dict(size=C1["frames"], coef=torch.tensor([0.0, C1["t"], 1.0]), num_inference_steps=C1["steps"],
                warmup_ratio=C1["warmup_ratio"])


class StockProcessor:
    """diffusers AttnProcessor2_0 for the UNet case (3-D input, no mask, no norms): q/k/v projections,
    scaled_dot_product_attention per head, output projection.  [restated: diffusers is not installed]"""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        b, h = hidden_states.shape[0], attn.heads
        q, k, v = attn.to_q(hidden_states), attn.to_k(ctx), attn.to_v(ctx)
        split = lambda t: t.view(b, -1, h, t.shape[-1] // h).transpose(1, 2)
        o = torch.nn.functional.scaled_dot_product_attention(split(q), split(k), split(v), scale=attn.scale)
        return attn.to_out[1](attn.to_out[0](o.transpose(1, 2).reshape(b, -1, q.shape[-1])))


def reference_processor_class():
    """The reference's OuterInterpolatedAttnProcessor with the two methods the N-frame step loop of this repo drives
    (set_coefs / bind_coef_buffer) mapped onto the reference's own ``activate(t)``; ``__call__`` is the reference's."""
    from gen_golden import import_reference
    ref = import_reference()

    class RefProc(ref.OuterInterpolatedAttnProcessor):
        def set_coefs(self, coef):
            assert coef.numel() == 3
            self.activate(float(coef[1]))           # interpolation.py:37-42

        def bind_coef_buffer(self, buf):
            pass
    return RefProc


def main():
    import paid_oracle as O
    from attention_interpolation_diffusion_b200.pipeline import InterpolationPipeline
    torch.set_num_threads(os.cpu_count())
    net = c1_unet_cpu()
    pipe = InterpolationPipeline(net, use_cuda_graphs=False)
    port = "--port" in sys.argv
    if port:
        from oracle_processor import OracleAttnProcessor
        procs = {name: OracleAttnProcessor(O.MODE_OUTER, True, C1["frames"], C1["t"]) for name in net.attn_processors}
    else:
        cls = reference_processor_class()
        procs = {name: cls(t=C1["t"], is_fused=True, original_attn=StockProcessor()) for name in net.attn_processors}
    net.set_attn_processor(procs)
    t0 = time.time()
    out = pipe.interpolate(**c1_inputs(), **c1_call_kwargs())
    print(f"C1 on CPU (fp32, {'oracle port' if port else 'UNMODIFIED reference processors'}): {time.time() - t0:.0f} s on "
          f"{os.cpu_count()} cores, latents {tuple(out.shape)}, rms {float(out.pow(2).mean().sqrt()):.4f}, "
          f"finite {bool(torch.isfinite(out).all())}")
    path = os.path.join(ROOT, "tests", "golden", "e2e_sd15_c1.npz")
    if os.path.exists(path):
        old = torch.from_numpy(np.load(path)["latents"])
        print(f"vs committed fixture: max abs diff {float((out - old).abs().max()):.3e}, rel-RMS "
              f"{float((out - old).pow(2).mean().sqrt() / old.pow(2).mean().sqrt()):.3e}")
    if port:
        return
    np.savez_compressed(path, latents=out.numpy().astype(np.float32), **{k: np.array(v) for k, v in C1.items() if k != "model"},
                        generator=np.array("unmodified /root/reference/interpolation.py OuterInterpolatedAttnProcessor"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
