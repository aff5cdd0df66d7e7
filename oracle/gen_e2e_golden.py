"""Writes tests/golden/e2e_sd15_c1.npz: final latents of BASELINE configs[0] (SD1.5 64x64 latent, 2 endpoint prompts,
3-frame AID, 10 steps) computed on the CPU in fp32 with the reference's processor semantics (OracleAttnProcessor) inside
the UNet harness.  TEST INFRASTRUCTURE: the GPU test test_e2e_sd15_c1_drift compares the CUDA pipeline (fp16) with it.

    python oracle/gen_e2e_golden.py            (about 10 minutes on 8 cores)

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


def main():
    import paid_oracle as O
    from oracle_processor import OracleAttnProcessor
    from attention_interpolation_diffusion_b200.pipeline import InterpolationPipeline
    torch.set_num_threads(os.cpu_count())
    net = c1_unet_cpu()
    pipe = InterpolationPipeline(net, use_cuda_graphs=False)
    net.set_attn_processor({name: OracleAttnProcessor(O.MODE_OUTER, True, C1["frames"], C1["t"]) for name in net.attn_processors})
    t0 = time.time()
    out = pipe.interpolate(**c1_inputs(), **c1_call_kwargs())
    print(f"C1 on CPU (fp32, reference processor semantics): {time.time() - t0:.0f} s, latents {tuple(out.shape)}, "
          f"rms {float(out.pow(2).mean().sqrt()):.4f}, finite {bool(torch.isfinite(out).all())}")
    path = os.path.join(ROOT, "tests", "golden", "e2e_sd15_c1.npz")
    np.savez_compressed(path, latents=out.numpy().astype(np.float32), **{k: np.array(v) for k, v in C1.items() if k != "model"})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
