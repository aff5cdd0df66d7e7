"""Generate tests/golden/explore_beta.json by running the UNMODIFIED reference exploration loop (prior.py:12-335).

Test infrastructure.  Run in the build container only (needs /root/reference):
    python oracle/gen_explore_golden.py

How the reference is hosted: ``/root/reference/prior.py`` is imported as-is; its imports of ``bayes_opt``, ``lpips``
(prior.py:3-4, not installed) and of the reference's ``utils`` (matplotlib / lpips) are satisfied by empty stub modules --
none of them is touched by ``BetaPriorPipeline``.  The pipeline object is built without ``__init__`` (which downloads
CLIP), its ``pipe`` is ``SyntheticPipe`` below and ``_get_feature`` returns the synthetic frame itself, so the run
exercises exactly the reference's candidate selection, distance bookkeeping, Beta fit and subset selection:
``explore_with_beta``, ``_add_next_point``, ``_update_alpha_beta``, ``extract_uniform_points``,
``extract_uniform_points_plus`` / ``find_minimal_spread_and_path`` / ``is_path_possible``.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "explore_beta.json")
CASES = [dict(name="beta_prior", seed=3, exploration_size=12, init_alpha=3, init_beta=3, uniform=False, interpolation_size=5),
         dict(name="beta_prior_fast_start", seed=8, exploration_size=10, init_alpha=2, init_beta=5, uniform=False, interpolation_size=6),
         dict(name="uniform", seed=5, exploration_size=9, init_alpha=3, init_beta=3, uniform=True, interpolation_size=4)]


def synthetic_frames(ts, seed: int, dim: int = 48) -> torch.Tensor:
    """A deterministic stand-in for "denoise the frame at t": a curve in feature space that moves unevenly in t
    (s = t^2.2) and bends away from the chord, fp64.  Row i is the frame at ts[i]."""
    g = torch.Generator("cpu").manual_seed(seed)
    a, b, c = (torch.randn(dim, generator=g, dtype=torch.float64) for _ in range(3))
    t = torch.as_tensor(list(ts), dtype=torch.float64)
    s = t ** 2.2
    return (1 - s)[:, None] * a + s[:, None] * b + 0.6 * torch.sin(np.pi * s)[:, None] * c


class SyntheticPipe:
    """``interpolate_single(t, ...)`` -> object with ``.images = [start, frame(t), end]`` (prior.py:95-104, 131-143)."""

    def __init__(self, seed: int):
        self.seed, self.calls = seed, []

    def interpolate_single(self, t, **kw):
        self.calls.append(float(t))
        fr = synthetic_frames([0.0, t, 1.0], self.seed)
        return types.SimpleNamespace(images=[fr[0], fr[1], fr[2]])


def import_reference_prior():
    for name, attrs in (("bayes_opt", ("BayesianOptimization", "SequentialDomainReductionTransformer")), ("lpips", ("LPIPS",)),
                        ("utils", ("compute_lpips", "compute_smoothness_and_consistency"))):
        stub = types.ModuleType(name)
        for a in attrs:
            setattr(stub, a, None)
        sys.modules[name] = stub
    sys.path.insert(0, REF)
    import prior  # the reference module, unmodified

    assert os.path.realpath(prior.__file__).startswith(REF)
    return prior


def main():
    prior = import_reference_prior()
    out = []
    for case in CASES:
        ref = object.__new__(prior.BetaPriorPipeline)          # no __init__: it loads CLIP from the hub
        ref.pipe = SyntheticPipe(case["seed"])
        ref._get_feature = lambda image: image.reshape(1, -1)
        images, features, ds, xs, alpha, beta = ref.explore_with_beta(
            None, None, None, None, None, num_inference_steps=4, exploration_size=case["exploration_size"],
            init_alpha=case["init_alpha"], init_beta=case["init_beta"], uniform=case["uniform"])
        out.append(dict(case, xs=[float(x) for x in xs], ds=[float(d) for d in ds], alpha=float(alpha), beta=float(beta),
                        requested_ts=ref.pipe.calls,
                        uniform_points=[int(i) for i in ref.extract_uniform_points(ds, case["interpolation_size"])],
                        uniform_points_plus=[int(i) for i in ref.extract_uniform_points_plus(features, case["interpolation_size"])]))
        print(case["name"], "xs", np.round(xs, 4).tolist(), "alpha, beta", alpha, beta, "subset", out[-1]["uniform_points_plus"])
    json.dump({"generator": "oracle/gen_explore_golden.py (unmodified reference prior.py, synthetic frames)", "cases": out},
              open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
