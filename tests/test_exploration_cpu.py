"""Beta-prior exploration loop against the UNMODIFIED reference loop.

tests/golden/explore_beta.json was produced by /root/reference/prior.py's ``BetaPriorPipeline`` (explore_with_beta,
_add_next_point, _update_alpha_beta, extract_uniform_points, extract_uniform_points_plus) on synthetic frames
(oracle/gen_explore_golden.py); ``BetaPriorExplorer`` must visit the same points, keep the same distances, fit the same
prior and pick the same subsets.  Host logic only: no GPU, no model.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gen_explore_golden import synthetic_frames  # noqa: E402

from attention_interpolation_diffusion_b200.exploration import (BetaPriorExplorer, fit_beta_prior, minimal_spread_path,  # noqa: E402
                                                                next_point, uniform_points)

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "explore_beta.json")))["cases"]


class SyntheticPipe:
    """interpolate_candidates(ts) -> frames [start, t_1, ..., t_K, end] of the synthetic curve."""

    def __init__(self, seed):
        self.seed, self.calls = seed, []

    def interpolate_candidates(self, ts, **kw):
        self.calls.append([float(t) for t in ts])
        return synthetic_frames([0.0, *ts, 1.0], self.seed)


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_explorer_reproduces_the_reference_loop(case):
    pipe = SyntheticPipe(case["seed"])
    ex = BetaPriorExplorer(pipe, feature_fn=lambda frames: frames)
    frames, features, ds, xs, alpha, beta = ex.explore_with_beta(
        exploration_size=case["exploration_size"], init_alpha=case["init_alpha"], init_beta=case["init_beta"],
        uniform=case["uniform"], batch=1)
    assert [c[0] for c in pipe.calls] == pytest.approx(case["requested_ts"], abs=1e-9)     # the same frames were asked for
    assert xs == pytest.approx(case["xs"], abs=1e-9)
    assert ds == pytest.approx(case["ds"], rel=1e-9, abs=1e-12)
    assert (alpha, beta) == pytest.approx((case["alpha"], case["beta"]), rel=1e-7)
    assert len(frames) == len(features) == len(xs) == case["exploration_size"]
    for f, x in zip(frames, xs):                       # frames stay aligned with their parameters
        assert torch.equal(f[0], synthetic_frames([x], case["seed"])[0])
    assert ex.extract_uniform_points(ds, case["interpolation_size"]) == case["uniform_points"]
    assert ex.extract_uniform_points_plus(features, case["interpolation_size"]) == case["uniform_points_plus"]


def test_batched_rounds_bisect_the_widest_gaps_in_one_call():
    """batch = B: one interpolate_candidates call per round carries the B widest gaps' midpoints (one sharded batch instead of
    B sequential 3-frame denoises); bookkeeping stays consistent: xs sorted, ds[i] is the distance between neighbours."""
    pipe = SyntheticPipe(3)
    ex = BetaPriorExplorer(pipe, feature_fn=lambda frames: frames)
    frames, features, ds, xs, alpha, beta = ex.explore_with_beta(exploration_size=12, batch=4)
    assert len(xs) == 12 and xs == sorted(xs) and len(set(xs)) == 12
    assert [len(c) for c in pipe.calls] == [1, 2, 4, 3]          # the first frame, then rounds bounded by the gaps that exist and by the points still wanted
    from attention_interpolation_diffusion_b200.exploration import feature_distance
    for i in range(11):
        assert ds[i] == pytest.approx(feature_distance(features[i], features[i + 1]), rel=1e-12)
        assert torch.equal(frames[i][0], synthetic_frames([xs[i]], 3)[0])
    assert alpha > 0 and beta > 0
    out = ex.generate_interpolation(interpolation_size=5, exploration_size=10, batch=3)
    assert out.shape == (5, 48) and torch.equal(out[0], synthetic_frames([0.0], 3)[0]) and torch.equal(out[-1], synthetic_frames([1.0], 3)[0])


def test_selection_helpers():
    # the Beta-CDF midpoint of the widest-distance gap; rank picks the next widest
    idx, t = next_point([0.0, 0.5, 1.0], [0.1, 0.3], 3, 3)
    assert idx == 1 and 0.5 < t < 1.0
    assert next_point([0.0, 0.5, 1.0], [0.1, 0.3], 3, 3, rank=1)[0] == 0
    assert next_point([0.0, 0.5, 1.0], [0.1, 0.3], 1, 1) == (1, pytest.approx(0.75))
    assert next_point([0.0, 0.25, 1.0], [0.3, 0.1], 3, 3, uniform=True) == (1, pytest.approx(0.625))
    # a prior fitted to distances that ARE a Beta CDF recovers its parameters
    from scipy.stats import beta as B
    xs = np.linspace(0, 1, 9)
    ds = np.diff(B.cdf(xs, 2.5, 4.0))
    assert fit_beta_prior(list(xs), list(ds)) == pytest.approx((2.5, 4.0), rel=1e-5)
    assert uniform_points([1, 1, 1, 1], 3) == [0, 1, 3]
    # min-spread path: exact optimum by brute force on random graphs
    import itertools
    rng = np.random.default_rng(0)
    for m, n in ((6, 3), (7, 4), (8, 5)):
        w = -np.ones((m, m))
        w[np.triu_indices(m, 1)] = rng.random(m * (m - 1) // 2)
        best = min(((max(e) - min(e)), list(p)) for p in ((0, *mid, m - 1) for mid in itertools.combinations(range(1, m - 1), n - 2))
                   for e in [[w[a, b] for a, b in zip(p[:-1], p[1:])]])
        spread, path = minimal_spread_path(w, n)
        assert spread == pytest.approx(best[0]) and path == best[1]
    assert minimal_spread_path(w, 9) == (None, None)            # more nodes than the graph has
