"""Beta-prior exploration of the interpolation parameter on top of the batched step loop.

Mirrors ``BetaPriorPipeline`` of the reference (prior.py:12-335): starting from ``[0, 0.5, 1]`` the loop repeatedly
bisects -- in the CDF of the current Beta prior -- the gap with the largest perceptual distance, denoises the new
frame, measures its distance to both neighbours and refits the prior to the cumulative distances; finally a subset
of the explored frames with evenly spread distances is picked.

What differs from the reference, deliberately:

* the frame for a new ``t`` comes from ``InterpolationPipeline.interpolate_candidates`` -- and ``batch > 1`` bisects
  the ``batch`` widest gaps of a round in ONE (frame-shardable) batch ``[start, t_1, ..., t_B, end]`` instead of ``B``
  sequential 3-frame denoises (frame i of that batch equals the middle frame of ``interpolate_single(t_i)``,
  SURVEY.md section 4 property 2).  ``batch = 1`` reproduces the reference's sequence of points exactly;
* the perceptual feature extractor is injected (``feature_fn``: latents ``(n, 4, H, W)`` -> features ``(n, F)``): the
  reference hard-wires CLIP ViT-B/32 on decoded images (prior.py:13-33), which needs the VAE and CLIP checkpoints that
  are out of scope here.  Distances are ``1 - cosine similarity`` of the features, computed on
  the device the features live on;
* the subset selection solves the min-spread path problem exactly (two-pointer over the sorted edge weights +
  reachability) instead of the reference's bisection on the spread to 1e-6 (prior.py:222-287); both return the same
  path whenever the optimum is unique by more than that tolerance.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
from scipy.optimize import curve_fit
from scipy.stats import beta as beta_distribution


def feature_distance(a: torch.Tensor, b: torch.Tensor) -> float:
    """``1 - cos(a, b)`` of two feature vectors (prior.py:17-21), in at least fp32."""
    up = lambda t: t.reshape(1, -1).float() if t.dtype in (torch.float16, torch.bfloat16) else t.reshape(1, -1)
    return float(1 - torch.nn.functional.cosine_similarity(up(a), up(b))[0])


def fit_beta_prior(xs: Sequence[float], ds: Sequence[float]) -> Tuple[float, float]:
    """(alpha, beta) whose CDF passes closest to the normalised cumulative distances at the explored points
    (prior.py:35-56): least squares from (1, 1), both parameters kept positive."""
    cum = np.concatenate([[0.0], np.cumsum(np.asarray(ds, dtype=np.float64) / float(sum(ds)))])
    params, _ = curve_fit(lambda x, a, b: beta_distribution.cdf(x, a, b), np.asarray(xs, dtype=np.float64), cum,
                          p0=[1.0, 1.0], bounds=([1e-6, 1e-6], [np.inf, np.inf]))
    return float(params[0]), float(params[1])


def next_point(xs: Sequence[float], ds: Sequence[float], alpha: float, beta: float, uniform: bool = False,
               rank: int = 0) -> Tuple[int, float]:
    """(gap index, t) of the next frame: the midpoint, in the prior's CDF, of the gap with the ``rank``-th largest distance
    (prior.py:74-86); ``uniform``: the plain midpoint of the widest gap in t (prior.py:88-90)."""
    order = np.argsort(-np.asarray(ds, dtype=np.float64), kind="stable")
    idx = int(order[rank])
    lo, hi = beta_distribution.cdf([xs[idx], xs[idx + 1]], alpha, beta)
    t = float(beta_distribution.ppf((lo + hi) / 2, alpha, beta))
    if uniform:
        widths = np.asarray(xs, dtype=np.float64) - np.asarray([0.0] + list(xs[:-1]), dtype=np.float64)
        idx = int(np.argmax(widths)) - 1          # widths[k] = xs[k] - xs[k-1] (widths[0] = 0): left end of the widest gap
        t = (xs[idx] + xs[idx + 1]) / 2
    return idx, t


def uniform_points(ds: Sequence[float], interpolation_size: int) -> List[int]:
    """Greedy subset with roughly equal accumulated distance (prior.py:200-210)."""
    expected = sum(ds) / (interpolation_size - 1)
    acc, out = 0.0, [0]
    for i, d in enumerate(ds):
        acc += d
        if acc >= expected:
            out.append(i)
            acc = 0.0
    return out


def minimal_spread_path(weights: np.ndarray, n: int) -> Tuple[Optional[float], Optional[List[int]]]:
    """Among the index-increasing paths ``0 = i_1 < ... < i_n = m - 1`` the one whose edge weights ``weights[i][j]`` have
    the smallest spread (max - min); returns (spread, path) or (None, None) when no path with n nodes exists
    (prior.py:222-287 searches the same optimum by bisection).  Two-pointer over the sorted distinct weights: for every
    lower bound the smallest feasible upper bound, feasibility = reachability with exactly n nodes inside the window."""
    m = weights.shape[0]
    iu = np.triu_indices(m, 1)
    W = np.unique(weights[iu])

    def path_in(lo: float, hi: float) -> Optional[List[int]]:
        ok = np.triu((weights >= lo) & (weights <= hi), 1)
        reach = np.zeros((n + 1, m), dtype=bool)      # reach[l][j]: a path with l nodes ends in j
        reach[1, 0] = True
        for l in range(1, n):
            reach[l + 1] = (reach[l][:, None] & ok).any(axis=0)
        if not reach[n, m - 1]:
            return None
        path, j = [m - 1], m - 1
        for l in range(n, 1, -1):                      # walk back, smallest predecessor first
            j = int(np.nonzero(reach[l - 1] & ok[:, j])[0][0])
            path.append(j)
        return path[::-1]

    best, best_path, hi_i = None, None, 0
    for lo_i, lo in enumerate(W):
        hi_i = max(hi_i, lo_i)
        found = None
        while hi_i < len(W):
            found = path_in(lo, W[hi_i])
            if found is not None:
                break
            hi_i += 1
        if found is None:
            break                                       # no window starting at or above lo admits a path
        if best is None or W[hi_i] - lo < best:
            best, best_path = float(W[hi_i] - lo), found
    return best, best_path


class BetaPriorExplorer:
    """``BetaPriorPipeline`` (prior.py:12-335) on the batched step loop.

    pipe        an ``InterpolationPipeline`` (or anything with ``interpolate_candidates(ts, **inputs) -> frames`` returning
                the frames ``[start, t_1, ..., t_K, end]``)
    feature_fn  frames ``(n, ...)`` -> features ``(n, F)`` (the reference: CLIP image features of the decoded frames)
    """

    def __init__(self, pipe, feature_fn: Callable[[torch.Tensor], torch.Tensor]):
        self.pipe = pipe
        self.feature_fn = feature_fn

    def _frames(self, ts, inputs):
        return self.pipe.interpolate_candidates(ts, **inputs)

    def explore_with_beta(self, exploration_size: int = 16, init_alpha: float = 3, init_beta: float = 3,
                          uniform: bool = False, batch: int = 1, **inputs):
        """Returns (frames, features, ds, xs, alpha, beta) like prior.py:119-199; ``inputs`` are passed to the step loop
        (latent_start, latent_end, embeds_start, embeds_end, negative_embeds, num_inference_steps, ...)."""
        first = self._frames([0.5], inputs)
        frames = [first[i:i + 1] for i in range(3)]
        features = [f for f in self.feature_fn(first)]
        xs = [0.0, 0.5, 1.0]
        ds = [feature_distance(features[0], features[1]), feature_distance(features[1], features[2])]
        alpha, beta = float(init_alpha), float(init_beta)
        while len(xs) < exploration_size:
            take = max(1, min(batch, exploration_size - len(xs), len(ds)))
            picks = [next_point(xs, ds, alpha, beta, uniform, rank=r) for r in range(1 if uniform else take)]
            if any(t < 0 or t > 1 or not np.isfinite(t) for _, t in picks):
                break                                   # prior.py:92-93
            new = self._frames([t for _, t in picks], inputs)[1:-1]
            new_features = self.feature_fn(new)
            # insert from the rightmost gap to the leftmost so earlier insertions do not shift the later indices
            for k in sorted(range(len(picks)), key=lambda k: -picks[k][0]):
                idx, t = picks[k]
                f = new_features[k]
                d1, d2 = feature_distance(features[idx], f), feature_distance(features[idx + 1], f)
                frames.insert(idx + 1, new[k:k + 1])
                features.insert(idx + 1, f)
                xs.insert(idx + 1, t)
                del ds[idx]
                ds.insert(idx, d1)
                ds.insert(idx + 1, d2)
            alpha, beta = (1.0, 1.0) if uniform else fit_beta_prior(xs, ds)
        return frames, features, ds, xs, alpha, beta

    def extract_uniform_points(self, ds, interpolation_size: int):
        return uniform_points(ds, interpolation_size)

    def extract_uniform_points_plus(self, features, interpolation_size: int):
        """Indices of the ``interpolation_size`` explored frames whose consecutive feature distances are most even
        (prior.py:212-221)."""
        m = len(features)
        weights = -np.ones((m, m))
        for i in range(m):
            for j in range(i + 1, m):
                weights[i, j] = feature_distance(features[i], features[j])
        return minimal_spread_path(weights, interpolation_size)[1]

    def generate_interpolation(self, interpolation_size: int = 7, exploration_size: int = 16, init_alpha: float = 3,
                               init_beta: float = 3, uniform: bool = False, batch: int = 1, **inputs):
        """prior.py:289-335: explore, then keep the evenly spread subset.  Returns the chosen frames ``(n, ...)``."""
        frames, features, ds, xs, alpha, beta = self.explore_with_beta(exploration_size, init_alpha, init_beta, uniform, batch,
                                                                       **inputs)
        self.frames, self.ds, self.xs, self.alpha, self.beta_param = frames, ds, xs, alpha, beta
        idx = self.extract_uniform_points_plus(features, interpolation_size)
        return torch.cat([frames[i] for i in idx], dim=0)
