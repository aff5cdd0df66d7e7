"""CPU oracle for the PAID / AID interpolated-attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never on the product path.

What it restates (reference = /root/reference, commit dfa4e13a):

* ``interpolation.py:548-679``  OuterInterpolatedAttnProcessor.__call__
* ``interpolation.py:682-804``  InnerInterpolatedAttnProcessor.__call__
* ``interpolation.py:581-584``  deactivated branch -> stock attention
* ``prior.py:481-502``          generate_beta_tensor
* ``interpolation.py:861-918``  slerp (harness input construction only)

Half of the arithmetic of that path lives in a third-party dependency that is
NOT vendored in the reference and NOT installed in this image:
``diffusers==0.27.1`` (requirements.txt:11) -- ``Attention.get_attention_scores``,
``head_to_batch_dim``, ``batch_to_head_dim``, the ``to_q/to_k/to_v/to_out``
Linear layers and ``AttnProcessor2_0``.  Their published algorithm
(softmax(scale * q k^T) v on (B*h, T, d) views; Linear without bias for
q/k/v and with bias for to_out[0]) is restated here from the reference's own
call sites (interpolation.py:613-667, 738-792).

Pinning: the reference ships NO tests and NO golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the
reference code itself: ``oracle/gen_golden.py`` imports the unmodified
``/root/reference/interpolation.py`` (with ``prior`` stubbed, because
prior.py:3-4 imports packages that are not installed) and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file
against those vectors, and the coefficient schedule against the known answers
saved in the reference notebooks (``play_sd.ipynb`` cell 5:
Beta.ppf(0.75;3,3)=0.6405638352103529).

Two independent formulations are given and cross-checked in the tests:

* ``forward_direct``  -- follows the reference's data flow (replicate endpoint
  K/V, concatenate along tokens, materialise the probabilities);
* ``forward_merged``  -- flash-style partial attentions + log-sum-exp merge
  (SURVEY.md Appendix D), which is the form the CUDA kernels implement.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

MODE_PLAIN, MODE_OUTER, MODE_INNER = 0, 1, 2
MODE_NAMES = {"plain": MODE_PLAIN, "outer": MODE_OUTER, "inner": MODE_INNER}


# ----------------------------------------------------------------------------
# coefficient schedule  (prior.py:481-502, ends forced by interpolation.py:21-22)
# ----------------------------------------------------------------------------
def generate_beta_tensor(size: int, alpha: float = 3, beta: float = 3) -> torch.Tensor:
    from scipy.stats import beta as beta_dist

    probs = np.arange(size, dtype=np.float64) / (size - 1)
    return torch.tensor(beta_dist.ppf(probs, alpha, beta), dtype=torch.float32)


def coefficients(size: int, alpha: float = 1, beta: float = 1, t: Optional[float] = None) -> torch.Tensor:
    """coef as the processor ctor builds it (interpolation.py:20-31)."""
    if t is not None:
        assert 0 < t < 1, "t must be between 0 and 1"
        return torch.tensor([0.0, float(t), 1.0])
    c = generate_beta_tensor(size, alpha, beta)
    c[0], c[-1] = 0.0, 1.0
    return c


# ----------------------------------------------------------------------------
# weights of one attention layer (diffusers Attention)
# ----------------------------------------------------------------------------
@dataclass
class LayerWeights:
    wq: torch.Tensor  # (C, C)    to_q.weight, no bias
    wk: torch.Tensor  # (C, Cc)   to_k.weight, no bias
    wv: torch.Tensor  # (C, Cc)   to_v.weight, no bias
    wo: torch.Tensor  # (C, C)    to_out[0].weight
    bo: torch.Tensor  # (C,)      to_out[0].bias
    heads: int

    def to(self, dtype):
        return LayerWeights(self.wq.to(dtype), self.wk.to(dtype), self.wv.to(dtype),
                            self.wo.to(dtype), self.bo.to(dtype), self.heads)


def make_layer(C: int, Cc: int, heads: int, seed: int, dtype=torch.float32) -> LayerWeights:
    """Seeded weights with nn.Linear's default scale (U(-1/sqrt(fan_in), +))."""
    rs = np.random.RandomState(seed)

    def lin(o, i):
        b = 1.0 / math.sqrt(i)
        return torch.from_numpy(rs.uniform(-b, b, size=(o, i))).to(dtype)

    wq, wk, wv, wo = lin(C, C), lin(C, Cc), lin(C, Cc), lin(C, C)
    bo = torch.from_numpy(rs.uniform(-1 / math.sqrt(C), 1 / math.sqrt(C), size=(C,))).to(dtype)
    return LayerWeights(wq, wk, wv, wo, bo, heads)


def make_inputs(N: int, S: int, C: int, L: Optional[int], Cc: int, seed: int, dtype=torch.float32):
    """x ~ N(0,1) (N,S,C); ctx None (self) or (N,L,Cc) ~ N(0,1)."""
    rs = np.random.RandomState(seed + 7919)
    x = torch.from_numpy(rs.standard_normal((N, S, C))).to(dtype)
    ctx = None if L is None else torch.from_numpy(rs.standard_normal((N, L, Cc))).to(dtype)
    return x, ctx


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------
def _heads(t: torch.Tensor, h: int) -> torch.Tensor:
    """(N,T,C) -> (N,h,T,d)   (head_to_batch_dim without flattening N*h)."""
    N, T, C = t.shape
    return t.reshape(N, T, h, C // h).permute(0, 2, 1, 3)


def _unheads(t: torch.Tensor) -> torch.Tensor:
    """(N,h,T,d) -> (N,T,C)   (batch_to_head_dim)."""
    N, h, T, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(N, T, h * d)


def _softmax_attn(q, k, v, scale):
    """get_attention_scores + bmm: probabilities fully materialised."""
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    return torch.matmul(torch.softmax(s, dim=-1), v)


def _project(x, ctx, w: LayerWeights):
    src = x if ctx is None else ctx
    return x @ w.wq.T, src @ w.wk.T, src @ w.wv.T


# ----------------------------------------------------------------------------
# formulation 1: the reference's data flow
# ----------------------------------------------------------------------------
def _direct_core(q, k, v, ends, coef, mode: int, fused: bool, scale: float, h: int) -> torch.Tensor:
    """Attention of projected q (N,R,C) over projected k,v (N,L,C) -> (N,R,C)
    (before the output projection).  ends = (k_begin, v_begin, k_end, v_end),
    each (L,C): interpolation.py:627-630."""
    N = q.shape[0]
    qh = _heads(q, h)
    if mode == MODE_PLAIN:
        return _unheads(_softmax_attn(qh, _heads(k, h), _heads(v, h), scale))
    kb, vb, ke, ve = ends
    c = coef.to(q.dtype).reshape(N, 1, 1)
    rep = lambda t: t.unsqueeze(0).expand(N, -1, -1)      # interpolation.py:632-635
    if mode == MODE_OUTER:
        Kb, Vb, Ke, Ve = rep(kb), rep(vb), rep(ke), rep(ve)
        if fused:                                         # interpolation.py:643-649
            Kb, Vb = torch.cat([k, Kb], 1), torch.cat([v, Vb], 1)
            Ke, Ve = torch.cat([k, Ke], 1), torch.cat([v, Ve], 1)
        hb = _unheads(_softmax_attn(qh, _heads(Kb, h), _heads(Vb, h), scale))
        he = _unheads(_softmax_attn(qh, _heads(Ke, h), _heads(Ve, h), scale))
        return (1 - c) * hb + c * he                      # interpolation.py:662-664
    if mode == MODE_INNER:
        Kx = (1 - c) * rep(kb) + c * rep(ke)              # interpolation.py:772-775
        Vx = (1 - c) * rep(vb) + c * rep(ve)
        if fused:                                         # interpolation.py:781-785
            Kx, Vx = torch.cat([k, Kx], 1), torch.cat([v, Vx], 1)
        return _unheads(_softmax_attn(qh, _heads(Kx, h), _heads(Vx, h), scale))
    raise ValueError(mode)


def forward_direct(x, ctx, w: LayerWeights, coef, mode: int, fused: bool,
                   scale: Optional[float] = None, kv_endpoints=None) -> torch.Tensor:
    """One processor call.  x (N,S,C); ctx None or (N,L,Cc); coef (N,).

    kv_endpoints: optional (k_begin, v_begin, k_end, v_end), each (L,C): use
    these instead of rows 0 / N-1 of this batch (frame-sharded execution).
    """
    h = w.heads
    scale = (x.shape[-1] // h) ** -0.5 if scale is None else scale
    q, k, v = _project(x, ctx, w)
    ends = (k[0], v[0], k[-1], v[-1]) if kv_endpoints is None else kv_endpoints
    hid = _direct_core(q, k, v, ends, coef, mode, fused, scale, h)
    return hid @ w.wo.T + w.bo                            # interpolation.py:666-667


def forward_chunked(x, ctx, w: LayerWeights, coef, mode: int, fused: bool,
                    scale: Optional[float] = None, rows: int = 512) -> torch.Tensor:
    """forward_direct evaluated frame by frame and in query-row chunks so the
    materialised probabilities stay small (exact: frames interact only through
    the endpoint K/V, query rows not at all).  Used for the CPU baseline timing
    and for larger-size checks."""
    N, S, _ = x.shape
    h = w.heads
    scale = (x.shape[-1] // h) ** -0.5 if scale is None else scale
    q, k, v = _project(x, ctx, w)
    ends = (k[0], v[0], k[-1], v[-1])
    hid = torch.empty_like(q)
    for n in range(N):
        for r0 in range(0, S, rows):
            hid[n:n + 1, r0:r0 + rows] = _direct_core(
                q[n:n + 1, r0:r0 + rows], k[n:n + 1], v[n:n + 1], ends, coef[n:n + 1], mode, fused, scale, h)
    return hid @ w.wo.T + w.bo


def forward_rows(x, ctx, w: LayerWeights, coef, mode: int, fused: bool, rows, scale: Optional[float] = None) -> torch.Tensor:
    """forward_direct restricted to the query rows ``rows`` (index tensor / list) of every frame: (N, len(rows), C).
    Exact (query rows do not interact); K / V are projected for all tokens, so a full-size layer (S = 4096, N = 16 / 32)
    is checked in seconds on a row sample."""
    N = x.shape[0]
    h = w.heads
    scale = (x.shape[-1] // h) ** -0.5 if scale is None else scale
    rows = torch.as_tensor(rows, dtype=torch.long)
    src = x if ctx is None else ctx
    q = x[:, rows] @ w.wq.T                                 # interpolation.py:613
    k, v = src @ w.wk.T, src @ w.wv.T                       # interpolation.py:623-624
    ends = (k[0], v[0], k[-1], v[-1])
    hid = torch.empty_like(q)
    for n in range(N):
        hid[n:n + 1] = _direct_core(q[n:n + 1], k[n:n + 1], v[n:n + 1], ends, None if coef is None else coef[n:n + 1],
                                    mode, fused, scale, h)
    return hid @ w.wo.T + w.bo


# ----------------------------------------------------------------------------
# formulation 2: partial attentions + log-sum-exp merge (what the kernels do)
# ----------------------------------------------------------------------------
def _partial(q, k, v, scale):
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    m = s.amax(dim=-1, keepdim=True)
    p = torch.exp(s - m)
    return torch.matmul(p, v), p.sum(-1, keepdim=True), m


def _merge(a, b):
    (oa, la, ma), (ob, lb, mb) = a, b
    m = torch.maximum(ma, mb)
    ea, eb = torch.exp(ma - m), torch.exp(mb - m)
    return (oa * ea + ob * eb) / (la * ea + lb * eb)


def forward_merged(x, ctx, w: LayerWeights, coef, mode: int, fused: bool,
                   scale: Optional[float] = None, kv_endpoints=None) -> torch.Tensor:
    N = x.shape[0]
    h = w.heads
    d = x.shape[-1] // h
    scale = d ** -0.5 if scale is None else scale
    q, k, v = _project(x, ctx, w)
    qh, kh, vh = _heads(q, h), _heads(k, h), _heads(v, h)
    norm = lambda p: p[0] / p[1]
    if mode == MODE_PLAIN:
        return _unheads(norm(_partial(qh, kh, vh, scale))) @ w.wo.T + w.bo
    if kv_endpoints is None:
        kb, vb, ke, ve = k[0], v[0], k[-1], v[-1]
    else:
        kb, vb, ke, ve = kv_endpoints
    c = coef.to(x.dtype).reshape(N, 1, 1, 1)
    eh = lambda t: _heads(t.unsqueeze(0), h)  # (1,h,L,d) broadcast over frames
    p_self = _partial(qh, kh, vh, scale) if fused else None
    if mode == MODE_OUTER:
        pb = _partial(qh, eh(kb), eh(vb), scale)
        pe = _partial(qh, eh(ke), eh(ve), scale)
        hb = _merge(p_self, pb) if fused else norm(pb)
        he = _merge(p_self, pe) if fused else norm(pe)
        hid = (1 - c) * hb + c * he
    else:
        c3 = coef.to(x.dtype).reshape(N, 1, 1)
        Kx = (1 - c3) * kb.unsqueeze(0) + c3 * ke.unsqueeze(0)
        Vx = (1 - c3) * vb.unsqueeze(0) + c3 * ve.unsqueeze(0)
        px = _partial(qh, _heads(Kx, h), _heads(Vx, h), scale)
        hid = _merge(p_self, px) if fused else norm(px)
    return _unheads(hid) @ w.wo.T + w.bo


# ----------------------------------------------------------------------------
# harness input helper  (interpolation.py:861-918)
# ----------------------------------------------------------------------------
def slerp(v0: torch.Tensor, v1: torch.Tensor, t: float, threshold: float = 0.9995) -> torch.Tensor:
    """Spherical interpolation over the last dim; lerp where the directions are
    colinear (|cos| > threshold) or undefined (NaN)."""
    n0 = v0 / torch.linalg.vector_norm(v0, dim=-1, keepdim=True)
    n1 = v1 / torch.linalg.vector_norm(v1, dim=-1, keepdim=True)
    dot = (n0 * n1).sum(-1, keepdim=True)
    use_lerp = dot.abs().isnan() | (dot.abs() > threshold)
    theta = torch.arccos(dot)
    s0 = torch.sin(theta - theta * t) / torch.sin(theta)
    s1 = torch.sin(theta * t) / torch.sin(theta)
    return torch.where(use_lerp, torch.lerp(v0, v1, t), s0 * v0 + s1 * v1)


# ----------------------------------------------------------------------------
# error metrics used by every parity test (tolerance: SURVEY.md section 8c)
# ----------------------------------------------------------------------------
REL_RMS_TOL = 2e-3        # kernel (fp16/bf16 in/out, fp32 accumulate) vs fp32 oracle
MAX_ABS_TOL_X_RMS = 2e-2  # max |err| <= 2e-2 * RMS(oracle output)


def error_metrics(test: torch.Tensor, ref: torch.Tensor):
    test, ref = test.double(), ref.double()
    rms = ref.pow(2).mean().sqrt().item()
    err = test - ref
    return {"rel_rms": err.pow(2).mean().sqrt().item() / max(rms, 1e-30),
            "max_abs_over_rms": err.abs().max().item() / max(rms, 1e-30),
            "ref_rms": rms}


def within_tolerance(test, ref, rel_rms=REL_RMS_TOL, max_abs=MAX_ABS_TOL_X_RMS):
    m = error_metrics(test, ref)
    return (m["rel_rms"] <= rel_rms and m["max_abs_over_rms"] <= max_abs), m


# ----------------------------------------------------------------------------
# IP-Adapter variants (interpolation.py:51-545).  ip: (N, T, Cc) image tokens of each frame (the reference
# feeds a 3x row-repeated tensor for its hard-coded batch of 3 and un-repeats it with [::3] / [6:9]).
# ----------------------------------------------------------------------------
def _ip_parts(x, ctx, ip, w, wk_ip, wv_ip):
    q, k, v = _project(x, ctx, w)
    return q, k, v, ip @ wk_ip.T, ip @ wv_ip.T


def forward_ip_outer(x, ctx, ip, w: LayerWeights, wk_ip, wv_ip, coef, fused: bool, ip_scale: float) -> torch.Tensor:
    """OuterInterpolatedIPAttnProcessor (interpolation.py:240-387): outer interpolation of the text attention plus
    ip_scale times the outer interpolation of the image-token attention, same queries."""
    h, scale = w.heads, (x.shape[-1] // w.heads) ** -0.5
    q, k, v, kip, vip = _ip_parts(x, ctx, ip, w, wk_ip, wv_ip)
    hid = _direct_core(q, k, v, (k[0], v[0], k[-1], v[-1]), coef, MODE_OUTER, fused, scale, h)
    hid = hid + ip_scale * _direct_core(q, kip, vip, (kip[0], vip[0], kip[-1], vip[-1]), coef, MODE_OUTER, fused, scale, h)
    return hid @ w.wo.T + w.bo


def forward_ip_inner(x, ctx, ip, w: LayerWeights, wk_ip, wv_ip, coef, fused: bool, ip_scale: float) -> torch.Tensor:
    """InnerInterpolatedIPAttnProcessor (interpolation.py:417-545).  Reference quirk kept: the image part attends
    with each frame's OWN image K/V only (:525-527; the lerped key_cross of :512-523 is computed but unused)."""
    h, scale = w.heads, (x.shape[-1] // w.heads) ** -0.5
    q, k, v, kip, vip = _ip_parts(x, ctx, ip, w, wk_ip, wv_ip)
    hid = _direct_core(q, k, v, (k[0], v[0], k[-1], v[-1]), coef, MODE_INNER, fused, scale, h)
    hid = hid + ip_scale * _direct_core(q, kip, vip, None, None, MODE_PLAIN, False, scale, h)
    return hid @ w.wo.T + w.bo


def forward_ip_scale_control(x, ctx, ip, w: LayerWeights, wk_ip, wv_ip, coef, fused: bool, activated: bool) -> torch.Tensor:
    """ScaleControlIPAttnProcessor (interpolation.py:76-211): text attention (outer-interpolated when activated,
    plain otherwise) plus coef[n] times the attention over the END frame's image tokens (ip[0][6:9], :187-196)."""
    h, scale = w.heads, (x.shape[-1] // w.heads) ** -0.5
    N = x.shape[0]
    q, k, v, kip, vip = _ip_parts(x, ctx, ip, w, wk_ip, wv_ip)
    if activated:
        hid = _direct_core(q, k, v, (k[0], v[0], k[-1], v[-1]), coef, MODE_OUTER, fused, scale, h)
    else:
        hid = _direct_core(q, k, v, None, None, MODE_PLAIN, False, scale, h)
    ke, ve = kip[-1:].expand(N, -1, -1), vip[-1:].expand(N, -1, -1)
    hid = hid + coef.to(x.dtype).reshape(N, 1, 1) * _direct_core(q, ke, ve, None, None, MODE_PLAIN, False, scale, h)
    return hid @ w.wo.T + w.bo


def forward_ip_stock(x, ctx, ip, w: LayerWeights, wk_ip, wv_ip, ip_scale: float) -> torch.Tensor:
    """Stock IP-Adapter attention (diffusers IPAdapterAttnProcessor2_0, the deactivated branch :248-251):
    plain text attention + ip_scale * plain image-token attention."""
    h, scale = w.heads, (x.shape[-1] // w.heads) ** -0.5
    q, k, v, kip, vip = _ip_parts(x, ctx, ip, w, wk_ip, wv_ip)
    hid = _direct_core(q, k, v, None, None, MODE_PLAIN, False, scale, h)
    hid = hid + ip_scale * _direct_core(q, kip, vip, None, None, MODE_PLAIN, False, scale, h)
    return hid @ w.wo.T + w.bo


def make_ip(N: int, T: int, C: int, Cc: int, seed: int, dtype=torch.float32):
    """Seeded image tokens (N,T,Cc) and to_k_ip / to_v_ip weights (C,Cc)."""
    rs = np.random.RandomState(seed + 104729)
    b = 1.0 / math.sqrt(Cc)
    ip = torch.from_numpy(rs.standard_normal((N, T, Cc))).to(dtype)
    wk = torch.from_numpy(rs.uniform(-b, b, size=(C, Cc))).to(dtype)
    wv = torch.from_numpy(rs.uniform(-b, b, size=(C, Cc))).to(dtype)
    return ip, wk, wv
