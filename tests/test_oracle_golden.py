"""CPU: the oracle restatement against the reference's own outputs (tests/golden)
and the reference's only saved known answers (coefficient schedule)."""
import pytest
import torch

import paid_oracle as O
from golden_util import IP_RUNS, MODES, case_names, ip_case_names, load_case, load_ip_case, oracle_ip


@pytest.mark.parametrize("name", case_names())
def test_oracle_matches_reference_vectors(name):
    c = load_case(name)
    for (m, fused), y_ref in c["outs"].items():
        mode = O.MODE_NAMES[m]
        for fn in (O.forward_direct, O.forward_merged):
            y = fn(c["x"], c["ctx"], c["w"], c["coef"], mode, fused)[:, ::c["stride"]]
            assert (y - y_ref).abs().max().item() < 2e-6, (name, m, fused, fn.__name__)
        y = O.forward_chunked(c["x"], c["ctx"], c["w"], c["coef"], mode, fused, rows=17)[:, ::c["stride"]]
        assert (y - y_ref).abs().max().item() < 2e-6


@pytest.mark.parametrize("name", ip_case_names())
def test_ip_oracle_matches_reference_vectors(name):
    c = load_ip_case(name)
    for run in IP_RUNS:
        assert (oracle_ip(c, run) - c["outs"][run]).abs().max().item() < 2e-6, (name, run)


def test_golden_set_is_complete():
    names = case_names()
    assert len(names) >= 11
    assert any("cross" in n for n in names) and any("d40" in n for n in names) and any("d160" in n for n in names)


def test_beta_known_answers():
    # play_sd.ipynb cell 5 / cell 12 saved stdout
    from scipy.stats import beta
    assert beta.ppf(0.75, 3, 3) == 0.6405638352103529
    assert beta.ppf(0.25, 3, 3) == 0.3594361647896471
    assert beta.ppf(0.75, 1, 1) == 0.75
    t = O.generate_beta_tensor(5, 3, 3)
    assert t.dtype == torch.float32
    assert torch.allclose(t, torch.tensor([0.0, 0.3594361647896471, 0.5, 0.6405638352103529, 1.0]))
    c = O.coefficients(7, 4, 4)
    assert c[0] == 0 and c[-1] == 1 and torch.all(c[1:] > c[:-1])
    assert torch.equal(O.coefficients(9, t=0.3), torch.tensor([0.0, 0.3, 1.0]))


@pytest.mark.parametrize("m,fused", MODES)
def test_properties_endpoints_and_sharding(m, fused):
    """SURVEY.md section 4 properties 1-3 in fp64: endpoints equal plain
    attention; an N-frame batch equals independent [0, i, N-1] batches."""
    mode = O.MODE_NAMES[m]
    N, S, C, h = 6, 48, 64, 4
    w = O.make_layer(C, C, h, 5, torch.float64)
    for L in (None, 11):
        x, ctx = O.make_inputs(N, S, C, L, C, 5, torch.float64)
        coef = O.coefficients(N, 2, 5).double()
        y = O.forward_direct(x, ctx, w, coef, mode, fused)
        plain = O.forward_direct(x, ctx, w, coef, O.MODE_PLAIN, False)
        assert (y[0] - plain[0]).abs().max() < 1e-13 and (y[-1] - plain[-1]).abs().max() < 1e-13
        for i in range(1, N - 1):
            idx = [0, i, N - 1]
            y3 = O.forward_direct(x[idx], None if ctx is None else ctx[idx], w, coef[idx], mode, fused)
            assert (y3[1] - y[i]).abs().max() < 1e-13
        # frame-sharded execution with externally supplied endpoint K/V
        src = x if ctx is None else ctx
        ends = ((src[0] @ w.wk.T), (src[0] @ w.wv.T), (src[-1] @ w.wk.T), (src[-1] @ w.wv.T))
        ys = O.forward_direct(x[2:4], None if ctx is None else ctx[2:4], w, coef[2:4], mode, fused, kv_endpoints=ends)
        assert (ys - y[2:4]).abs().max() < 1e-13


def test_slerp_matches_definition():
    torch.manual_seed(0)
    a, b = torch.randn(1, 4, 8, 8, dtype=torch.float64), torch.randn(1, 4, 8, 8, dtype=torch.float64)
    assert torch.allclose(O.slerp(a, b, 0.0), a) and torch.allclose(O.slerp(a, b, 1.0), b)
    assert torch.allclose(O.slerp(a, a * 2, 0.25), torch.lerp(a, a * 2, 0.25))  # colinear -> lerp


def test_slerp_matches_the_reference(golden_dir):
    """tests/golden/aux_slerp.npz: outputs of the unmodified reference slerp (interpolation.py:861-918; generic rows, a colinear
    row and an all-zero row) -- the oracle's restatement and the pipeline's own must reproduce them."""
    import os

    import numpy as np
    from attention_interpolation_diffusion_b200.pipeline import slerp
    blob = np.load(os.path.join(golden_dir, "aux_slerp.npz"))
    a, b, out = (torch.from_numpy(blob[k]) for k in ("a", "b", "out"))
    for i, t in enumerate(blob["ts"].tolist()):
        assert (O.slerp(a, b, t) - out[i]).abs().max() < 1e-12, t
        assert (slerp(a, b, t) - out[i]).abs().max() < 1e-12, t
