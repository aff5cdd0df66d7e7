"""Generate tests/golden/*.npz by running the UNMODIFIED reference processors.

Run in the build container only (needs /root/reference; the GPU box never runs
this):   python oracle/gen_golden.py

How the reference is hosted:
* ``/root/reference/interpolation.py`` is imported as-is.  Its ``from prior
  import generate_beta_tensor`` (interpolation.py:7) is satisfied by a stub
  module, because prior.py:3-4 imports bayes_opt / lpips which are not
  installed; the stub restates prior.py:498-502 with the installed scipy.
* the ``attn`` argument is ``RefAttention`` below: a duck-typed stand-in for
  diffusers 0.27 ``Attention`` (not installed) implementing exactly the members
  the processors touch.

Every vector is produced by the reference code in fp32 (and checked against the
fp64 run); the oracle restatement is compared with it here too, so a generator
run fails loudly if the two ever disagree.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import paid_oracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference():
    stub = types.ModuleType("prior")
    stub.generate_beta_tensor = O.generate_beta_tensor
    sys.modules["prior"] = stub
    sys.path.insert(0, REF)
    import interpolation  # the reference module, unmodified

    assert os.path.realpath(interpolation.__file__).startswith(REF)
    return interpolation


class RefAttention(nn.Module):
    """Members of diffusers ``Attention`` used at interpolation.py:588-677."""

    def __init__(self, w: O.LayerWeights):
        super().__init__()
        C, Cc = w.wq.shape[0], w.wk.shape[1]
        self.heads = w.heads
        self.scale = (C // w.heads) ** -0.5
        self.to_q = nn.Linear(C, C, bias=False)
        self.to_k = nn.Linear(Cc, C, bias=False)
        self.to_v = nn.Linear(Cc, C, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(C, C, bias=True), nn.Dropout(0.0)])
        with torch.no_grad():
            self.to_q.weight.copy_(w.wq), self.to_k.weight.copy_(w.wk), self.to_v.weight.copy_(w.wv)
            self.to_out[0].weight.copy_(w.wo), self.to_out[0].bias.copy_(w.bo)
        self.to(w.wq.dtype)
        self.spatial_norm = self.group_norm = self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.upcast_attention = self.upcast_softmax = False

    def prepare_attention_mask(self, mask, target_length, batch_size):
        assert mask is None
        return None

    def head_to_batch_dim(self, t):
        B, T, C = t.shape
        h = self.heads
        return t.reshape(B, T, h, C // h).permute(0, 2, 1, 3).reshape(B * h, T, C // h)

    def batch_to_head_dim(self, t):
        Bh, T, d = t.shape
        h = self.heads
        return t.reshape(Bh // h, h, T, d).permute(0, 2, 1, 3).reshape(Bh // h, T, h * d)

    def get_attention_scores(self, q, k, attention_mask=None):
        assert attention_mask is None
        s = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype),
                          q, k.transpose(-1, -2), beta=0, alpha=self.scale)
        return s.softmax(dim=-1).to(q.dtype)


def run_reference(ref, w, x, ctx, coef, mode, fused):
    cls = ref.OuterInterpolatedAttnProcessor if mode == O.MODE_OUTER else ref.InnerInterpolatedAttnProcessor
    proc = cls(size=x.shape[0], is_fused=fused)
    proc.coef = coef.clone()      # the ctor's Beta schedule is replaced by the case's coefficients
    with torch.no_grad():
        return proc(RefAttention(w), x, encoder_hidden_states=ctx)


# name, C, Cc, heads, S, L(None=self), N, coefficient spec
SMALL = [
    ("d64_self_n3", 128, 128, 2, 40, None, 3, ("t", 0.37)),
    ("d64_self_n5", 128, 128, 2, 40, None, 5, ("beta", 2.0, 3.0)),
    ("d64_cross_n5", 128, 96, 2, 40, 13, 5, ("beta", 4.0, 4.0)),
    ("d64_cross77_n3", 64, 48, 1, 33, 77, 3, ("t", 0.5)),
    ("d40_self_n4", 80, 80, 2, 24, None, 4, ("beta", 1.0, 1.0)),
    ("d40_cross_n4", 80, 48, 2, 24, 9, 4, ("beta", 3.0, 3.0)),
    ("d80_self_n3", 160, 160, 2, 20, None, 3, ("t", 0.8)),
    ("d160_self_n3", 160, 160, 1, 16, None, 3, ("t", 0.25)),
]
# seeded cases at UNet-like geometry: inputs are regenerated from the seed
# (numpy RandomState streams are frozen), only a row subsample of the output is stored.
SEEDED = [
    ("sdxl32_like_self", 640, 640, 10, 256, None, 4, ("beta", 4.0, 4.0), 8),
    ("sdxl32_like_cross", 640, 2048, 10, 256, 77, 4, ("beta", 4.0, 4.0), 8),
    ("sd15_mid_self", 1280, 1280, 8, 64, None, 3, ("t", 0.5), 4),
]
MODES = [("outer", False), ("outer", True), ("inner", False), ("inner", True)]


def coef_of(spec, N):
    return O.coefficients(N, t=spec[1]) if spec[0] == "t" else O.coefficients(N, spec[1], spec[2])


def main():
    ref = import_reference()
    os.makedirs(OUT, exist_ok=True)
    worst = 0.0
    for seed, case in enumerate(SMALL + SEEDED):
        name, C, Cc, h, S, L, N, cspec = case[:8]
        stride = case[8] if len(case) > 8 else None
        w = O.make_layer(C, Cc, h, seed=100 + seed)
        x, ctx = O.make_inputs(N, S, C, L, Cc, seed=100 + seed)
        coef = coef_of(cspec, N)
        blob = {"meta": np.array([C, Cc, h, S, -1 if L is None else L, N, 100 + seed, stride or 0], dtype=np.int64),
                "coef": coef.numpy()}
        if stride is None:
            blob.update(x=x.numpy(), wq=w.wq.numpy(), wk=w.wk.numpy(), wv=w.wv.numpy(), wo=w.wo.numpy(), bo=w.bo.numpy())
            if ctx is not None:
                blob["ctx"] = ctx.numpy()
        for mname, fused in MODES:
            mode = O.MODE_NAMES[mname]
            y = run_reference(ref, w, x, ctx, coef, mode, fused)
            y64 = run_reference(ref, w.to(torch.float64), x.double(), None if ctx is None else ctx.double(),
                                coef.double(), mode, fused)
            e_ref = (y.double() - y64).abs().max().item()
            yo = O.forward_direct(x, ctx, w, coef, mode, fused)
            ym = O.forward_merged(x.double(), None if ctx is None else ctx.double(), w.to(torch.float64),
                                  coef.double(), mode, fused)
            e_dir = (yo.double() - y64).abs().max().item()
            e_mrg = (ym - y64).abs().max().item()
            worst = max(worst, e_dir, e_mrg)
            assert e_dir < 5e-6 and e_mrg < 1e-12, (name, mname, fused, e_dir, e_mrg)
            key = f"y_{mname}_{'fused' if fused else 'pure'}"
            blob[key] = (y if stride is None else y[:, ::stride]).numpy()
            print(f"{name:22s} {mname:5s} fused={int(fused)} ref32-ref64={e_ref:.2e} "
                  f"oracle32-ref64={e_dir:.2e} merged64-ref64={e_mrg:.2e}")
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
    # coefficient known answers (play_sd.ipynb cell 5 / cell 12 saved stdout)
    from scipy.stats import beta as B
    assert B.ppf(0.75, 3, 3) == 0.6405638352103529 and B.ppf(0.75, 1, 1) == 0.75
    print("worst oracle-vs-reference abs diff:", worst)


class RefIPAttn(nn.Module):
    """Fields of diffusers IPAdapterAttnProcessor2_0 read by the IP variants (interpolation.py:70-74, 137-138)."""

    def __init__(self, wk_ip, wv_ip, T, scale):
        super().__init__()
        C, Cc = wk_ip.shape
        self.num_tokens, self.scale = (T,), [scale]
        self.to_k_ip = nn.ModuleList([nn.Linear(Cc, C, bias=False)])
        self.to_v_ip = nn.ModuleList([nn.Linear(Cc, C, bias=False)])
        with torch.no_grad():
            self.to_k_ip[0].weight.copy_(wk_ip), self.to_v_ip[0].weight.copy_(wv_ip)


# name, C, Cc, heads, S, L, T(image tokens)   -- the reference IP processors are hard-coded to a batch of 3
IP_CASES = [("ip_d64", 128, 96, 2, 40, 13, 4), ("ip_d40_t16", 80, 48, 2, 24, 77, 16)]


def main_ip():
    ref = sys.modules["interpolation"]
    N, ip_scale = 3, 0.7
    for seed, (name, C, Cc, h, S, L, T) in enumerate(IP_CASES):
        w = O.make_layer(C, Cc, h, seed=300 + seed)
        x, ctx = O.make_inputs(N, S, C, L, Cc, seed=300 + seed)
        ip, wk_ip, wv_ip = O.make_ip(N, T, C, Cc, seed=300 + seed)
        coef = O.coefficients(N, t=0.35)
        ip9 = ip.repeat_interleave(3, dim=0)        # the reference's row layout [s,s,s,t,t,t,e,e,e] (sdxl:2146-2185)
        blob = dict(meta=np.array([C, Cc, h, S, L, N, T], dtype=np.int64), coef=coef.numpy(), ip_scale=np.float32(ip_scale),
                    x=x.numpy(), ctx=ctx.numpy(), ip=ip.numpy(), wk_ip=wk_ip.numpy(), wv_ip=wv_ip.numpy(),
                    wq=w.wq.numpy(), wk=w.wk.numpy(), wv=w.wv.numpy(), wo=w.wo.numpy(), bo=w.bo.numpy())
        runs = {
            "outer_fused": (ref.OuterInterpolatedIPAttnProcessor, True, lambda: O.forward_ip_outer(x, ctx, ip, w, wk_ip, wv_ip, coef, True, ip_scale)),
            "outer_pure": (ref.OuterInterpolatedIPAttnProcessor, False, lambda: O.forward_ip_outer(x, ctx, ip, w, wk_ip, wv_ip, coef, False, ip_scale)),
            "inner_fused": (ref.InnerInterpolatedIPAttnProcessor, True, lambda: O.forward_ip_inner(x, ctx, ip, w, wk_ip, wv_ip, coef, True, ip_scale)),
            "scale_fused": (ref.ScaleControlIPAttnProcessor, True, lambda: O.forward_ip_scale_control(x, ctx, ip, w, wk_ip, wv_ip, coef, True, True)),
            "scale_pure": (ref.ScaleControlIPAttnProcessor, False, lambda: O.forward_ip_scale_control(x, ctx, ip, w, wk_ip, wv_ip, coef, False, True)),
        }
        for key, (cls, fused, orc) in runs.items():
            proc = cls(t=0.35, is_fused=fused, ip_attn=RefIPAttn(wk_ip, wv_ip, T, ip_scale))
            with torch.no_grad():
                y = proc(RefAttention(w), x, encoder_hidden_states=(ctx, [ip9]))
            err = (y - orc()).abs().max().item()
            assert err < 5e-6, (name, key, err)
            blob["y_" + key] = y.numpy()
            print(f"{name:12s} {key:12s} oracle-vs-reference {err:.2e}")
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)


def main_slerp():
    """Latent interpolation used to build the frames (reference interpolation.py:861-918): generic rows, a colinear pair
    (falls back to lerp, :887-900) and an all-zero row (NaN cosine, same fallback), fp64."""
    ref = sys.modules["interpolation"]
    g = torch.Generator("cpu").manual_seed(404)
    a = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64)
    b = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64)
    b[0, 1, 2] = 1.7 * a[0, 1, 2]          # a colinear row
    a[0, 3, 5] = 0                         # a zero row
    ts = [0.0, 0.1, 0.35, 0.5, 0.9, 1.0]
    out = torch.stack([ref.slerp(a, b, t) for t in ts])
    for i, t in enumerate(ts):
        err = (O.slerp(a, b, t) - out[i]).abs().max().item()
        assert err < 1e-12, (t, err)
    np.savez_compressed(os.path.join(OUT, "aux_slerp.npz"), a=a.numpy(), b=b.numpy(), ts=np.array(ts), out=out.numpy())
    print("slerp: oracle == reference on", len(ts), "parameters")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    if "--only-slerp" in sys.argv:          # adds tests/golden/aux_slerp.npz without rewriting the attention vectors
        import_reference()
        main_slerp()
        sys.exit(0)
    main()
    main_ip()
    main_slerp()
