"""Loader for tests/golden/*.npz (written by oracle/gen_golden.py from the real reference)."""
import glob
import os

import numpy as np
import torch

import paid_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODES = [("outer", False), ("outer", True), ("inner", False), ("inner", True)]


def case_names():
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
                  if not n.startswith(("ip_", "e2e_", "aux_")))      # aux_: vectors of helpers next to the path (slerp)


def ip_case_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "ip_*.npz")))


IP_RUNS = ["outer_fused", "outer_pure", "inner_fused", "scale_fused", "scale_pure"]


def load_ip_case(name):
    """IP-Adapter vectors written by the reference's IP processors (batch of 3)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    C, Cc, h, S, L, N, T = (int(v) for v in z["meta"])
    t = lambda k: torch.from_numpy(z[k])
    w = O.LayerWeights(t("wq"), t("wk"), t("wv"), t("wo"), t("bo"), heads=h)
    return dict(name=name, w=w, x=t("x"), ctx=t("ctx"), ip=t("ip"), wk_ip=t("wk_ip"), wv_ip=t("wv_ip"), coef=t("coef"),
                ip_scale=float(z["ip_scale"]), outs={k: t("y_" + k) for k in IP_RUNS}, h=h, T=T, N=N)


def oracle_ip(c, run, x=None, ctx=None, ip=None, w=None, wk_ip=None, wv_ip=None):
    x, ctx, ip = (c["x"] if x is None else x), (c["ctx"] if ctx is None else ctx), (c["ip"] if ip is None else ip)
    w, wk_ip, wv_ip = (c["w"] if w is None else w), (c["wk_ip"] if wk_ip is None else wk_ip), (c["wv_ip"] if wv_ip is None else wv_ip)
    kind, fused = run.split("_")[0], run.endswith("fused")
    if kind == "outer":
        return O.forward_ip_outer(x, ctx, ip, w, wk_ip, wv_ip, c["coef"], fused, c["ip_scale"])
    if kind == "inner":
        return O.forward_ip_inner(x, ctx, ip, w, wk_ip, wv_ip, c["coef"], fused, c["ip_scale"])
    return O.forward_ip_scale_control(x, ctx, ip, w, wk_ip, wv_ip, c["coef"], fused, True)


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    C, Cc, h, S, L, N, seed, stride = (int(v) for v in z["meta"])
    L = None if L < 0 else L
    coef = torch.from_numpy(z["coef"])
    if stride == 0:
        w = O.LayerWeights(*(torch.from_numpy(z[k]) for k in ("wq", "wk", "wv", "wo", "bo")), heads=h)
        x = torch.from_numpy(z["x"])
        ctx = torch.from_numpy(z["ctx"]) if "ctx" in z.files else None
    else:  # seeded case: regenerate the inputs exactly as the generator did
        w = O.make_layer(C, Cc, h, seed=seed)
        x, ctx = O.make_inputs(N, S, C, L, Cc, seed=seed)
    outs = {(m, f): torch.from_numpy(z[f"y_{m}_{'fused' if f else 'pure'}"]) for m, f in MODES}
    return dict(name=name, w=w, x=x, ctx=ctx, coef=coef, outs=outs, stride=stride or 1,
                C=C, Cc=Cc, h=h, S=S, L=L, N=N)
