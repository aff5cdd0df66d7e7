"""Loader for tests/golden/*.npz (written by oracle/gen_golden.py from the real reference)."""
import glob
import os

import numpy as np
import torch

import paid_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODES = [("outer", False), ("outer", True), ("inner", False), ("inner", True)]


def case_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


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
