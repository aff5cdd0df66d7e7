"""CPU: the C-ABI library loads, exports every symbol include/paid_attn.h declares, agrees with the
ctypes structs, and rejects bad arguments with a status + message (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "paid_attn.h")


@pytest.fixture(scope="module")
def cabi():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    from attention_interpolation_diffusion_b200 import _cabi
    _cabi.load_library()
    return _cabi


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(paid_[a-z_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(cabi):
    lib = cabi.load_library()
    names = declared_functions()
    assert set(names) == set(cabi.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.paid_attn_abi_version() == 1


def test_struct_layout_matches_header(cabi):
    # the library checks struct_size itself: a correct size passes validation, a wrong one is EINVAL
    lib = cabi.load_library()
    p = cabi.PaidAttnParams()
    p.struct_size = C.sizeof(cabi.PaidAttnParams)
    p.dtype, p.mode, p.N, p.S, p.L, p.C, p.Cc, p.heads = 0, 1, 3, 64, 64, 128, 128, 2
    assert lib.paid_attn_workspace_bytes(C.byref(p)) == 4 * 3 * 64 * 128 * 2
    p.mode = 2
    assert lib.paid_attn_workspace_bytes(C.byref(p)) == 6 * 3 * 64 * 128 * 2
    p.struct_size -= 8
    assert lib.paid_attn_workspace_bytes(C.byref(p)) == 0
    assert "ABI mismatch" in cabi.last_error()


def test_argument_errors_are_reported_not_thrown(cabi):
    lib = cabi.load_library()
    p = cabi.PaidAttnParams()
    p.struct_size = C.sizeof(cabi.PaidAttnParams)
    p.dtype, p.mode, p.N, p.S, p.L, p.C, p.Cc, p.heads = 0, 1, 3, 64, 64, 130, 130, 4
    assert lib.paid_attn_forward(C.byref(p), None) == cabi.PAID_EINVAL
    assert "multiple of heads" in cabi.last_error()
    p.C = p.Cc = 128
    assert lib.paid_attn_forward(C.byref(p), None) == cabi.PAID_EINVAL      # null tensors
    p.L = 77                                                                 # self-attention needs L == S
    assert lib.paid_attn_forward(C.byref(p), None) == cabi.PAID_EINVAL
    assert lib.paid_attn_forward(None, None) == cabi.PAID_EINVAL
    assert lib.paid_linear(None, None, None, None, 1, 1, 1, 0, 0, None) == cabi.PAID_EINVAL
    c = cabi.PaidCoreParams()
    assert lib.paid_attn_core(C.byref(c), None) == cabi.PAID_EINVAL


def test_no_cpu_fallback(cabi):
    x = torch.zeros(3, 8, 64, dtype=torch.float16)
    w = torch.zeros(64, 64, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cabi.attn_forward(x, None, w, w, w, w, None, None, 1, cabi.PAID_PLAIN, False)
    with pytest.raises((RuntimeError, NotImplementedError)):
        cabi.attn_forward(x.float(), None, w, w, w, w, None, None, 1, cabi.PAID_PLAIN, False)


def test_processor_api_mirrors_reference():
    from attention_interpolation_diffusion_b200 import (InnerInterpolatedAttnProcessor,
                                                        OuterInterpolatedAttnProcessor, generate_beta_tensor)
    p = OuterInterpolatedAttnProcessor(t=0.3, is_fused=True)
    assert p.size == 3 and p.activated and p.is_fused and torch.allclose(p.coef, torch.tensor([0, 0.3, 1.0]))
    p.deactivate()
    assert not p.activated
    p.activate(0.7)
    assert p.activated and torch.allclose(p.coef, torch.tensor([0, 0.7, 1.0]))
    with pytest.raises(AssertionError):
        p.activate(1.0)
    with pytest.raises(AssertionError):
        InnerInterpolatedAttnProcessor(t=0.0)
    q = InnerInterpolatedAttnProcessor(size=7, alpha=3, beta=3)
    ref = generate_beta_tensor(7, 3, 3)
    ref[0], ref[-1] = 0, 1
    assert q.size == 7 and torch.equal(q.coef, ref) and q.original_attn is None
    q.set_coefs(torch.tensor([0.2, 0.4, 0.6, 0.9]))
    assert q.size == 4 and q.coef[0] == 0 and q.coef[-1] == 1
    import paid_oracle as O
    assert torch.equal(generate_beta_tensor(9, 2, 5), O.generate_beta_tensor(9, 2, 5))
