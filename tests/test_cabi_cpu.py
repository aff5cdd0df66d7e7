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
    assert lib.paid_attn_abi_version() == 3


def test_struct_layout_matches_header(cabi):
    # the library checks struct_size itself: a correct size passes validation, a wrong one is EINVAL
    lib = cabi.load_library()
    p = cabi.PaidAttnParams()
    p.struct_size = C.sizeof(cabi.PaidAttnParams)
    p.dtype, p.mode, p.N, p.S, p.L, p.C, p.Cc, p.heads = 0, 1, 3, 64, 64, 128, 128, 2
    assert lib.paid_attn_workspace_bytes(C.byref(p)) == 4 * 3 * 64 * 128 * 2
    p.mode = 2
    assert lib.paid_attn_workspace_bytes(C.byref(p)) == 6 * 3 * 64 * 128 * 2
    p.mode, p.plain_tail = 1, 3          # classifier-free-guidance rows ride in the same call: workspace for N + plain_tail frames
    assert lib.paid_attn_workspace_bytes(C.byref(p)) == 4 * 6 * 64 * 128 * 2
    p.plain_tail = -1
    assert lib.paid_attn_workspace_bytes(C.byref(p)) == 0 and "plain_tail" in cabi.last_error()
    p.plain_tail = 0
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
    # guidance rows (plain_tail) need per-frame K / V: one shared (L, C) matrix cannot serve the interpolated frames
    p.L, p.mode = 64, 0
    buf = torch.zeros(16, dtype=torch.float16)          # validation runs before anything touches the pointers
    p.x = p.wq = p.wo = p.y = p.k_pre = p.v_pre = buf.data_ptr()
    p.kv_pre_broadcast, p.plain_tail = 1, 3
    assert lib.paid_attn_forward(C.byref(p), None) == cabi.PAID_EINVAL and "plain_tail" in cabi.last_error()
    with pytest.raises(ValueError, match="plain_tail"):
        cabi.make_params(torch.zeros(3, 8, 64, dtype=torch.float16), None, None, None, None, None, None, None, 1, 1, True, plain_tail=3)
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


def test_glue_entry_points_validate_arguments(cabi):
    """paid_add_layer_norm / paid_group_norm_nhwc / paid_residual_bias_add / paid_geglu: argument errors come back as
    a status + message before anything touches the device (so they can be checked on a GPU-less host)."""
    lib = cabi.load_library()
    buf = C.create_string_buffer(4096 + 16)
    ok = C.c_void_p((C.addressof(buf) + 15) & ~15)            # 16-byte aligned host address: only validated, never used
    odd = C.c_void_p(ok.value + 2)
    # LayerNorm
    assert lib.paid_add_layer_norm(None, None, ok, ok, None, ok, 4, 64, 1e-5, 0, None) == cabi.PAID_EINVAL
    assert lib.paid_add_layer_norm(ok, ok, ok, ok, None, ok, 4, 64, 1e-5, 0, None) == cabi.PAID_EINVAL     # delta without x_out
    assert "x_out" in cabi.last_error()
    assert lib.paid_add_layer_norm(ok, None, ok, ok, None, ok, 4, 60, 1e-5, 0, None) == cabi.PAID_EUNSUPPORTED
    assert lib.paid_add_layer_norm(ok, None, ok, ok, None, ok, 4, 4096, 1e-5, 0, None) == cabi.PAID_EUNSUPPORTED
    assert lib.paid_add_layer_norm(odd, None, ok, ok, None, ok, 4, 64, 1e-5, 0, None) == cabi.PAID_EINVAL
    assert "aligned" in cabi.last_error()
    assert lib.paid_add_layer_norm(ok, None, ok, ok, None, ok, 4, 64, 1e-5, 7, None) == cabi.PAID_EINVAL   # dtype
    # GroupNorm
    need = lib.paid_group_norm_workspace_bytes(7, 128 * 128, 320, 32)
    assert need > 0 and need % 8 == 0 and need <= 1 << 20
    assert lib.paid_group_norm_workspace_bytes(7, 1024, 100, 25) == 0                                      # C % 8
    assert lib.paid_group_norm_workspace_bytes(7, 1024, 320, 33) == 0                                      # C % groups
    gn = lambda x, ws, wsb, N, HW, Cc, G: lib.paid_group_norm_nhwc(x, None, ok, ok, ok, ws, wsb, N, HW, Cc, G, 1e-5, 1, 0, None)
    assert gn(None, ok, need, 7, 128 * 128, 320, 32) == cabi.PAID_EINVAL
    assert gn(ok, ok, need, 7, 128 * 128, 100, 25) == cabi.PAID_EUNSUPPORTED
    assert gn(ok, ok, need - 8, 7, 128 * 128, 320, 32) == cabi.PAID_EWORKSPACE
    assert "workspace" in cabi.last_error()
    assert gn(ok, None, 0, 7, 128 * 128, 320, 32) == cabi.PAID_EWORKSPACE
    assert gn(ok, ok, need, 70000, 16, 320, 32) == cabi.PAID_EINVAL                                        # grid.y limit
    # residual + bias add, GEGLU
    assert lib.paid_residual_bias_add(ok, ok, None, ok, 4, 64, 0, None) == cabi.PAID_EINVAL
    assert lib.paid_residual_bias_add(ok, ok, ok, ok, 4, 60, 0, None) == cabi.PAID_EINVAL
    assert lib.paid_residual_bias_add(ok, odd, ok, ok, 4, 64, 0, None) == cabi.PAID_EINVAL
    assert lib.paid_geglu(ok, None, 4, 64, 0, None) == cabi.PAID_EINVAL
    assert lib.paid_geglu(ok, ok, 4, 60, 0, None) == cabi.PAID_EINVAL
    # Python shims refuse CPU tensors (no CPU path)
    x = torch.zeros(2, 8, 64, dtype=torch.float16)
    w = torch.ones(64, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cabi.add_layer_norm(x, None, w, w)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cabi.group_norm_nhwc(torch.zeros(2, 64, 4, 4, dtype=torch.float16).contiguous(memory_format=torch.channels_last), w, w, 32)
    with pytest.raises((RuntimeError, ValueError)):
        cabi.residual_bias_add(torch.zeros(2, 64, 4, 4, dtype=torch.float16), torch.zeros(2, 64, 4, 4, dtype=torch.float16), w)


def test_harness_block_restructure_equals_textbook_block():
    """The harness' BasicTransformerBlock carries each sub-layer's output as a pending residual into the next norm
    (add + LayerNorm fused on the GPU); on CPU tensors the same control flow must equal the textbook
    x += attn1(norm1 x); x += attn2(norm2 x, ctx); x += ff(norm3 x), chained over blocks."""
    from attention_interpolation_diffusion_b200 import unet_harness as U

    class Stub:   # any deterministic processor
        def __call__(self, attn, x, encoder_hidden_states=None, attention_mask=None, **kw):
            c = x if encoder_hidden_states is None else encoder_hidden_states
            return attn.to_out[0](attn.to_q(x) * attn.to_k(c).mean(1, keepdim=True))

    torch.manual_seed(0)
    blocks = []
    for _ in range(3):
        b = U.BasicTransformerBlock(64, 2, 32)
        b.attn1.set_processor(Stub()), b.attn2.set_processor(Stub())
        blocks.append(b)
    x, ctx = torch.randn(2, 10, 64), torch.randn(2, 7, 32)
    with torch.no_grad():
        ref = x
        for b in blocks:
            ref = ref + b.attn1(b.norm1(ref))
            ref = ref + b.attn2(b.norm2(ref), encoder_hidden_states=ctx)
            ref = ref + b.ff(b.norm3(ref))
        h, pending = x, None
        for b in blocks:
            h, pending = b(h, ctx, pending)
        assert torch.allclose(h + pending, ref, atol=1e-5)
        # ResNet block on CPU: the plain composition (conv biases, time embedding add, skip add)
        r = U.ResnetBlock2D(32, 64, 16)
        xi, temb = torch.randn(2, 32, 5, 5), torch.randn(2, 16)
        F = torch.nn.functional
        hh = r.conv1(F.silu(r.norm1(xi)))
        hh = hh + r.time_emb_proj(F.silu(temb))[:, :, None, None]
        hh = r.conv2(F.silu(r.norm2(hh)))
        assert torch.allclose(r(xi, temb), r.conv_shortcut(xi) + hh, atol=1e-5)


def test_integration_stub_struct_matches_the_library(cabi):
    """The ctypes struct a maintainer would copy out of INTEGRATION.md must be the struct the library validates
    (struct_size guard): same fields, same types, same order as _cabi.PaidAttnParams."""
    text = open(os.path.join(os.path.dirname(HEADER), "..", "INTEGRATION.md")).read()
    block = text[text.index("class PaidAttnParams(ctypes.Structure)"):text.index("def outer_call")]
    fields = re.findall(r'\("(\w+)",\s*ctypes\.(c_\w+)\)', block)
    assert [(n, getattr(C, t)) for n, t in fields] == list(cabi.PaidAttnParams._fields_)
