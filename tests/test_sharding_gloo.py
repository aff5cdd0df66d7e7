"""CPU, world_size 2 over gloo: the frame-shard plan and the endpoint broadcast -- the only collective of the
path.  The per-rank math is done by the ORACLE here (checker); the GPU path is covered by test_parity_gpu."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import paid_oracle as O
from attention_interpolation_diffusion_b200.sharding import FrameShard, plan_frame_shards


def test_plan():
    assert plan_frame_shards(7, 2) == [[0, 6, 1, 2], [3, 4, 5]]
    assert plan_frame_shards(16, 8) == [[0, 15]] + [[2 * i - 1, 2 * i] for i in range(1, 8)]
    assert plan_frame_shards(7, 1) == [[0, 6, 1, 2, 3, 4, 5]]
    for bad in ((3, 4), (8, 8), (2, 2), (1, 1)):
        with pytest.raises(ValueError):        # rank 0 holds both endpoints, every other rank at least one frame
            plan_frame_shards(*bad)
    for n in range(2, 40):
        for w in (1, 2, 4, 8):
            if w > 1 and w > n - 1:
                continue
            sh = plan_frame_shards(n, w)
            assert sorted(f for part in sh for f in part) == list(range(n))
            assert sh[0][:2] == [0, n - 1]                                   # both endpoints on rank 0, local frames 0 and 1
            assert max(map(len, sh)) - min(map(len, sh)) <= 1 and min(map(len, sh)) >= 1
    s0 = FrameShard(0, 2, 7)
    x = torch.arange(7.0).reshape(7, 1)
    parts = [FrameShard(r, 2, 7).local(x) for r in range(2)]
    assert parts[0].flatten().tolist() == [0, 6, 1, 2] and torch.equal(s0.unshard(parts), x)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, S, C, h = 7, 24, 64, 2
        w = O.make_layer(C, C, h, 3, torch.float64)
        x, _ = O.make_inputs(N, S, C, None, C, 3, torch.float64)
        coef = O.coefficients(N, 4, 4).double()
        full = O.forward_direct(x, None, w, coef, O.MODE_OUTER, True)
        sh = FrameShard(rank, world, N)
        xl = sh.local(x)
        # the product's exchange: rank 0 fills the layer's buffer from its two endpoint frames, ONE broadcast
        kv = sh.kv_buffer("layer0", S, C, xl)
        if sh.owns_endpoints:
            kv[0], kv[1], kv[2], kv[3] = xl[0] @ w.wk.T, xl[0] @ w.wv.T, xl[1] @ w.wk.T, xl[1] @ w.wv.T
        assert sh.exchange(kv, ready_on_main=sh.owns_endpoints) is None          # CPU tensors: ordered, no event
        assert sh.broadcasts == 1 and sh.kv_buffer("layer0", S, C, xl) is kv      # persistent per layer
        y = O.forward_direct(xl, None, w, sh.local(coef), O.MODE_OUTER, True, kv_endpoints=tuple(kv))
        q.put((rank, float((y - sh.local(full)).abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_endpoint_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=100) for _ in procs)
    [p.join(30) for p in procs]
    assert [r for r, _ in res] == [0, 1]
    assert all(e < 1e-12 for _, e in res), res


def _ip_worker(rank, world, port, q):
    """Frame-sharded IP-Adapter call: the PRODUCT's endpoint routing (_InterpolatedIPAttnProcessor._endpoints: rank 0's
    rows broadcast for self-attention, locally projected endpoint contexts for cross-attention) feeds the oracle's
    per-rank math."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from attention_interpolation_diffusion_b200 import OuterInterpolatedIPAttnProcessor, PaidIPAdapterAttnProcessor
        N, S, C, h, Cc, L, T = 6, 20, 64, 2, 48, 9, 4
        dt = torch.float64
        w = O.make_layer(C, Cc, h, 5, dt)
        x, ctx = O.make_inputs(N, S, C, L, Cc, 5, dt)
        ip, wk_ip, wv_ip = O.make_ip(N, T, C, Cc, 5, dt)
        coef = O.coefficients(N, 3, 3).double()
        scale = (C // h) ** -0.5
        sh = FrameShard(rank, world, N)
        ipa = PaidIPAdapterAttnProcessor(C, Cc, num_tokens=(T,), scale=0.6)
        ql, kl, vl, kipl, vipl = O._ip_parts(sh.local(x), sh.local(ctx), sh.local(ip), w, wk_ip, wv_ip)
        cl = sh.local(coef)
        errs = []
        # outer-IP self-attention form: the endpoints are rows of rank 0's K / V, broadcast (the one collective)
        proc = OuterInterpolatedIPAttnProcessor(size=N, is_fused=True, alpha=3, beta=3, ip_attn=ipa)
        proc.shard = sh
        layer = object()
        et = proc._endpoints(layer, kl, vl, {}, "kv_ext", cross=False)
        assert (et["begin_frame"], et["end_frame"]) == ((0, 1) if sh.owns_endpoints else (-1, -1))
        ends_t = (kl[0], vl[0], kl[1], vl[1]) if sh.owns_endpoints else tuple(et["kv_ext"])
        # cross-attention form (text and image tokens): every rank projects the endpoint contexts itself, no collective
        ce, ie = torch.stack([ctx[0], ctx[-1]]), torch.stack([ip[0], ip[-1]])
        st = {"kv_ext": torch.stack([ce[0] @ w.wk.T, ce[0] @ w.wv.T, ce[1] @ w.wk.T, ce[1] @ w.wv.T]),
              "ip_kv_ext": torch.stack([ie[0] @ wk_ip.T, ie[0] @ wv_ip.T, ie[1] @ wk_ip.T, ie[1] @ wv_ip.T])}
        before = sh.broadcasts
        ec, ei = proc._endpoints(layer, kl, vl, st, "kv_ext", cross=True), proc._endpoints(layer, kipl, vipl, st, "ip_kv_ext", cross=True)
        assert sh.broadcasts == before
        ends_c = (kl[0], vl[0], kl[1], vl[1]) if sh.owns_endpoints else tuple(ec["kv_ext"])
        ends_i = (kipl[0], vipl[0], kipl[1], vipl[1]) if sh.owns_endpoints else tuple(ei["kv_ext"])
        assert max(float((a - b).abs().max()) for a, b in zip(ends_t, ends_c)) < 1e-12       # both routes agree
        hid = O._direct_core(ql, kl, vl, ends_c, cl, O.MODE_OUTER, True, scale, h)
        hid = hid + 0.6 * O._direct_core(ql, kipl, vipl, ends_i, cl, O.MODE_OUTER, True, scale, h)
        full = O.forward_ip_outer(x, ctx, ip, w, wk_ip, wv_ip, coef, True, 0.6)
        errs.append(float((hid @ w.wo.T + w.bo - sh.local(full)).abs().max()))
        # scale control: the END frame's image-token K/V for every frame, from the end-frame context every rank holds
        n = sh.local_frames
        kend, vend = (ip[-1:] @ wk_ip.T), (ip[-1:] @ wv_ip.T)
        hid = O._direct_core(ql, kl, vl, ends_c, cl, O.MODE_OUTER, True, scale, h)
        hid = hid + cl.reshape(n, 1, 1) * O._direct_core(ql, kend.expand(n, -1, -1), vend.expand(n, -1, -1), None, None,
                                                         O.MODE_PLAIN, False, scale, h)
        full = O.forward_ip_scale_control(x, ctx, ip, w, wk_ip, wv_ip, coef, True, True)
        errs.append(float((hid @ w.wo.T + w.bo - sh.local(full)).abs().max()))
        q.put((rank, max(errs)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_ip_adapter_endpoint_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ip_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=100) for _ in procs)
    [p.join(30) for p in procs]
    assert [r for r, _ in res] == [0, 1]
    assert all(e < 1e-12 for _, e in res), res
