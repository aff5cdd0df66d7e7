"""CPU, world_size 2 over gloo: the frame-shard plan and the endpoint broadcast -- the only collective of the
path.  The per-rank math is done by the ORACLE here (checker); the GPU path is covered by test_parity_gpu."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import paid_oracle as O
from attention_interpolation_diffusion_b200.sharding import (FrameShard, broadcast_endpoints, endpoint_owners,
                                                             plan_frame_shards)


def test_plan():
    assert plan_frame_shards(7, 2) == [(0, 4), (4, 7)]
    assert plan_frame_shards(16, 8) == [(2 * i, 2 * i + 2) for i in range(8)]
    assert plan_frame_shards(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert endpoint_owners(plan_frame_shards(3, 4), 3) == (0, 2)
    assert endpoint_owners(plan_frame_shards(7, 1), 7) == (0, 0)
    for n in range(2, 40):
        for w in (1, 2, 4, 8):
            sh = plan_frame_shards(n, w)
            assert sh[0][0] == 0 and sh[-1][1] == n and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            assert max(h - l for l, h in sh) - min(h - l for l, h in sh) <= 1


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
        kv = torch.zeros(4, S, C, dtype=torch.float64)
        if rank == sh.begin_owner:
            kv[0], kv[1] = xl[0] @ w.wk.T, xl[0] @ w.wv.T
        if rank == sh.end_owner:
            kv[2], kv[3] = xl[-1] @ w.wk.T, xl[-1] @ w.wv.T
        broadcast_endpoints(kv, sh.begin_owner, sh.end_owner)
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
    """Frame-sharded IP-Adapter call: the PRODUCT's endpoint exchange (_InterpolatedIPAttnProcessor._endpoints: rows of
    the local K/V on the owner ranks, broadcast over the process group) feeds the oracle's per-rank math."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from attention_interpolation_diffusion_b200 import (OuterInterpolatedIPAttnProcessor, PaidIPAdapterAttnProcessor,
                                                            ScaleControlIPAttnProcessor)
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
        # outer-IP: text and image-token endpoints both exchanged
        proc = OuterInterpolatedIPAttnProcessor(size=N, is_fused=True, alpha=3, beta=3, ip_attn=ipa)
        proc.shard = sh
        et, ei = proc._endpoints(kl, vl), proc._endpoints(kipl, vipl)
        assert et["begin_frame"] == (0 if rank == sh.begin_owner else -1)
        assert et["end_frame"] == (sh.local_frames - 1 if rank == sh.end_owner else -1)
        hid = O._direct_core(ql, kl, vl, tuple(et["kv_ext"]), cl, O.MODE_OUTER, True, scale, h)
        hid = hid + 0.6 * O._direct_core(ql, kipl, vipl, tuple(ei["kv_ext"]), cl, O.MODE_OUTER, True, scale, h)
        full = O.forward_ip_outer(x, ctx, ip, w, wk_ip, wv_ip, coef, True, 0.6)
        errs.append(float((hid @ w.wo.T + w.bo - sh.local(full)).abs().max()))
        # scale control: only the END frame's image-token K/V travel
        proc = ScaleControlIPAttnProcessor(size=N, is_fused=True, alpha=3, beta=3, ip_attn=ipa)
        proc.shard = sh
        kv = proc._endpoints(kipl[-1:], vipl[-1:], need_begin=False)["kv_ext"]
        n = sh.local_frames
        hid = O._direct_core(ql, kl, vl, tuple(et["kv_ext"]), cl, O.MODE_OUTER, True, scale, h)
        hid = hid + cl.reshape(n, 1, 1) * O._direct_core(ql, kv[2:3].expand(n, -1, -1), kv[3:4].expand(n, -1, -1), None, None,
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
