#!/usr/bin/env python
"""bench.py -- interpolation-frames/sec of an N-frame SDXL PAID sequence (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (own arm: the sm_100a CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (reference arm: CPU port of the reference path)

One "step" = one whole 50-step denoise of the frame sequence through the reference-shaped pipeline
(`InterpolationPipeline.interpolate`): per denoising step one conditional UNet pass with fused-outer AID in every
one of the 140 attention layers (first int(50*0.5)=25 steps) and one unconditional pass with plain attention.
Workload at N=1: BASELINE.json configs[2] "SDXL 128x128 latent, 7-frame PAID (guidance prompt), 50 steps";
at R GPUs the sequence has 7*R frames, frame-sharded (weak scaling) with one NCCL broadcast of the endpoint K/V per
interpolated attention call.  Synthetic latents / embeddings (seed 1002) and random-init UNet weights.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_GPU = 7
STEPS_PER_SEQUENCE = 50
WARMUP_RATIO = 0.5
METRIC = "interpolation-frames/sec (SDXL UNet, 50 steps)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--model", default="sdxl", choices=["sdxl", "sd15", "tiny"])
    ap.add_argument("--frames", type=int, default=0, help="total frames of the sequence (default 7 per GPU)")
    ap.add_argument("--atype", default="fused_outer", choices=["fused_outer", "fused_inner"])
    ap.add_argument("--denoise-steps", type=int, default=STEPS_PER_SEQUENCE)
    ap.add_argument("--ip-tokens", type=int, default=0,
                    help="image-conditioned morphing (BASELINE configs[4]): IP-Adapter processors with this many image "
                         "tokens per frame (16 = ip-adapter-plus); 0 = text-only PAID")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly (for ncu launch lists)")
    ap.add_argument("--ncu-range", action="store_true", help="bracket the timed region with cudaProfilerStart/Stop (ncu --profile-from-start off: the launch list of the timed steps only)")
    ap.add_argument("--watchdog-s", type=float, default=420.0,
                    help="abort (exit code 17, diagnostic on stderr) when one phase of the run makes no progress for this long: a "
                         "stuck collective must not hold the GPUs until the caller's own limit")
    return ap.parse_args()


class Watchdog:
    """A multi-rank run that stops making progress (a collective some rank never joins) would otherwise spin on the GPUs
    until the caller kills it.  Every phase of the bench announces itself here; a phase that exceeds its limit ends the
    process with a diagnostic instead (torchrun then tears the other ranks down)."""

    def __init__(self, limit_s: float, rank: int, world: int):
        import threading
        self.limit, self.rank, self.world = limit_s, rank, world
        self.name, self.since, self.budget = "start", time.time(), limit_s
        self._stop = threading.Event()
        if limit_s > 0:
            threading.Thread(target=self._run, daemon=True).start()

    def phase(self, name: str, extra_s: float = 0.0):
        self.name, self.since, self.budget = name, time.time(), self.limit + extra_s

    def stop(self):
        self._stop.set()

    def _run(self):
        while not self._stop.wait(5.0):
            idle = time.time() - self.since
            if idle > self.budget:
                sys.stderr.write(json.dumps({"error": "bench watchdog", "phase": self.name, "no_progress_s": round(idle, 1),
                                             "rank": self.rank, "world": self.world}) + "\n")
                sys.stderr.flush()
                try:                                  # where every thread of this rank stands (the next session's first clue)
                    import faulthandler
                    faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                    sys.stderr.flush()
                finally:
                    os._exit(17)


# ---------------------------------------------------------------------------------------------------------
# clocks: sampled with nvidia-smi DURING the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# synthetic inputs: seed 1002, N(0,1) latents and embeddings
# ---------------------------------------------------------------------------------------------------------
def make_host_inputs(cfg, dtype, ip_tokens=0):
    import torch
    g = torch.Generator("cpu").manual_seed(1002)
    r = lambda *s: torch.randn(*s, generator=g).to(dtype).pin_memory() if torch.cuda.is_available() else torch.randn(*s, generator=g).to(dtype)
    side, cc = cfg.sample_size, cfg.cross_attention_dim
    d = dict(latent_start=r(1, 4, side, side), latent_end=r(1, 4, side, side), embeds_start=r(1, 77, cc),
             embeds_end=r(1, 77, cc), negative_embeds=r(1, 77, cc), guide_embeds=r(1, 77, cc))
    if cfg.text_time:
        d.update(pooled_start=r(1, 1280), pooled_end=r(1, 1280), pooled_negative=r(1, 1280), pooled_guide=r(1, 1280))
    if ip_tokens:
        d.update(ip_start=r(1, ip_tokens, cc), ip_end=r(1, ip_tokens, cc))
    return d


def kernel_rooflines(net, frames, atype, peak_tf):
    """The attention-core kernel in isolation, per attention-layer class of the UNet: median of 7 launches timed with
    CUDA events on the launching stream, L2 flushed (256 MB memset) between launches, random q/k/v of the layer's
    geometry.  Algorithmic flops per SURVEY.md 8d."""
    import torch
    from attention_interpolation_diffusion_b200 import _cabi
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    classes = {}
    for g in net.attention_geometry():
        key = (g["S"], g["L"], g["C"], g["heads"])
        classes[key] = classes.get(key, 0) + 1
    coef = torch.linspace(0, 1, frames, device="cuda")
    mode, mult = (_cabi.PAID_OUTER, 6) if atype == "fused_outer" else (_cabi.PAID_INNER, 4)
    out = []
    for (S, L, C, h), count in sorted(classes.items(), key=lambda kv: -kv[0][0] * kv[0][1]):
        q = torch.randn(frames, S, C, device="cuda").half()
        k = torch.randn(frames, L, C, device="cuda").half()
        v = torch.randn(frames, L, C, device="cuda").half()
        A = 2.0 * frames * S * L * C
        for name, m, fused, mul in ((atype, mode, True, mult), ("plain", _cabi.PAID_PLAIN, False, 2)):
            ts = []
            for i in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _cabi.attn_core(q, k, v, coef, h, m, fused)
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            ms = statistics.median(ts)
            tf = mul * A / ms / 1e9
            out.append({"S": S, "L": L, "C": C, "heads": h, "layers": count, "mode": name, "ms": round(ms, 4),
                        "achieved_tflops": round(tf, 1), "frac": round(tf / peak_tf, 3)})
    return out


def attention_traffic(net, mix):
    """roofline.traffic: ncu DRAM bytes (read + write) per attention-core launch, averaged over the launch mix of one
    sequence, from the committed `ncu --set full` capture (profiles/attn_traffic.json: one row per SDXL attention-layer
    class and mode at 7 frames; scaled linearly with the frame count).  None if no capture covers the geometry.
    mix: (mode, UNet forwards of that kind per sequence, frames per forward) -- the merged conditional + unconditional
    forwards after the warm-up steps carry twice the frames."""
    p = os.path.join(ROOT, "profiles", "attn_traffic.json")
    if not os.path.exists(p):
        return None, None, "no ncu capture committed"
    table = json.load(open(p))
    if table.get("kernel_source_sha") != attention_kernel_sha():
        return None, None, ("the committed ncu capture (profiles/attn_traffic.json) was taken on a different build of the attention "
                            "kernels (source hash mismatch): re-run tools/ncu_core.py + tools/attn_traffic.py")
    rows = {(r["S"], r["L"], r["heads"], r["mode"]): r for r in table["rows"]}
    total = launches = alg = 0.0
    for g in net.attention_geometry():
        for mode, reps, frames in mix:
            if reps == 0:
                continue
            r = rows.get((g["S"], g["L"], g["heads"], mode))
            if r is None:
                return None, None, "ncu capture does not cover this geometry"
            total += reps * (r["dram_read_bytes"] + r["dram_write_bytes"]) * frames / r["frames"]
            alg += reps * 2.0 * (2 * frames * g["S"] * g["C"] + 2 * frames * g["L"] * g["C"])   # Q, K, V read + H written
            launches += reps
    return total / launches, alg / launches, (f"launch-weighted mean over {int(launches)} launches of one sequence; per-shape rows in "
                              "profiles/attn_traffic.json (ncu --set full, tools/ncu_core.py)")


def attention_kernel_sha() -> str:
    """Hash of the attention-kernel sources: profiles/attn_traffic.json is only valid for the build it was captured on."""
    import hashlib
    h = hashlib.sha256()
    for f in ("attn_tc.cu", "attn_dw.cu", "sm100_ptx.cuh", "paid_common.cuh"):
        with open(os.path.join(ROOT, "attention_interpolation_diffusion_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["bf16_tflops_sustained"], j["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained bf16 GEMM)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference processors, timed on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_frames_per_sec(model: str, frames: int, atype: str, denoise_steps: int, budget_rows: int = 256):
    """Times the reference's attention path (oracle port of interpolation.py:573-804 hosted on the Attention
    stand-in; diffusers itself is not installed) on a BOUNDED sample and extrapolates to one whole sequence:
    per distinct attention-layer geometry of the UNet, one interpolated call and one plain call on `budget_rows`
    query rows of every frame (full keys), scaled by S / rows and by the number of such layers and forwards.
    Non-attention UNet blocks are NOT included (they are outside the hot path), which favours the CPU number."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import paid_oracle as O
    from attention_interpolation_diffusion_b200.unet_harness import CONFIGS, UNetHarness

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    with torch.device("meta"):
        geo = UNetHarness(CONFIGS[model]).attention_geometry()
    classes = {}
    for g in geo:
        key = (g["S"], g["L"], g["C"], g["Cc"], g["heads"], g["self_attn"])
        classes[key] = classes.get(key, 0) + 1
    mode = O.MODE_OUTER if atype == "fused_outer" else O.MODE_INNER
    coef = O.coefficients(frames, 4, 4)
    t_aid = t_plain = 0.0
    for (S, L, C, Cc, h, is_self), count in classes.items():
        w = O.make_layer(C, Cc, h, seed=1)
        x, ctx = O.make_inputs(frames, S, C, None if is_self else L, Cc, seed=1)
        rows = min(S, budget_rows)
        for m, fused in ((mode, True), (O.MODE_PLAIN, False)):
            t0 = time.perf_counter()
            q, k, v = O._project(x, ctx, w)                       # q/k/v projections, full size
            t1 = time.perf_counter()
            ends = (k[0], v[0], k[-1], v[-1])
            hid = torch.zeros_like(q)
            for n in range(frames):                               # attention core on the row sample
                hid[n:n + 1, :rows] = O._direct_core(q[n:n + 1, :rows], k[n:n + 1], v[n:n + 1], ends, coef[n:n + 1],
                                                     m, fused, (C // h) ** -0.5, h)
            t2 = time.perf_counter()
            _ = hid @ w.wo.T + w.bo                               # output projection, full size
            t3 = time.perf_counter()
            t_full = (t1 - t0) + (t3 - t2) + (t2 - t1) * (S / rows)
            if m == O.MODE_PLAIN:
                t_plain += count * t_full
            else:
                t_aid += count * t_full
    n_aid = int(denoise_steps * WARMUP_RATIO)
    n_plain = 2 * denoise_steps - n_aid
    seq_s = n_aid * t_aid + n_plain * t_plain
    sample = (f"oracle port, attention stack only: per layer geometry {budget_rows} query rows of each of {frames} frames "
              f"(full keys), 1 AID + 1 plain call, extrapolated x S/rows x layer count x ({n_aid} AID + {n_plain} plain forwards)")
    return frames / seq_s, cores, sample, {"t_aid_forward_s": t_aid, "t_plain_forward_s": t_plain}


def _reference_layer_runner(model: str, frames: int, atype: str):
    """(kind, geometry list, run(layer geometry, aid) -> seconds) for the reference arm.  kind "reference": the UNMODIFIED
    processors of /root/reference/interpolation.py on the Attention stand-in (only where that tree exists: the build
    container); kind "port": the oracle restatement of the same path (oracle/paid_oracle.py forward_chunked), which is what
    runs on the GPU box.  One call = one whole attention layer: q/k/v projections, attention of ALL frames and ALL query
    rows, output projection -- no row sampling."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import paid_oracle as O
    from attention_interpolation_diffusion_b200.unet_harness import CONFIGS, UNetHarness
    torch.set_num_threads(os.cpu_count())
    with torch.device("meta"):
        geo = UNetHarness(CONFIGS[model]).attention_geometry()
    mode = O.MODE_OUTER if atype == "fused_outer" else O.MODE_INNER
    coef = O.coefficients(frames, 4, 4)
    cache = {}

    def tensors(g):
        key = (g["S"], g["L"], g["C"], g["Cc"], g["heads"], g["self_attn"])
        if key not in cache:      # layers of one geometry share weights and inputs (the time does not depend on the values)
            w = O.make_layer(g["C"], g["Cc"], g["heads"], seed=1)
            x, ctx = O.make_inputs(frames, g["S"], g["C"], None if g["self_attn"] else g["L"], g["Cc"], seed=1)
            cache[key] = (w, x, ctx)
        return cache[key]

    kind = "port"
    ref_mod = None
    if os.path.isdir("/root/reference") and os.environ.get("PAID_BENCH_FORCE_PORT") != "1":
        try:
            from gen_golden import RefAttention, import_reference
            from gen_e2e_golden import StockProcessor
            ref_mod = import_reference()
            kind = "reference"
        except Exception:      # the reference tree is not importable here: fall back to the restatement
            ref_mod = None

    def run(g, aid: bool) -> float:
        w, x, ctx = tensors(g)
        t0 = time.perf_counter()
        with torch.no_grad():
            if ref_mod is not None:
                cls = ref_mod.OuterInterpolatedAttnProcessor if mode == O.MODE_OUTER else ref_mod.InnerInterpolatedAttnProcessor
                proc = cls(size=frames, is_fused=True, original_attn=StockProcessor())
                proc.coef = coef.clone()
                if not aid:
                    proc.deactivate()
                proc(RefAttention(w), x, encoder_hidden_states=ctx)
            elif aid:
                O.forward_chunked(x, ctx, w, coef, mode, True, rows=512)
            else:
                O.forward_chunked(x, ctx, w, coef, O.MODE_PLAIN, False, rows=512)
        return time.perf_counter() - t0

    return kind, geo, run


def run_reference_arm(args):
    """Reference arm: the reference's CPU implementation of the path on the box's host cores.  It cannot run K whole
    sequences inside a few minutes (one SDXL sequence is about 1.3 hours of CPU attention), so the K timed steps
    PARTITION one full pass over the attention stack: step k runs the layers k, k + K, k + 2K, ... of the UNet's 140 (32)
    attention layers, each once interpolated (AID) and once deactivated (plain), complete (all frames, all rows).  Summed
    over the steps that is exactly one AID forward and one plain forward of the attention stack, measured; the sequence
    time is n_aid * T_aid + n_plain * T_plain (the only extrapolation: x forwards per sequence, and x frames / 7 when the
    sequence has more than 7 frames -- the reference's cost per frame does not depend on the batch).  ms_per_step is the
    measured wall time of a step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = args.frames or FRAMES_PER_GPU * args.gpus
    timed_frames = min(frames, 7)
    kind, geo, run = _reference_layer_runner(args.model, timed_frames, args.atype)
    K, W = max(args.steps, 1), args.warmup
    small = min(geo, key=lambda g: g["S"] * g["L"])
    for _ in range(W):                                   # warm-up: thread pool, allocator, the cheapest layer
        run(small, True), run(small, False)
    t_aid = t_plain = 0.0
    step_s = []
    for k in range(K):
        t0 = time.perf_counter()
        for g in geo[k::K]:
            t_aid += run(g, True)
            t_plain += run(g, False)
        step_s.append(time.perf_counter() - t0)
    n_aid = int(args.denoise_steps * WARMUP_RATIO)
    n_plain = 2 * args.denoise_steps - n_aid
    seq_s = (n_aid * t_aid + n_plain * t_plain) * frames / timed_frames
    value = frames / seq_s
    cores = os.cpu_count()
    sample = (f"{'unmodified reference processors' if kind == 'reference' else 'oracle port of the reference processors'}, attention "
              f"stack only ({len(geo)} layers): every layer once interpolated + once deactivated, all {timed_frames} frames and all "
              f"query rows, partitioned over the {K} timed steps = one measured AID forward ({t_aid:.1f} s) + one measured plain forward "
              f"({t_plain:.1f} s); sequence = {n_aid} x AID + {n_plain} x plain forwards"
              + (f" x {frames}/{timed_frames} frames" if frames != timed_frames else "") + " (the only extrapolation)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * statistics.mean(step_s), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, frames),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample,
                             "t_aid_forward_s": t_aid, "t_plain_forward_s": t_plain, "measured_s": sum(step_s),
                             "sequence_s_extrapolated": seq_s},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, frames):
    name = {"sdxl": "SDXL 128x128 latent", "sd15": "SD1.5 64x64 latent", "tiny": "tiny test UNet"}[args.model]
    ip = f" + IP-Adapter image morphing ({args.ip_tokens} image tokens per frame)" if args.ip_tokens else ""
    if args.ip_tokens:
        cfg = "CFG (conditional + unconditional UNet pass per step)"
    elif args.gpus > 2:
        cfg = ("CFG (warm-up steps: an interpolated conditional and a stock unconditional UNet pass; afterwards both run stock "
               "attention as one UNet call with 2 n frames)")
    else:
        cfg = ("CFG (the conditional and the unconditional frames of a step run as one UNet call with 2 n frames: during the warm-up "
               "steps the attention layers interpolate the first n and run stock attention on the last n, afterwards all run stock "
               "attention)")
    return {"workload": f"{name}, {frames}-frame PAID (guide prompt){ip}, {args.atype} AID in all attention layers, "
                        f"{args.denoise_steps} steps, warmup_ratio {WARMUP_RATIO}, {cfg}",
            "frames": frames, "frames_per_gpu": frames // max(args.gpus, 1), "denoise_steps": args.denoise_steps,
            "parallelism": f"frame-sharded x{args.gpus}" if args.gpus > 1 else "single GPU",
            "l2": "working set (5.1 GB fp16 UNet weights + activations) is far larger than the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------------
def run_own_arm(args):
    import torch
    import torch.distributed as dist

    from attention_interpolation_diffusion_b200 import _cabi
    from attention_interpolation_diffusion_b200.pipeline import InterpolationPipeline
    from attention_interpolation_diffusion_b200.sharding import FrameShard
    from attention_interpolation_diffusion_b200.unet_harness import build_unet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wd = Watchdog(args.watchdog_s, rank, world)
    wd.phase("process group + UNet build")
    assert torch.cuda.is_available(), "bench.py (own arm) needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.load_library()
    torch.backends.cudnn.benchmark = True
    frames = args.frames or FRAMES_PER_GPU * world
    dtype = torch.float16
    net = build_unet(args.model, dev, dtype, seed=1002)
    shard = FrameShard(rank, world, frames, None) if world > 1 else None
    pipe = InterpolationPipeline(net, shard=shard, use_cuda_graphs=not args.no_graphs)

    def install(p_):
        if args.ip_tokens:
            torch.manual_seed(1002)
            p_.load_aid_ip_adapter(num_tokens=args.ip_tokens, scale=1.0, t=None, is_fused=True, early=args.atype, size=frames,
                                   alpha=4, beta=4)
        else:
            p_.load_aid(t=None, is_fused=True, atype=args.atype, size=frames, alpha=4, beta=4)

    install(pipe)
    host = make_host_inputs(net.cfg, dtype, args.ip_tokens)
    devin = {k: v.to(dev) for k, v in host.items()}
    kw = dict(size=frames, alpha=4.0, beta=4.0, num_inference_steps=args.denoise_steps, warmup_ratio=WARMUP_RATIO)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    step_dev = lambda: pipe.interpolate(**devin, **kw)

    def step_e2e():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        return pipe.interpolate(**d, **kw).float().cpu()

    for i in range(args.warmup):
        wd.phase(f"warm-up sequence {i}")
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:           # one nvidia-smi poller per job (rank 0's GPU), not one per rank
        sampler.start()
    launches0 = _cabi.launch_count() + pipe.graph_kernel_launches
    if args.ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    wd.phase("timed sequences", extra_s=30.0 * args.steps + (1e6 if args.ncu_range else 0))
    ms = timed(step_dev, args.steps)
    if args.ncu_range:
        torch.cuda.cudart().cudaProfilerStop()
    launches = _cabi.launch_count() + pipe.graph_kernel_launches - launches0
    clocks = sampler.stop()
    value = frames * args.steps / (ms / 1000.0)

    e2e = None
    if not args.no_e2e:
        wd.phase("end-to-end sequences", extra_s=30.0 * args.steps)
        out = step_e2e()
        ms_e = timed(step_e2e, args.steps)
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        e2e = {"value": frames * args.steps / (ms_e / 1000.0), "unit": "frames/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": out.numel() * 4, "finite": bool(out.isfinite().all())}

    # roofline of the dominant kernel: CUDA events around every attention-core launch of ONE more sequence, run
    # eagerly (kernels inside a CUDA-graph replay cannot be bracketed by events), same inputs, same launches
    wd.phase("eager sequence with the measurement hook")
    pipe.use_cuda_graphs = False
    _cabi.profile_read(reset=True)
    _cabi.profile_enable(True)
    ms_eager = timed(step_dev, 1)
    _cabi.profile_enable(False)
    prof_rows = _cabi.profile_rows(reset=True)
    attn_rows = [r for r in prof_rows if r["kind"] == "attention"]
    k_ms, k_launches, k_flops = (sum(r["ms"] for r in attn_rows), sum(r["launches"] for r in attn_rows),
                                 sum(r["flops"] for r in attn_rows))
    pipe.use_cuda_graphs = not args.no_graphs

    # frame-sharded run against the single-GPU run of the same sequence, once, outside the timed region (4 denoising steps)
    sharded_parity = None
    wd.phase("sharded parity + report", extra_s=600.0)     # rank 0 also times the isolated kernels and the CPU baseline here
    if world > 1:
        kw_p = dict(kw, num_inference_steps=4)
        local_out = pipe.interpolate(**devin, **kw_p)
        parts = [torch.empty(len(ids), *local_out.shape[1:], dtype=local_out.dtype, device=dev) for ids in shard.shards]
        for rk in range(world):
            if rk == rank:
                parts[rk].copy_(local_out)
            dist.broadcast(parts[rk], src=rk)
        if rank == 0:
            # the unsharded run: the reference's schedule (two UNet calls per step), launched eagerly -- the plainest path of
            # the pipeline, so that a problem in this check can never take the measured line with it
            try:
                gathered = shard.unshard(parts).float()
                single = InterpolationPipeline(net, shard=None, use_cuda_graphs=False, merge_plain_passes=False)
                install(single)
                ref_out = single.interpolate(**devin, **kw_p).float()
                rms = ref_out.pow(2).mean().sqrt()
                sharded_parity = {"rel_rms": float((gathered - ref_out).pow(2).mean().sqrt() / rms),
                                  "max_abs_over_rms": float((gathered - ref_out).abs().max() / rms),
                                  "bit_identical": bool(torch.equal(gathered, ref_out)), "frames": frames, "denoise_steps": 4,
                                  "broadcasts_per_aid_forward": sum(1 for g_ in net.attention_geometry() if g_["self_attn"]),
                                  "how": "all ranks' frames gathered on rank 0 vs the unsharded run of the same sequence on GPU 0 (one "
                                         "batch of all frames, two UNet calls per step, eager launches: cuDNN picks other convolution "
                                         "kernels for that batch size, so the two runs are not bit-identical here; "
                                         "tests/test_multirank_gpu.py shows bit-identity on equal conv batches)"}
                del single
            except Exception as e:      # noqa: BLE001 -- reported in the line, never fatal for the measurement
                sharded_parity = {"error": f"{type(e).__name__}: {e}"[:400], "frames": frames}
        dist.barrier()

    if rank == 0:
        peak_tf, _, peak_src = peaks()
        achieved = k_flops / (k_ms / 1000.0) / 1e12 if k_ms > 0 else None
        by_shape = kernel_rooflines(net, frames // world, args.atype, peak_tf)
        n_aid, fl = int(args.denoise_steps * WARMUP_RATIO), frames // world
        merged = pipe.merge_plain_passes and not args.ip_tokens      # the two plain passes of a post-warm-up step run as one
        mix = ([("interpolated", n_aid, fl), ("plain", n_aid, fl), ("plain", args.denoise_steps - n_aid, 2 * fl)] if merged else
               [("interpolated", n_aid, fl), ("plain", 2 * args.denoise_steps - n_aid, fl)])
        traffic, alg_bytes, traffic_how = attention_traffic(net, mix)
        roofline = {"kernel": f"attention core ({_cabi.last_kernel()})", "bound": "tensor", "achieved": achieved,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if achieved else None,
                    "peak_source": peak_src, "traffic": traffic, "traffic_unit": "bytes per launch",
                    "traffic_how": traffic_how, "launches": k_launches,
                    "avg_launch_ms": k_ms / max(k_launches, 1), "share_of_step": k_ms / (ms / args.steps),
                    "eager_step_ms": ms_eager, "by_shape_isolated": by_shape,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "how": "CUDA events around every attention-core launch of one extra, eagerly launched sequence "
                           "after the timed region (the timed steps replay CUDA graphs); share_of_step = summed "
                           "kernel time / timed step; algorithmic flops per SURVEY.md 8d (fused-outer 6A, "
                           "fused-inner 4A, plain 2A; A = 2 N S L C)"}
        # every tcgen05 kernel of the path, per shape, in situ (same eager sequence, same events)
        modes = {0: "plain", 16: "outer_pure", 17: "outer_fused", 32: "inner_pure", 33: "inner_fused"}
        in_situ = []
        for r in sorted(prof_rows, key=lambda r: -r["ms"]):
            d = r["d"]
            shape = ({"S": d[0], "L": d[1], "C": d[2], "mode": modes.get(d[3], str(d[3]))} if r["kind"] == "attention"
                     else {"M": d[0], "N": d[1], "K": d[2], "groups": d[3]})
            tf = r["flops"] / (r["ms"] / 1000.0) / 1e12 if r["ms"] > 0 else None
            in_situ.append({"kernel": r["kind"], **shape, "launches": r["launches"], "ms": round(r["ms"], 3),
                            "avg_us": round(1000.0 * r["ms"] / max(r["launches"], 1), 2), "tflops": round(tf, 1) if tf else None,
                            "frac": round(tf / peak_tf, 3) if tf else None})
        gemm = [r for r in prof_rows if r["kind"] != "attention"]
        g_ms, g_fl = sum(r["ms"] for r in gemm), sum(r["flops"] for r in gemm)
        roofline["gemm"] = {"kernel": "projection / feed-forward GEMMs (linear_tc_pair_kernel, GEGLU epilogue)", "bound": "tensor",
                            "achieved": g_fl / (g_ms / 1000.0) / 1e12 if g_ms > 0 else None, "peak": peak_tf, "unit": "TFLOP/s",
                            "frac": g_fl / (g_ms / 1000.0) / 1e12 / peak_tf if g_ms > 0 else None,
                            "launches": sum(r["launches"] for r in gemm), "share_of_step": g_ms / (ms / args.steps)}
        roofline["in_situ_by_shape"] = in_situ
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if (args.frames and world > 1) else "weak",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": workload_config(args, frames),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline}
        if sharded_parity is not None:
            line["sharded_parity"] = sharded_parity
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample, extra = cpu_reference_frames_per_sec(args.model, frames, args.atype, args.denoise_steps)
            line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample, **extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured forwards hold NCCL work on the communicator: drop them before tearing the process group down, and never
        # let a stuck teardown keep the GPUs (the line above is already out)
        threading.Timer(30.0, lambda: os._exit(0)).start()
        pipe._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_own_arm(a)
