"""CPU: the parts of bench.py that do not need a GPU -- the reference arm's JSON line (on the tiny geometry so it takes
seconds) and the committed ncu traffic table covering every attention-layer class of the headline workload."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny", "--frames", "3",
                          "--steps", "1", "--warmup", "0", "--denoise-steps", "4"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["vs_baseline"] is None
    # the unmodified reference processors where /root/reference exists (this container), the oracle port elsewhere (GPU box)
    assert line["cpu_baseline"]["kind"] == ("reference" if os.path.isdir("/root/reference") else "port")
    assert line["cpu_baseline"]["cores"] == os.cpu_count()
    # the line reports what it timed: one AID + one plain forward of the attention stack, partitioned over the steps
    cb = line["cpu_baseline"]
    assert abs(cb["measured_s"] - line["ms_per_step"] * line["steps"] / 1000.0) < 1e-6 * max(cb["measured_s"], 1)
    assert cb["t_aid_forward_s"] + cb["t_plain_forward_s"] <= cb["measured_s"] * 1.001
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_port_path():
    """The path the GPU box takes (no /root/reference there): the oracle port."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny", "--frames", "3",
                          "--steps", "2", "--warmup", "1", "--denoise-steps", "4"], capture_output=True, text=True, timeout=300,
                         env={**os.environ, "PAID_BENCH_FORCE_PORT": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0 and line["steps"] == 2


def test_committed_ncu_traffic_covers_the_headline_workload():
    sys.path.insert(0, ROOT)
    import bench
    from attention_interpolation_diffusion_b200.unet_harness import CONFIGS, UNetHarness
    with torch.device("meta"):
        net = UNetHarness(CONFIGS["sdxl"])
    traffic, alg, how = bench.attention_traffic(net, [("interpolated", 25, 7), ("plain", 25, 7), ("plain", 25, 14)])
    if traffic is None and "source hash mismatch" in how:
        import pytest
        pytest.skip("profiles/attn_traffic.json is stale for this build of the kernels (bench.py then reports traffic: null)")
    assert traffic is not None and alg is not None, how
    assert 0.3 * alg < traffic < 3 * alg, (traffic, alg)      # DRAM traffic of the same order as the algorithmic bytes


def test_watchdog_ends_a_stalled_run():
    """bench.Watchdog: a phase without progress ends the process with exit code 17 and a diagnostic on stderr (a stuck
    collective must not hold the GPUs until the caller's own limit); a stopped watchdog does nothing."""
    import subprocess
    code = ("import sys, time; sys.path.insert(0, %r); import bench; w = bench.Watchdog(0.5, 3, 8); w.phase('warm-up sequence 0'); "
            "time.sleep(30)") % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=25)
    assert res.returncode == 17, (res.returncode, res.stderr[-500:])
    msg = json.loads(next(l for l in res.stderr.splitlines() if l.startswith('{')))      # followed by the threads' stack dump
    assert "Thread" in res.stderr or "File" in res.stderr
    assert msg["error"] == "bench watchdog" and msg["phase"] == "warm-up sequence 0" and msg["rank"] == 3 and msg["world"] == 8
    code = ("import sys, time; sys.path.insert(0, %r); import bench; w = bench.Watchdog(0.5, 0, 1); w.stop(); time.sleep(7); "
            "print('alive')") % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=25)
    assert res.returncode == 0 and "alive" in res.stdout
