"""Per-shape DRAM traffic of the attention core from an `ncu --set full` capture of tools/ncu_core.py (run here, no GPU):
    python tools/attn_traffic.py gpurun_out/x_attn_core.ncu-rep profiles/attn_traffic.json
bench.py weights these by the launch mix of the timed sequence to report roofline.traffic (bytes per launch)."""
import csv, json, subprocess, sys
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from ncu_core import CLASSES
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def val(d, m):
    f = float(d[ix[m]].replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "%": 1}.get(units[ix[m]], 1)
data = [d for d in data if "attn_tc" in d[ix["Kernel Name"]] or "attn_dw" in d[ix["Kernel Name"]]]
assert len(data) == 2 * len(CLASSES), len(data)
table = []
for i, d in enumerate(data):
    S, L, h = CLASSES[i // 2]
    table.append({"S": S, "L": L, "heads": h, "mode": "interpolated" if i % 2 == 0 else "plain", "frames": 7,
                  "kernel": "attn_dw_kernel" if "attn_dw" in d[ix["Kernel Name"]] else "attn_tc_kernel",
                  "dram_read_bytes": val(d, "dram__bytes_read.sum"), "dram_write_bytes": val(d, "dram__bytes_write.sum"),
                  "duration_us_under_ncu": val(d, "gpu__time_duration.sum"),
                  "tensor_pipe_pct": val(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  "xu_pipe_pct": val(d, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                  "dram_pct": val(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench
json.dump({"source": rep.rsplit("/", 1)[-1], "kernel_source_sha": bench.attention_kernel_sha(), "how": "ncu --set full --clock-control none, one launch per row (tools/ncu_core.py)",
           "rows": table}, open(out, "w"), indent=1)
print(json.dumps(table, indent=1))
