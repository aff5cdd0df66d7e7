"""Compact per-launch table from an .ncu-rep (run here, no GPU needed):  python tools/ncu_summary.py rep [out.txt]"""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "dur_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("lts__t_sector_hit_rate.pct", "l2hit%")]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
lines = ["kernel".ljust(28) + " ".join(n.rjust(10) for _, n in COLS)]
for d in data:
    name = d[ix["Kernel Name"]]
    name = "attn_tc_kernel" if "attn_tc" in name else ("linear_tc_kernel" if "linear_tc" in name else name[:26])
    vals = []
    for m, _ in COLS:
        v = d[ix[m]] if m in ix else ""
        try:
            f = float(v.replace(",", ""))
            if units[ix[m]] == "byte": f /= 1e6
            if units[ix[m]] == "Kbyte": f /= 1e3
            if units[ix[m]] == "Gbyte": f *= 1e3
            if units[ix[m]] == "ms": f *= 1e3
            if units[ix[m]] == "ns": f /= 1e3
            vals.append(f"{f:10.2f}")
        except ValueError:
            vals.append(v[:10].rjust(10))
    lines.append(name.ljust(28) + " ".join(vals))
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(f"# {rep}: ncu --set full --clock-control none, one row per captured launch\n" + out + "\n")
