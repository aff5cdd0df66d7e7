"""Top stalled SASS instructions of one kernel from an .ncu-rep source page.
usage: python tools/ncu_stalls.py rep kernel_regex [top_n]"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) >= len(rows[h]) and r[0].startswith("0x")]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: 0 for s in stalls}
recs = []
for d in data:
    n = int(d[ix["# Samples"]] or 0)
    st = {s: int(d[ix[s]] or 0) for s in stalls}
    for s in stalls:
        tot[s] += st[s]
    recs.append((n, d[ix["Source"]].strip(), st, int(d[ix["Instructions Executed"]] or 0)))
total = sum(r[0] for r in recs)
print(f"# {kre}: {len(recs)} SASS instructions, {total} samples, {sum(r[3] for r in recs)} warp-instructions executed")
print("# stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(tot.items(), key=lambda x: -x[1])[:9]))
top = sorted(range(len(recs)), key=lambda i: -recs[i][0])[:topn]
for i in sorted(top):
    n, src, st, ex = recs[i]
    main = max(st.items(), key=lambda x: x[1])
    print(f"{i:5d} {100*n/total:5.1f}% ex={ex:8d} {main[0][6:]:16s} {src[:90]}")
