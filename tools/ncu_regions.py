"""Buckets an ncu source page by source-line ranges of kernels.cuh (given as name:lo-hi,... ; unmatched lines are listed).
usage: python tools/ncu_regions.py report.ncu-rep "name:lo-hi,lo-hi;name2:lo-hi" """
import csv, io, subprocess, sys
rep = sys.argv[1]
regions = []
for part in sys.argv[2].split(";"):
    name, rs = part.split(":")
    regions.append((name, [tuple(map(int, r.split("-"))) for r in rs.split(",")]))
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout  # a .csv = an exported source page
rows = list(csv.reader(io.StringIO(out)))
ci = None
acc = {n: [0.0, 0.0, 0.0] for n, _ in regions}
acc["other"] = [0.0, 0.0, 0.0]
others = []
for r in rows:
    if "Instructions Executed" in r and r[0] == "Line No":
        ci = {n: r.index(n) for n in ("# Samples", "Instructions Executed", "Thread Instructions Executed")}; continue
    if ci is None or len(r) < 10 or r[0] in ("", "Line No"):
        continue
    try:
        ln = int(r[0]); inst = float(r[ci["Instructions Executed"]]); th = float(r[ci["Thread Instructions Executed"]]); smp = float(r[ci["# Samples"]])
    except ValueError:
        continue
    for name, rs in regions:
        if any(lo <= ln <= hi for lo, hi in rs):
            a = acc[name]; break
    else:
        a = acc["other"]; others.append((inst, ln, r[1].strip()[:80]))
    a[0] += inst; a[1] += th; a[2] += smp
ti = sum(a[0] for a in acc.values()); ts = sum(a[2] for a in acc.values())
print(f"total warp instructions {ti:.4e}")
for n, a in acc.items():
    print(f"{n:14s} {a[0]/1e6:8.1f} M inst {a[0]/ti*100:5.1f}%  thr {a[1]/max(a[0],1):4.1f}  samples {a[2]/max(ts,1)*100:5.1f}%")
for inst, ln, src in sorted(others, reverse=True)[:12]:
    print(f"   other {inst/1e6:6.1f} M  line {ln}: {src}")
