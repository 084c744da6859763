"""Aggregates an ncu source page (needs -lineinfo and --import-source on) per CUDA source line: executed warp
instructions, average active threads per instruction, stall samples.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout  # a .csv = an exported source page
rows = list(csv.reader(io.StringIO(out)))
allrows, fname, ci = [], "?", None
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if "Instructions Executed" in r and r[0] == "Line No":
        ci = {n: r.index(n) for n in ("# Samples", "Instructions Executed", "Thread Instructions Executed")}; continue
    if ci is None or len(r) < 10 or r[0] in ("", "Line No"):
        continue
    try:
        inst = float(r[ci["Instructions Executed"]]); th = float(r[ci["Thread Instructions Executed"]]); smp = float(r[ci["# Samples"]])
    except ValueError:
        continue
    allrows.append((inst, th, smp, fname, r[0], r[1].strip()[:105]))
tot_inst = sum(r[0] for r in allrows); tot_s = sum(r[2] for r in allrows); tot_th = sum(r[1] for r in allrows)
print(f"total warp instructions {tot_inst:.3e}, avg threads/inst {tot_th/tot_inst:.1f}, samples {tot_s:.0f}")
for inst, th, smp, f, ln, src in sorted(allrows, key=lambda r: -r[0])[:top]:
    print(f"{inst/tot_inst*100:5.1f}% inst {smp/max(tot_s,1)*100:5.1f}% smp thr {th/max(inst,1):4.1f} {f}:{ln:>4} {src}")
