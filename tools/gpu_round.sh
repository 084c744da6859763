#!/bin/bash
# One GPU session: full parity suite, smoke, bench lines for all workloads, CPU arm, ncu launch list + full capture.
TAG=${TAG:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for wl in c4 c1 c2 c3; do
  python bench.py --workload $wl > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
done
python bench.py --workload c5 --icons 1024 --steps 5 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_c4.json 2>&1
python - <<PY
import json
for wl in ("c4","c1","c2","c3","c5"):
    try:
        d=json.load(open("gpurun_out/${TAG}_bench_%s.json"%wl))
        cb=d.get("cpu_baseline",{})
        print(wl, "value=%.1f ms=%.3f e2e=%.1f frac=%.4f cpu=%s cores=%s launches=%d stages=%s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], cb.get("value"), cb.get("cores"), d["gpu_launches"], {k: round(v,4) for k,v in d["stage_ms_per_launch"].items() if v}))
    except Exception as ex:
        print(wl, "FAILED", ex); print(open("gpurun_out/${TAG}_bench_%s.err"%wl).read()[-800:])
print(open("gpurun_out/${TAG}_bench_ref_c4.json").read()[:300])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_c4.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_raster_c4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:setup_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_setup_c4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls gpurun_out | head -40
