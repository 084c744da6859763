#!/bin/bash
# parity suite + quick bench lines (default driver command included)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_default.json"))
    print("c4 fps=%.1f ms=%.3f e2e=%.1f graph=%s waits=%s crc=%s frac=%.4f kern=%s stages=%s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("cuda_graph_ms_per_step"), d.get("host_waits_in_timed_region"), d.get("frame_matches_golden_crc"), d["roofline"]["frac"], d["roofline"]["kernel"], {k: round(v,4) for k,v in d["stage_ms_per_launch"].items() if v}))
    ib=d.get("icon_batch"); print("icon_batch", {k: ib[k] for k in ("value","ms_per_batch","crc_ok","icons_checked","icons_depth_crc_mismatch","icons_colour_crc_mismatch")}, ib["e2e"]["value"], ib["stage_ms_per_launch"])
    print("cpu", d.get("cpu_baseline"))
except Exception as ex:
    print("default bench failed", ex); print(open("gpurun_out/r2_bench_default.err").read()[-3000:])
PY
for wl in ${WLS:-c1 c2 c3}; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_$wl.json"))
    print("$wl", "fps=%.1f ms=%.4f e2e=%.1f graph=%s waits=%s ok=%s frac=%.4f kern=%s stages=%s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("cuda_graph_ms_per_step"), d.get("host_waits_in_timed_region"), d.get("frame_matches_golden_crc"), d["roofline"]["frac"], d["roofline"]["kernel"], {k: round(v,4) for k,v in d["stage_ms_per_launch"].items() if v}))
except Exception as ex:
    print("$wl failed", ex); print(open("gpurun_out/r2_bench_$wl.err").read()[-2000:])
PY
done
