#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; grep -E "Error|passed|failed|FAILED" gpurun_out/t.log | head -6
for r in 0 1; do for wl in c4 c5; do
  extra=""; [ $wl = c5 ] && extra="--icons 1024 --steps 5"
  EUC_RASTER2=$r python bench.py --workload $wl --no-cpu-baseline --no-icon-batch $extra > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - $r $wl <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("raster2=%s %s value=%.1f ms=%.4f ok=%s frags=%s stages=%s"%(sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], d.get("frame_matches_golden_crc"), d.get("fragments_per_step"), {k: round(v,4) for k,v in d["stage_ms_per_launch"].items() if v}))
except Exception as ex:
    print("FAILED", sys.argv[1:], ex); print(open("gpurun_out/ab.err").read()[-800:])
PY
done; done
