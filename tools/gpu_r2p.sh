#!/bin/bash
# the ride-along icon batch with per-call host times (EUC_BENCH_TRACE) after the collector was taken out of the timed loops
mkdir -p gpurun_out
for i in 1 2 3 4; do
  EUC_BENCH_TRACE=1 python bench.py --no-cpu-baseline > gpurun_out/p$i.json 2> gpurun_out/p$i.err
  python - gpurun_out/p$i.json <<PY
import json,sys
d=json.load(open(sys.argv[1])); ib=d["icon_batch"]
print("c4 value=%.1f icon_batch=%.0f ms=%.4f waits=%s"%(d["value"], ib["value"], ib["ms_per_batch"], ib.get("host_waits_in_timed_region")))
PY
  grep "trace" gpurun_out/p$i.err | grep "c5" | cut -c1-400
done
