#!/bin/bash
# runs bench (C4 by default) for every variant in build/ab plus the in-tree library; WLS="c4 c5" selects workloads
mkdir -p gpurun_out
for wl in ${WLS:-c4}; do
  extra=""; [ $wl = c5 ] && extra="--icons 1024 --steps 5"
  for lib in euc_b200/csrc/libeuc_b200.so build/ab/*.so; do
    [ -f $lib ] || continue
    EUC_B200_LIB=$PWD/$lib python bench.py --workload $wl --no-cpu-baseline $extra > gpurun_out/ab.json 2> gpurun_out/ab.err
    python - "$lib" "$wl" <<PY
import json,sys
try:
    d=json.load(open("gpurun_out/ab.json"))
    print(sys.argv[2], sys.argv[1].split("/")[-1], "value=%.1f ms=%.4f golden=%s frags=%s stages=%s"%(d["value"], d["ms_per_step"], d.get("frame_matches_golden_crc"), d.get("fragments_per_step"), {k: round(v,4) for k,v in d["stage_ms_per_launch"].items() if v}))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex); print(open("gpurun_out/ab.err").read()[-600:])
PY
  done
done
