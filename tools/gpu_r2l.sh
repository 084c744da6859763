#!/bin/bash
# follow-up of gpu_r2k.sh: the group test that failed (in-tree library, then the previous commit's), C4 and the 4096-icon batch for the one-tile-per-ticket variant
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_async_group.py -x -q -k "gathers" 2>&1 | tail -60
echo "---- base library"
EUC_B200_LIB=$PWD/build/ab/libeuc_base.so timeout 300 python -m pytest tests/test_async_group.py -x -q -k "gathers" 2>&1 | tail -15
run() {  # run <lib> <workload> <bench args...>
  local lib=$1 wl=$2; shift 2
  [ -f $lib ] || return
  EUC_B200_LIB=$PWD/$lib timeout 240 python bench.py --workload $wl --no-cpu-baseline "$@" > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$lib" "$wl" <<PY
import json,sys
try:
    d=json.load(open("gpurun_out/ab.json"))
    print(sys.argv[2], sys.argv[1].split("/")[-1], "value=%.1f ms=%.4f graph=%s golden=%s stages=%s"%(d["value"], d["ms_per_step"], d.get("cuda_graph_ms_per_step"), d.get("frame_matches_golden_crc"), {k: round(v,4) for k,v in d["stage_ms_per_launch"].items() if v}))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex); print(open("gpurun_out/ab.err").read()[-600:])
PY
}
for lib in build/ab/libeuc_chunk1.so build/ab/libeuc_base.so build/ab/libeuc_chunk1.so; do run $lib c4 --no-icon-batch; done
for lib in build/ab/libeuc_base.so build/ab/libeuc_chunk1.so; do run $lib c5 --icons 4096 --steps 5; done
