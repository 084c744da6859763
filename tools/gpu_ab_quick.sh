#!/bin/bash
# parity suite on the in-tree library, then C4 / icon-batch bench lines for the in-tree library and every variant in build/ab
# (variants: one switch of kernels.cuh flipped with -D, built with the flags of __graft_entry__.py)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -${TAIL:-25}
run() {  # run <lib> <workload> <bench args...>
  local lib=$1 wl=$2; shift 2
  [ -f $lib ] || return
  EUC_B200_LIB=$PWD/$lib timeout 240 python bench.py --workload $wl --no-cpu-baseline "$@" > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$lib" "$wl" "$*" <<PY
import json,sys
try:
    d=json.load(open("gpurun_out/ab.json"))
    print(sys.argv[2], sys.argv[3], sys.argv[1].split("/")[-1], "value=%.1f ms=%.4f golden=%s stages=%s"%(d["value"], d["ms_per_step"], d.get("frame_matches_golden_crc"), {k: round(v,4) for k,v in d["stage_ms_per_launch"].items() if v}))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex); print(open("gpurun_out/ab.err").read()[-600:])
PY
}
IN=euc_b200/csrc/libeuc_b200.so
for lib in $IN build/ab/*.so $IN; do run $lib c4 --no-icon-batch; done
for lib in $IN build/ab/*.so $IN; do run $lib c5 --icons 4096 --steps 5; done
for wl in ${EXTRA_WLS:-}; do for lib in $IN build/ab/*.so; do run $lib $wl; done; done
