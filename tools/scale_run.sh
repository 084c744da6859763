#!/bin/bash
# usage: tools/scale_run.sh N [extra bench args]   -- prints one summary line for an N-GPU C4 bench
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N "$@" 2>&1 | grep -E "^\{" > gpurun_out/scale_$N.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_$N.json"))
    print("${EUC_GATHER:-p2p} N=$N fps=%.0f ms=%.3f e2e=%.0f golden=%s stages=%s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("frame_matches_golden_crc"), {k: round(v,4) for k,v in d["stage_ms_per_launch"].items() if v}))
except Exception as ex:
    print("N=$N FAILED", ex)
PY
