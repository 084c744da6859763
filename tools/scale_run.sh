#!/bin/bash
# scaling sweep on one box: N in $NS (default "1 2"), driver-style launch; TAG names the outputs
mkdir -p gpurun_out
for n in ${NS:-1 2}; do
  if [ $n = 1 ]; then
    timeout 900 python bench.py --gpus 1 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${TAG:-r2}_scale_n$n.json 2> gpurun_out/${TAG:-r2}_scale_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${TAG:-r2}_scale_n$n.json 2> gpurun_out/${TAG:-r2}_scale_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG:-r2}_scale_n$n.json") if l.startswith("{")][-1])
    ib=d.get("icon_batch") or {}
    print("N=$n c4 fps=%.1f ms=%.4f e2e=%.1f crc=%s launches=%s stages=%s | icons/s=%s e2e=%s crc=%s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("frame_matches_golden_crc"), d["gpu_launches"], {k: round(v,4) for k,v in d["stage_ms_per_launch"].items() if v}, ib.get("value"), (ib.get("e2e") or {}).get("value"), ib.get("crc_ok")))
except Exception as ex:
    print("N=$n failed", ex); print(open("gpurun_out/${TAG:-r2}_scale_n$n.err").read()[-3000:])
PY
done
