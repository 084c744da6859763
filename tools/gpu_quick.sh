#!/bin/bash
# quick iteration: parity tests + bench lines (no ncu)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for wl in ${WLS:-c4 c1 c2 c3}; do
  python bench.py --workload $wl --no-cpu-baseline > gpurun_out/q_$wl.json 2> gpurun_out/q_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/q_$wl.json"))
    print("$wl", "fps=%.1f ms=%.3f e2e=%.1f mfrag=%.0f frac=%.4f stages=%s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["mfrag_per_s"], d["roofline"]["frac"], {k: round(v,4) for k,v in d["stage_ms_per_launch"].items()}))
except Exception as ex:
    print("$wl failed", ex); print(open("gpurun_out/q_$wl.err").read()[-1500:])
PY
done
python bench.py --workload c5 --icons 1024 --steps 5 --no-cpu-baseline > gpurun_out/q_c5.json 2> gpurun_out/q_c5.err
python -c "
import json
d=json.load(open('gpurun_out/q_c5.json')); print('c5 icons/s=%.0f ms=%.3f'%(d['value'], d['ms_per_step']), {k: round(v,4) for k,v in d['stage_ms_per_launch'].items()})" || tail -5 gpurun_out/q_c5.err
