#!/bin/bash
# the icon batch that rides along with the default C4 line: repeatability, with / without the clock sampler, 5 / 20 / 40 batches
mkdir -p gpurun_out
show() { python - "$1" "$2" <<PY
import json,sys
try:
    d=json.load(open(sys.argv[1])); ib=d.get("icon_batch")
    print(sys.argv[2], "c4 value=%.1f"%d["value"], "icon_batch=%.0f ms=%.4f stages=%s"%(ib["value"], ib["ms_per_batch"], {k: round(v,4) for k,v in ib["stage_ms_per_launch"].items() if v}) if ib else "")
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
python bench.py --no-cpu-baseline > gpurun_out/o1.json 2>/dev/null; show gpurun_out/o1.json default-1
python bench.py --no-cpu-baseline > gpurun_out/o2.json 2>/dev/null; show gpurun_out/o2.json default-2
EUC_BENCH_NOSAMPLER=1 python bench.py --no-cpu-baseline > gpurun_out/o3.json 2>/dev/null; show gpurun_out/o3.json nosampler
python bench.py --no-cpu-baseline --icon-steps 5 > gpurun_out/o4.json 2>/dev/null; show gpurun_out/o4.json icon-steps-5
python bench.py --no-cpu-baseline --icon-steps 40 > gpurun_out/o5.json 2>/dev/null; show gpurun_out/o5.json icon-steps-40
python bench.py --workload c5 --steps 20 --no-cpu-baseline > gpurun_out/o6.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/o6.json')); print('c5 standalone 20 steps', d['value'], d['ms_per_step'], d['stage_ms_per_launch'])"
EUC_B200_LIB=$PWD/build/ab/libeuc_base.so python bench.py --no-cpu-baseline > gpurun_out/o7.json 2>/dev/null; show gpurun_out/o7.json base-lib
