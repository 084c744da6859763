#!/bin/bash
# A/B bench of build/ab variants + one full ncu capture of the C4 raster kernel of the in-tree build
TAG=${TAG:-r2a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
WLS="${WLS:-c4}" bash tools/ab_run.sh
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_raster_c4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_*
