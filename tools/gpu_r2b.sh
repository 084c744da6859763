#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/${TAG:-r2b}_prof_raster_c4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-icon-batch > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:setup_kernel -s 3 -c 1 -o gpurun_out/${TAG:-r2b}_prof_setup_c4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-icon-batch > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/${TAG:-r2b}_*
