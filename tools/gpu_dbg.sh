#!/bin/bash
# debugging session: one failing test with full output, then a full ncu capture of the C4 raster kernel
mkdir -p gpurun_out
python -m pytest tests/test_runtime_pipeline.py -m gpu -x -q 2>&1 | tail -60 > gpurun_out/dbg_test.log
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/${TAG:-r1d}_prof_raster_c4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
