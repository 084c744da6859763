#!/bin/bash
# full ncu captures of the C4 raster and setup kernels (current in-tree build) + GPU parity suite
TAG=${TAG:-r1e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_raster_c4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:setup_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_setup_c4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/${TAG}_*
