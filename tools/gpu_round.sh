#!/bin/bash
# One GPU session: parity tests, smoke, bench lines for all workloads, ncu launch list + full capture of the raster kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for wl in c4 c1 c2 c3; do
  python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -c 2500 gpurun_out/bench_$wl.json; tail -5 gpurun_out/bench_$wl.err
done
python bench.py --workload c5 --icons 1024 --steps 5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 2500 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c4.json 2>&1; cat gpurun_out/bench_ref_c4.json
nproc
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 2 -o gpurun_out/prof_raster_c4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
