#!/bin/bash
# Round evidence on one GPU: parity suite, smoke, bench lines of every workload (default driver command first), CPU arm,
# ncu launch list + full captures.  TAG names the outputs under gpurun_out/.
TAG=${TAG:-r2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
for wl in c1 c2 c3; do
  timeout 600 python bench.py --workload $wl > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
done
timeout 600 python bench.py --workload c5 --icons 4096 --steps 5 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_c4.json 2>&1
python - <<PY
import json
for wl in ("c4","c1","c2","c3","c5"):
    try:
        d=json.load(open("gpurun_out/${TAG}_bench_%s.json"%wl))
        cb=d.get("cpu_baseline",{})
        print(wl, "value=%.1f ms=%.4f e2e=%.1f graph=%s waits=%s ok=%s frac=%.4f kern=%s cpu=%s cores=%s launches=%d stages=%s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("cuda_graph_ms_per_step"), d.get("host_waits_in_timed_region"), d.get("frame_matches_golden_crc"), d["roofline"]["frac"], d["roofline"]["kernel"], cb.get("value"), cb.get("cores"), d["gpu_launches"], {k: round(v,4) for k,v in d["stage_ms_per_launch"].items() if v}))
        if wl == "c4": print("   icon_batch", {k: d["icon_batch"][k] for k in ("value","ms_per_batch","crc_ok","icons_checked")}, "e2e", d["icon_batch"]["e2e"]["value"])
    except Exception as ex:
        print(wl, "FAILED", ex); print(open("gpurun_out/${TAG}_bench_%s.err"%wl).read()[-800:])
print(open("gpurun_out/${TAG}_bench_ref_c4.json").read()[:400])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches_c4.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-icon-batch > gpurun_out/ncu_b.log 2>&1
capture() {  # capture <kernel regex> <name> <bench args...>: full ncu capture, exported as raw + source CSV pages (the .ncu-rep files exceed what comes back)
  local k=$1 name=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/${TAG}_$name -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_full_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_${name}_source.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$name.ncu-rep
}
CAPS=${CAPS:-"raster_c4 setup_c4 resolve_c3 raster_c5"}  # which captures to take (each costs about a minute of GPU time)
for c in $CAPS; do
  case $c in
    raster_c4) capture raster_kernel raster_c4 --no-icon-batch ;;
    setup_c4) capture setup_kernel setup_c4 --no-icon-batch ;;
    resolve_c3) capture resolve_kernel resolve_c3 --workload c3 ;;
    raster_c5) capture raster_kernel raster_c5 --workload c5 --icons 1024 ;;
  esac
done
ls -la gpurun_out | grep ${TAG}_ | head -40
