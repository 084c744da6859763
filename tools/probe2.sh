for v in "" "EUC_BENCH_NOPROF=1" "EUC_BENCH_NOSAMPLER=1"; do
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 --no-cpu-baseline --no-icon-batch --steps 50 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['stage_ms_per_launch'], d['e2e']['ms_per_step'])"
done
