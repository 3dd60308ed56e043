#!/bin/bash
set -u
N=${1:-2}
for bind in 0 1; do
  if [ $bind = 0 ]; then export NPB_BENCH_NO_BIND=1; else unset NPB_BENCH_NO_BIND; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$bind \
    bench.py --gpus $N --steps 100 --warmup 5 --no-extra --no-api --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); e=d['e2e']; print('bind=$bind', round(d['value']), 'e2e', round(e['value']), 'h2d/rank', round(e['h2d_gbs_per_rank'],1), 'ceiling', round(e['h2d_ceiling_gbs_per_rank'],1), 'cores', e['rank_bound_to_gpu_cores'])"
done
nvidia-smi topo -m 2>/dev/null | head -12; nproc; numactl -H 2>/dev/null | head -5
