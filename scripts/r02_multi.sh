#!/bin/bash
# the driver's scaling command on N GPUs of one box (run under gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps ${STEPS:-300} --warmup 5 ) > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -4 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['config']['workload'], 'n_gpus', d['n_gpus'], round(d['value']), 'frames/s', round(d['ms_per_step']*1e3, 1), 'us/step path', round(d['roofline_path']['frac'], 3))
print('e2e', d['e2e']); print('api', d['value_api'])
for k, v in (d['extra'] or {}).get('configs', {}).items(): print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != 'quality'})
PY
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 ) 2>&1 | grep -E "^\{|real" | cut -c1-200
