#!/bin/bash
# the driver's two bench commands on one B200 (default arguments except the step count)
set -u
mkdir -p gpurun_out
( time python bench.py --steps ${STEPS:-300} --warmup 5 ) > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -5 gpurun_out/r02_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print(d['config']['workload'], round(d['value']), 'frames/s', round(d['ms_per_step']*1e3, 1), 'us/step path', round(d['roofline_path']['frac'], 3), 'kernel', round(d['roofline']['frac'], 3), d['clocks'])
print('e2e', d['e2e']); print('api', d['value_api']); print('cpu', d['cpu_baseline']); print('port', d['cpu_baseline_port'])
for k, v in (d['extra'] or {}).get('configs', {}).items(): print(k, v)
PY
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
cut -c1-300 gpurun_out/r02_bench_ref.json; tail -4 gpurun_out/r02_bench_ref.err; nproc
