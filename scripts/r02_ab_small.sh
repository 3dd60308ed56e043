#!/bin/bash
# A/B on one B200: resident pixel-pass CTAs per SM for batches of up to 32 frames
# (NPB_PAIR_SMALL_CTAS = 1 | 2 | 3; the library default is 2), headline step, every setting twice
for rep in 1 2; do for n in 1 2 3; do
NPB_PAIR_SMALL_CTAS=$n timeout 300 python bench.py --config nyuv2 --steps 300 --warmup 5 --no-e2e --no-cpu-baseline --no-api --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('small_ctas $n', round(d['value']), 'frames/s', round(d['ms_per_step']*1e3,1),'us')"
done; done
