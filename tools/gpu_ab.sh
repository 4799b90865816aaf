#!/bin/bash
# A/B of the BatchNorm-sums-in-the-data-gradient-epilogue choices now that the reduction pass could hide behind a weight gradient
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline"
run() { tag=$1; shift; "$@" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/ab_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['ms_per_step'],2), 'e2e', round(64e3/d['e2e']['value'],2), d['step_ms'], d['clocks']['sm_mhz'], d['roofline']['elementwise_hbm_frac'])
"; }
run all $B
EEL_BNSUMS_64=0 run no64 $B
EEL_BNSUMS_WIDE=0 run nowide $B
EEL_BNSUMS_64=0 EEL_BNSUMS_WIDE=0 run none $B
run all2 $B
