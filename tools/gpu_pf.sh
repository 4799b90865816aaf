#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_data_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -3
B="timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline"
run() { tag=$1; shift; "$@" > gpurun_out/pf_$tag.json 2> gpurun_out/pf_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/pf_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['ms_per_step'],2), 'e2e', round(64e3/d['e2e']['value'],2), d['e2e']['value'], d['step_ms'], d['clocks']['sm_mhz'])
"; }
run a1 $B
run a2 $B
