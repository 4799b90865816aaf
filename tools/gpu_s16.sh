#!/bin/bash
# which NVML query stalls a step?
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-eager-baseline --clock-sampler nvml"
for cfg in 50,cr 200,cr 50,c 50,r 1000,cr; do
  EEL_BENCH_NVML=$cfg $B > gpurun_out/s16_$cfg.json 2> gpurun_out/s16_$cfg.err
  python -c "
import json
d=json.loads(open('gpurun_out/s16_$cfg.json').read().strip().splitlines()[-1])
print('$cfg', round(d['ms_per_step'],2), 'e2e', round(64e3/d['e2e']['value'],2), d['step_ms'], d['clocks'])
"
done
