#!/bin/bash
# which clock sampler perturbs the timed region? per-step spread with nvml thread / nvidia-smi child / none
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline"
for m in nvml smi none nvml smi none; do
  $B --clock-sampler $m > gpurun_out/s15_$m.json 2> gpurun_out/s15_$m.err
  python -c "
import json
d=json.loads(open('gpurun_out/s15_$m.json').read().strip().splitlines()[-1])
print('$m', round(d['ms_per_step'],2), 'e2e', round(64e3/d['e2e']['value'],2), d['step_ms'], d['clocks'])
"
done
