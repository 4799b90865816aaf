#!/bin/bash
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-eager-baseline --clock-sampler nvml"
run() { tag=$1; shift; "$@" > gpurun_out/s17_$tag.json 2> gpurun_out/s17_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/s17_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['ms_per_step'],2), 'e2e', round(64e3/d['e2e']['value'],2), d['step_ms'], d['allocator_in_timed_region'], d['peak_mem_gb'])
"; }
run a1 $B
run a2 $B
run a3 $B
run nostream1 $B --no-wgrad-stream
run nostream2 $B --no-wgrad-stream
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True run exp1 $B
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True run exp2 $B
