#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parallel_gpu.py tests/test_golden_gpu.py -m gpu -q --tb=short -x > gpurun_out/s18_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s18_pytest.log; tail -5 gpurun_out/s18_pytest.log
B="timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-eager-baseline --clock-sampler nvml"
run() { tag=$1; shift; "$@" > gpurun_out/s18_$tag.json 2> gpurun_out/s18_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/s18_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['ms_per_step'],2), 'e2e', round(64e3/d['e2e']['value'],2), d['step_ms'], d['allocator_in_timed_region'], d['peak_mem_gb'])
"; }
run a1 $B
run a2 $B
run a3 $B
run a4 $B
run smi1 $B --clock-sampler smi
run smi2 $B --clock-sampler smi
