#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r2check_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2check_pytest.log; tail -5 gpurun_out/r2check_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline"
run() { tag=$1; shift; "$@" > gpurun_out/r2check_$tag.json 2> gpurun_out/r2check_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2check_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['ms_per_step'],2), 'e2e', round(64e3/d['e2e']['value'],2), d['step_ms'], d['allocator_in_timed_region'], d['peak_mem_gb'], d['loss'], d['clocks'], d['roofline']['elementwise_hbm_frac'])
"; }
run a1 $B --profile-out gpurun_out/r2check_breakdown.csv --profile-shapes gpurun_out/r2check_shapes.csv
run a2 $B
run off $B --no-wgrad-stream
