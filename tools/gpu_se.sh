#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k "test_se or to_patch_bias" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py tests/test_parity_configs_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -3
B="timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline"
run() { tag=$1; shift; "$@" > gpurun_out/se_$tag.json 2> gpurun_out/se_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/se_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['ms_per_step'],2), 'e2e', round(64e3/d['e2e']['value'],2), d['step_ms'], d['clocks']['sm_mhz'], d['roofline']['elementwise_hbm_frac'], d['gpu_launches'])
"; }
run a1 $B --profile-out gpurun_out/se_breakdown.csv
run a2 $B
grep -E "^se_|^head" gpurun_out/se_breakdown.csv
