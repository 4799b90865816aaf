#!/bin/bash
# what the driver runs at round end, in the same order: smoke(), pytest -m gpu, bench.py (both arms) with default flags
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/final_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/final_pytest.log; tail -3 gpurun_out/final_pytest.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; cut -c1-300 gpurun_out/final_bench_reference.json
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), d['step_ms'], d['clocks'])
print('roofline', r['kernel'], round(r['achieved'],1), round(r['frac'],3), 'plain', r.get('plain_conv_launches'), 'bn', r.get('dgrad_with_bn_sums_launches'), 'elem', round(r['elementwise_hbm_frac'],3), 'traffic', r['traffic'], r['traffic_algorithmic'])
print('cpu', d['cpu_baseline'])
e=d['gpu_eager_baseline']; print('eager', {k:(round(v['value'],1), v['ours_over_eager']) for k,v in e.items() if isinstance(v,dict)})
print('launches', d['gpu_launches'], 'alloc', d['allocator_in_timed_region'], 'mem', d['peak_mem_gb'])
"
