#!/bin/bash
# what the driver runs at round end, in the same order: smoke(), pytest -m gpu, bench.py (both arms) with default flags;
# then the end-of-round kernel breakdown and (argument "ncu") the ncu capture of the dominant family for roofline.traffic
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/final_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/final_pytest.log; tail -3 gpurun_out/final_pytest.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; cut -c1-200 gpurun_out/final_bench_reference.json
timeout 900 python bench.py --profile-out gpurun_out/r02_kernel_breakdown_b64_256_bf16.csv --profile-shapes gpurun_out/r02_kernel_shapes_b64_256_bf16.csv > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), d['step_ms'], d['clocks'])
print('roofline', r['kernel'], round(r['achieved'],1), round(r['frac'],3), 'plain', r.get('plain_conv_launches'), 'bn', r.get('dgrad_with_bn_sums_launches'), 'elem', round(r['elementwise_hbm_frac'],3), 'traffic', r['traffic'], r['traffic_algorithmic'])
print('cpu', d['cpu_baseline'])
e=d['gpu_eager_baseline']; print('eager', {k:(round(v['value'],1), v['ours_over_eager']) for k,v in e.items() if isinstance(v,dict)})
print('launches', d['gpu_launches'], 'alloc', d['allocator_in_timed_region'], 'mem', d['peak_mem_gb'])
"
if [ "$1" = "ncu" ]; then
  mkdir -p /tmp/ncu
  BENCH="python bench.py --no-cpu-baseline --no-eager-baseline --clock-sampler none"
  timeout 900 ncu --clock-control none --set full -k regex:"tc_conv_kernel" -c 36 -o /tmp/ncu/r02_tc_conv_ncu_full_b64 -f $BENCH --steps 1 --warmup 0 > /dev/null 2>&1
  ncu -i /tmp/ncu/r02_tc_conv_ncu_full_b64.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02_tc_conv_ncu_full_b64.csv; wc -l gpurun_out/r02_tc_conv_ncu_full_b64.csv
  python tools/ncu_aggregate.py traffic gpurun_out/r02_tc_conv_ncu_full_b64.csv tc_conv3x3 > gpurun_out/r02_traffic.json; cat gpurun_out/r02_traffic.json
fi
