#!/bin/bash
# fused bridge backward (eel_add_interleave_bwd_bnsums) + BatchNorm sums from every conv data gradient: tests, then A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/s14_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s14_pytest.log; tail -15 gpurun_out/s14_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline"
$B --profile-out gpurun_out/s14_breakdown_all.csv --profile-shapes gpurun_out/s14_shapes_all.csv > gpurun_out/s14_bench_all.json 2> gpurun_out/s14_bench_all.err; cut -c1-200 gpurun_out/s14_bench_all.json
EEL_BNSUMS_WIDE=0 $B --profile-out gpurun_out/s14_breakdown_nowide.csv > gpurun_out/s14_bench_nowide.json 2> gpurun_out/s14_bench_nowide.err; cut -c1-200 gpurun_out/s14_bench_nowide.json
EEL_BNSUMS_WIDE=0 EEL_BRIDGE_BNSUMS=0 $B > gpurun_out/s14_bench_none.json 2> gpurun_out/s14_bench_none.err; cut -c1-200 gpurun_out/s14_bench_none.json
$B > gpurun_out/s14_bench_all2.json 2> gpurun_out/s14_bench_all2.err; cut -c1-200 gpurun_out/s14_bench_all2.json
