#!/bin/bash
# A/B of the weight-gradient stream (ops._Wgrad): full GPU test suite with it on, then bench with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/s13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s13_pytest.log; tail -4 gpurun_out/s13_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out gpurun_out/s13_breakdown_on.csv > gpurun_out/s13_bench_on.json 2> gpurun_out/s13_bench_on.err; cut -c1-200 gpurun_out/s13_bench_on.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-wgrad-stream > gpurun_out/s13_bench_off.json 2> gpurun_out/s13_bench_off.err; cut -c1-200 gpurun_out/s13_bench_off.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/s13_bench_on2.json 2> gpurun_out/s13_bench_on2.err; cut -c1-200 gpurun_out/s13_bench_on2.json
