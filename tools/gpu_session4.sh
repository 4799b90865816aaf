#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/s12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s12_pytest.log; tail -4 gpurun_out/s12_pytest.log
for i in 1 2; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out gpurun_out/s12_breakdown_$i.csv --profile-shapes gpurun_out/s12_shapes_$i.csv > gpurun_out/s12_bench_$i.json 2> gpurun_out/s12_bench_$i.err; cut -c1-200 gpurun_out/s12_bench_$i.json; done
EEL_FUSED_MLP=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out gpurun_out/s12_breakdown_nofuse.csv > gpurun_out/s12_bench_nofuse.json 2> gpurun_out/s12_bench_nofuse.err; cut -c1-200 gpurun_out/s12_bench_nofuse.json
