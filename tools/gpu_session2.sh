#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -k "hft or permute or packed or fold or prune or conv3x3 or convt or linear" > gpurun_out/s2_ops.log 2>&1; echo "rc=$?" >> gpurun_out/s2_ops.log
tail -15 gpurun_out/s2_ops.log
timeout 600 python tools/op_bench.py --only hft > gpurun_out/s2_opbench_hft.txt 2>&1; tail -12 gpurun_out/s2_opbench_hft.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/s2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s2_pytest.log; tail -5 gpurun_out/s2_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out gpurun_out/s2_breakdown.csv --profile-shapes gpurun_out/s2_shapes.csv > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; cut -c1-300 gpurun_out/s2_bench.json
