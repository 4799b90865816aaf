#!/bin/bash
# round-2 first GPU session: all GPU tests, parity at the benchmarked shapes (verbose), bench with the eager baseline, bf16 conditioning curve
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s1_smoke.log
timeout 900 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_parity_configs_gpu.py > gpurun_out/s1_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s1_pytest.log
timeout 1200 python -m pytest tests/test_parity_configs_gpu.py -m gpu -q -s --tb=short > gpurun_out/s1_parity.log 2>&1; echo "rc=$?" >> gpurun_out/s1_parity.log
timeout 600 python tools/bf16_conditioning.py --autocast --steps 0,5,20,60,120,250 > gpurun_out/s1_cond.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/s1_breakdown.csv --profile-shapes gpurun_out/s1_shapes.csv > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -3 gpurun_out/s1_smoke.log; tail -5 gpurun_out/s1_pytest.log; tail -5 gpurun_out/s1_parity.log; cat gpurun_out/s1_cond.log | tail -8; cut -c1-600 gpurun_out/s1_bench.json
