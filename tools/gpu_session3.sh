#!/bin/bash
mkdir -p gpurun_out
python tools/canny_bench.py > gpurun_out/s5_canny.txt 2>&1; cat gpurun_out/s5_canny.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s5_canny_launches.csv python tools/canny_bench.py --once > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/s5_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s5_pytest.log; tail -4 gpurun_out/s5_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out gpurun_out/s5_breakdown.csv --profile-shapes gpurun_out/s5_shapes.csv > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; cut -c1-260 gpurun_out/s5_bench.json
