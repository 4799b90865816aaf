#!/bin/bash
# 2 GPUs: NCCL parity of the data-parallel model (tests/test_parallel_gpu.py) + one N = 2 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -s --tb=short > gpurun_out/r02_nccl_parity_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_nccl_parity_2gpu.log; tail -8 gpurun_out/r02_nccl_parity_2gpu.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; cut -c1-400 gpurun_out/r02_bench_n2.json; tail -3 gpurun_out/r02_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['step_ms'], d['allocator_in_timed_region'])
"
