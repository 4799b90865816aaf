#!/bin/bash
# 8 GPUs of one box: the bench line at 256^2 and at 512^2 (BASELINE config 4) with the end-of-round code
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --no-cpu-baseline --no-eager-baseline "$@" > gpurun_out/r02_n8_$tag.json 2> gpurun_out/r02_n8_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02_n8_$tag.json').read().strip().splitlines()[-1])
print('$tag', 'N', d['n_gpus'], round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value'],1), d['step_ms'], d['allocator_in_timed_region'], d['peak_mem_gb'])
" || tail -5 gpurun_out/r02_n8_$tag.err; }
run 256 --size 256 --batch 64 --steps 10 --warmup 3
run 512 --size 512 --batch 64 --steps 5 --warmup 3
