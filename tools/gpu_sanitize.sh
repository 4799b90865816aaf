#!/bin/bash
# compute-sanitizer passes (SURVEY.md section 5 "race detection"): memcheck + racecheck over the kernels with hand-rolled
# synchronisation (tcgen05/TMA/mbarrier pipelines, swizzled smem transposes, split-K red.global, lock-free union-find)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
K='conv3x3 or hft or wgrad or canny or linear or convt'
for tool in memcheck racecheck; do
  timeout 1500 $SAN --tool $tool --print-limit 20 --error-exitcode 3 \
    python -m pytest tests/test_ops_gpu.py tests/test_edges_gpu.py -m gpu -q -x -k "$K" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit=$?" >> gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
