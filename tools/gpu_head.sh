#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k "head" 2>&1 | tail -3
timeout 300 python tools/op_bench.py --only head 2>&1 | tail -3
