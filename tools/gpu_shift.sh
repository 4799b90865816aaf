#!/bin/bash
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k "linear or shift or capmlp or mlp" 2>&1 | tail -3
timeout 300 python tools/op_bench.py --only shift 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -2
