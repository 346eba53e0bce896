#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s 2>&1 | grep -E "passed|failed|FAILED|Error|assert |ReLU net" | tail -40 > gpurun_out/engine_test.log
cat gpurun_out/engine_test.log
