#!/bin/bash
# GPU unit tests under a hard timeout
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -60 > gpurun_out/kernels_test.log
tail -30 gpurun_out/kernels_test.log
