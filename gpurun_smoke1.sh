#!/bin/bash
# first GPU contact: unit tests of the tcgen05 kernels under a hard timeout
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_pw_gemm_gpu.py -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pw_gemm_test.log
