#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu -x 2>&1 | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline 2>&1 | tail -3 > gpurun_out/bench_nograph.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_nograph.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k in d["kernels"][:8]: print(k["kernel"], k["ms_per_step"])
for k, v in list(d["kernel_stage_ms"].items())[:14]: print(k, v)
PY
