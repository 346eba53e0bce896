#!/bin/bash
# Round 2, call P (1 GPU): warp-specialised tensor-core depthwise: tests, per-shape sweep A/B, engine tests, bench.
tag=${1:-r02p}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== pytest dw kernels (default dispatch + forced tensor-core path)"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "dw" > gpurun_out/${tag}_pytest_dw.log 2>&1
tail -25 gpurun_out/${tag}_pytest_dw.log | cut -c1-300
for sel in 0 357; do for ws in 1 0; do
  if [ "$sel" = "0" ] && [ "$ws" = "0" ]; then continue; fi
  echo "== dw sweep AN_DW_MMA=$sel AN_DWM_WS=$ws"
  AN_DW_MMA=$sel AN_DWM_WS=$ws timeout 300 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_dw_mma${sel}_ws$ws.json > gpurun_out/${tag}_sweep_dw_mma${sel}_ws$ws.log 2>&1
  tail -1 gpurun_out/${tag}_sweep_dw_mma${sel}_ws$ws.log | cut -c1-300
done; done
echo "== pytest kernels + engine"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
tail -12 gpurun_out/${tag}_pytest.log | cut -c1-300
echo "== bench n1"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 300 gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
    for k in d["kernel_classes"]: print(k)
    for k, v in list(d["kernel_stage_ms"].items())[:30]: print(k, v)
except Exception as e:
    print("bench parse failed", e)
PY
