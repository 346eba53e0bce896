#!/bin/bash
# Round 2, call I (1 GPU): HMMA issue-rate probe; nbuf heuristic fix + unrolled partial reductions: quick tests + bench.
tag=${1:-r02i}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== hmma probe"
(cd scripts/probe && timeout 120 ./hmma_probe) > gpurun_out/${tag}_hmma_probe.txt 2>&1; cat gpurun_out/${tag}_hmma_probe.txt
echo "== pytest kernels + engine"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
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
