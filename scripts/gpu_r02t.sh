#!/bin/bash
# Round 2, call T (1 GPU): state check after the tensor-core depthwise work: full GPU test-suite, smoke, bench.
tag=${1:-r02t}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
tail -8 gpurun_out/${tag}_pytest.log | cut -c1-300
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== dw sweep default dispatch"
timeout 300 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_dw.json > gpurun_out/${tag}_sweep_dw.log 2>&1; tail -1 gpurun_out/${tag}_sweep_dw.log | cut -c1-300
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
