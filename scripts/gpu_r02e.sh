#!/bin/bash
# Round 2, call E (1 GPU): full test suite, dw sweep with / without the two-image tail mode, bench (C, C+).
tag=${1:-r02e}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
tail -12 gpurun_out/${tag}_pytest.log
cat gpurun_out/parity_fullsize.json 2>/dev/null | tr -d '\n ' | cut -c1-1500; echo
echo "== dw sweep"
timeout 600 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_dw.json > gpurun_out/${tag}_sweep_dw.log 2>&1
tail -1 gpurun_out/${tag}_sweep_dw.log
AN_DW_NOTAIL=1 timeout 600 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_dw_notail.json > gpurun_out/${tag}_sweep_dw_notail.log 2>&1
tail -1 gpurun_out/${tag}_sweep_dw_notail.log
echo "== bench n1"
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 300 gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
    print("cpu", d.get("cpu_baseline")); print("roofline", d.get("roofline"))
    for k in d["kernel_classes"]: print(k)
    for k in d["kernels"][:12]: print(k)
except Exception as e:
    print("bench parse failed", e)
PY
echo "== bench atomnas_c_plus"
timeout 600 python bench.py --model atomnas_c_plus --no-cpu-baseline > gpurun_out/${tag}_bench_cplus.json 2> gpurun_out/${tag}_bench_cplus.err
tail -c 300 gpurun_out/${tag}_bench_cplus.err
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench_cplus.json').read().strip().splitlines()[-1]); print('C+ value', round(d['value'],1), 'ms', round(d['ms_per_step'],2)); [print(k) for k in d['kernels'][:10]]" 2>&1 | tail -12
