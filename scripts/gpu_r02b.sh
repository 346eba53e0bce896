#!/bin/bash
# Round 2, call B (1 GPU): new depthwise kernels -- parity, per-shape sweep (both occupancy builds),
# bench, ncu --set full of the dominant shapes.   bash scripts/gpu_r02b.sh r02b
tag=${1:-r02b}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out /tmp/prof
echo "== pytest kernels + engine"
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
tail -8 gpurun_out/${tag}_pytest.log
echo "== dw sweep (default occupancy)"
timeout 600 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_dw.json > gpurun_out/${tag}_sweep_dw.log 2>&1
tail -2 gpurun_out/${tag}_sweep_dw.log
echo "== dw sweep (AN_DW_OCC=0)"
AN_DW_OCC=0 timeout 600 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_dw_occ0.json > gpurun_out/${tag}_sweep_dw_occ0.log 2>&1
tail -2 gpurun_out/${tag}_sweep_dw_occ0.log
echo "== bench n1"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 600 gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
    for k in d["kernel_classes"]: print(k)
    for k in d["kernels"][:10]: print(k)
except Exception as e:
    print("bench parse failed", e)
PY
echo "== ncu full on depthwise shapes"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dw2_|dw_dgrad2' --launch-skip 15 -c 15 \
    -f -o /tmp/prof/${tag}_dw python scripts/prof_dw.py 64 > gpurun_out/${tag}_ncu_dw.log 2>&1
tail -2 gpurun_out/${tag}_ncu_dw.log
ncu -i /tmp/prof/${tag}_dw.ncu-rep --page raw --csv > gpurun_out/${tag}_dw.raw.csv 2>/dev/null
ncu -i /tmp/prof/${tag}_dw.ncu-rep --page source --csv > gpurun_out/${tag}_dw.src.csv 2>/dev/null
ls -la gpurun_out | tail -8
