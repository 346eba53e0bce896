#!/bin/bash
# Round 2, call H (1 GPU): multi-segment finalize / reduce + 4 staged GEMM tiles: tests, bench A/B, pw sweep A/B.
tag=${1:-r02h}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
tail -8 gpurun_out/${tag}_pytest.log
for nb in 4 2; do
  echo "== bench n1 AN_PW_NBUF=$nb"
  AN_PW_NBUF=$nb timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_nbuf$nb.json 2> gpurun_out/${tag}_bench_nbuf$nb.err
  tail -c 300 gpurun_out/${tag}_bench_nbuf$nb.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_nbuf$nb.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
    for k in d["kernel_classes"]: print(k)
    for k, v in list(d["kernel_stage_ms"].items())[:12]: print(k, v)
except Exception as e:
    print("bench parse failed", e)
PY
done
echo "== pw sweep nbuf 4 / 2"
AN_PW_NBUF=4 timeout 300 python scripts/bench_kernels.py pw 256 --json gpurun_out/${tag}_sweep_pw_nbuf4.json > gpurun_out/${tag}_sweep_pw_nbuf4.log 2>&1; tail -1 gpurun_out/${tag}_sweep_pw_nbuf4.log
AN_PW_NBUF=2 timeout 300 python scripts/bench_kernels.py pw 256 --json gpurun_out/${tag}_sweep_pw_nbuf2.json > gpurun_out/${tag}_sweep_pw_nbuf2.log 2>&1; tail -1 gpurun_out/${tag}_sweep_pw_nbuf2.log
