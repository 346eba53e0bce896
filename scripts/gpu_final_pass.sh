#!/bin/bash
# Round 2, final single-GPU pass: full GPU test-suite, smoke, headline bench (with the CPU reference arm inside), series lines.
tag=${1:-r02f}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log | cut -c1-300
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for spec in "n1" "atomnas_a --model atomnas_a --no-cpu-baseline" "atomnas_c_plus --model atomnas_c_plus --no-cpu-baseline" \
            "masks_random --masks random --no-cpu-baseline" "masks_searched --masks searched --no-cpu-baseline" "accum8 --accum 8 --no-cpu-baseline"; do
  set -- $spec; name=$1; shift
  echo "== bench $name: $*"
  timeout 900 python bench.py "$@" > gpurun_out/${tag}_bench_${name}.json 2> gpurun_out/${tag}_bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${name}.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches_per_step")}, "e2e", (d.get("e2e") or {}).get("value"),
          "frac", (d.get("roofline") or {}).get("frac"), "traffic", (d.get("roofline") or {}).get("traffic"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
except Exception as e:
    print("bench parse failed", e)
PY
done
