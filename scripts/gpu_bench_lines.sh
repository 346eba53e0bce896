#!/bin/bash
# Round 2, evidence call 1 (1 GPU): bench lines of the round -- headline, AtomNAS-A, AtomNAS-C+, shrink series, (S) series,
# cuDNN bar on the same GPU, CPU reference arm.
tag=${1:-r02}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # name, args...
  name=$1; shift
  echo "== bench $name: $*"
  timeout 900 python bench.py "$@" > gpurun_out/${tag}_bench_${name}.json 2> gpurun_out/${tag}_bench_${name}.err
  tail -c 200 gpurun_out/${tag}_bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${name}.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("impl", "value", "unit", "ms_per_step", "gpu_launches_per_step", "unavailable")}, "e2e", (d.get("e2e") or {}).get("value"),
          "roofline", d.get("roofline"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
    for k in d.get("kernel_classes", []): print("   ", k)
except Exception as e:
    print("bench parse failed", e)
PY
}
run n1
run atomnas_a --model atomnas_a --no-cpu-baseline
run atomnas_c_plus --model atomnas_c_plus --no-cpu-baseline
run masks_random --masks random --no-cpu-baseline
run masks_searched --masks searched --no-cpu-baseline
run accum8 --accum 8 --no-cpu-baseline
run refgpu_bf16 --impl reference-gpu --ref-mode bf16
run refgpu_fp32 --impl reference-gpu --ref-mode fp32
run reference --impl reference --steps 2 --warmup 1
