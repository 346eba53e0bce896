#!/bin/bash
# Round 2, call A (1 GPU): full GPU test suite, FMA probe, bench (graph), reference-gpu arms,
# AtomNAS-A line, mask fixtures.  Usage under gpurun: bash scripts/gpu_r02a.sh r02a
tag=${1:-r02a}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt
echo "== fma probe"; timeout 120 scripts/probe/fma_probe | tee gpurun_out/${tag}_fma_probe.txt
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
tail -15 gpurun_out/${tag}_pytest.log
echo "== bench n1"
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 800 gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
    print("clocks", d["clocks"]); print("roofline", d.get("roofline")); print("cpu", d.get("cpu_baseline"))
    for k in d["kernel_classes"]: print(k)
    for k in d["kernels"][:12]: print(k)
except Exception as e:
    print("bench parse failed", e)
PY
for mode in bf16 fp32; do
  echo "== reference-gpu $mode"
  timeout 600 python bench.py --impl reference-gpu --ref-mode $mode --steps 5 --warmup 3 > gpurun_out/${tag}_refgpu_$mode.json 2> gpurun_out/${tag}_refgpu_$mode.err
  tail -c 600 gpurun_out/${tag}_refgpu_$mode.json; tail -c 300 gpurun_out/${tag}_refgpu_$mode.err
done
echo "== atomnas_a"
timeout 600 python bench.py --model atomnas_a --no-cpu-baseline > gpurun_out/${tag}_bench_a.json 2> gpurun_out/${tag}_bench_a.err
tail -c 300 gpurun_out/${tag}_bench_a.err; cut -c1-400 gpurun_out/${tag}_bench_a.json
for m in random searched; do
  echo "== masks $m"
  timeout 600 python bench.py --masks $m --no-cpu-baseline > gpurun_out/${tag}_bench_$m.json 2> gpurun_out/${tag}_bench_$m.err
  tail -c 300 gpurun_out/${tag}_bench_$m.err; cut -c1-300 gpurun_out/${tag}_bench_$m.json
  python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench_$m.json').read().strip().splitlines()[-1]); print(d.get('shrink'), d['ms_per_step'])" 2>&1 | tail -1
done
echo "== accum 8"
timeout 600 python bench.py --accum 8 --steps 5 --no-cpu-baseline > gpurun_out/${tag}_bench_accum8.json 2> gpurun_out/${tag}_bench_accum8.err
tail -c 300 gpurun_out/${tag}_bench_accum8.err; cut -c1-300 gpurun_out/${tag}_bench_accum8.json
