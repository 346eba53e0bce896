#!/bin/bash
# One measurement round on the GPU box: bench (both arms), ncu launch list, ncu --set full of the
# conv kernels.  Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh r01 [full]
# Outputs land in gpurun_out/<tag>_*; scripts/ncu_summary.py turns the raw csv into profiles/.
tag=${1:-r01}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out /tmp/prof
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 1500 gpurun_out/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
print("clocks", d["clocks"]); print("roofline", d.get("roofline")); print("cpu", d.get("cpu_baseline"))
for k in d["kernels"][:12]: print(k)
for k, v in list(d["kernel_stage_ms"].items())[:20]: print(k, v)
r = json.loads(open("gpurun_out/${tag}_bench_ref.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"])
PY
# launch list of one eager step (same network / batch as the bench), one metric, no clock control
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python scripts/profile_step.py 256 2 > gpurun_out/${tag}_ncu_launch.log 2>&1
tail -2 gpurun_out/${tag}_ncu_launch.log
if [ "$2" == "full" ]; then
  # the high-resolution blocks dominate the step: the first conv launches of the forward pass and
  # the last ones of the backward pass.  T = conv launches per step, from the launch list.
  RE='pw_gemm_kernel|pw_wgrad_kernel|dw_conv_stream|dw_dgrad2_stream|dw_wgrad_stream|affine_act'
  T=$(python - <<PY
import csv, re
rows = list(csv.reader(l for l in open("gpurun_out/${tag}_launches.csv") if l.startswith('"')))
k = rows[0].index("Kernel Name")
print(sum(1 for r in rows[1:] if re.search("$RE", r[k])) // 2)
PY
)
  echo "conv launches per step: $T"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -c 56 \
      -f -o /tmp/prof/${tag}_fwd python scripts/profile_step.py 64 1 > gpurun_out/${tag}_ncu_full.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" --launch-skip $((T-100)) -c 100 \
      -f -o /tmp/prof/${tag}_bwd python scripts/profile_step.py 64 1 >> gpurun_out/${tag}_ncu_full.log 2>&1
  tail -3 gpurun_out/${tag}_ncu_full.log
  for w in fwd bwd; do
    ncu -i /tmp/prof/${tag}_$w.ncu-rep --page raw --csv > gpurun_out/${tag}_full_$w.raw.csv 2>/dev/null
  done
  ls -la /tmp/prof gpurun_out | tail -12
fi
