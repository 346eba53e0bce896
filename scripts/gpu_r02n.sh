#!/bin/bash
# Round 2, call N (1 GPU): v2 tensor-core depthwise: timing experiments (pieces switched off, ring depth) + ncu.
tag=${1:-r02n}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out /tmp/prof
for e in 0 1 2 4 7; do
  echo "== AN_DWM_EXP=$e"
  AN_DWM_EXP=$e AN_DW_MMA=57 timeout 300 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_exp$e.json > gpurun_out/${tag}_sweep_exp$e.log 2>&1
  grep -E "'C': 144, 'H': 56, 'k': 7, 'stride': 1" gpurun_out/${tag}_sweep_exp$e.log | cut -c1-140
done
for d in 3 5; do for e in 0 7; do
  echo "== AN_DWM_DEPTH=$d AN_DWM_EXP=$e"
  AN_DWM_DEPTH=$d AN_DWM_EXP=$e AN_DW_MMA=57 timeout 300 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_d${d}_exp$e.json > gpurun_out/${tag}_sweep_d${d}_exp$e.log 2>&1
  grep -E "'C': 144, 'H': 56, 'k': 7, 'stride': 1" gpurun_out/${tag}_sweep_d${d}_exp$e.log | cut -c1-140
done; done
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dwm_' --launch-skip 9 -c 3 \
    -f -o /tmp/prof/${tag}_dwm python scripts/prof_dw.py 64 > gpurun_out/${tag}_ncu_dwm.log 2>&1
tail -2 gpurun_out/${tag}_ncu_dwm.log
ncu -i /tmp/prof/${tag}_dwm.ncu-rep --page raw --csv > gpurun_out/${tag}_dwm.raw.csv 2>/dev/null
ncu -i /tmp/prof/${tag}_dwm.ncu-rep --page source --csv > gpurun_out/${tag}_dwm.src.csv 2>/dev/null
ls -la gpurun_out | grep ${tag}_dwm
