#!/bin/bash
# Round 2, call L (1 GPU): TMA row-rate experiment: 64-byte rows (one column tile x 32 channels) vs 32-byte rows.
tag=${1:-r02l}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for nt in 0 1; do for e in 0 7; do
  echo "== dw sweep AN_DWM_NT=$nt AN_DWM_EXP=$e"
  AN_DWM_NT=$nt AN_DWM_EXP=$e AN_DW_MMA=357 timeout 300 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_nt${nt}_exp$e.json > gpurun_out/${tag}_sweep_nt${nt}_exp$e.log 2>&1
  tail -1 gpurun_out/${tag}_sweep_nt${nt}_exp$e.log | cut -c1-300
  grep -E "'dw_fwd', 'C': (144|480), 'H': (56|14), 'k': 7, 'stride': 1" gpurun_out/${tag}_sweep_nt${nt}_exp$e.log | cut -c1-200
done; done
