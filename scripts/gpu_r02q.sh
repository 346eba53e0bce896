#!/bin/bash
tag=${1:-r02q}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for e in 8 15; do
echo "== trace exp=$e k7"
AN_DWM_EXP=$e AN_DWM_WS=0 AN_DW_MMA=7 timeout 120 python scripts/trace_dwm.py 7 2>&1 | tail -64 > gpurun_out/${tag}_trace_k7_exp$e.txt
cat gpurun_out/${tag}_trace_k7_exp$e.txt | head -70
done
