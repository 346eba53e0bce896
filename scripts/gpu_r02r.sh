#!/bin/bash
tag=${1:-r02r}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== pytest dw kernels (default dispatch + forced tensor-core path)"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "dw" > gpurun_out/${tag}_pytest_dw.log 2>&1
tail -5 gpurun_out/${tag}_pytest_dw.log | cut -c1-300
echo "== trace k7"
AN_DWM_EXP=8 AN_DW_MMA=7 timeout 120 python scripts/trace_dwm.py 7 2>&1 | tail -64 > gpurun_out/${tag}_trace_k7.txt
sed -n '10,18p;42,50p' gpurun_out/${tag}_trace_k7.txt
for sel in 0 357; do
  echo "== dw sweep AN_DW_MMA=$sel"
  AN_DW_MMA=$sel timeout 300 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_dw_mma${sel}.json > gpurun_out/${tag}_sweep_dw_mma${sel}.log 2>&1
  tail -1 gpurun_out/${tag}_sweep_dw_mma${sel}.log | cut -c1-300
done
