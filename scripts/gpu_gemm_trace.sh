#!/bin/bash
# Per-tile clock trace of the pointwise GEMM.  Needs the diagnostic library next to the product one:
#   nvcc <flags of atomnas_b200/_build.py> -DAN_GEMM_TRACE -c atomnas_b200/csrc/pw_gemm.cu -o /tmp/t/pw_gemm.o   (other objects from atomnas_b200/build/)
#   nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o atomnas_b200/libatomnas_b200_trace.so /tmp/t/*.o
tag=${1:-r02z}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
[ -f atomnas_b200/libatomnas_b200_trace.so ] || { echo "build atomnas_b200/libatomnas_b200_trace.so first (see header)"; exit 1; }
cp atomnas_b200/libatomnas_b200.so /tmp/orig.so
cp atomnas_b200/libatomnas_b200_trace.so atomnas_b200/libatomnas_b200.so
for w in ${2:-project_dgrad expand_fwd expand_fwd_nostats project_fwd expand_dgrad}; do
  echo "== trace $w"
  timeout 120 python scripts/trace_gemm.py $w 2>&1 | tail -20 > gpurun_out/${tag}_gemm_trace_$w.txt
  sed -n '6,10p' gpurun_out/${tag}_gemm_trace_$w.txt
done
cp /tmp/orig.so atomnas_b200/libatomnas_b200.so
