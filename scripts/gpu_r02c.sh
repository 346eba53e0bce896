#!/bin/bash
# Round 2, call C (N GPUs, gpurun --gpus N): multi-rank NCCL parity test + the driver's torchrun bench command.
tag=${1:-r02c}; N=${2:-2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${tag}_gpus.txt
echo "== multirank pytest"
timeout 600 python -m pytest tests/test_multirank_gpu.py -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/${tag}_pytest_multirank.log 2>&1
tail -12 gpurun_out/${tag}_pytest_multirank.log
for n in $(seq 2 2 $N) ; do
  [ $n -eq 6 ] && continue
  echo "== torchrun bench N=$n"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
  echo "rc=$?"; tail -c 400 gpurun_out/${tag}_bench_n$n.err
  python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench_n$n.json').read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'graph', d['cuda_graph'], 'launches', d['gpu_launches_per_step'])" 2>&1 | tail -1
done
echo "== reference arm under torchrun (rank 0 only prints)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref_n2.json 2> gpurun_out/${tag}_bench_ref_n2.err
echo "rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_ref_n2.json
