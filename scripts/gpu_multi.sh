#!/bin/bash
# Round 2, multi-GPU pass: bench.py under torchrun exactly as the driver launches it, both arms.   bash scripts/gpu_r02x.sh N tag
N=${1:-2}; tag=${2:-r02f}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${tag}_gpus_n$N.txt
echo "== bench b200 n=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N \
    > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "rc=$?"; tail -c 300 gpurun_out/${tag}_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n$N.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "scaling", "gpu_launches_per_step")}, "e2e", (d.get("e2e") or {}).get("value"), d.get("clocks"))
except Exception as e:
    print("bench parse failed", e)
PY
echo "== bench reference arm n=$N (rank 0 only, bounded)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --impl reference --steps 2 --warmup 1 \
    > gpurun_out/${tag}_bench_ref_n$N.json 2> gpurun_out/${tag}_bench_ref_n$N.err
echo "rc=$?"; tail -c 200 gpurun_out/${tag}_bench_ref_n$N.json
