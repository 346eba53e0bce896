#!/bin/bash
# The reference's own train.py (unmodified, staged under oracle/_ref/AtomNAS) on the B200 engine through
# dropin/launch.py: apps/slimming/shrink/atomnas_c.yml, synthetic ImageNet (torchvision FakeData, the reference's
# `imagenet1k_fake`), one process, stopped by `timeout` after SECONDS of training.  Evidence run, not a benchmark:
# the step time is bounded by FakeData + PIL transforms in the loader workers.
#   bash scripts/run_reference_train.sh [seconds] [per_gpu_batch] [out.log]
secs=${1:-150}
bs=${2:-128}
out=${3:-gpurun_out/r02_reference_train_py_dropin.log}
root="${GRAFT_REPO_ROOT:-$(pwd)}"
cd "$root/oracle/_ref/AtomNAS" || { echo "oracle/_ref/AtomNAS missing (python oracle/make_ref.py)"; exit 1; }
export ARNOLD_OUTPUT=/tmp/atomnas_out DATA_ROOT=/tmp/none PYTHONPATH="$root/oracle/_ref/stubs"
export RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29533
mkdir -p "$(dirname "$root/$out")"
timeout "$secs" python "$root/dropin/launch.py" train.py app:apps/slimming/shrink/atomnas_c.yml \
    --dataset imagenet1k_fake --per_gpu_batch_size "$bs" --data_loader_workers 32 --log_interval 10 \
    --bn_calibration_per_gpu_batch_size "$bs" > "$root/$out" 2>&1
echo "exit code $? (124 = stopped by timeout as planned)" >> "$root/$out"
grep -E "Epoch|loss|images|Train|params|MAC|FLOP|Error|error" "$root/$out" | tail -25
