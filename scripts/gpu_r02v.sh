#!/bin/bash
# Round 2, call V (1 GPU): stem weight gradient on the tensor cores: tests + bench; reference train.py through dropin.
tag=${1:-r02v}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== pytest stem + engine + trainer"
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_trainer_gpu.py tests/test_dropin_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -k "not full_size" > gpurun_out/${tag}_pytest.log 2>&1
tail -6 gpurun_out/${tag}_pytest.log | cut -c1-300
echo "== bench n1"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 300 gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"])
    for k in d["kernel_classes"]: print(k)
    for k, v in list(d["kernel_stage_ms"].items())[:26]: print(k, v)
except Exception as e:
    print("bench parse failed", e)
PY
echo "== reference train.py through dropin/launch.py"
bash scripts/run_reference_train.sh 150 128 gpurun_out/${tag}_reference_train_py_dropin.log
