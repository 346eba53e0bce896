#!/bin/bash
# Round 2, call D (1 GPU): the whole GPU test suite, AtomNAS-C+ (config 4) bench line, FMA probe.
tag=${1:-r02d}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 60 scripts/probe/fma_probe > gpurun_out/${tag}_fma_probe.txt 2>&1; head -6 gpurun_out/${tag}_fma_probe.txt
echo "== pytest -m gpu (no -x)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
tail -25 gpurun_out/${tag}_pytest.log
cat gpurun_out/parity_fullsize.json 2>/dev/null
echo "== bench atomnas_c_plus"
timeout 600 python bench.py --model atomnas_c_plus --no-cpu-baseline > gpurun_out/${tag}_bench_cplus.json 2> gpurun_out/${tag}_bench_cplus.err
tail -c 500 gpurun_out/${tag}_bench_cplus.err; cut -c1-250 gpurun_out/${tag}_bench_cplus.json
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench_cplus.json').read().strip().splitlines()[-1]); print('C+ value', round(d['value'],1), 'ms', round(d['ms_per_step'],2)); [print(k) for k in d['kernel_classes']]" 2>&1 | tail -8
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
