#!/bin/bash
# Round 2, evidence call 2 (1 GPU): ncu launch list with DRAM bytes of eager steps at 256 images (roofline.traffic),
# ncu --set full of the depthwise kernels, per-shape sweeps at 224 and at the config-5 resolutions.
tag=${1:-r02}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out /tmp/prof
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python scripts/profile_step.py 256 2 > gpurun_out/${tag}_ncu_launch.log 2>&1
tail -2 gpurun_out/${tag}_ncu_launch.log
python scripts/ncu_by_class.py gpurun_out/${tag}_launches.csv gpurun_out/${tag}_dram_by_class.json atomnas_c 256 2
gzip -f gpurun_out/${tag}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dw2_|dw_dgrad2|dwm_' --launch-skip 15 -c 15 \
    -f -o /tmp/prof/${tag}_dw python scripts/prof_dw.py 64 > gpurun_out/${tag}_ncu_dw.log 2>&1
tail -2 gpurun_out/${tag}_ncu_dw.log
ncu -i /tmp/prof/${tag}_dw.ncu-rep --page raw --csv > gpurun_out/${tag}_dw.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_dw.raw.csv > gpurun_out/${tag}_ncu_full_dw_summary.txt 2>&1
head -4 gpurun_out/${tag}_ncu_full_dw_summary.txt | cut -c1-300
for what in dw pw; do
  timeout 300 python scripts/bench_kernels.py $what 256 --json gpurun_out/${tag}_sweep_${what}_224.json > gpurun_out/${tag}_sweep_${what}_224.log 2>&1
  tail -1 gpurun_out/${tag}_sweep_${what}_224.log | cut -c1-300
done
for res in 128 160 192; do
  timeout 300 python scripts/bench_kernels.py dw 256 --res $res --json gpurun_out/${tag}_sweep_dw_$res.json > gpurun_out/${tag}_sweep_dw_$res.log 2>&1
  tail -1 gpurun_out/${tag}_sweep_dw_$res.log | cut -c1-300
done
AN_DW_MMA=0 timeout 300 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_dw_224_ffma_only.json > gpurun_out/${tag}_sweep_dw_224_ffma_only.log 2>&1
tail -1 gpurun_out/${tag}_sweep_dw_224_ffma_only.log | cut -c1-300
AN_DW_MMA=357 timeout 300 python scripts/bench_kernels.py dw 256 --json gpurun_out/${tag}_sweep_dw_224_mma_forced.json > gpurun_out/${tag}_sweep_dw_224_mma_forced.log 2>&1
tail -1 gpurun_out/${tag}_sweep_dw_224_mma_forced.log | cut -c1-300
