#!/bin/bash
# Round 2, call G (1 GPU): ncu launch list with DRAM bytes of one eager step at 256 images (roofline.traffic),
# ncu --set full of the dominant depthwise shapes, SASS evidence.
tag=${1:-r02g}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out /tmp/prof
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python scripts/profile_step.py 256 2 > gpurun_out/${tag}_ncu_launch.log 2>&1
tail -2 gpurun_out/${tag}_ncu_launch.log
python scripts/ncu_by_class.py gpurun_out/${tag}_launches.csv gpurun_out/${tag}_dram_by_class.json atomnas_c 256 2
gzip -f gpurun_out/${tag}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dw2_|dw_dgrad2' --launch-skip 15 -c 15 \
    -f -o /tmp/prof/${tag}_dw python scripts/prof_dw.py 64 > gpurun_out/${tag}_ncu_dw.log 2>&1
tail -2 gpurun_out/${tag}_ncu_dw.log
ncu -i /tmp/prof/${tag}_dw.ncu-rep --page raw --csv > gpurun_out/${tag}_dw.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_dw.raw.csv > gpurun_out/${tag}_ncu_full_dw_summary.txt 2>&1
head -3 gpurun_out/${tag}_ncu_full_dw_summary.txt | cut -c1-300
