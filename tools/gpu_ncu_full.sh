#!/bin/bash
# ncu --set full of the kernels matching $1 (regex) in tools/time_stage_kernels.py; $2 = tag; $3 = launches to capture
mkdir -p gpurun_out
python tools/time_stage_kernels.py > gpurun_out/plain_$2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -c ${3:-2} -o gpurun_out/prof_$2 python tools/time_stage_kernels.py > gpurun_out/ncufull_$2.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_$2.ncu-rep --page details --csv > gpurun_out/prof_$2.details.csv 2>/dev/null
python tools/ncu_details_summary.py gpurun_out/prof_$2.details.csv
