#!/bin/bash
# launch list (per-kernel device time) of tools/time_stage_kernels.py; $1 = tag, $2 = optional arg (u8)
mkdir -p gpurun_out
python tools/time_stage_kernels.py $2 > gpurun_out/plain_$1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$1.csv python tools/time_stage_kernels.py $2 > gpurun_out/ncu_$1.log 2>&1
echo "ncu exit $?"
python tools/ncu_list_summary.py gpurun_out/launches_$1.csv
