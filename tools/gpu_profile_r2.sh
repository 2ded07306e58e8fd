#!/bin/bash
# Round-2 ncu evidence of one matcher step (batch 64 pairs, 480x640, K=512): launch list + `--set full` of every kernel of a step.
#   $1 = workload (dense|sparse|angle)   $2 = tag   $3 = launches to skip   $4 = launches to capture
# The report is summarised ON THE BOX (raw page -> one line per launch, details page -> text) and then deleted: two full
# reports exceed what gpurun copies back.
mkdir -p gpurun_out
WL=${1:-dense}; TAG=${2:-r2_$WL}
python tools/profile_step.py $WL 64 3 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py $WL 64 3 > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu list exit $?"
python tools/ncu_list_summary.py gpurun_out/launches_$TAG.csv
python tools/profile_step.py $WL 64 2 > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'^(band_|integral_|prefix_|score|nms|topk|dense_at|sparse_|pack_|sinkhorn|cost_|xd_)' -s ${3:-13} -c ${4:-14} -o gpurun_out/prof_$TAG python tools/profile_step.py $WL 64 2 > gpurun_out/ncufull_$TAG.log 2>&1
echo "ncu full exit $?"
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep > gpurun_out/ncu_full_summary_$TAG.txt
ncu -i gpurun_out/prof_$TAG.ncu-rep --page details > gpurun_out/ncu_details_$TAG.txt 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$TAG.csv 2>/dev/null
rm -f gpurun_out/prof_$TAG.ncu-rep
cat gpurun_out/ncu_full_summary_$TAG.txt
