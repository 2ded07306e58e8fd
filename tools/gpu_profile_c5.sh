#!/bin/bash
# ncu evidence of BASELINE config 5 (1080x1920, K=2048, batch 8): launch list of two steps + `--set full` of the streaming
# Sinkhorn kernels (sinkhorn_xl.cu), summarised on the box.
mkdir -p gpurun_out
python tools/profile_config5.py 8 > gpurun_out/plain_c5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_config5_b8.csv python tools/profile_config5.py 8 > gpurun_out/ncu_c5.log 2>&1
echo "ncu list exit $?"
python tools/ncu_list_summary.py gpurun_out/launches_r2_config5_b8.csv
ncu --set full --clock-control none --import-source on -k regex:'^(cost_pk|xs_sweep|xs_col|xs_finalize|pack_f16|topk)' -s 12 -c 10 -o gpurun_out/prof_c5 -f python tools/profile_config5.py 8 > gpurun_out/ncufull_c5.log 2>&1
echo "ncu full exit $?"
python tools/ncu_summary.py gpurun_out/prof_c5.ncu-rep > gpurun_out/ncu_full_summary_r2_config5_b8.txt
ncu -i gpurun_out/prof_c5.ncu-rep --page details > gpurun_out/ncu_details_r2_config5_b8.txt 2>/dev/null
rm -f gpurun_out/prof_c5.ncu-rep
cat gpurun_out/ncu_full_summary_r2_config5_b8.txt
