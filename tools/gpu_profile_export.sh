#!/bin/bash
# ncu evidence of the export-default configuration (K=1024, 512 hard-binarised pairs, eps 0.05, NMS radius 5; batch 64):
# launch list of two steps + `--set full` of the hybrid Sinkhorn kernel with 8-bit operands and of pack_bits_kernel.
mkdir -p gpurun_out
python tools/export_once.py > gpurun_out/plain_export.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_export_b64.csv python tools/export_once.py > gpurun_out/ncu_export.log 2>&1
echo "ncu list exit $?"
python tools/ncu_list_summary.py gpurun_out/launches_r2_export_b64.csv | grep -v "at::"
ncu --set full --clock-control none --import-source on -k regex:'^(pack_bits|sinkhorn_hy)' -s 2 -c 3 -o gpurun_out/prof_export -f python tools/export_once.py > gpurun_out/ncufull_export.log 2>&1
echo "ncu full exit $?"
python tools/ncu_summary.py gpurun_out/prof_export.ncu-rep > gpurun_out/ncu_full_summary_r2_export_b64.txt
ncu -i gpurun_out/prof_export.ncu-rep --page details > gpurun_out/ncu_details_r2_export_b64.txt 2>/dev/null
rm -f gpurun_out/prof_export.ncu-rep
cat gpurun_out/ncu_full_summary_r2_export_b64.txt
