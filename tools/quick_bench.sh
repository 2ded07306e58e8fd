#!/bin/bash
# per-kernel CUDA-event times + device-resident step of one workload, no CPU baseline / e2e / other configs
python bench.py --workload ${1:-dense} --no-e2e --no-cpu-baseline --no-configs --steps 30 > gpurun_out/quick_${1:-dense}.json 2>gpurun_out/quick.err || tail -5 gpurun_out/quick.err
python - <<PY
import json
j=json.load(open("gpurun_out/quick_${1:-dense}.json"))
print("${1:-dense}: value",round(j["value"]),"ms",round(j["ms_per_step"],4),"two",round(j["two_caller_streams"]["value"]))
print("   "+"  ".join(f"{n.split('_kernel')[0]}={k['ms']*1e3:.1f}" for n,k in j["kernels"].items()), " stage_sum_us=%.1f"%(j["stage_roofline"]["sum_kernel_ms"]*1e3))
PY
