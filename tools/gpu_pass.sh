#!/bin/bash
# One GPU-box pass of round 2: parity tests (all, no -x), measured parity figures, smoke, benches.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=12 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/measure_parity.py gpurun_out/parity_measured.json > gpurun_out/measure_parity.log 2>&1; echo "measure exit $?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
for wl in "$@"; do
  timeout 900 python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"
  tail -c 400 gpurun_out/bench_$wl.json
done
