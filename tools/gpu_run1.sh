#!/bin/bash
# First GPU bring-up: each test file under its own timeout so a hang cannot take the whole call down.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_ops_gpu test_spectral_gpu test_vit_gpu; do
  timeout 600 python -m pytest tests/$f.py -m gpu -x -q -s --timeout=240 --timeout-method=thread > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
timeout 600 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1
echo "perf_probe exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/test_ops_gpu.log
