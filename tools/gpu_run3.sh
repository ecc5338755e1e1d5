#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -m gpu -x -q -s --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1; echo "perf_probe exit $?" >> gpurun_out/summary.txt
nproc > gpurun_out/nproc.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/perf_probe.log; head -c 3000 gpurun_out/bench.json
