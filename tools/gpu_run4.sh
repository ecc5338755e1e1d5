#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -m gpu -q -s --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1; echo "perf_probe exit $?" >> gpurun_out/summary.txt
timeout 200 python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1; echo "gemm_probe exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3; cat gpurun_out/perf_probe.log; cat gpurun_out/gemm_probe.log; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print('value',d['value'],'e2e',d['e2e']['value'])
    for k in d['kernels']: print(k)
except Exception as e: print('bench parse failed', e)
PY
