#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "bench_fused exit $?" >> gpurun_out/summary.txt
DSS_VIT_FUSED_LN=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench_unfused exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_fused2.json 2> gpurun_out/bench_fused2.err
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300; grep -n "Saved eigs to.*2048\|extract_all:" gpurun_out/pytest_gpu.log | tail -3 | cut -c1-420
for f in bench_fused bench_unfused bench_fused2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/4,2), k.get("bound","")[:1], k.get("frac")) for k in d["kernels"][:11]])
except Exception as e: print("$f", e)
PY
done
