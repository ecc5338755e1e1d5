#!/bin/bash
# closing run of round 2: full GPU suite + the bench lines that go to profiles/ (after the LayerNorm accounting fix)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread -s > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?"
tail -2 gpurun_out/pytest_gpu_final.log | cut -c1-200; grep -n "extract_all:" gpurun_out/pytest_gpu_final.log | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
DSS_VIT_FUSED_LN=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_unfused_final.json 2>/dev/null; echo "bench unfused exit $?"
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/bench_c4_final.json 2>/dev/null; echo "bench c4 exit $?"
python - <<PY
import json
for f in ("bench_final","bench_unfused_final","bench_c4_final"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/d["steps"],2), k.get("frac")) for k in d["kernels"][:14]])
PY
