#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread > gpurun_out/pytest_gpu_final3.log 2>&1; echo "pytest exit $?"
tail -2 gpurun_out/pytest_gpu_final3.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final3.json 2>/dev/null; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_final3.json").read().strip().splitlines()[-1])
print("bench", round(d["value"]), round(d["e2e"]["value"]), d["cpu_baseline"]["value"], [(k["kernel"], round(k["total_ms"]/d["steps"],2), k.get("frac")) for k in d["kernels"][:6]])
PY
