#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 -k "fused_layernorm or vit or features or forward" > gpurun_out/pytest_mma1.log 2>&1
echo "pytest rc $?"; tail -1 gpurun_out/pytest_mma1.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mma1_$i.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_mma1_$i.json").read().strip().splitlines()[-1])
print("bench", round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/d["steps"],2), k.get("frac")) for k in d["kernels"][:6]])
PY
done
