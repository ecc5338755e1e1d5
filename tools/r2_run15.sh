#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 -k "fused_layernorm or vit or forward or features" > gpurun_out/pytest_ln.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_ln.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_epi2_$i.json 2> gpurun_out/bench_epi2_$i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_epi2_$i.json").read().strip().splitlines()[-1])
print("bench", round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/d["steps"],2), k["frac"]) for k in d["kernels"][:8]])
PY
done
