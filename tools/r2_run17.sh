#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 -k "fused_layernorm or vit or forward or features" > gpurun_out/pytest_ln_direct.log 2>&1
echo "pytest direct rc $?"; tail -1 gpurun_out/pytest_ln_direct.log
for dv in 1 0 1 0; do
DSS_LN_DIRECT=$dv timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_lndirect$dv.json 2> gpurun_out/bench_lndirect$dv.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_lndirect$dv.json").read().strip().splitlines()[-1])
print("bench direct=$dv", round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/d["steps"],2), k["frac"]) for k in d["kernels"][:6]])
PY
done
