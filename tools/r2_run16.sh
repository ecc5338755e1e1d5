#!/bin/bash
mkdir -p gpurun_out
for cl in 2 1; do
DSS_LN_CLUSTER=$cl timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 -k "fused_layernorm or vit or forward or features" > gpurun_out/pytest_ln_cl$cl.log 2>&1
echo "pytest cl=$cl rc $?"; tail -1 gpurun_out/pytest_ln_cl$cl.log
done
for cl in 2 1 2 1; do
DSS_LN_CLUSTER=$cl timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_lncl$cl.json 2> gpurun_out/bench_lncl$cl.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_lncl$cl.json").read().strip().splitlines()[-1])
print("bench cl=$cl", round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/d["steps"],2), k["frac"]) for k in d["kernels"][:6]])
PY
done
