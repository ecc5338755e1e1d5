#!/bin/bash
# fused-LN GEMM: 256-column tiles for fc1 (N = 1536), two accumulator stages, three 32 KB weight stages
mkdir -p gpurun_out
DSS_LN_BN=256 timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 -k "fused_layernorm or vit or features or forward" > gpurun_out/pytest_bn256.log 2>&1
echo "pytest bn256 rc $?"; tail -1 gpurun_out/pytest_bn256.log
for bn in 256 192 256 192; do
DSS_LN_BN=$bn timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lnbn$bn.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_lnbn$bn.json").read().strip().splitlines()[-1])
print("bench BN=$bn", round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/d["steps"],2), k.get("frac")) for k in d["kernels"][:6]])
PY
done
