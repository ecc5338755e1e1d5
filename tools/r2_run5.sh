#!/bin/bash
# attention A/B (item rotation, lazy reference maximum), each in its own process on the same box
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in 0 1 2 3 0 3; do DSS_ATTN_FLAGS=$f timeout 120 python tools/attn_probe.py 296 2>&1 | sed "s/^/flags=$f /" >> gpurun_out/attn_ab.log; done
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_round2_gpu.py -m gpu -q -p no:cacheprovider -k "attention" --timeout=300 > gpurun_out/pytest_attn.log 2>&1; echo "pytest_attn exit $?" >> gpurun_out/summary.txt
DSS_ATTN_FLAGS=0 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_round2_gpu.py -m gpu -q -p no:cacheprovider -k "attention" --timeout=300 > gpurun_out/pytest_attn0.log 2>&1; echo "pytest_attn0 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err
cat gpurun_out/summary.txt gpurun_out/attn_ab.log; tail -2 gpurun_out/pytest_attn.log gpurun_out/pytest_attn0.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_fused.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/4,2), k.get("frac")) for k in d["kernels"][:11]])
PY
