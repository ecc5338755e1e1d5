#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 77 --print-limit 40 python -m pytest tests/test_ops_gpu.py tests/test_round2_gpu.py tests/test_spectral_gpu.py -q -m gpu -x --timeout 800 -k "(gemm and 128) or epilogues or (attention and (64 or 130 or 197)) or (fused_layernorm and (128 or 300 or 77)) or segmentations or two_cliques or affinity_symmetric" > gpurun_out/sanitize_sync2.log 2>&1
echo "synccheck rc $?"; tail -3 gpurun_out/sanitize_sync2.log | cut -c1-200
grep "=========     at " gpurun_out/sanitize_sync2.log | sort | uniq -c | head
timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 -k "attention or gemm or vit or fused" > gpurun_out/pytest_after_syncwarp.log 2>&1; echo "pytest rc $?"; tail -1 gpurun_out/pytest_after_syncwarp.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_after_syncwarp.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_after_syncwarp.json").read().strip().splitlines()[-1])
print("bench", round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/d["steps"],2), k.get("frac")) for k in d["kernels"][:7]])
PY
