#!/bin/bash
# GPU call: eigsh pairing variant (tests + A/B) and the extract_all throughput test
mkdir -p gpurun_out
DSS_EIG_VARIANT=3 timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 -k "eig or golden or spectral or lap or cluster or config or drop_in" > gpurun_out/pytest_eig3.log 2>&1
echo "pytest eig3 rc $?"; tail -3 gpurun_out/pytest_eig3.log
for v in 0 3 0 3; do
  DSS_EIG_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_eigv$v.json 2> gpurun_out/bench_eigv$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_eigv$v.json").read().strip().splitlines()[-1])
k=[x for x in d["kernels"] if x["name"]=="eigsh"][0]
print("variant $v", round(d["value"]), round(d["e2e"]["value"]), k)
PY
done
DSS_EIG_VARIANT=3 timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4_eig3.json 2>/dev/null
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4_eig0.json 2>/dev/null
python - <<PY
import json
for v in (0,3):
    d=json.loads(open(f"gpurun_out/bench_c4_eig{v}.json").read().strip().splitlines()[-1])
    print("c4 variant", v, round(d["value"]), [ (x["name"], x["ms_per_step"]) for x in d["kernels"] if x["name"]=="eigsh"])
PY
timeout 420 python -m pytest tests/test_round2_gpu.py -q -m gpu -k "extract_all" -s --timeout 200 > gpurun_out/pytest_extract_all.log 2>&1
echo "pytest extract_all rc $?"
grep -n "extract_all:\|passed\|failed\|Timeout\|Error" gpurun_out/pytest_extract_all.log | head -20
