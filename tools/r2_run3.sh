#!/bin/bash
# Round 2, GPU call 3: tests, ncu of affinity / fused-LN GEMM / eigsh (demangled names), LN fusion A/B, bench.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
prof() {  # name, kernel regex (demangled), skip
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/ncu_target.py 296 2 296 > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?" >> gpurun_out/summary.txt
  python tools/ncu_summary.py report gpurun_out/prof_$1.ncu-rep > gpurun_out/ncu_$1.txt 2>&1
}
prof affinity 'gemm_f16_tcgen05_kernel<\(int\)100' 1
prof eigsh 'lanczos_laplacian_kernel' 1
prof gemm_ln_fc1 'gemm_ln_f16_tcgen05_kernel<\(bool\)1' 12
prof gemm_ln_qkv 'gemm_ln_f16_tcgen05_kernel<\(bool\)0' 12
prof gemm_fc2 'gemm_f16_tcgen05_kernel<\(int\)2, \(int\)192, \(int\)0, \(int\)2' 12
DSS_VIT_FUSED_LN=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench_unfused exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --no-cpu-baseline --images-per-step 294 > gpurun_out/bench_fused294.json 2> gpurun_out/bench_fused294.err; echo "bench_fused294 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -12 gpurun_out/pytest_gpu.log; grep -n "Saved eigs\|extract_all:" gpurun_out/pytest_gpu.log | tail -4
for f in bench_unfused bench_fused294 bench; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/4,2), k.get("frac")) for k in d["kernels"][:11]])
    if "parity" in d: print(d["parity"])
except Exception as e: print("$f", e)
PY
done
for f in affinity eigsh gemm_ln_fc1 gemm_ln_qkv gemm_fc2; do echo "## $f"; head -20 gpurun_out/ncu_$f.txt; done
