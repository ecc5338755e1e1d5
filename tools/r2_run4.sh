#!/bin/bash
# Round 2, GPU call: tests + A/B benches (fused LN on/off, eigensolver variant) + ncu of attention / affinity / eigsh.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "bench_fused exit $?" >> gpurun_out/summary.txt
DSS_VIT_FUSED_LN=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench_unfused exit $?" >> gpurun_out/summary.txt
DSS_EIG_VARIANT=2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_eig41.json 2> gpurun_out/bench_eig41.err; echo "bench_eig41 exit $?" >> gpurun_out/summary.txt
prof() {  # name, kernel regex (demangled), skip
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/ncu_target.py 296 2 296 > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?" >> gpurun_out/summary.txt
  python tools/ncu_summary.py report gpurun_out/prof_$1.ncu-rep > gpurun_out/ncu_$1.txt 2>&1
}
prof attention 'attention_tcgen05_kernel' 12
prof affinity 'gemm_f16_tcgen05_kernel<\(int\)100' 1
prof eigsh 'lanczos_laplacian_kernel' 1
prof gemm_ln_fc1 'gemm_ln_f16_tcgen05_kernel<\(bool\)1' 12
cat gpurun_out/summary.txt; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300; grep -n "extract_all:" gpurun_out/pytest_gpu.log | tail -2
for f in bench_fused bench_unfused bench_eig41; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["e2e"]["value"]), [(k["kernel"], round(k["total_ms"]/4,2), k.get("frac")) for k in d["kernels"][:11]])
except Exception as e: print("$f", e)
PY
done
for f in attention affinity eigsh gemm_ln_fc1; do echo "## $f"; grep -E "Kernel|void|duration|tensor|pipe_xu|issue_active|dram_throughput" gpurun_out/ncu_$f.txt | cut -c1-200; done
