#!/bin/bash
# Round 2 record run (1 GPU): smoke, GPU test suite, both bench arms, the other BASELINE configs, ncu launch list of the
# bench command and --set full captures of every kernel class >= 1 % of the step. Summaries go to profiles/ by hand.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/ncu_*.txt gpurun_out/prof_*.ncu-rep
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1; nproc > gpurun_out/nproc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
DSS_VIT_FUSED_LN=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench_unfused exit $?" >> gpurun_out/summary.txt
for w in c3 c4 c5; do
  timeout 1500 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench_$w exit $?" >> gpurun_out/summary.txt
done
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary.txt
python tools/ncu_summary.py launches gpurun_out/launches_bench.csv > gpurun_out/launches_bench.txt 2>&1
prof() {  # name, kernel regex (demangled), skip
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/ncu_target.py 296 2 296 > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?" >> gpurun_out/summary.txt
  python tools/ncu_summary.py report gpurun_out/prof_$1.ncu-rep > gpurun_out/ncu_$1.txt 2>&1
}
prof attention 'attention_tcgen05_kernel' 12
prof gemm_ln_fc1 'gemm_ln_f16_tcgen05_kernel<\(bool\)1' 12
prof gemm_ln_qkv 'gemm_ln_f16_tcgen05_kernel<\(bool\)0' 12
prof gemm_fc2 'gemm_f16_tcgen05_kernel<\(int\)2, \(int\)192, \(int\)0, \(int\)2' 12
prof gemm_proj 'gemm_f16_tcgen05_kernel<\(int\)2, \(int\)192, \(int\)4, \(int\)1' 12
prof gemm_patch 'gemm_f16_tcgen05_kernel<\(int\)4' 1
prof gemm_kproj 'gemm_f16_tcgen05_kernel<\(int\)5' 1
prof affinity 'gemm_f16_tcgen05_kernel<\(int\)100' 1
prof eigsh 'lanczos_laplacian_kernel' 1
prof im2col 'im2col_f16_kernel' 1
prof rownorm 'rownorm_split_kernel' 1
prof layernorm 'layernorm_f16_kernel' 1
python tools/make_traffic.py gpurun_out 296 > gpurun_out/traffic.json 2>&1
rm -f gpurun_out/prof_gemm_patch.ncu-rep gpurun_out/prof_gemm_kproj.ncu-rep gpurun_out/prof_im2col.ncu-rep gpurun_out/prof_rownorm.ncu-rep gpurun_out/prof_layernorm.ncu-rep
cat gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-200; grep -n "extract_all:\|clustered spectrum" gpurun_out/pytest_gpu.log | tail -5
for f in bench bench_unfused bench_c3 bench_c4 bench_c5 bench_ref; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", d["metric"], round(d["value"],1), d.get("e2e",{}).get("value"), [(k["kernel"], round(k["total_ms"]/d["steps"],2), k.get("frac")) for k in d.get("kernels",[])[:9]])
    if "parity" in d: print("   parity", {k:v for k,v in d["parity"].items() if k not in ("reference_singular_lu_fallback",)})
    if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("as_shipped",{}).get("value"))
except Exception as e: print("$f ERR", e)
PY
done
head -20 gpurun_out/launches_bench.txt
