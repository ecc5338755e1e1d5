#!/bin/bash
# first run of the c3 / c4 / c5 bench workloads and the reference arm
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for w in c3 c4 c5; do
  timeout 1200 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench_$w exit $?" >> gpurun_out/summary.txt
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for w in c3 c4 c5 ref; do echo "== $w"; tail -3 gpurun_out/bench_$w.err | cut -c1-300; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
    print(d["metric"], round(d["value"],1), d.get("e2e",{}).get("value"), d.get("config",{}).get("images_per_step_per_gpu"))
    print([(k["kernel"], round(k["total_ms"]/d["steps"],2), k.get("frac")) for k in d.get("kernels",[])[:9]])
    print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='sample'}) for k,v in d.get("cpu_baseline",{}).items() if k!='sample'})
    print(d.get("parity")); 
    for r in d.get("sweep",[]): print({k:v for k,v in r.items() if k!='kernels'})
except Exception as e: print("ERR", e)
PY
done
