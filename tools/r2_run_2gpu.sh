#!/bin/bash
# 2-GPU check of the sharded path (one rank per GPU, NCCL for the weight broadcast / barrier / max-over-ranks only)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "2gpu c2 rc $?"; tail -c 600 gpurun_out/bench_2gpu.json | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_c4.json 2> gpurun_out/bench_2gpu_c4.err
echo "2gpu c4 rc $?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_samebox.json 2>/dev/null
python - <<PY
import json
for f in ("bench_1gpu_samebox","bench_2gpu","bench_2gpu_c4"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d["scaling"], d["ms_per_step"])
PY
