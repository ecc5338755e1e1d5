#!/bin/bash
# attention timing ablations + polynomial share (library built with -DDSS_ATTN_ABLATION, tools/README.md)
mkdir -p gpurun_out
for abl in 0 1024 256 512 768 0 1 2 4 32 39 8 16 24 64 63 127; do
  DSS_ATTN_ABL=$abl timeout 120 python tools/attn_probe.py 296 2>&1 | tail -1
done | tee gpurun_out/attn_ablations_r2.txt
