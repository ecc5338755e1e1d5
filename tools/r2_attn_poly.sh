#!/bin/bash
# polynomial share of the exponentials in the attention kernel (library built with -DDSS_ATTN_ABLATION)
mkdir -p gpurun_out
for rep in 1 2; do
for abl in 0 1024 256 512 768; do
  DSS_ATTN_ABL=$abl timeout 120 python tools/attn_probe.py 296 2>&1 | tail -1
done
done | tee gpurun_out/attn_poly_share.txt
