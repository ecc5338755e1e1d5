"""Decode the clock64 trace written by an ablation build (DSS_ATTN_TRACE=path): per-phase cycles of CTA 0."""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.int64)
t0 = t[t > 0].min()
names = ["S ready", "ld done", "max done", "o_full ok", "exps done", "arrived"]
for wi in range(8):
    print(f"--- softmax warp {4 + wi} (group {wi // 4}, quarter {wi % 4}), cycles relative to the first stamp")
    prev_end = None
    for m in range(0, 12):
        row = t[wi * 1024 + m * 8: wi * 1024 + m * 8 + 6]
        if row[0] == 0: continue
        rel = row - t0
        d = np.diff(rel)
        gap = "" if prev_end is None else f" gap {rel[0] - prev_end}"
        print(f"tile {m:2d}: start {rel[0]:7d}{gap}  ld {d[0]:5d} max {d[1]:5d} owait {d[2]:5d} exps {d[3]:5d} fence {d[4]:5d}  total {rel[5]-rel[0]:5d}")
        prev_end = rel[5]
for g in range(2):
    s = t[8192 + g * 2048: 8192 + g * 2048 + 4 * 12].reshape(-1, 4) - t0
    print(f"--- MMA warp, query tile g={g}: (S(n+1): s_free seen, issued | PV(n): p_full seen, issued)")
    for n, r in enumerate(s):
        print(f"  n={n:2d}  S {int(r[0]):7d} +{int(r[1]-r[0]):5d}   PV {int(r[2]):7d} +{int(r[3]-r[2]):5d}")
