"""profiles/r2_traffic.json from the `ncu --set full` summaries (tools/ncu_summary.py report ...): DRAM bytes read + written
per launch and per image for every captured kernel class, next to the algorithmic bytes bench.py divides by.

    python tools/make_traffic.py gpurun_out 296 > profiles/r2_traffic.json
"""
import json
import re
import sys
from pathlib import Path

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
# ncu summary file stem -> kernel class name used by bench.py
CLASS = {"attention": "attention", "affinity": "affinity", "eigsh": "eigsh", "gemm_ln_fc1": "gemm_fc1", "gemm_ln_qkv": "gemm_qkv",
         "gemm_fc2": "gemm_fc2", "gemm_proj": "gemm_proj", "gemm_patch": "gemm_patch", "gemm_kproj": "gemm_kproj",
         "im2col": "im2col", "rownorm": "rownorm", "layernorm": "layernorm", "degree": "degree_reduce"}


def parse(path):
    out = {}
    for line in Path(path).read_text().splitlines():
        m = re.match(r"\s+(\S+)\s+([\d.,]+)\s*(\S*)", line)
        if not m:
            if line.startswith("void") or " grid " in line:
                out["kernel"] = line.strip()
            continue
        key, val, unit = m.group(1), float(m.group(2).replace(",", "")), m.group(3)
        out[key] = val * UNIT.get(unit, 1.0) if "bytes" in key else (val, unit)
    return out


def main():
    d, images = Path(sys.argv[1]), int(sys.argv[2])
    res = {"images_per_launch": images, "source": "ncu --set full --clock-control none, one launch of the 296-image step per class "
           "(tools/r2_final.sh); dram__bytes_read.sum + dram__bytes_write.sum", "per_image_bytes": {}, "per_launch": {}}
    for f in sorted(d.glob("ncu_*.txt")):
        stem = f.stem[4:]
        p = parse(f)
        if "dram__bytes_read.sum" not in p:
            continue
        tot = p["dram__bytes_read.sum"] + p["dram__bytes_write.sum"]
        name = CLASS.get(stem, stem)
        res["per_image_bytes"][name] = tot / images
        dur = p.get("gpu__time_duration.sum", (None, ""))
        res["per_launch"][name] = {"kernel": p.get("kernel", ""), "dram_read_bytes": p["dram__bytes_read.sum"],
                                   "dram_write_bytes": p["dram__bytes_write.sum"],
                                   "duration": f"{dur[0]} {dur[1]}" if dur[0] is not None else None,
                                   "tensor_pipe_pct": p.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", (None,))[0],
                                   "xu_pipe_pct": p.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", (None,))[0],
                                   "dram_pct_of_peak": p.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", (None,))[0],
                                   "issue_active_pct": p.get("smsp__issue_active.avg.pct_of_peak_sustained_active", (None,))[0]}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
