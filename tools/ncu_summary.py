"""Summaries of ncu artefacts for profiles/: a launch list (gpu__time_duration per launch -> per-kernel shares)
and the headline metrics of a --set full report.

    python tools/ncu_summary.py launches gpurun_out/launches.csv [images]
    python tools/ncu_summary.py report gpurun_out/prof_x.ncu-rep
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg"]


def launches(path, images=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':48s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:48s} {v[0]:8d} {v[1]:10.1f} {v[1] / v[0]:9.1f} {v[1] / tot * 100:6.1f}%")
    print(f"{'total':48s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}" + (f"   = {tot / images:.1f} us/image over {images} images" if images else ""))
    print("(cold-cache, serialised per-launch times under ncu: compare shares, not absolutes)")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        print("=" * 100)
        print(d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k:
                    print(f"  {k:72s} {row[i]:>16s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
    else:
        report(sys.argv[2])
