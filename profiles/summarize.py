"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_v1.csv  > profiles/r01_v1_launches.txt
    python profiles/summarize.py full     gpurun_out/prof_v1.ncu-rep  > profiles/r01_v1_ncu_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {path}")
    print(f"# total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches; compare SHARES, not absolutes")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e6:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in FULL_METRICS if m in hdr]
    kn = hdr.index("Kernel Name")
    print(f"# ncu --set full --clock-control none: {path}")
    print("# units: " + ", ".join(f"{m}[{units[i]}]" for m, i in cols))
    for r in rows[2:]:
        print(re.sub(r"\(.*", "", r[kn])[:70])
        for m, i in cols:
            print(f"    {m:70s} {r[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
