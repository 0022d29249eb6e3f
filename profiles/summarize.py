"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

  python profiles/summarize.py launches gpurun_out/launches_X.csv  > profiles/launches_X.txt
  python profiles/summarize.py raw gpurun_out/X.ncu-rep            > profiles/X.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sectors_op_red.sum", "l1tex__t_bytes.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        name = row["Kernel Name"].split("(")[0][-70:]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"# {path}: per-kernel device time (ncu gpu__time_duration.sum, cold-cache, serialised: compare SHARES)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} us {c:5d} x {t / c:9.1f} us/launch {100 * t / tot:5.1f}%  {k}")
    print(f"{tot:10.1f} us total")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: ncu --set full, selected metrics per captured launch")
    for r in rows[2:]:
        print("-----", r[idx["Kernel Name"]][:100])
        for k in KEYS:
            if k in idx:
                print(f"  {k:70s} {r[idx[k]]:>18s} {units[idx[k]]}")
        stalls = sorted(((float(r[i].replace(',', '')), h) for h, i in idx.items() if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct") and r[i]), reverse=True)[:6]
        for v, h in stalls:
            print(f"  stall {h:64s} {v:18.2f} %")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
