#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/rNN_launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof_X.ncu-rep profiles/rNN_kernel_X.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[hdr + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except (ValueError, IndexError):
            continue
        tot[r[ki][:90]] += v
        cnt[r[ki][:90]] += 1
    s = sum(tot.values())
    with open(dst, "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        fh.write("# source: %s\n" % src)
        fh.write("%-92s %6s %14s %10s %8s\n" % ("kernel", "n", "total_us", "avg_us", "share"))
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            fh.write("%-92s %6d %14.1f %10.1f %8.4f\n" % (k, cnt[k], v / 1e3, v / 1e3 / cnt[k], v / s))
    print(open(dst).read())


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    with open(dst, "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on; source: %s\n" % src)
        for V in rows[2:]:
            fh.write("## kernel: %s  grid %s block %s\n" % (V[H.index("Kernel Name")], V[H.index("Grid Size")], V[H.index("Block Size")]))
            for k in KEEP:
                if k in H:
                    i = H.index(k)
                    fh.write("%-75s %-18s %s\n" % (k, U[i], V[i]))
            # stall reasons (warp-level, per issue)
            st = [(float(V[i]), h) for i, h in enumerate(H)
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and V[i]]
            for v, h in sorted(st, reverse=True)[:8]:
                fh.write("%-75s %-18s %.3f\n" % (h, "warps/issue", v))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
