#!/usr/bin/env python
"""Static and dynamic share of the FAST LM kernel's source regions in an ncu report (needs -lineinfo, --import-source on):
    python tools/ncu_regions.py gpurun_out/prof_X.ncu-rep
Per region of the tick (bucketed by the section markers of fsq_lmwarp.cu): SASS instructions (static footprint), share of
executed warp instructions, lanes active per instruction, share of stall samples and of no_instruction samples."""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    num = lambda v: float(v) if v not in ("", "-") else 0.0
    fpath, cur, hdr = "", None, None
    st = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}; continue
        if r[0].strip():
            cur = (fpath, int(r[0])); continue
        a = st[cur]
        a[0] += 1; a[1] += num(r[hdr["Instructions Executed"]]); a[2] += num(r[hdr["Thread Instructions Executed"]])
        a[3] += num(r[hdr["# Samples"]]); a[4] += num(r[hdr["stall_no_inst"]]) if "stall_no_inst" in hdr else 0.0
    src = open(os.path.join(ROOT, "fluorosequencingimageanalysis_b200", "csrc", "fsq_lmwarp.cu")).read().split("\n")

    def find(s):
        for i, l in enumerate(src):
            if s in l:
                return i + 1
        raise KeyError(s)
    marks = [("kernel head", find("lmwarp_kernel(const WarpArgs a) {")), ("refill", find("refill idle lanes")),
             ("pass (call site, reductions)", find("pass at the trial point")), ("trial bookkeeping", find("trial bookkeeping (:1245-1335)")),
             ("park", find("park: a long fit leaves the lane")), ("new linearisation", find("new linearisation at x")),
             ("lmpar search", find("lmpar (:2077-2190), FP32")), ("bounds, next trial point", find("bounds (:1184-1231)")),
             ("results", find("results (pflib.py:461-477)"))]
    k0 = find("lmwarp_kernel(const WarpArgs a) {")

    def region(key):
        f, l = key
        if f == "fsq_chol7.cuh":
            return "Cholesky (fsq_chol7.cuh)"
        if f != "fsq_lmwarp.cu":
            return "pass (w_pass incl. exp / sincos)" if "sm_100" in f else "other (" + f + ")"
        if l < k0:
            return "pass (w_pass incl. exp / sincos)"
        name = "kernel head"
        for n, m in marks:
            if l >= m:
                name = n
        return name
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for k, a in st.items():
        g = agg[region(k)]
        for i in range(5):
            g[i] += a[i]
    ts, td, tsm, tn = (sum(a[i] for a in agg.values()) for i in (0, 1, 3, 4))
    tt = sum(a[2] for a in agg.values())
    print("SASS instructions %d (%.1f KB), executed warp instructions %.4g, lanes per instruction %.2f, samples %d" % (ts, ts * 16 / 1024.0, td, tt / td, tsm))
    for rg, a in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("%-34s static %5d (%4.1f %%)  executed %5.1f %%  lanes %5.1f  samples %5.1f %%  no_instruction samples %5.1f %%"
              % (rg, a[0], 100 * a[0] / ts, 100 * a[1] / td, a[2] / max(a[1], 1), 100 * a[3] / tsm, 100 * a[4] / max(tn, 1)))


if __name__ == "__main__":
    main()
