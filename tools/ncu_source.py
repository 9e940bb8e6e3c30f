#!/usr/bin/env python
"""Per-source-line view of an ncu report (needs -lineinfo + --import-source on):
    python tools/ncu_source.py gpurun_out/prof_X.ncu-rep [top_n] [sort: inst|samples|noinst]
Prints, per CUDA source line: warp instructions executed, average active threads per instruction,
stall samples (total / no_instruction / wait / long_sb / math)."""
import csv, subprocess, sys

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    key = sys.argv[3] if len(sys.argv) > 3 else "inst"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fpath = ""; hdr = None; agg = []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
        if hdr is None or not r[0].strip(): continue
        g = lambda k: float(r[hdr[k]]) if r[hdr[k]] not in ("", "-") else 0.0
        i = g("Instructions Executed")
        if i <= 0: continue
        agg.append(dict(inst=i, tpi=g("Thread Instructions Executed") / i, samples=g("# Samples"), noinst=g("stall_no_inst"),
                        wait=g("stall_wait"), lsb=g("stall_long_sb"), math=g("stall_math"), ssb=g("stall_short_sb"),
                        file=fpath, line=r[0], src=r[1].strip()[:100]))
    ti = sum(a["inst"] for a in agg); ts = sum(a["samples"] for a in agg)
    tt = sum(a["inst"] * a["tpi"] for a in agg)
    print("total warp-inst %.4g  avg threads/inst %.2f  samples %d" % (ti, tt / ti, ts))
    agg.sort(key=lambda a: -a[key])
    print("%9s %5s %5s | %7s %6s %6s %6s %6s %6s | line" % ("inst", "%", "thr", "smp", "noinst", "wait", "lsb", "ssb", "math"))
    for a in agg[:top]:
        print("%9.3g %5.1f %5.1f | %7d %6d %6d %6d %6d %6d | %s:%s %s" % (a["inst"], 100 * a["inst"] / ti, a["tpi"], a["samples"], a["noinst"],
              a["wait"], a["lsb"], a["ssb"], a["math"], a["file"], a["line"], a["src"]))

if __name__ == "__main__":
    main()
