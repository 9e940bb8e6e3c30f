"""Developer diagnostic (GPU): the FAST solvers against the goldens + throughput probe."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fluorosequencingimageanalysis_b200 import engine, pflib, synth
from oracle.stability import agree
G = os.path.join(ROOT, "tests", "golden")
g = np.load(os.path.join(G, "fits5_seed0.npz")); st = np.load(os.path.join(G, "stable5_seed0.npz"))
img = synth.synth_frame(0)
out = {}
for solver in ("minpack", "fast64", "fast"):
    for faithful in ((True, False) if solver == "minpack" else (False,)):
        res = engine.find_peptides_batch(img, faithful=faithful, want_fit_img=True, solver=solver)
        P = res.fit[:, [2, 3, 0, 1, 4, 5, 6]].copy()
        P[:, 2] = res.fit[:, 0] - res.cand_hw[:, 0] + 2.5; P[:, 3] = res.fit[:, 1] - res.cand_hw[:, 1] + 2.5
        tag = solver + ("-faithful" if faithful else "")
        out[tag] = (P, res.fit, res.ints)
        line = "%-18s" % tag
        for key in ("ref", "clean"):
            ok = agree(P, g[key + "_params"]); s = st["stable_" + key]
            line += " | %s all %.4f stable %.4f" % (key, ok.mean(), ok[s].mean())
        rob = st["stable_ref"] & (g["n_qrsolv"] == 0); conv = st["stable_ref"] & (g["ref_status"] == 1)
        okr = agree(P, g["ref_params"])
        line += " | ref stable&robust %.4f (n=%d) stable&status1 %.4f (n=%d)" % (okr[rob].mean(), rob.sum(), okr[conv].mean(), conv.sum())
        chi = res.fit[:, 10]
        line += " | chi2<=ref(1+1e-6) %.4f | niter %.2f nfev %.2f | status %s" % (
            np.mean(chi <= g["ref_fnorm"] * (1 + 1e-6)), res.ints[:, 1].mean(), res.ints[:, 2].mean(),
            dict(zip(*[x.tolist() for x in np.unique(res.ints[:, 0], return_counts=True)])))
        print(line, flush=True)
# fast64 vs fast (mixed) vs numpy prototype
for a_, b_ in (("fast64", "fast"), ("fast64", "minpack")):
    A, B = out[a_][0], out[b_][0]
    m = (np.abs(A[:, :6] - B[:, :6]) / np.maximum(np.abs(A[:, :6]), 1e-300)).max(axis=1)
    print("%s vs %s: max-rel-diff pct 50/90/99 %s frac<1e-4 %.4f <1e-6 %.4f" % (a_, b_, np.percentile(m, [50, 90, 99]), np.mean(m < 1e-4), np.mean(m < 1e-6)))
print("metrics fast vs minpack-clean where params agree:",)
A, B = out["fast"], out["minpack"]
same = agree(A[0], B[0], tol=1e-8, ctol=1e-8)
print("  n same %d  r2 maxdiff %.3g rmse maxrel %.3g s_n maxrel %.3g" % (same.sum(), np.abs(A[1][same, 8] - B[1][same, 8]).max(),
      (np.abs(A[1][same, 7] - B[1][same, 7]) / B[1][same, 7]).max(), (np.abs(A[1][:, 9] - B[1][:, 9]) / np.abs(B[1][:, 9])).max()))
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "fast_gpu.npz"), **{k: v[0] for k, v in out.items()}, **{k + "_ints": v[2] for k, v in out.items()})

# throughput: 40-frame stack
fr = synth.synth_timetrace(1, n_frames=40)
frd = engine.to_device_frames(fr)
det = engine.detect_batch(frd)
print("candidates", det.total)
for solver, faithful in (("minpack", False), ("fast64", False), ("fast", False)):
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fit, ints, _ = engine.fit_candidates(frd, det.cand_hw, det.cand_frame, det.total, faithful=faithful, solver=solver)
        e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    print("%-8s faithful=%s: %.3f ms -> %.4g fits/s (niter %.2f)" % (solver, faithful, min(ts[1:]), det.total / (min(ts[1:]) * 1e-3), ints[:, 1].double().mean().item()), flush=True)
