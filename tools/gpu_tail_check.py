"""Developer diagnostic (GPU): how much of the FAST kernel's time is the tail of long fits?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fluorosequencingimageanalysis_b200 import engine, synth, _lib
for nfr in (40, 160):
    fr = synth.synth_timetrace(1, n_frames=nfr)
    frd = engine.to_device_frames(fr)
    det = engine.detect_batch(frd)
    for maxiter, park in ((200, 0), (200, 32), (60, 0), (30, 0), (15, 0)):
        o = _lib.default_opts(faithful=False, solver="fast", park_after=park, maxiter=maxiter)
        ts = []
        for rep in range(4):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fit, ints, _ = engine.fit_candidates(frd, det.cand_hw, det.cand_frame, det.total, opts=o)
            e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        nfev = ints[:, 2].double().sum().item()
        print("frames %d cands %d maxiter %3d park %2d: %.3f ms -> %.4g fits/s, %.4g passes/s (%.2f passes/fit)" % (
            nfr, det.total, maxiter, park, min(ts[1:]), det.total / (min(ts[1:]) * 1e-3), nfev / (min(ts[1:]) * 1e-3), nfev / det.total), flush=True)
