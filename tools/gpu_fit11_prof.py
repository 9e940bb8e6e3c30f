"""Developer diagnostic (GPU): launches of one arrangement of the 11x11 FAST kernel for ncu -- either 200 000 windows or
only the fits that run to maxiter (the launch's critical path).
    python tools/gpu_fit11_prof.py [thread|hybrid8|g8|g4|...] [isolated|dense] [all|long]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine
from gpu_fit11_variants import windows, run, VARIANTS
variant = sys.argv[1] if len(sys.argv) > 1 else "g8"
kind = sys.argv[2] if len(sys.argv) > 2 else "isolated"
which = sys.argv[3] if len(sys.argv) > 3 else "all"
w = windows(kind, 200000)
lo, hi, lmin, lmax = engine.GAUSSFIT_DEFAULT_LIMITS
wd = engine._device_windows(w, torch.device("cuda"))
p0 = engine.moments_batch(wd, lo, hi, lmin, lmax)
if which == "long":
    r, _ = run(wd, p0, -1, reps=1)
    idx = np.nonzero(r.status.cpu().numpy() == 5)[0][:1184]
    wd = engine._device_windows(w[idx], torch.device("cuda"))
    p0 = p0[torch.from_numpy(idx).cuda()].contiguous()
r, ms = run(wd, p0, VARIANTS[variant], reps=2)
print(variant, kind, which, wd.shape[0], "windows", ms, "ms")
