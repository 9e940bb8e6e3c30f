"""Developer diagnostic (GPU): the 11x11 FAST kernel on batches of growing size -- how much of the 200 000-window time
is the tail of the maxiter fits (tools/gpu_fit11_coop.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine
from gpu_fit11_variants import windows, run
for kind in ("isolated", "dense"):
    for n in (200000, 1000000, 3000000):
        w = windows(kind, n)
        wd = torch.from_numpy(w).cuda()
        lo, hi, lmin, lmax = engine.GAUSSFIT_DEFAULT_LIMITS
        p0 = engine.moments_batch(wd, lo, hi, lmin, lmax)
        r, ms = run(wd, p0, -1)
        print("%-8s %8d windows %9.3f ms  %.3e fits/s  mean niter %.1f" % (kind, n, ms, n / ms * 1e3, r.niter.double().mean().item()))
        del wd, p0, r
        torch.cuda.empty_cache()
