"""Developer diagnostic (GPU): one launch of the 11x11 FAST kernel on 200 000 windows, for ncu.
    python tools/gpu_fit11_prof.py [coop|thread] [isolated|dense]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine, _lib
from gpu_fit11_coop import windows, run  # noqa  (runs its own comparison first when imported as a script: guarded below)
variant = sys.argv[1] if len(sys.argv) > 1 else "coop"
kind = sys.argv[2] if len(sys.argv) > 2 else "isolated"
w = windows(kind, 200000)
wd = torch.from_numpy(w).cuda()
lo, hi, lmin, lmax = engine.GAUSSFIT_DEFAULT_LIMITS
p0 = engine.moments_batch(wd, lo, hi, lmin, lmax)
r, ms = run(wd, p0, variant)
print(variant, kind, ms, "ms")
