"""Developer diagnostic (GPU): what the long 11x11 fits cost -- pass counts (nfev) of the fits that run to maxiter, and the
time per tick of each kernel arrangement when ONLY such fits run (the critical path of a 200 000-window launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine
from gpu_fit11_variants import windows, run, VARIANTS

for kind in ("isolated", "dense"):
    w = windows(kind, 200000)
    lo, hi, lmin, lmax = engine.GAUSSFIT_DEFAULT_LIMITS
    wd = engine._device_windows(w.astype("uint16"), torch.device("cuda"))
    p0 = engine.moments_batch(wd, lo, hi, lmin, lmax)
    r, ms = run(wd, p0, -1, reps=1)
    st, nfev, niter = r.status.cpu().numpy(), r.nfev.cpu().numpy(), r.niter.cpu().numpy()
    long_ = st == 5
    print("%s: %d fits at maxiter; their nfev: min %d median %d p90 %d max %d; all fits: nfev mean %.1f p95 %d p99 %d; share of passes in fits with nfev > 32: %.3f"
          % (kind, long_.sum(), nfev[long_].min(), np.median(nfev[long_]), np.percentile(nfev[long_], 90), nfev[long_].max(),
             nfev.mean(), np.percentile(nfev, 95), np.percentile(nfev, 99), nfev[nfev > 32].sum() / nfev.sum()), flush=True)
    idx = np.nonzero(long_)[0]
    for nsel in (len(idx), 1184, 148):
        sel = torch.from_numpy(idx[:nsel]).cuda()
        ws, ps = engine._device_windows(w[idx[:nsel]].astype("uint16"), torch.device("cuda")), p0[sel].contiguous()
        tick = float(nfev[idx[:nsel]].max())
        for name, wps in VARIANTS.items():
            if name.startswith("hybrid"):
                continue
            r2, ms2 = run(ws, ps, wps, reps=2)
            print("   %5d long fits only, %-7s %8.3f ms = %.2f us per tick of the longest fit (%d passes)" % (nsel, name, ms2, ms2 * 1e3 / tick, tick), flush=True)
