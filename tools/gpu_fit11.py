"""Developer diagnostic (GPU): throughput of the generic-window entry (fsq_gaussfit_batch) on 11x11 and 5x5
windows cut around the true spots of synthetic frames, MINPACK vs FAST."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine, synth, gaussfitter
rng = np.random.default_rng(0)
for win, n in ((11, 200000), (5, 200000)):
    h = win // 2
    # synthetic windows: Gaussian + Poisson noise (same generator family as synth_frame)
    r, c = np.indices((win, win))
    cy, cx = rng.uniform(h - 0.5, h + 0.5, n), rng.uniform(h - 0.5, h + 0.5, n)
    A, s = rng.uniform(800, 4000, n), 1.5
    clean = 400.0 + A[:, None, None] * np.exp(-((r[None] - cy[:, None, None]) ** 2 + (c[None] - cx[:, None, None]) ** 2) / (2 * s * s))
    wins = rng.poisson(clean).astype(np.float64) + rng.normal(0, 10, clean.shape)
    p0 = np.stack([np.full(n, 400.0), A, cx, cy, np.full(n, 1.5), np.full(n, 1.5), np.zeros(n)], axis=1) * rng.uniform(0.9, 1.1, (n, 7))
    lo = np.zeros((n, 7)); hi = np.tile(np.array([0, 0, 0, 0, 0, 0, 360.]), (n, 1))
    lmin = np.tile(np.array([0, 0, 0, 0, 1, 1, 1], dtype=np.uint8), (n, 1)); lmax = np.tile(np.array([0, 0, 0, 0, 0, 0, 1], dtype=np.uint8), (n, 1))
    dev = lambda a, dt: torch.as_tensor(a, dtype=dt).cuda()
    W, P0, LO, HI, LMIN, LMAX = dev(wins, torch.float64), dev(p0, torch.float64), dev(lo, torch.float64), dev(hi, torch.float64), dev(lmin, torch.uint8), dev(lmax, torch.uint8)
    for solver, m in (("minpack", 20000), ("fast", n)):
        ts = []
        for rep in range(3):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            res = engine.gaussfit_batch(W[:m], P0[:m], LO[:m], HI[:m], LMIN[:m], LMAX[:m], faithful=False, solver=solver)
            e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        print("%2dx%-2d %-8s n=%6d: %.3f ms -> %.4g fits/s (mean niter %.1f, status>0 %.4f)" % (
            win, win, solver, m, min(ts[1:]), m / (min(ts[1:]) * 1e-3), res.niter.double().mean().item(), (res.status > 0).double().mean().item()), flush=True)
