"""Developer diagnostic (GPU): the arrangements of the 11x11 FAST kernel (fsq_lm_opts.warps_per_sm: 0 = thread-per-window
bulk + lane-group finish of the parked fits (8 or 4 lanes per window by their number), -1 thread per window only, -2 / -3 / -4 =
4 / 8 / 2 lanes per window for every fit, -5 / -6 = bulk + 4- / 8-lane finish) on isolated-spot windows (configs[0]) and on windows cut from the dense 2048x2048 frame
(configs[3]), with float64 and uint16 window data: time, and results against the thread-per-window kernel.
    python tools/gpu_fit11_variants.py [n_windows] [park_after ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine, synth, _lib
from test_gpu_fit import agree

VARIANTS = {"thread": -1, "hybrid": 0, "hybrid8": -6, "hybrid4": -5, "g4": -2, "g8": -3, "g2": -4}


def windows(kind, n):
    if kind == "isolated":
        out = []
        seed = 100
        while sum(len(o) for o in out) < n and seed <= 110:
            img, cr, cc, amp = synth.synth_frame_with_truth(seed)
            out.append(synth.cut_windows(img, cr, cc, 11))
            seed += 1
        w = np.concatenate(out)
    else:
        img, cr, cc, amp = synth.synth_frame_with_truth(40, H=2048, W=2048, n_spots=20000)
        w = synth.cut_windows(img, cr, cc, 11)
    reps = -(-n // len(w))
    return np.concatenate([w] * reps)[:n]


def run(wd, p0, wps, park=0, reps=3):
    n = wd.shape[0]
    lo, hi, lmin, lmax = engine.GAUSSFIT_DEFAULT_LIMITS
    dev = wd.device
    ex = lambda v, dt: torch.as_tensor(v, dtype=dt).to(dev).expand(n, 7).contiguous()
    o = _lib.default_opts(faithful=False, solver="fast", warps_per_sm=wps, park_after=park)
    args = (wd, p0, ex(lo, torch.float64), ex(hi, torch.float64), ex(lmin, torch.uint8), ex(lmax, torch.uint8))
    r = engine.gaussfit_batch(*args, opts=o, solver="fast", rescue=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = engine.gaussfit_batch(*args, opts=o, solver="fast", rescue=False)
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return r, min(ts)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    parks = [int(v) for v in sys.argv[2:]] or [0]
    for kind in ("isolated", "dense"):
        w = windows(kind, n)
        lo, hi, lmin, lmax = engine.GAUSSFIT_DEFAULT_LIMITS
        for dt in ("float64", "uint16"):
            wd = engine._device_windows(w.astype(dt), torch.device("cuda"))
            p0 = engine.moments_batch(wd, lo, hi, lmin, lmax)
            base = None
            for name, wps in VARIANTS.items():
                for park in (parks if name.startswith("hybrid") else [0]):
                    r, ms = run(wd, p0, wps, park)
                    st = r.status.cpu().numpy()
                    line = "%-8s %-7s %-8s park %3d %8.3f ms  %.3e fits/s  status>0 %.4f  -16: %d  maxiter: %d  mean niter %.2f" % (
                        kind, dt, name, park, ms, n / ms * 1e3, (st > 0).mean(), (st == -16).sum(), (st == 5).sum(), r.niter.double().mean().item())
                    if base is None:
                        base = r
                    else:
                        Pa, Pb = base.params.cpu().numpy(), r.params.cpu().numpy()
                        ca, cb = base.chi2.cpu().numpy(), r.chi2.cpu().numpy()
                        line += "  | vs thread: identical %.4f, within tol %.4f, status equal %.4f, chi2 1e-6 %.4f" % (
                            (Pa == Pb).all(axis=1).mean(), agree(Pa, Pb).mean(), (base.status == r.status).float().mean().item(),
                            (np.abs(ca - cb) <= 1e-6 * np.abs(ca)).mean())
                    print(line, flush=True)


if __name__ == "__main__":
    main()
