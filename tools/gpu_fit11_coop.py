"""Developer diagnostic (GPU): the cooperative 11x11 FAST kernel against the thread-per-window one (results and time) on
isolated-spot windows (configs[0]) and on windows cut from the dense 2048x2048 frame (configs[3])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine, synth, _lib
from test_gpu_fit import agree


def windows(kind, n):
    if kind == "isolated":
        out = []
        seed = 100
        while sum(len(o) for o in out) < n:
            img, cr, cc, amp = synth.synth_frame_with_truth(seed)
            out.append(synth.cut_windows(img, cr, cc, 11))
            seed += 1
            if seed > 110:
                break
        w = np.concatenate(out)
    else:
        img, cr, cc, amp = synth.synth_frame_with_truth(40, H=2048, W=2048, n_spots=20000)
        w = synth.cut_windows(img, cr, cc, 11)
    reps = -(-n // len(w))
    return np.concatenate([w] * reps)[:n]


def run(wd, p0, variant):
    n = wd.shape[0]
    lo, hi, lmin, lmax = engine.GAUSSFIT_DEFAULT_LIMITS
    dev = wd.device
    ex = lambda v, dt: torch.as_tensor(v, dtype=dt).to(dev).expand(n, 7).contiguous()
    o = _lib.default_opts(faithful=False, solver="fast", warps_per_sm=(-2 if variant == "coop" else 0))
    args = (wd, p0, ex(lo, torch.float64), ex(hi, torch.float64), ex(lmin, torch.uint8), ex(lmax, torch.uint8))
    r = engine.gaussfit_batch(*args, opts=o, solver="fast", rescue=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = engine.gaussfit_batch(*args, opts=o, solver="fast", rescue=False)
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return r, min(ts)


def main():
  for kind in ("isolated", "dense"):
      w = windows(kind, 200000)
      wd = torch.from_numpy(w).cuda()
      lo, hi, lmin, lmax = engine.GAUSSFIT_DEFAULT_LIMITS
      p0 = engine.moments_batch(wd, lo, hi, lmin, lmax)
      res = {}
      for variant in ("thread", "coop"):
          r, ms = run(wd, p0, variant)
          res[variant] = r
          st = r.status.cpu().numpy()
          ni = r.niter.cpu().numpy()
          print("%-8s %-6s %8.3f ms  %.3e fits/s  status>0 %.4f  -16: %d  maxiter: %d  mean niter %.1f"
                % (kind, variant, ms, len(w) / ms * 1e3, (st > 0).mean(), (st == -16).sum(), (st == 5).sum(), ni.mean()))
      a, b = res["thread"], res["coop"]
      Pa, Pb = a.params.cpu().numpy(), b.params.cpu().numpy()
      ca, cb = a.chi2.cpu().numpy(), b.chi2.cpu().numpy()
      ok = agree(Pa, Pb)
      print("%-8s coop vs thread: parameters within tolerance %.4f, status equal %.4f, chi2 within 1e-6 %.4f, coop chi2 <= thread (1+1e-6) %.4f"
            % (kind, ok.mean(), (a.status == b.status).float().mean().item(), (np.abs(ca - cb) <= 1e-6 * np.abs(ca)).mean(),
               (cb <= ca * (1 + 1e-6)).mean()))


if __name__ == "__main__":
    main()
