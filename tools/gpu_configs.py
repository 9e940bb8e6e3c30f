"""Developer measurement (GPU): the BASELINE.json configurations other than the bench's configs[1], plus the
time-trace path.  Prints one line per measurement; kept under profiles/ as a text file."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine, synth, phase_correlate as pc


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)


# ---- config 4: 2048x2048 frames, 20 000 spots
fr = np.stack([synth.synth_frame(40 + k, H=2048, W=2048, n_spots=20000) for k in range(4)])
frd = engine.to_device_frames(fr)
pipe = engine.FieldPipeline(4, 2048, 2048, dtype=torch.uint16, faithful=False, solver="fast", warps_per_sm=0, consolidate=True,
                            cap_per_frame=400000, cap_psf_per_frame=60000)
ms = timed(lambda: pipe.run(frd))
n, m = pipe.total(), pipe.total_psfs()
print("config 4: 4 frames 2048x2048 x 20k spots: %d candidates, %d final PSFs; detection + fit + metrics + consolidation %.3f ms "
      "-> %.4g fits/s, %.1f frames/s" % (n, m, ms, n / ms * 1e3, 4 / ms * 1e3))
ms_d = timed(lambda: pipe.run_detect_only(frd))
print("config 4: detection alone %.3f ms (%.1f Mpixel/s)" % (ms_d, 4 * 2048 * 2048 / ms_d / 1e3))
# 11x11 windows around the true spots, default gaussfit arguments, start values on the device
rng = np.random.default_rng(0)
_, cr, cc, _ = synth.spot_layout(40, 2048, 2048, 20000)
ok = (cr > 8) & (cr < 2040) & (cc > 8) & (cc < 2040)
r0, c0 = np.rint(cr[ok]).astype(int), np.rint(cc[ok]).astype(int)
win = np.stack([fr[0][a - 5:a + 6, b - 5:b + 6] for a, b in zip(r0, c0)]).astype(np.float64)
win = np.concatenate([win] * 10)[:200000]
wd = torch.from_numpy(win).cuda()
for solver in ("fast", "minpack"):
    w = wd if solver == "fast" else wd[:20000]
    ms = timed(lambda: engine.gaussfit_default_batch(w, solver=solver, faithful=False), reps=3)
    r, _ = engine.gaussfit_default_batch(w, solver=solver, faithful=False)
    print("config 4: %d windows 11x11, gaussfit default arguments (moments on the device), solver %s: %.3f ms -> %.4g fits/s, "
          "status>0 %.4f, mean niter %.1f" % (w.shape[0], solver, ms, w.shape[0] / ms * 1e3, (r.status > 0).double().mean().item(), r.niter.double().mean().item()))
ms = timed(lambda: engine.moments_batch(wd))
print("          moments kernel alone: %.3f ms for %d windows" % (ms, wd.shape[0]))
del pipe, frd, wd

# ---- config 3: 10 cycles x 100 fields, 1000 spots per field (here: 10 fields x 10 cycles per step, 4 steps cycled)
ex = synth.synth_experiment(7, n_fields=10, n_cycles=10)             # [10,10,512,512]
stack = torch.from_numpy(ex.reshape(100, 512, 512).view(np.int16)).view(torch.uint16).pin_memory()
fs = engine.FieldStream(100, 512, 512, dtype=torch.uint16, depth=4, host_io=True, fetch="psfs", faithful=False, solver="fast", warps_per_sm=4)
def run(nsteps):
    tick, tot, psf = [], 0, 0
    for k in range(nsteps):
        tick.append(fs.submit(stack))
        if k >= 2: fs.begin_fetch(tick[k - 2])
        if k >= 3:
            r = fs.end_fetch(tick[k - 3]); tot += r[0]; psf += r[1]
    for t in tick[max(0, nsteps - 3):]:
        r = fs.end_fetch(t); tot += r[0]; psf += r[1]
    return tot, psf
run(4); torch.cuda.synchronize()
t0 = time.perf_counter(); tot, psf = run(20); dt = time.perf_counter() - t0
print("config 3: 100-frame steps (10 fields x 10 cycles, 1000 spots/field), host frames in, final PSFs out: %.3f ms/step, "
      "%.4g fits/s, %.0f frames/s, %.1f candidates and %.1f PSFs per frame" % (dt / 20 * 1e3, tot / dt, 2000 / dt, tot / 2000, psf / 2000))
del fs

# ---- time-trace path: 40-frame movie, frame 0 peak-fitted, PSFs tracked, photometry per frame
mv = synth.synth_timetrace(1, n_frames=40)
mvd = engine.to_device_frames(mv)
tt = engine.timetrace_batch(mvd)
m = len(tt["psf_int"])
spots = torch.from_numpy(tt["psf_int"][:, 1:3].copy()).cuda()
ms_t = timed(lambda: engine.track_centroid_batch(mvd, spots))
hw = torch.from_numpy(tt["track_hw"].reshape(-1, 2).copy()).cuda()
live = torch.from_numpy((tt["track_state"].reshape(-1) != 0)).cuda()
fidx = torch.arange(40, dtype=torch.int32, device="cuda").repeat(m)
ms_p = timed(lambda: engine.photometry_batch(mvd, hw[live], fidx[live]))
t0 = time.perf_counter(); engine.timetrace_batch(mvd); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("time trace: 40 frames 512x512, %d frame-0 PSFs: tracking %.3f ms (%.4g spot-frames/s), mexican-hat photometry of %d spot-frames "
      "%.3f ms, whole path incl. frame-0 find_peptides and host copies %.1f ms" % (m, ms_t, m * 39 / ms_t * 1e3, int(live.sum()), ms_p, dt * 1e3))
ms_r = timed(lambda: pc.phase_correlate_batch(mvd[:-1], mvd[1:], 20))
print("registration: 39 consecutive pairs of 512x512 frames, upsample 20: %.3f ms (cuFFT + cuBLAS)" % ms_r)
