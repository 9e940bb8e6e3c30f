"""Developer/profiling driver (GPU): run the candidate-fit kernel of one solver a few times on a
40-frame stack (the bench workload) -- small enough for `ncu --set full`.
    python tools/gpu_fit_prof.py <solver> [reps] [n_frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fluorosequencingimageanalysis_b200 import engine, synth, _lib
solver = sys.argv[1] if len(sys.argv) > 1 else "fast64"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nfr = int(sys.argv[3]) if len(sys.argv) > 3 else 40
faithful = solver.endswith("-faithful")
solver = solver.replace("-faithful", "")
fr = synth.synth_timetrace(1, n_frames=nfr)
frd = engine.to_device_frames(fr)
det = engine.detect_batch(frd)
ts = []
for rep in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    o = _lib.default_opts(faithful=faithful, solver=solver, warps_per_sm=int(os.environ.get("FSQ_WARPS", "0")),
                          park_after=int(os.environ.get("FSQ_PARK", "0")), maxiter=int(os.environ.get("FSQ_MAXITER", "200")))
    fit, ints, _ = engine.fit_candidates(frd, det.cand_hw, det.cand_frame, det.total, opts=o)
    e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print("%s: %d candidates, %.3f ms best -> %.4g fits/s; mean niter %.2f nfev %.2f" % (
    solver, det.total, min(ts), det.total / (min(ts) * 1e-3), ints[:, 1].double().mean().item(), ints[:, 2].double().mean().item()))
