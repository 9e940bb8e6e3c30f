"""Developer diagnostic (GPU): the FAST solver's park/resume scheduling must not change any result;
timing for several park_after values."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fluorosequencingimageanalysis_b200 import engine, synth, _lib
fr = synth.synth_timetrace(1, n_frames=int(sys.argv[1]) if len(sys.argv) > 1 else 40)
frd = engine.to_device_frames(fr)
det = engine.detect_batch(frd)
print("candidates", det.total)
ref = None
for park in (0, 16, 24, 32, 48, 64):
    o = _lib.default_opts(faithful=False, solver="fast", park_after=park)
    ts = []
    for rep in range(5):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fit, ints, _ = engine.fit_candidates(frd, det.cand_hw, det.cand_frame, det.total, opts=o)
        e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    same = ""
    if ref is None: ref = (fit.clone(), ints.clone())
    else: same = "identical to park_after=0: fit %s ints %s" % (torch.equal(ref[0].view(torch.int64), fit.view(torch.int64)), torch.equal(ref[1], ints))
    print("park_after %3d: %.3f ms best -> %.4g fits/s  %s" % (park, min(ts[1:]), det.total / (min(ts[1:]) * 1e-3), same), flush=True)
