"""Developer/profiling driver (GPU): time the detection launches on the bench batch (40 frames 512x512 u16)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fluorosequencingimageanalysis_b200 import engine, synth
nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 40
stacks = [engine.to_device_frames(synth.synth_timetrace(1 + v, n_frames=nfr)) for v in range(8)]
pipe = engine.FieldPipeline(nfr, 512, 512, dtype=torch.uint16)
ts = []
for rep in range(12):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); pipe.run_detect_only(stacks[rep % 8]); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
n = pipe.total()
b = nfr * 512 * 512 * 2 + 8 * n
t = sum(ts[4:]) / len(ts[4:])
print("detect: %d frames, %d candidates, %.4f ms avg -> %.1f GB/s algorithmic, %.0f frames/s" % (nfr, n, t, b / t / 1e6, nfr / t * 1e3))
