"""Developer A/B helper (GPU): detection launches of a 200-frame batch timed alone, for the build that is loaded."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine, synth
pool = synth.experiment_field_pool(5, 8, n_cycles=10)
frames = np.concatenate([pool.reshape(-1, 512, 512)] * 3)[:200]
fd = torch.from_numpy(frames.view(np.int16)).view(torch.uint16).cuda()
p = engine.FieldPipeline(200, 512, 512, dtype=torch.uint16, solver="fast", faithful=False)
p.run_detect_only(fd); torch.cuda.synchronize()
ts = []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); p.run_detect_only(fd); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
print("%s: detection of 200 frames %.4f ms (min of 6), candidates %d" % (sys.argv[1] if len(sys.argv) > 1 else "", min(ts), p.total()))
