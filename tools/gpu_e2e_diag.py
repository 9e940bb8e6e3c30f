"""Developer diagnostic (GPU): where does the e2e region lose time against the resident region?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine, synth
F, H, W, depth, steps = 40, 512, 512, 6, 200
host = []
for v in range(8):
    st = synth.synth_timetrace(1 + v, n_frames=F)
    host.append(torch.from_numpy(st.view(np.int16)).view(torch.uint16).pin_memory())
dev = [t.cuda() for t in host]
for fetch in ("candidates", "psfs"):
    for h2d in (False, True):
        for d2h in (False, True):
            fs = engine.FieldStream(F, H, W, dtype=torch.uint16, depth=depth, host_io=True, fetch=fetch, faithful=False, solver="fast", warps_per_sm=4)
            src = host if h2d else dev
            def run(n):
                tick = []
                for k in range(n):
                    tick.append(fs.submit(src[k % 8]))
                    if d2h:
                        if k >= depth - 2: fs.begin_fetch(tick[k - depth + 2])
                        if k >= depth - 1: fs.end_fetch(tick[k - depth + 1])
                    elif k >= depth - 1:
                        tick[k - depth + 1]["ev_count"].synchronize()
                if d2h:
                    for t in tick[max(0, n - depth + 1):]: fs.end_fetch(t)
                fs.synchronize()
            run(8); torch.cuda.synchronize()
            t0 = time.perf_counter(); run(steps); dt = time.perf_counter() - t0
            print("fetch=%-10s h2d=%d d2h=%d: %.3f ms/step" % (fetch, h2d, d2h, dt / steps * 1e3), flush=True)
            del fs
