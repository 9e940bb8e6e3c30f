"""Developer A/B helper (GPU): results hash + timing of the FAST frame-path LM kernel for the build that is loaded.
    python tools/gpu_lm_ab.py <tag>      -> gpurun_out/lm_ab_<tag>.npz (fit bits of a 3-frame batch) and one timing line"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine, synth

tag = sys.argv[1]
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 3          # launches in flight
stack = np.stack([synth.synth_frame(0), synth.synth_frame(3), synth.synth_frame(11, n_spots=1000)])
res = engine.find_peptides_batch(stack, solver="fast", faithful=False)
h = hashlib.sha256(res.fit.tobytes() + res.ints.tobytes()).hexdigest()[:16]
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "lm_ab_%s.npz" % tag), fit=res.fit, ints=res.ints)
# timing: 200-frame launches, fit only, lone and 3 in flight
pool = synth.experiment_field_pool(5, 8, n_cycles=10)
frames = np.concatenate([pool.reshape(-1, 512, 512)] * 3)[:200]
fd = torch.from_numpy(frames.view(np.int16)).view(torch.uint16).cuda()
pipes = [engine.FieldPipeline(200, 512, 512, dtype=torch.uint16, solver="fast", faithful=False, warps_per_sm=4) for _ in range(NP)]
streams = [torch.cuda.Stream() for _ in range(NP)]
for p in pipes:
    p.run(fd)
torch.cuda.synchronize()
n = pipes[0].total()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    pipes[0].run_fit_only(fd)
e1.record(); e1.synchronize()
lone = e0.elapsed_time(e1) / 3
cur = torch.cuda.current_stream()
for s in streams:
    s.wait_stream(cur)
e0.record()
for rep in range(4):
    for p, s in zip(pipes, streams):
        with torch.cuda.stream(s):
            p.run_fit_only(fd)
for s in streams:
    cur.wait_stream(s)
e1.record(); e1.synchronize()
conc = e0.elapsed_time(e1) / (4 * NP)
print("%-10s hash %s  fits/launch %d  lone %.3f ms (%.3e fits/s)  %d in flight %.3f ms per launch (%.3e fits/s)  mean niter %.2f"
      % (tag, h, n, lone, n / lone * 1e3, NP, conc, n / conc * 1e3, pipes[0].out_int[:n, 1].double().mean().item()))
