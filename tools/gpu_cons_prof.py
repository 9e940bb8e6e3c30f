"""Developer diagnostic (GPU): time fsq_consolidate / fsq_pack_psfs on the bench batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fluorosequencingimageanalysis_b200 import engine, synth
fr = synth.synth_timetrace(1, n_frames=40)
res = engine.find_peptides_batch(fr, faithful=False, solver="fast", to_host=False)
n = int(res.cand_hw.shape[0])
c = engine.consolidate_batch(res.cand_hw, res.cand_frame, res.fit, n, 40)
st = c.state.cpu().numpy()
print("candidates %d  accepted %d  final %d" % (n, (st > 0).sum(), (st >= 2).sum()))
for rep in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    e0.record()
    c = engine.consolidate_batch(res.cand_hw, res.cand_frame, res.fit, n, 40, out=c)
    e1.record()
    pk = engine.pack_psfs_batch(c, res.cand_frame, res.fit, n, 40)
    e2.record(); e2.synchronize()
    print("consolidate %.3f ms  pack %.3f ms" % (e0.elapsed_time(e1), e1.elapsed_time(e2)))
