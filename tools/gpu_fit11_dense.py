"""Developer diagnostic (GPU): FAST vs MINPACK on 11x11 windows cut from a dense 2048x2048 frame (config 4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fluorosequencingimageanalysis_b200 import engine, synth
fr = synth.synth_frame(40, H=2048, W=2048, n_spots=20000)
_, cr, cc, _ = synth.spot_layout(40, 2048, 2048, 20000)
ok = (cr > 8) & (cr < 2040) & (cc > 8) & (cc < 2040)
r0, c0 = np.rint(cr[ok]).astype(int), np.rint(cc[ok]).astype(int)
win = np.stack([fr[a - 5:a + 6, b - 5:b + 6] for a, b in zip(r0, c0)]).astype(np.float64)
wd = torch.from_numpy(win).cuda()
rf, p0 = engine.gaussfit_default_batch(wd, solver="fast", faithful=False)
rm, _ = engine.gaussfit_default_batch(wd, solver="minpack", faithful=False)
sf, sm = rf.status.cpu().numpy(), rm.status.cpu().numpy()
print("n", len(sf), "FAST status", dict(zip(*[x.tolist() for x in np.unique(sf, return_counts=True)])))
print("MINPACK status", dict(zip(*[x.tolist() for x in np.unique(sm, return_counts=True)])))
nf, nm = rf.niter.cpu().numpy(), rm.niter.cpu().numpy()
print("niter FAST pct 50/90/99/max", np.percentile(nf, [50, 90, 99]), nf.max(), " MINPACK", np.percentile(nm, [50, 90, 99]), nm.max())
cf, cm = rf.chi2.cpu().numpy(), rm.chi2.cpu().numpy()
good = sf > 0
print("chi2 FAST <= MINPACK(1+1e-6): %.4f of converged; ratio median %.4f" % (np.mean(cf[good] <= cm[good] * (1 + 1e-6)), np.median(cf[good] / cm[good])))
bad = np.nonzero(sf <= 0)[0][:5]
P0 = p0.cpu().numpy(); PF = rf.params.cpu().numpy(); PM = rm.params.cpu().numpy()
np.set_printoptions(precision=4, suppress=True, linewidth=200)
for i in bad:
    print("bad", i, "status", sf[i], "niter", nf[i], "\n  p0 ", P0[i], "\n  fast", PF[i], "\n  mpk ", PM[i], "status", sm[i])
