"""Developer diagnostic (GPU): dump per-fit results and per-step traces for offline comparison with the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fluorosequencingimageanalysis_b200 import engine, pflib, synth
G = os.path.join(ROOT, "tests", "golden")
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)
g = np.load(os.path.join(G, "fits5_seed0.npz"))
img = synth.synth_frame(0)
cands = g["cands"]
subs = np.stack([img[h - 2:h + 3, w - 2:w + 3].astype(np.int64) for h, w in cands])
p0, lo, hi, lmin, lmax = pflib._pflib_limits(subs)
save = {}
for faithful in (True, False):
    tag = "f" if faithful else "c"
    res = engine.find_peptides_batch(img, faithful=faithful)
    save["pf_fit_" + tag] = res.fit; save["pf_int_" + tag] = res.ints
    r = engine.gaussfit_batch(subs, p0, lo, hi, lmin, lmax, faithful=faithful)
    P = r.params.cpu().numpy()
    save["gen_params_" + tag] = P; save["gen_status_" + tag] = r.status.cpu().numpy(); save["gen_niter_" + tag] = r.niter.cpu().numpy()
    save["gen_nfev_" + tag] = r.nfev.cpu().numpy(); save["gen_chi2_" + tag] = r.chi2.cpu().numpy(); save["gen_nq_" + tag] = r.n_qrsolv.cpu().numpy()
    PF = res.fit[:, [2, 3, 0, 1, 4, 5, 6]].copy()
    PF[:, 2] = res.fit[:, 0] - cands[:, 0] + 2.5; PF[:, 3] = res.fit[:, 1] - cands[:, 1] + 2.5
    print(tag, "pflib-kernel vs generic-kernel: status equal %.4f niter equal %.4f params maxabs %.3g" % (
        np.mean(res.ints[:, 0] == save["gen_status_" + tag]), np.mean(res.ints[:, 1] == save["gen_niter_" + tag]), np.max(np.abs(PF - P))))
    n_tr = 600
    rt, trace = engine.gaussfit_batch_trace(subs[:n_tr].astype(np.float64), p0[:n_tr], lo[:n_tr], hi[:n_tr], lmin[:n_tr], lmax[:n_tr], faithful=faithful, trace_steps=400)
    save["trace_" + tag] = trace.astype(np.float64)
    print(tag, "trace run status equal generic", np.mean(rt.status.cpu().numpy() == save["gen_status_" + tag][:n_tr]))
np.savez_compressed(os.path.join(OUT, "fits5_gpu.npz"), **save)
print("saved")
