"""Developer diagnostic (GPU): compares the CUDA path with the committed goldens and prints a summary."""
import glob, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fluorosequencingimageanalysis_b200 import engine, pflib, gaussfitter, synth
G = os.path.join(ROOT, "tests", "golden")

def relerr(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)

print("device", torch.cuda.get_device_name(0))
# ---------------- detection
for f in sorted(glob.glob(os.path.join(G, "detect_*.npz"))):
    d = np.load(f)
    kw = {k[3:]: d[k] for k in d.files if k.startswith("kw_")}
    kw = {k: (v if v.ndim else v.item()) for k, v in kw.items()}
    det = engine.detect_batch(d["img"], **kw)
    hw = det.cand_hw[:det.total].cpu().numpy()
    ok = hw.shape == d["cands"].shape and np.array_equal(hw, d["cands"])
    print("detect %-22s n=%6d golden=%6d equal=%s thr=%r golden=%r" % (os.path.basename(f)[7:-4], det.total, len(d["cands"]), ok, float(det.thr[0].item()), float(d["thr"])))
# batch of frames: each frame equals its single-frame result
fr = np.stack([synth.synth_frame(s) for s in (0, 3, 5)])
det = engine.detect_batch(fr)
pf = det.per_frame()
g0 = np.load(os.path.join(G, "detect_c1_seed0.npz"))["cands"]; g3 = np.load(os.path.join(G, "detect_c1_seed3.npz"))["cands"]
print("batch detect: frame0 ok", np.array_equal(pf[0], g0), "frame1 ok", np.array_equal(pf[1], g3), det.n_cand.cpu().numpy())

# ---------------- KAT-1
k = np.load(os.path.join(G, "kat1_fit.npz"))
out = pflib._fit_2d_gaussian(k["sub"])
print("KAT1 gpu ", [repr(v) for v in out[:7]])
print("KAT1 gold", [repr(float(v)) for v in k["fit7"]])
mp = gaussfitter.gaussfit(k["sub"], params=(np.median(k["sub"]), np.amax(k["sub"]), 2.5, 2.5, 1, 1, 0), limitedmin=[True]*7,
        limitedmax=[False,False,True,True,True,True,True], minpars=np.array([0,(np.amax(k["sub"])-np.mean(k["sub"]))/3.0,2,2,.75,.75,0]),
        maxpars=np.array([0,0,3,3,2,2,360.]), returnmp=True)
print("KAT1 status/niter/nfev/fnorm", mp.status, mp.niter, mp.nfev, mp.fnorm, "gold", int(k["status"]), int(k["niter"]), int(k["nfev"]), float(k["fnorm"]))
print("KAT1 perror", mp.perror, "gold", k["perror"])

# ---------------- fits5 (all candidates of the config-1 frame)
g = np.load(os.path.join(G, "fits5_seed0.npz"))
img = synth.synth_frame(0)
for faithful in (True, False):
    torch.cuda.synchronize(); t = time.time()
    res = engine.find_peptides_batch(img, faithful=faithful, want_fit_img=True)
    torch.cuda.synchronize(); dt = time.time() - t
    assert np.array_equal(res.cand_hw, g["cands"])
    P = res.fit[:, [2, 3, 0, 1, 4, 5, 6]].copy()      # H, A, h0, w0, sh, sw, th  -> mpfit order needs window coords
    P[:, 2] = res.fit[:, 0] - res.cand_hw[:, 0] + 2.5
    P[:, 3] = res.fit[:, 1] - res.cand_hw[:, 1] + 2.5
    ref = g["ref_params"] if faithful else g["clean_params"]
    rst = g["ref_status"] if faithful else g["clean_status"]
    rchi = g["ref_fnorm"] if faithful else g["clean_fnorm"]
    st = res.ints[:, 0]
    e = np.max(relerr(P[:, :6], ref[:, :6]), axis=1)
    robust = g["n_qrsolv"] == 0
    print("fits5 faithful=%s: %.3fs  status agree %.4f  params<1e-4 all %.4f robust %.4f (n_robust %d)  niter agree %.4f  gpu_nqrsolv0 == robust %.4f" % (
        faithful, dt, np.mean(st == rst), np.mean(e < 1e-4), np.mean(e[robust] < 1e-4), robust.sum(),
        np.mean(res.ints[:, 1] == (g["ref_niter"] if faithful else g["clean_niter"])), np.mean((res.ints[:, 3] == 0) == robust)))
    print("   status hist gpu", dict(zip(*np.unique(st, return_counts=True))), "ref", dict(zip(*np.unique(rst, return_counts=True))))
    chi = res.fit[:, 10]
    print("   chi2_gpu <= chi2_ref(1+1e-9): %.4f ; vs faithful-ref: %.4f ; mean niter gpu %.2f ref %.2f" % (np.mean(chi <= rchi * (1 + 1e-9)), np.mean(chi <= g["ref_fnorm"] * (1 + 1e-9)), res.ints[:, 1].mean(), (g["ref_niter"] if faithful else g["clean_niter"]).mean()))
    if faithful:
        print("   r2 maxerr (robust)", np.max(np.abs(res.fit[robust, 8] - g["r_2"][robust])), " s_n maxrel", np.max(relerr(res.fit[:, 9], g["s_n"])), "rmse maxrel(robust)", np.max(relerr(res.fit[robust, 7], g["rmse"][robust])))
        bad = np.nonzero(robust & (e >= 1e-4))[0][:5]
        for i in bad:
            print("   BAD robust", i, "gpu", P[i], st[i], res.ints[i], "ref", ref[i], rst[i], g["ref_niter"][i])
        keys, idx = pflib.consolidate_packed(res.cand_hw, res.fit, img.shape)
        gk = set(map(tuple, g["final_keys"])); mk = set(map(tuple, keys))
        print("   final PSFs gpu %d golden %d overlap %d" % (len(mk), len(gk), len(mk & gk)))

# ---------------- fits11
g = np.load(os.path.join(G, "fits11_seed0.npz"))
n = len(g["windows"])
lo = np.tile(np.array([0, 0, 0, 0, 0, 0, 0.]), (n, 1)); hi = np.tile(np.array([0, 0, 0, 0, 0, 0, 360.]), (n, 1))
lmin = np.tile(np.array([0, 0, 0, 0, 1, 1, 1], dtype=np.uint8), (n, 1)); lmax = np.tile(np.array([0, 0, 0, 0, 0, 0, 1], dtype=np.uint8), (n, 1))
for faithful in (True, False):
    r = engine.gaussfit_batch(g["windows"], g["p0"], lo, hi, lmin, lmax, faithful=faithful)
    P = r.params.cpu().numpy(); st = r.status.cpu().numpy()
    ref = g["ref_params"] if faithful else g["clean_params"]; rst = g["ref_status"] if faithful else g["clean_status"]
    e = np.max(relerr(P[:, :6], ref[:, :6]), axis=1); robust = g["n_qrsolv"] == 0
    print("fits11 faithful=%s status agree %.3f params<1e-4 all %.3f robust %.3f (%d) niter agree %.3f" % (faithful, np.mean(st == rst), np.mean(e < 1e-4), np.mean(e[robust] < 1e-4), robust.sum(), np.mean(r.niter.cpu().numpy() == (g["ref_niter"] if faithful else g["clean_niter"]))))

# ---------------- throughput probe
fr = np.stack([synth.synth_frame(100 + s) for s in range(16)])
frd = engine.to_device_frames(fr)
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    det = engine.detect_batch(frd)
    torch.cuda.synchronize(); t1 = time.time()
    fit, ints, _ = engine.fit_candidates(frd, det.cand_hw, det.cand_frame, det.total, faithful=True)
    torch.cuda.synchronize(); t2 = time.time()
    fit2, ints2, _ = engine.fit_candidates(frd, det.cand_hw, det.cand_frame, det.total, faithful=False)
    torch.cuda.synchronize(); t3 = time.time()
    print("16 frames: detect %.2f ms (%d cands); fit faithful %.2f ms = %.3g fits/s; clean %.2f ms = %.3g fits/s" % ((t1 - t) * 1e3, det.total, (t2 - t1) * 1e3, det.total / (t2 - t1), (t3 - t2) * 1e3, det.total / (t3 - t2)))
