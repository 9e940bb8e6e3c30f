#!/usr/bin/env python
"""
TEST INFRASTRUCTURE.  Generates the committed golden vectors under tests/golden/ by running
the REFERENCE ITSELF (oracle/_ref = mechanical py3 transliteration built by
oracle/build_ref.py from /root/reference) on seeded synthetic inputs, together with the two
flavours of the restated oracle (oracle/lm_oracle.py faithful / clean) so that the robust set
 is stored next to the reference answers.

Run in the build container (needs /root/reference):
    python oracle/make_golden.py [--procs 8]

Outputs (all small .npz):
  tests/golden/kat1_fit.npz              SURVEY.md App. D KAT-1
  tests/golden/detect_<name>.npz         candidate lists + thresholds for several frames/params
  tests/golden/fits5_seed0.npz           every candidate of the config-1 frame (seed 0):
                                         reference params/status/niter/nfev/fnorm/perror,
                                         oracle n_qrsolv / n_reject, clean-oracle params/status/fnorm,
                                         metrics (r_2, rmse, s_n), final PSF keys after consolidation
  tests/golden/fits11_seed0.npz          200 windows 11x11, default gaussfit arguments
"""
import argparse
import hashlib
import multiprocessing
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import build_ref                                  # noqa: E402
from oracle import pflib_oracle as po                         # noqa: E402
from fluorosequencingimageanalysis_b200 import synth          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
_REF = None


def ref():
    global _REF
    if _REF is None:
        build_ref.build(quiet=True)
        _REF = build_ref.load()
    return _REF


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------ per-window workers
def _fit5(sub):
    pflib, gaussfitter, _ = ref()
    p, lmin, lmax, mn, mx = po.pflib_fit_args(sub)
    mp = gaussfitter.gaussfit(sub, params=p, limitedmin=lmin, limitedmax=lmax, minpars=mn,
                              maxpars=mx, returnmp=True)
    fa, fit_fa = po.gaussfit(sub, params=p, limitedmin=lmin, limitedmax=lmax, minpars=mn,
                             maxpars=mx, faithful=True)
    cl, fit_cl = po.gaussfit(sub, params=p, limitedmin=lmin, limitedmax=lmax, minpars=mn,
                             maxpars=mx, faithful=False)
    same = (np.array_equal(mp.params, fa.params) and mp.status == fa.status and
            mp.niter == fa.niter and mp.nfev == fa.nfev and mp.fnorm == fa.fnorm)
    fitimg = gaussfitter.twodgaussian(mp.params, 0, 1, 1)(*np.indices(sub.shape))
    r_2, rmse, s_n = po.fit_metrics(sub, fitimg)
    perr = mp.perror if mp.perror is not None else np.full(7, np.nan)
    return (mp.params, perr, mp.status, mp.niter, mp.nfev, mp.fnorm, fa.n_qrsolv, fa.n_reject,
            cl.params, cl.status, cl.niter, cl.nfev, cl.fnorm, cl.n_qrsolv, same, r_2, rmse, s_n)


def _fit11(win):
    pflib, gaussfitter, _ = ref()
    mp = gaussfitter.gaussfit(win, returnmp=True)
    fa, _ = po.gaussfit(win, faithful=True)
    cl, _ = po.gaussfit(win, faithful=False)
    same = (np.array_equal(mp.params, fa.params) and mp.status == fa.status and
            mp.niter == fa.niter and mp.nfev == fa.nfev and mp.fnorm == fa.fnorm)
    perr = mp.perror if mp.perror is not None else np.full(7, np.nan)
    p0 = np.array(po.moments(win), dtype=float)
    return (mp.params, perr, mp.status, mp.niter, mp.nfev, mp.fnorm, fa.n_qrsolv, fa.n_reject,
            cl.params, cl.status, cl.niter, cl.nfev, cl.fnorm, same, p0)


def _stack(rows, names):
    out = {}
    for i, n in enumerate(names):
        out[n] = np.array([r[i] for r in rows])
    return out


# ------------------------------------------------------------------ generators
def make_kat1():
    pflib, gaussfitter, _ = ref()
    r, c = np.indices((5, 5))
    sub = np.random.default_rng(0).poisson(
        100 + 2000 * np.exp(-((r - 2.3) ** 2 + (c - 2.7) ** 2) / (2 * 1.1 ** 2))).astype(np.int64)
    out = pflib._fit_2d_gaussian(sub)
    p, lmin, lmax, mn, mx = po.pflib_fit_args(sub)
    mp = gaussfitter.gaussfit(sub, params=p, limitedmin=lmin, limitedmax=lmax, minpars=mn,
                              maxpars=mx, returnmp=True)
    np.savez(os.path.join(GOLD, "kat1_fit.npz"), sub=sub,
             fit7=np.array([float(v) for v in out[:7]]), fit_img=out[7],
             status=mp.status, niter=mp.niter, nfev=mp.nfev, fnorm=mp.fnorm, perror=mp.perror,
             covar=mp.covar, s_n=pflib.illumina_s_n(sub))
    print("kat1", out[:7], mp.status, mp.niter, mp.nfev, mp.fnorm)


def make_detect():
    pflib, _, _ = ref()
    cases = {
        "c1_seed0": dict(img=synth.synth_frame(0), kw={}),
        "c1_seed3": dict(img=synth.synth_frame(3), kw={}),
        "dense_1000": dict(img=synth.synth_frame(11, n_spots=1000), kw={}),
        "rect_200x333": dict(img=synth.synth_frame(5, H=200, W=333, n_spots=120), kw={}),
        "cstd3_mf7": dict(img=synth.synth_frame(7, H=256, W=256, n_spots=150),
                          kw=dict(median_filter_size=7, c_std=3)),
        "mf3": dict(img=synth.synth_frame(8, H=128, W=160, n_spots=60),
                    kw=dict(median_filter_size=3)),
        "k3x3": dict(img=synth.synth_frame(9, H=128, W=128, n_spots=50),
                     kw=dict(correlation_matrix=np.array([[-1, -1, -1], [-1, 8, -1], [-1, -1, -1]]))),
        "k7x7": dict(img=synth.synth_frame(10, H=96, W=120, n_spots=40),
                     kw=dict(correlation_matrix=(np.arange(49).reshape(7, 7) % 5 - 2) * 100 +
                             np.pad(np.full((3, 3), 900), 2))),
        "full16bit": dict(img=(np.random.default_rng(21).integers(0, 65536, (96, 96))
                               .astype(np.uint16)), kw={}),
        "flat": dict(img=np.full((64, 64), 400, dtype=np.uint16), kw={}),
        "tiny_6x7": dict(img=synth.synth_frame(2, H=24, W=24, n_spots=2)[:6, :7].copy(), kw={}),
    }
    for name, c in cases.items():
        img = c["img"]
        cands = pflib._psf_candidates(img, **c["kw"])
        _, cm, thr = po.detect_maps(img, **c["kw"])
        assert cands == po.psf_candidates(img, **c["kw"]), name
        kw = {("kw_" + k): np.asarray(v) for k, v in c["kw"].items()}
        np.savez_compressed(os.path.join(GOLD, "detect_%s.npz" % name), img=img,
                            cands=np.array(cands, dtype=np.int32).reshape(-1, 2), thr=thr,
                            cm_sum=np.int64(cm.sum()), cm_max=np.int64(cm.max()), **kw)
        print("detect", name, img.shape, len(cands), repr(float(thr)))


FITS5_CASES = {
    # name: (frame factory, candidate filter or None) -- every BASELINE config's spot density has reference fits
    "seed0": (lambda: synth.synth_frame(0), None),                                  # configs[0]/[1]
    "seed3": (lambda: synth.synth_frame(3), None),                                  # second seed
    "dense1000": (lambda: synth.synth_frame(11, n_spots=1000), None),               # configs[2]/[4] density
    # configs[3]: the 2048x2048 / 20 000-spot frame; every candidate inside a 320x320 region (the frame is
    # regenerated by the tests from the seed and checked against img_sha, 8 MB is too large to commit)
    "d2048": (lambda: synth.synth_frame(4, H=2048, W=2048, n_spots=20000),
              lambda h, w: 800 <= h < 1120 and 800 <= w < 1120),
}


def make_fits5(procs, name="seed0"):
    pflib, _, _ = ref()
    factory, keep = FITS5_CASES[name]
    img = factory()
    all_cands = pflib._psf_candidates(img)
    cands = [c for c in all_cands if keep is None or keep(*c)]
    subs = [img[h - 2:h + 3, w - 2:w + 3].astype(np.int64) for h, w in cands]
    t = time.time()
    with multiprocessing.Pool(procs) as pool:
        rows = pool.map(_fit5, subs, chunksize=16)
    dt = time.time() - t
    names = ["ref_params", "ref_perror", "ref_status", "ref_niter", "ref_nfev", "ref_fnorm",
             "n_qrsolv", "n_reject", "clean_params", "clean_status", "clean_niter", "clean_nfev",
             "clean_fnorm", "clean_n_qrsolv", "oracle_equals_ref", "r_2", "rmse", "s_n"]
    d = _stack(rows, names)
    # the reference's full pipeline on the same frame (consolidation + re-key) -- one more
    # pass would cost 500 s, so rebuild it from the per-candidate reference answers with the
    # restated consolidation (checked equal to the reference's on a sub-frame below).  For a region
    # sample the consolidation runs over the sampled candidates only (the tests do the same).
    bins = {}
    for (h, w), row in zip(cands, rows):
        p = row[0]
        if row[15] < 0.7:
            continue
        bins[(h, w)] = (p[2] + h - 2.5, p[3] + w - 2.5, p[0], p[1], p[4], p[5], p[6], None, None,
                        row[16], row[15], row[17])
    final = po.consolidate(bins, img.shape, 4)
    keys = np.array(sorted(final.keys()), dtype=np.int32).reshape(-1, 2)
    np.savez_compressed(os.path.join(GOLD, "fits5_%s.npz" % name), img_sha=sha(img),
                        cands=np.array(cands, dtype=np.int32), n_cands_frame=len(all_cands),
                        final_keys=keys,
                        final_h0=np.array([final[tuple(k)][0] for k in keys]),
                        final_w0=np.array([final[tuple(k)][1] for k in keys]),
                        ref_seconds_total=dt, procs=procs, **d)
    st, cnt = np.unique(d["ref_status"], return_counts=True)
    print("fits5_%s: %d fits in %.1f s on %d procs; status %s; oracle==ref %d/%d; robust(n_qrsolv==0) %d; "
          "accepted %d; final %d" % (name, len(rows), dt, procs, dict(zip(st.tolist(), cnt.tolist())),
                                     int(d["oracle_equals_ref"].sum()), len(rows),
                                     int((d["n_qrsolv"] == 0).sum()), len(bins), len(final)), flush=True)


def make_pipeline_small():
    """Reference find_peptides end to end on a small crop (bounded cost) -- pins consolidation."""
    pflib, _, _ = ref()
    img = synth.synth_frame(4, H=96, W=96, n_spots=30)
    t = time.time()
    out = pflib.find_peptides(img)
    dt = time.time() - t
    mine = po.find_peptides(img, faithful=True)
    assert sorted(out.keys()) == sorted(mine.keys())
    for k in out:
        for a, b in zip(out[k], mine[k]):
            assert np.array_equal(np.asarray(a), np.asarray(b)), k
    keys = np.array(sorted(out.keys()), dtype=np.int32).reshape(-1, 2)
    vals = np.array([[float(v) for v in out[tuple(k)][:7]] + [float(out[tuple(k)][i]) for i in (9, 10, 11)]
                     for k in keys])
    np.savez_compressed(os.path.join(GOLD, "pipeline_small.npz"), img=img, keys=keys, vals=vals,
                        sub_imgs=np.array([out[tuple(k)][7] for k in keys]),
                        fit_imgs=np.array([out[tuple(k)][8] for k in keys]))
    print("pipeline_small: %d psfs in %.1f s; oracle find_peptides identical" % (len(keys), dt))


def _fits11_windows(name):
    if name == "seed0":                       # configs[0]: windows around isolated spots
        img, cr, cc, amp = synth.synth_frame_with_truth(0)
        return synth.cut_windows(img, cr, cc, 11)[:200]
    if name == "d2048":                       # configs[3]: windows cut from the dense 2048^2 / 20 000-spot frame
        img, cr, cc, amp = synth.synth_frame_with_truth(4, H=2048, W=2048, n_spots=20000)
        return synth.cut_windows(img, cr, cc, 11)[:400]
    raise KeyError(name)


def make_fits11(procs, name="seed0"):
    wins = _fits11_windows(name)
    with multiprocessing.Pool(procs) as pool:
        rows = pool.map(_fit11, list(wins), chunksize=8)
    names = ["ref_params", "ref_perror", "ref_status", "ref_niter", "ref_nfev", "ref_fnorm",
             "n_qrsolv", "n_reject", "clean_params", "clean_status", "clean_niter", "clean_nfev",
             "clean_fnorm", "oracle_equals_ref", "p0"]
    d = _stack(rows, names)
    np.savez_compressed(os.path.join(GOLD, "fits11_%s.npz" % name), windows=wins, **d)
    st, cnt = np.unique(d["ref_status"], return_counts=True)
    print("fits11_%s: %d; status %s; oracle==ref %d; robust %d" % (
        name, len(rows), dict(zip(st.tolist(), cnt.tolist())), int(d["oracle_equals_ref"].sum()),
        int((d["n_qrsolv"] == 0).sum())), flush=True)


# ----------------------------------------------------------------------------- the rest of gaussfit's surface
def variant_cases():
    """name -> (window side, gaussfit keyword arguments as a function of the window)"""
    F = lambda *idx: np.array([i in idx for i in range(7)])          # noqa: E731  fresh `fixed` array per call
    pf = dict(limitedmin=[True] * 7, limitedmax=[False, False, True, True, True, True, True],
              minpars=[0, 0, 2, 2, .75, .75, 0], maxpars=[0, 0, 3, 3, 2, 2, 360])
    return {
        "default11": (11, lambda w: dict(fixed=F())),
        "circle11": (11, lambda w: dict(circle=1, fixed=F())),
        "norotate11": (11, lambda w: dict(rotate=0, fixed=F())),
        "noheight11": (11, lambda w: dict(vheight=0, fixed=F())),
        "fixed_theta11": (11, lambda w: dict(fixed=F(6))),
        "fixed_centre11": (11, lambda w: dict(fixed=F(2, 3), params=[np.median(w), w.max() - np.median(w), 5., 5., 1.5, 1.5, 0.])),
        "err11": (11, lambda w: dict(err=np.sqrt(np.maximum(w, 1.0)), fixed=F())),
        "circle_noheight_err11": (11, lambda w: dict(circle=1, vheight=0, err=np.sqrt(np.maximum(w, 1.0)), fixed=F())),
        "pflib_fixed_theta5": (5, lambda w: dict(params=[np.median(w), w.max(), 2.5, 2.5, 1, 1, 0], fixed=F(6), **pf)),
        "pflib_circle5": (5, lambda w: dict(params=[np.median(w), w.max(), 2.5, 2.5, 1], circle=1, fixed=F(), **pf)),
    }


def _variant_fit(args):
    name, win = args
    _, gaussfitter, _ = ref()
    side, kwf = variant_cases()[name]
    kw = kwf(win)
    mp = gaussfitter.gaussfit(win, returnmp=True, **kw)
    m = len(mp.params)
    P = np.full(7, np.nan)
    P[:m] = mp.params
    E = np.full(7, np.nan)
    C = np.full((7, 7), np.nan)
    if mp.perror is not None:
        E[:m] = mp.perror
        C[:m, :m] = mp.covar
    return P, E, C, mp.status, mp.niter, mp.nfev, mp.fnorm, mp.dof, m


def make_gaussfit_variants(procs):
    """tests/golden/gaussfit_variants.npz: the reference's own gaussfit (oracle/_ref) with circle / rotate=0 / vheight=0
    / fixed / err on 11x11 and 5x5 windows: params, perror, covar, status, niter, nfev, fnorm, dof per window."""
    img, cr, cc, amp = synth.synth_frame_with_truth(0)
    w11 = synth.cut_windows(img, cr, cc, 11)[:48]
    w5 = synth.cut_windows(img, cr, cc, 5)[:48]
    out = {"w11": w11, "w5": w5}
    with multiprocessing.Pool(procs) as pool:
        for name, (side, _) in variant_cases().items():
            wins = w11 if side == 11 else w5
            rows = pool.map(_variant_fit, [(name, w) for w in wins], chunksize=4)
            for i, key in enumerate(("params", "perror", "covar", "status", "niter", "nfev", "fnorm", "dof", "npar")):
                out["%s_%s" % (name, key)] = np.array([r[i] for r in rows])
            st, cnt = np.unique(out[name + "_status"], return_counts=True)
            print("variant %s: status %s" % (name, dict(zip(st.tolist(), cnt.tolist()))), flush=True)
    np.savez_compressed(os.path.join(GOLD, "gaussfit_variants.npz"), **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    todo = a.only.split(",") if a.only else ["kat1", "detect", "pipeline_small", "fits11", "fits5"]
    if "kat1" in todo:
        make_kat1()
    if "detect" in todo:
        make_detect()
    if "pipeline_small" in todo:
        make_pipeline_small()
    if "variants" in todo:
        make_gaussfit_variants(a.procs)
    for t_ in todo:                            # fits11[:case], fits5[:case]
        kind, _, case = t_.partition(":")
        if kind == "fits11":
            make_fits11(a.procs, case or "seed0")
        if kind == "fits5":
            make_fits5(a.procs, case or "seed0")


# ----------------------------------------------------------------------------- phase_correlate (reference run)
PHASE_CASES = [(1, 256, 256, 0.0, 0.0, 150), (2, 256, 256, 2.0, -3.0, 150), (3, 256, 256, 0.35, -1.6, 150),
               (4, 200, 333, -1.25, 0.8, 200), (5, 512, 512, 0.45, 0.15, 500), (6, 128, 96, 3.7, 2.2, 60)]


def make_phase_correlate_golden(path):
    """tests/golden/phase_correlate.npz: outputs of the reference's own phase_correlate (oracle/_ref, built from
    /root/reference/phase_correlate.py) on synth.shifted_pair inputs, upsample_factor 1 and 20."""
    import numpy as np
    from oracle import build_ref
    from fluorosequencingimageanalysis_b200 import synth
    m = build_ref.load_phase_correlate()
    out1, out20 = [], []
    for seed, H, W, dy, dx, n in PHASE_CASES:
        a, b = synth.shifted_pair(seed, H, W, dy, dx, n)
        out1.append([float(np.real(v)) for v in m.phase_correlate(a, b, 1)])
        out20.append([float(np.real(v)) for v in m.phase_correlate(a, b, 20)])
    np.savez_compressed(path, cases=np.array(PHASE_CASES, dtype=float), out1=np.array(out1), out20=np.array(out20))
