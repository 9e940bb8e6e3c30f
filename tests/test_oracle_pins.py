"""CPU: the oracle (oracle/) against the golden vectors generated from the reference itself
(oracle/make_golden.py -> tests/golden/), and against the reference (oracle/_ref) directly when
/root/reference is present in this container.  SURVEY.md App. D KAT-1 / KAT-2."""
import glob
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden, detect_kwargs
from oracle import pflib_oracle as po
from oracle import build_ref

HAVE_REF = os.path.isdir("/root/reference")


def test_kat1_known_answer():
    k = golden("kat1_fit.npz")
    # the literal numbers of SURVEY.md App. D
    assert k["sub"].tolist() == [[115, 135, 289, 319, 227], [164, 415, 950, 1039, 607],
                                 [199, 698, 1696, 1901, 1061], [180, 642, 1442, 1685, 903],
                                 [120, 255, 617, 719, 392]]
    (h0, w0, H, A, sh, sw, th, fit_img), res = po.fit_2d_gaussian(k["sub"], faithful=True, return_result=True)
    got = np.array([h0, w0, H, A, sh, sw, th])
    want = np.array([2.6810246650240313, 2.3069781155524147, 83.85864237125972, 1988.188217197878,
                     1.12742146986278, 1.1262983268602713, 0.0])
    assert np.allclose(got, want, rtol=1e-9, atol=0)
    assert np.array_equal(got, k["fit7"])                    # bit-for-bit with the reference run
    assert (res.status, res.niter, res.nfev) == (1, 6, 42)
    assert res.fnorm == pytest.approx(7247.585538787288, rel=1e-9)
    assert np.allclose(res.perror, k["perror"], rtol=1e-9)
    assert np.allclose(res.covar, k["covar"], rtol=1e-8, atol=1e-300)
    assert po.illumina_s_n(k["sub"]) == pytest.approx(5.236681248076466, rel=1e-12)
    r2, rmse, sn = po.fit_metrics(k["sub"], fit_img)
    assert r2 == pytest.approx(0.9989676657001576, rel=1e-10)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "detect_*.npz"))),
                         ids=lambda p: os.path.basename(p)[7:-4])
def test_detection_oracle_vs_reference_golden(path):
    d = np.load(path)
    kw = detect_kwargs(d)
    cands = po.psf_candidates(d["img"], **kw)
    assert np.array_equal(np.array(cands, dtype=np.int32).reshape(-1, 2), d["cands"])
    _, cm, thr = po.detect_maps(d["img"], **kw)
    assert thr == float(d["thr"])
    assert int(cm.sum()) == int(d["cm_sum"]) and int(cm.max()) == int(d["cm_max"])


def test_kat2_counts():
    assert len(golden("detect_c1_seed0.npz")["cands"]) == 4988
    assert float(golden("detect_c1_seed0.npz")["thr"]) == 12744327.49987743
    assert len(golden("detect_c1_seed3.npz")["cands"]) == 5036
    assert float(golden("detect_c1_seed3.npz")["thr"]) == 12943142.349889603


def test_even_median_size_matches_scipy():
    from scipy.ndimage import median_filter
    from fluorosequencingimageanalysis_b200 import synth
    img = synth.synth_frame(12, H=40, W=52, n_spots=10).astype(np.int64)
    for s in (2, 4, 6):
        win = po._windows(img, s, "symmetric")
        med = np.partition(win, (s * s) // 2, axis=0)[(s * s) // 2]
        assert np.array_equal(med, median_filter(img, s))


def test_fits5_oracle_vs_reference_golden(fits5, frame0):
    """Both oracle flavours reproduce the stored reference / clean answers bit-for-bit on a
    seeded sample of the 4988 candidates (the whole set was checked when the golden was made:
    oracle_equals_ref is all True)."""
    assert bool(fits5["oracle_equals_ref"].all())
    rng = np.random.default_rng(7)
    for i in rng.choice(len(fits5["cands"]), 24, replace=False):
        h, w = fits5["cands"][i]
        sub = frame0[h - 2:h + 3, w - 2:w + 3].astype(np.int64)
        _, res = po.fit_2d_gaussian(sub, faithful=True, return_result=True)
        assert np.array_equal(res.params, fits5["ref_params"][i])
        assert (res.status, res.niter, res.nfev) == (fits5["ref_status"][i], fits5["ref_niter"][i], fits5["ref_nfev"][i])
        assert res.fnorm == fits5["ref_fnorm"][i]
        assert res.n_qrsolv == fits5["n_qrsolv"][i]
        _, cl = po.fit_2d_gaussian(sub, faithful=False, return_result=True)
        assert np.array_equal(cl.params, fits5["clean_params"][i])
        assert cl.status == fits5["clean_status"][i]


def test_population_statistics(fits5):
    """status mix, acceptance and final PSF count of the config-1 frame."""
    st, cnt = np.unique(fits5["ref_status"], return_counts=True)
    assert dict(zip(st.tolist(), cnt.tolist())) == {1: 2125, 2: 2200, 3: 474, 5: 189}
    assert int((fits5["r_2"] >= 0.7).sum()) == 1957
    assert len(fits5["final_keys"]) == 468
    assert (fits5["clean_status"] == 1).mean() > 0.99
    robust = fits5["n_qrsolv"] == 0
    # robust set == fits where the faithful and the clean solver agree (section 8(c))
    agree = np.max(np.abs(fits5["ref_params"][:, :6] - fits5["clean_params"][:, :6]) /
                   np.maximum(np.abs(fits5["clean_params"][:, :6]), 1e-300), axis=1) < 1e-6
    assert agree[robust].mean() == 1.0
    assert agree[~robust].mean() < 0.1


def test_fits11_oracle_vs_reference_golden():
    g = golden("fits11_seed0.npz")
    assert bool(g["oracle_equals_ref"].all())
    for i in (0, 17, 101):
        res, _ = po.gaussfit(g["windows"][i], faithful=True)
        assert np.array_equal(res.params, g["ref_params"][i])
        assert res.status == g["ref_status"][i] and res.niter == g["ref_niter"][i]


def test_pipeline_small_oracle_vs_reference_golden():
    g = golden("pipeline_small.npz")
    out = po.find_peptides(g["img"], faithful=True)
    assert sorted(out.keys()) == [tuple(k) for k in g["keys"].tolist()]
    for k, v in zip(g["keys"], g["vals"]):
        psf = out[tuple(k)]
        got = np.array([float(x) for x in psf[:7]] + [float(psf[9]), float(psf[10]), float(psf[11])])
        assert np.array_equal(got, v)


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present (GPU box)")
def test_oracle_bitwise_equals_reference_live(frame0, fits5):
    """Runs the REFERENCE ITSELF (oracle/_ref) next to the restated oracle."""
    build_ref.build(quiet=True)
    pflib, gaussfitter, _ = build_ref.load()
    assert pflib._psf_candidates(frame0) == po.psf_candidates(frame0)
    rng = np.random.default_rng(11)
    for i in rng.choice(len(fits5["cands"]), 12, replace=False):
        h, w = fits5["cands"][i]
        sub = frame0[h - 2:h + 3, w - 2:w + 3].astype(np.int64)
        ref = pflib._fit_2d_gaussian(sub)
        mine = po.fit_2d_gaussian(sub, faithful=True)
        for a, b in zip(ref, mine):
            assert np.array_equal(np.asarray(a), np.asarray(b))


def test_phase_correlate_golden_is_the_reference_output():
    """tests/golden/phase_correlate.npz was written by the reference's own phase_correlate.py (oracle/_ref); where
    oracle/_ref exists the run is repeated and must reproduce the file bit for bit."""
    from oracle import build_ref
    from fluorosequencingimageanalysis_b200 import synth
    m = build_ref.load_phase_correlate()
    if m is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    g = golden("phase_correlate.npz")
    for k, (seed, H, W, dy, dx, n) in enumerate(g["cases"]):
        a, b = synth.shifted_pair(int(seed), int(H), int(W), dy, dx, int(n))
        assert [float(np.real(v)) for v in m.phase_correlate(a, b, 1)] == g["out1"][k].tolist()
        assert [float(np.real(v)) for v in m.phase_correlate(a, b, 20)] == g["out20"][k].tolist()
    # sub-pixel drifts come back at the 1/20 px resolution the caller asks for (flexlibrary.py:1717)
    assert g["out20"][2][:2].tolist() == [-0.35, 1.6] and g["out20"][4][:2].tolist() == [-0.45, -0.15]


# ------------------------------------------------------------------ flexlibrary: photometry and the two trackers
def _flex():
    from oracle import build_ref
    fl = build_ref.load_flexlibrary()
    if fl is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    return fl


def test_photometry_oracle_equals_reference_spot_methods():
    """oracle photometry_* == the reference's own Spot.photometry (flexlibrary.py:160-317, run from oracle/_ref),
    interior spots and spots whose radius-9 / radius-5 slices are cut by the border."""
    fl = _flex()
    from fluorosequencingimageanalysis_b200 import synth
    # int64 pixels: under the numpy of the reference's time Python's sum() over uint16 scalars ran in int64 (0 + uint16
    # -> int64); numpy 2 keeps uint16 and wraps, which is not the reference's behaviour
    img = synth.synth_frame(3, H=64, W=72, n_spots=25).astype(np.int64)
    im = fl.Image(image=img)
    rng = np.random.default_rng(1)
    pts = [(2, 2), (61, 69), (2, 40), (30, 2), (61, 5), (9, 9), (40, 40)] + \
          [(int(h), int(w)) for h, w in zip(rng.integers(2, 62, 40), rng.integers(2, 70, 40))]
    for h, w in pts:
        sp = fl.Spot(im, h, w, 5)
        assert sp.photometry(method='mexican_hat') == po.photometry_mexican_hat(img.astype(np.int64), h, w)
        assert sp.photometry(method='mexican_hat', radius=5, brim_size=2) == \
            po.photometry_mexican_hat(img.astype(np.int64), h, w, brim_size=2, radius=5)
        assert sp.photometry(method='simple') == po.photometry_simple(img, h, w)
        assert sp.photometry(method='maximum') == po.photometry_maximum(img, h, w)
        assert sp.illumina_s_n() == po.illumina_s_n(img[h - 2:h + 3, w - 2:w + 3])
    with pytest.raises(AttributeError):
        fl.Spot(im, 1, 10, 5)                                   # the 5x5 square leaves the image (flexlibrary.py:100-118)


def _bleaching_movie(seed, F=8, H=72, W=64, n_spots=25):
    from fluorosequencingimageanalysis_b200 import synth
    rng, cr, cc, amp = synth.spot_layout(seed, H, W, n_spots)
    off_at = rng.integers(2, F, n_spots)
    out = np.empty((F, H, W), dtype=np.uint16)
    for f in range(F):
        on = off_at > f
        out[f] = synth.add_noise(synth.render_clean(cr[on], cc[on], amp[on], H, W, 1.5, 400.0), np.random.default_rng(900 + 17 * seed + f))
    return out, [(int(round(a)), int(round(b))) for a, b in zip(cr, cc)]


def test_centroid_tracking_oracle_equals_reference():
    """oracle track() == Experiment.luminosity_centroid_particle_tracking of the reference itself
    (flexlibrary.py:1173-1317): bleaching spots, border spots, integer drift offsets."""
    fl = _flex()
    from oracle import track_oracle as tro
    for seed in (1, 2):
        mv, spots = _bleaching_movie(seed)
        F, H, W = mv.shape
        spots = spots + [(2, 2), (3, 3), (H - 3, W - 3), (4, 30), (40, W - 3), (6, 6), (50, 50)]
        frames = [fl.Image(image=mv[f]) for f in range(F)]
        for offsets in (None, [(f % 3 - 1, (f // 2) % 3 - 1) for f in range(F)]):
            for radius, cutoff in ((3, 3.0), (2, 5.0)):
                init = [fl.Spot(frames[0], h, w, 5) for h, w in spots]
                ref = fl.Experiment.luminosity_centroid_particle_tracking(frames, init, search_radius=radius,
                                                                          s_n_cutoff=cutoff, offsets=offsets)
                hw, st, _ = tro.track(mv, spots, size=5, search_radius=radius, s_n_cutoff=cutoff, offsets=offsets)
                for i, trace in enumerate(ref):
                    got = [(-1, -1) if s is None else (s.h, s.w) for s in trace]
                    assert got == [tuple(v) for v in hw[i].tolist()], (seed, radius, i)
                if radius == 3:
                    assert (st == 2).any() and (st == 0).any() and (st == 1).any()      # every branch is exercised


def test_greedy_tracking_oracle_equals_reference():
    """oracle greedy_particle_tracking == Experiment.greedy_particle_tracking of the reference itself
    (flexlibrary.py:680-1027): same traces in the same order, same drop-out count -- ties in distance, skipped frames,
    sub-pixel drift offsets.  (The reference cannot run with offsets=None: `for f in len(frame_spots)`, :787.)"""
    fl = _flex()
    from oracle import track_oracle as tro
    shape = (48, 56)
    rng = np.random.default_rng(4)
    base = np.stack([rng.integers(4, 44, 45), rng.integers(4, 52, 45)], axis=1)
    frames = []
    for f in range(5):
        sel = rng.uniform(size=len(base)) > 0.25
        pts = np.concatenate([base[sel] + rng.integers(-1, 2, (int(sel.sum()), 2)), np.stack([rng.integers(4, 44, 4), rng.integers(4, 52, 4)], axis=1)])
        keep = []
        for h, w in pts.tolist():
            if all(abs(h - a) > 1 or abs(w - b) > 1 for a, b in keep):
                keep.append((h, w))
        frames.append(keep)
    dummy = fl.Image(image=np.zeros(shape, dtype=np.uint16))
    for radius in (2, 3):
        for offsets in ([(0, 0)] * 5, [(0, 0), (0.35, -1.6), (-0.35, 1.6), (2.0, 0.0), (-2.0, 0.5)]):
            spots = [[fl.Spot(dummy, h, w, 5) for h, w in fr] for fr in frames]
            ref, nd = fl.Experiment.greedy_particle_tracking(spots, shape, candidate_radius=radius, offsets=offsets, spot_radius=0)
            want, wnd = tro.greedy_particle_tracking(frames, shape, candidate_radius=radius, offsets=offsets, spot_radius=0)
            assert nd == wnd
            as_idx = [[None if s is None else spots[f].index(s) for f, s in enumerate(tr)] for tr in ref]
            assert as_idx == want, (radius, offsets[1])
    assert any(t[0] is not None and t[1] is None and any(v is not None for v in t[2:]) for t in want)      # a trace that skips a frame


@pytest.mark.parametrize("case", ["seed3", "dense1000", "d2048"])
def test_new_fit_goldens_are_pinned_to_the_oracle(case):
    """The round-2 reference-fit goldens (second seed, 1000-spot density, the dense 2048^2 frame): the frame regenerates to
    the stored hash, the candidate list equals the oracle's, and both oracle flavours reproduce the stored reference /
    clean answers bit for bit on a seeded sample (the whole set was checked when the golden was made)."""
    import hashlib
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_parity_table import _frame
    g = golden("fits5_%s.npz" % case)
    img = _frame(case)
    assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == str(g["img_sha"])
    assert bool(g["oracle_equals_ref"].all())
    if case != "d2048":                                  # (the oracle's detection of the 2048^2 frame runs in the GPU suite)
        assert np.array_equal(np.array(po.psf_candidates(img), dtype=np.int32), g["cands"])
    rng = np.random.default_rng(5)
    for i in rng.choice(len(g["cands"]), 10, replace=False):
        h, w = g["cands"][i]
        sub = img[h - 2:h + 3, w - 2:w + 3].astype(np.int64)
        _, res = po.fit_2d_gaussian(sub, faithful=True, return_result=True)
        assert np.array_equal(res.params, g["ref_params"][i])
        assert (res.status, res.niter, res.nfev, res.n_qrsolv) == (g["ref_status"][i], g["ref_niter"][i], g["ref_nfev"][i], g["n_qrsolv"][i])
        _, cl = po.fit_2d_gaussian(sub, faithful=False, return_result=True)
        assert np.array_equal(cl.params, g["clean_params"][i]) and cl.status == g["clean_status"][i]
        r_2, rmse, s_n = po.fit_metrics(sub, po.gauss2d(res.params, (5, 5)))
        assert (r_2, rmse, s_n) == (g["r_2"][i], g["rmse"][i], g["s_n"][i])
    st = golden("stable5_%s.npz" % case)
    assert st["stable_ref"].shape == (len(g["cands"]),) and 0.2 < st["stable_ref"].mean() < 0.5


def test_fits11_dense_golden_is_pinned_to_the_oracle():
    g = golden("fits11_d2048.npz")
    assert bool(g["oracle_equals_ref"].all()) and len(g["windows"]) == 400
    for i in (3, 150, 399):
        res, _ = po.gaussfit(g["windows"][i], faithful=True)
        assert np.array_equal(res.params, g["ref_params"][i])
        assert res.status == g["ref_status"][i] and res.niter == g["ref_niter"][i]
        assert np.array_equal(np.array(po.moments(g["windows"][i]), dtype=float), g["p0"][i])


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present (GPU box)")
def test_gaussfit_variant_golden_is_the_reference_output():
    """tests/golden/gaussfit_variants.npz holds what the reference's own gaussfit returns (spot check, live)."""
    from oracle import make_golden
    build_ref.build(quiet=True)
    _, gaussfitter, _ = build_ref.load()
    g = golden("gaussfit_variants.npz")
    cases = make_golden.variant_cases()
    for name in ("circle11", "fixed_centre11", "circle_noheight_err11", "pflib_fixed_theta5"):
        side, kwf = cases[name]
        w = (g["w11"] if side == 11 else g["w5"])[7]
        mp = gaussfitter.gaussfit(w, returnmp=True, **kwf(w))
        m = int(g[name + "_npar"][7])
        assert len(mp.params) == m and np.array_equal(mp.params, g[name + "_params"][7][:m])
        assert mp.status == g[name + "_status"][7] and mp.nfev == g[name + "_nfev"][7] and mp.dof == g[name + "_dof"][7]
        if mp.perror is not None:
            assert np.array_equal(mp.covar, g[name + "_covar"][7][:m, :m])
