"""CPU: the oracle (oracle/) against the golden vectors generated from the reference itself
(oracle/make_golden.py -> tests/golden/), and against the reference (oracle/_ref) directly when
/root/reference is present in this container.  SURVEY.md App. D KAT-1 / KAT-2."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, detect_kwargs
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
