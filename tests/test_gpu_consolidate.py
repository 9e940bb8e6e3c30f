"""GPU parity: fsq_consolidate (R^2 gate + rival consolidation + re-key, pflib.py:466-468, 479-519) against the
oracle's dictionary logic -- bit-exact keys, survivors and dictionary order, including ties in R^2, exact .5
centres (python-2 rounding), the reference's key-collision assert, and frame boundaries inside a batch."""
import numpy as np
import pytest

from conftest import golden
from oracle import pflib_oracle as po
from test_host_logic import _packed_from_golden

pytestmark = pytest.mark.gpu


def _mods():
    from fluorosequencingimageanalysis_b200 import engine, pflib, synth
    return engine, pflib, synth


def test_consolidate_reproduces_reference_psf_keys(fits5):
    """Fed with the REFERENCE's own per-candidate fits of the config-1 frame, the device kernel returns exactly
    the reference's final PSF dictionary (keys, members, order)."""
    engine, pflib, _ = _mods()
    cands, fit = _packed_from_golden(fits5)
    keys, idx = pflib.consolidate_packed(cands, fit, (512, 512))
    wk, wi = po.consolidate_packed(cands, fit, (512, 512))
    assert np.array_equal(keys, wk) and np.array_equal(idx, wi)
    order = np.lexsort((keys[:, 1], keys[:, 0]))                  # the golden stores the keys sorted
    assert np.array_equal(keys[order], fits5["final_keys"])
    assert np.array_equal(fit[idx[order], 0], fits5["final_h0"])


def test_consolidate_random_ties_half_pixels_and_collisions():
    engine, pflib, _ = _mods()
    rng = np.random.default_rng(3)
    n_assert = 0
    for trial in range(40):
        H = W = 60
        n = int(rng.integers(20, 400))
        hw = np.unique(np.stack([rng.integers(2, H - 2, n), rng.integers(2, W - 2, n)], axis=1), axis=0)
        n = len(hw)
        fit = np.zeros((n, 12))
        spread = rng.choice([0.5, 1.5, 3.0])                            # centres up to `spread` px off the pixel
        fit[:, 0] = hw[:, 0] + rng.uniform(-spread, spread, n).round(1)  # many exact .5 ties
        fit[:, 1] = hw[:, 1] + rng.uniform(-spread, spread, n).round(1)
        fit[:, 8] = rng.choice([0.5, 0.71, 0.8, 0.9, 0.95, np.nan], n, p=[.2, .2, .2, .2, .19, .01])   # ties in r_2, NaN kept (:466)
        radius = int(rng.integers(2, 6))
        try:
            wk, wi = po.consolidate_packed(hw, fit, (H, W), 0.7, radius)
        except AssertionError:
            n_assert += 1
            if spread == 0.5:            # the device detects collisions within the rival reach (fsq.h)
                with pytest.raises(AssertionError):
                    pflib.consolidate_packed(hw, fit, (H, W), 0.7, radius)
            continue
        keys, idx = pflib.consolidate_packed(hw, fit, (H, W), 0.7, radius)
        assert np.array_equal(idx, wi), (trial, radius)
        assert np.array_equal(keys, wk), (trial, radius)
    print("collision-assert trials: %d" % n_assert)
    with pytest.raises(ValueError):
        pflib.consolidate_packed(hw, fit, (H, W), 0.7, 1)               # pflib.py:431-432
    k0, i0 = pflib.consolidate_packed(np.zeros((0, 2), dtype=np.int32), np.zeros((0, 12)), (H, W))
    assert k0.shape == (0, 2) and i0.shape == (0,)


def test_consolidate_batch_equals_frame_by_frame(fits5):
    """A whole stack in one call (device-resident count, capacity > n): frames never interact, per-frame results
    equal the one-frame calls; n_psf counts the final PSFs; find_peptides' dictionary is built from it."""
    import torch
    engine, pflib, synth = _mods()
    frames = synth.synth_timetrace(7, n_frames=3, H=128, W=128, n_spots=40)
    res = engine.find_peptides_batch(frames, faithful=False, solver="fast", to_host=False)
    n = int(res.cand_hw.shape[0])
    cap = n + 1000
    hw = torch.zeros((cap, 2), dtype=torch.int32, device="cuda"); hw[:n] = res.cand_hw
    fr = torch.zeros(cap, dtype=torch.int32, device="cuda"); fr[:n] = res.cand_frame
    fit = torch.zeros((cap, 12), dtype=torch.float64, device="cuda"); fit[:n] = res.fit
    n_dev = torch.tensor([n], dtype=torch.int64, device="cuda")
    c = engine.consolidate_batch(hw, fr, fit, cap, 3, n_dev=n_dev)
    c.check()
    state = c.state.cpu().numpy()[:n]
    key = c.key.cpu().numpy()[:n]
    fr_h, hw_h, fit_h = res.cand_frame.cpu().numpy(), res.cand_hw.cpu().numpy(), res.fit.cpu().numpy()
    order = engine.psf_dict_order(state, fr_h)
    n_psf = c.n_psf.cpu().numpy()
    off = 0
    for f in range(3):
        sel = np.nonzero(fr_h == f)[0]
        wk, wi = po.consolidate_packed(hw_h[sel], fit_h[sel], (128, 128))
        got = order[off:off + len(wi)]
        assert np.array_equal(got, sel[wi])
        assert np.array_equal(key[got], wk)
        assert n_psf[f] == len(wi) and len(wi) > 10
        off += len(wi)
    assert off == len(order)
    gated = ~(fit_h[:, 8] >= 0.7)
    assert np.array_equal(state == engine.PSF_GATED, gated)


def test_packed_psfs_are_the_dictionary_in_order():
    """fsq_pack_psfs == psf_dict_order applied on the host; the streaming pipeline with fetch="psfs" delivers the
    same records, and they are what pflib.find_peptides puts in its dictionary (same solver)."""
    import torch
    engine, pflib, synth = _mods()
    frames = synth.synth_timetrace(11, n_frames=4, H=160, W=128, n_spots=60)
    res = engine.find_peptides_batch(frames, faithful=False, solver="fast", to_host=False)
    n = int(res.cand_hw.shape[0])
    c = engine.consolidate_batch(res.cand_hw, res.cand_frame, res.fit, n, 4)
    pk = engine.pack_psfs_batch(c, res.cand_frame, res.fit, n, 4)
    base = pk.base.cpu().numpy()
    m = int(base[4])
    order = engine.psf_dict_order(c.state.cpu().numpy(), res.cand_frame.cpu().numpy())
    ints, fit = pk.ints.cpu().numpy()[:m], pk.fit.cpu().numpy()[:m]
    assert m == len(order) and m > 100
    assert np.array_equal(ints[:, 3], order)
    assert np.array_equal(ints[:, 0], res.cand_frame.cpu().numpy()[order])
    assert np.array_equal(ints[:, 1:3], c.key.cpu().numpy()[order])
    assert np.array_equal(fit.view(np.int64), res.fit.cpu().numpy()[order].view(np.int64))
    assert np.array_equal(np.diff(base), c.n_psf.cpu().numpy())
    # capacity protocol: rows beyond cap_psf are dropped, the total still tells
    small = engine.pack_psfs_batch(c, res.cand_frame, res.fit, n, 4, cap_psf=10)
    assert int(small.base[4].item()) == m
    assert np.array_equal(small.ints.cpu().numpy()[:, 3], order[:10])
    # streaming pipeline, final PSFs only on the host
    fs = engine.FieldStream(4, 160, 128, dtype=torch.uint16, depth=2, host_io=True, fetch="psfs", faithful=False, solver="fast")
    host = torch.from_numpy(frames.view(np.int16)).view(torch.uint16).pin_memory()
    for rep in range(3):
        t = fs.submit(host)
        nn, mm, pint, pfit, pbase = fs.end_fetch(t)
        assert nn == n and mm == m
        assert np.array_equal(pint.numpy(), ints) and np.array_equal(pfit.numpy().view(np.int64), fit.view(np.int64))
        assert np.array_equal(pbase.numpy(), base)
    # the drop-in dictionary of frame 2 holds exactly these records, in this order
    old = pflib.SOLVER, pflib.FAITHFUL
    pflib.SOLVER, pflib.FAITHFUL = "fast", False
    try:
        d = pflib.find_peptides(frames[2])
    finally:
        pflib.SOLVER, pflib.FAITHFUL = old
    a, b = base[2], base[3]
    assert [tuple(k) for k in ints[a:b, 1:3].tolist()] == list(d.keys())
    assert np.array_equal(np.array([v[0] for v in d.values()]), fit[a:b, 0])
    assert np.array_equal(np.array([v[10] for v in d.values()]), fit[a:b, 8])
