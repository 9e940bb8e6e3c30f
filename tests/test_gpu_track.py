"""GPU parity: fsq_track_centroid (Experiment.luminosity_centroid_particle_tracking, flexlibrary.py:1173-1317)
against the oracle's restatement (oracle/track_oracle.py, pinned to the reference's own functions by
tests/test_oracle_pins.py).
Positions and states exact; the Illumina S/N bit for bit (integer sums, numpy's summation order for the std)."""
import numpy as np
import pytest

from oracle import track_oracle as tro
from oracle import pflib_oracle as po

pytestmark = pytest.mark.gpu


def _movie(seed, F=12, H=96, W=80, n_spots=40, bleach=True):
    from fluorosequencingimageanalysis_b200 import synth
    rng, cr, cc, amp = synth.spot_layout(seed, H, W, n_spots)
    out = np.empty((F, H, W), dtype=np.uint16)
    off_at = rng.integers(2, F, n_spots) if bleach else np.full(n_spots, F + 1)     # photobleaching step per spot
    for f in range(F):
        on = off_at > f
        clean = synth.render_clean(cr[on], cc[on], amp[on], H, W, 1.5, 400.0)
        out[f] = synth.add_noise(clean, np.random.default_rng(5000 + 31 * seed + f))
    return out, cr, cc


def test_centroid_tracking_matches_oracle():
    from fluorosequencingimageanalysis_b200 import engine
    rng = np.random.default_rng(9)
    for seed, dtype in ((1, np.uint16), (2, np.int32)):
        frames, cr, cc = _movie(seed)
        F, H, W = frames.shape
        spots = [(int(round(a)), int(round(b))) for a, b in zip(cr, cc)]
        # spots hugging the border (window cut -> None; square cut -> None) and spots on empty background
        spots += [(2, 2), (3, 3), (H - 3, W - 3), (H - 4, 5), (2, 40), (50, W - 3), (5, 5), (60, 60)]
        spots += [(int(h), int(w)) for h, w in zip(rng.integers(2, H - 2, 30), rng.integers(2, W - 2, 30))]
        fr = frames.astype(dtype)
        for offsets in (None, np.stack([np.arange(F) % 3 - 1, (np.arange(F) // 2) % 3 - 1], axis=1)):
            for size, radius, cutoff in ((5, 3, 3.0), (3, 2, 4.0), (5, 4, 0.0)):
                hw, st, sn = engine.track_centroid_batch(fr, spots, offsets=offsets, size=size, search_radius=radius, s_n_cutoff=cutoff)
                hw, st, sn = hw.cpu().numpy(), st.cpu().numpy(), sn.cpu().numpy()
                whw, wst, wsn = tro.track(fr, spots, size=size, search_radius=radius, s_n_cutoff=cutoff, offsets=offsets)
                assert np.array_equal(st, wst), (seed, size, radius)
                assert np.array_equal(hw, whw), (seed, size, radius)
                ok = ~np.isnan(wsn)
                assert np.array_equal(np.isnan(sn[:, 1:]), np.isnan(wsn[:, 1:]))
                assert np.array_equal(sn[:, 1:][ok[:, 1:]], wsn[:, 1:][ok[:, 1:]]), "S/N must match bit for bit"
                if size == 5 and offsets is None and cutoff == 3.0:
                    assert (st == engine.TRACK_STAYED).sum() > 20 and (st == engine.TRACK_NONE).sum() > 10 \
                        and (st == engine.TRACK_CENTROID).sum() > 200          # every branch is exercised
                    # frame 0: the S/N of the initial square, as Spot.illumina_s_n (flexlibrary.py:319-320) gives it
                    for i in (0, 5, 17):
                        h, w = spots[i]
                        assert sn[i, 0] == po.illumina_s_n(fr[0][h - 2:h + 3, w - 2:w + 3])
    with pytest.raises(AttributeError):
        engine.track_centroid_batch(fr, spots, size=4)                          # flexlibrary.py:98-99


def test_tracking_batches_fields_and_timetrace_path():
    """Several fields in one call equal the per-field calls; the time-trace path (frame 0 peak-fitted, PSFs followed
    by the luminosity centroid, photometry per frame) returns consistent arrays."""
    from fluorosequencingimageanalysis_b200 import engine
    m1, cr1, cc1 = _movie(3, F=8)
    m2, cr2, cc2 = _movie(4, F=8)
    s1 = [(int(round(a)), int(round(b))) for a, b in zip(cr1, cc1)]
    s2 = [(int(round(a)), int(round(b))) for a, b in zip(cr2, cc2)]
    both = np.stack([m1, m2])
    hw, st, sn = engine.track_centroid_batch(both, s1 + s2, spot_field=[0] * len(s1) + [1] * len(s2))
    a = engine.track_centroid_batch(m1, s1)
    b = engine.track_centroid_batch(m2, s2)
    assert np.array_equal(hw.cpu().numpy(), np.concatenate([a[0].cpu().numpy(), b[0].cpu().numpy()]))
    assert np.array_equal(st.cpu().numpy(), np.concatenate([a[1].cpu().numpy(), b[1].cpu().numpy()]))
    tt = engine.timetrace_batch(m1)
    m = len(tt["psf_int"])
    assert m > 20 and tt["track_hw"].shape == (m, 8, 2) and tt["photometry"].shape == (m, 8)
    assert np.array_equal(tt["track_hw"][:, 0], tt["psf_int"][:, 1:3])          # traces start at the frame-0 PSF keys
    whw, wst, _ = tro.track(m1, [tuple(k) for k in tt["psf_int"][:, 1:3].tolist()])
    assert np.array_equal(tt["track_hw"], whw) and np.array_equal(tt["track_state"], wst)
    live = tt["track_state"] != engine.TRACK_NONE
    assert np.array_equal(np.isnan(tt["photometry"]), ~live)
    i, f = np.argwhere(live)[7]
    h, w = tt["track_hw"][i, f]
    assert tt["photometry"][i, f] == po.photometry_mexican_hat(m1[f].astype(np.int64), h, w)


# ------------------------------------------------------------------------------------ greedy cross-frame tracking
def _cycles(seed, F=6, H=96, W=96, n=120, p_off=0.25, jitter=1):
    """spot lists of F frames of one field: a fixed population observed with +-jitter px noise and drop-outs,
    plus a few spurious spots per frame; no two spots of a frame closer than 2 px (the reference's precondition)"""
    rng = np.random.default_rng(seed)
    base = np.stack([rng.integers(4, H - 4, n), rng.integers(4, W - 4, n)], axis=1)
    frames = []
    for f in range(F):
        sel = rng.uniform(size=n) > p_off
        pts = base[sel] + rng.integers(-jitter, jitter + 1, (int(sel.sum()), 2))
        pts = np.concatenate([pts, np.stack([rng.integers(4, H - 4, 8), rng.integers(4, W - 4, 8)], axis=1)])
        keep, seen = [], set()
        for h, w in pts.tolist():
            if all(abs(h - a) > 1 or abs(w - b) > 1 for a, b in seen):
                seen.add((h, w)); keep.append((h, w))
        frames.append(keep)
    return frames


def _traces_as_array(traces, F):
    return np.array([[-1 if v is None else v for v in t] for t in traces], dtype=np.int64).reshape(-1, F)


def test_greedy_tracking_matches_oracle():
    """fsq_track_greedy against the oracle's restatement of Experiment.greedy_particle_tracking (pinned to the
    reference's own function by tests/test_oracle_pins.py): identical traces in identical order -- ties in distance, skipped frames, drift offsets
    with sub-pixel parts, drop-outs at the border, several fields per launch."""
    from fluorosequencingimageanalysis_b200 import engine
    shape = (96, 96)
    offs = [(0, 0), (0.35, -1.6), (-0.35, 1.6), (2.0, 0.0), (-2.0, 0.5), (0.0, -0.5)]
    fields = [_cycles(s) for s in (1, 2, 3)]
    for radius in (2, 3):
        for offsets in (None, offs):
            got = engine.track_greedy_batch(fields, shape, candidate_radius=radius, offsets=offsets, spot_radius=0)
            for fld, (tr, nd) in zip(fields, got):
                want, wnd = tro.greedy_particle_tracking(fld, shape, candidate_radius=radius, offsets=offsets)
                assert nd == wnd
                assert np.array_equal(tr, _traces_as_array(want, 6)), (radius, offsets is None)
            if offsets is None and radius == 2:
                tr = got[0][0]
                assert ((tr >= 0).sum(axis=1) >= 4).sum() > 30                 # long traces exist
                gaps = [(t >= 0) for t in tr]
                assert any(g[0] and not g[1] and g[2:].any() for g in gaps)    # ... and traces that skip a frame
    one = engine.track_greedy_batch(fields[0], shape, spot_radius=3)
    want, wnd = tro.greedy_particle_tracking(fields[0], shape, spot_radius=3)
    assert one[1] == wnd and np.array_equal(one[0], _traces_as_array(want, 6))
    with pytest.raises(ValueError):
        engine.track_greedy_batch(fields[0], shape, offsets=[(1, 0)] * 6)       # flexlibrary.py:582-584
    with pytest.raises(AssertionError):
        engine.track_greedy_batch([[(10, 10), (10.2, 10.1)], [(10, 10)]], shape)   # :853-858
