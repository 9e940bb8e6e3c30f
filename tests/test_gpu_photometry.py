"""GPU parity: fsq_photometry (flexlibrary.Spot.photometry family, flexlibrary.py:160-210, 264-284) against the
oracle's restatement -- interior spots, spots whose slice is truncated by the image border (the reference then
uses slice-LOCAL indices for the crown/brim split), all pixel dtypes.  Exact: integer sums, float64 median."""
import numpy as np
import pytest

from oracle import pflib_oracle as po

pytestmark = pytest.mark.gpu


def test_photometry_matches_oracle_including_truncated_slices():
    from fluorosequencingimageanalysis_b200 import engine, synth
    rng = np.random.default_rng(5)
    frames = synth.synth_timetrace(3, n_frames=3, H=96, W=80, n_spots=40)
    n = 300
    hw = np.stack([rng.integers(0, 96, n), rng.integers(0, 80, n)], axis=1).astype(np.int32)
    hw[:8] = [[0, 0], [95, 79], [0, 40], [50, 0], [95, 3], [4, 79], [9, 9], [86, 70]]     # corners, edges, first fully-inside
    fr = rng.integers(0, 3, n).astype(np.int32)
    for dtype in (np.uint16, np.int32, np.uint8, np.int16):
        f = (frames % 251).astype(dtype) if dtype == np.uint8 else frames.astype(dtype)
        got = engine.photometry_batch(f, hw, fr, method="mexican_hat").cpu().numpy()
        want = np.array([po.photometry_mexican_hat(f[k].astype(np.int64), h, w) for (h, w), k in zip(hw, fr)], dtype=float)
        assert np.array_equal(got, want), dtype
        got = engine.photometry_batch(f, hw, fr, method="mexican_hat", radius=5, brim_size=2).cpu().numpy()
        want = np.array([po.photometry_mexican_hat(f[k].astype(np.int64), h, w, brim_size=2, radius=5) for (h, w), k in zip(hw, fr)], dtype=float)
        assert np.array_equal(got, want), dtype
        got = engine.photometry_batch(f, hw, fr, method="simple").cpu().numpy()
        want = np.array([po.photometry_simple(f[k].astype(np.int64), h, w) for (h, w), k in zip(hw, fr)], dtype=float)
        assert np.array_equal(got, want), dtype
        got = engine.photometry_batch(f, hw, fr, method="maximum").cpu().numpy()
        want = np.array([po.photometry_maximum(f[k].astype(np.int64), h, w) for (h, w), k in zip(hw, fr)], dtype=float)
        assert np.array_equal(got, want), dtype
    with pytest.raises(ValueError):
        engine.photometry_batch(frames, hw, fr, method="sextractor")          # flexlibrary.py:315 (photutils path: out of scope)
    fit = np.abs(rng.normal(1.0, 0.3, (10, 12)))
    assert np.array_equal(engine.photometry_from_fit(fit, "gaussian_volume"),
                          np.array([po.photometry_gaussian_volume(r[3], r[4], r[5]) for r in fit]))
    assert np.array_equal(engine.photometry_from_fit(fit, "sigmas"), 1e6 * fit[:, 4] * fit[:, 5])
