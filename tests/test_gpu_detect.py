"""GPU parity: candidate detection (fsq_detect, called through the C-ABI by engine.detect_batch)
against the reference-generated goldens and the CPU oracle.  Integer path: BIT-EXACT candidate
lists; the float64 threshold within 1e-13 relative (the kernel derives it from exact integer
moments, numpy from a float64 pairwise sum)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, detect_kwargs
from oracle import pflib_oracle as po

pytestmark = pytest.mark.gpu

THR_RTOL = 1e-13


def _engine():
    from fluorosequencingimageanalysis_b200 import engine
    return engine


def _cands(det):
    return det.cand_hw[:det.total].cpu().numpy()


def _want(img, **kw):
    return np.array(po.psf_candidates(img, **kw), dtype=np.int32).reshape(-1, 2)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "detect_*.npz"))),
                         ids=lambda p: os.path.basename(p)[7:-4])
def test_detect_equals_reference_golden(path):
    d = np.load(path)
    det = _engine().detect_batch(d["img"], **detect_kwargs(d))
    assert det.total == len(d["cands"])
    assert np.array_equal(_cands(det), d["cands"])
    assert float(det.thr[0].item()) == pytest.approx(float(d["thr"]), rel=THR_RTOL, abs=0.0)


def test_correlation_map_bit_exact():
    from fluorosequencingimageanalysis_b200 import synth
    eng = _engine()
    img = synth.synth_frame(21, H=200, W=333, n_spots=150)
    det = eng.detect_batch(img, keep_scratch=True)
    cm32 = eng.detect_cm32(det)[0]
    _, cm, thr = po.detect_maps(img)
    assert np.array_equal(cm32.astype(np.int64), np.minimum(cm, 0xFFFFFFFF))
    assert float(det.thr[0].item()) == pytest.approx(thr, rel=THR_RTOL)


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.int32, np.int64, np.uint16])
def test_pixel_dtypes(dtype):
    from fluorosequencingimageanalysis_b200 import synth
    img = synth.synth_frame(31, H=96, W=130, n_spots=25)
    if dtype == np.uint8:
        img = (img >> 4).clip(0, 255)
    img = img.astype(dtype)
    det = _engine().detect_batch(img)
    assert np.array_equal(_cands(det), _want(img))


@pytest.mark.parametrize("mf,k,c_std", [(1, 5, 2), (2, 5, 2), (3, 3, 1.5), (4, 5, 2), (6, 7, 2.5), (7, 9, 2),
                                        (8, 1, 0.5), (9, 9, 3)])
def test_non_default_parameters(mf, k, c_std):
    from fluorosequencingimageanalysis_b200 import synth
    rng = np.random.default_rng(mf * 100 + k)
    img = synth.synth_frame(40 + mf, H=120, W=97, n_spots=40)
    K = rng.integers(-3000, 9000, (k, k)).astype(np.int64)
    det = _engine().detect_batch(img, median_filter_size=mf, correlation_matrix=K, c_std=c_std)
    assert np.array_equal(_cands(det), _want(img, median_filter_size=mf, correlation_matrix=K, c_std=c_std))


@pytest.mark.parametrize("shape", [(1, 1), (4, 40), (40, 4), (5, 5), (6, 7), (33, 65), (64, 64), (95, 129)])
def test_edge_shapes(shape):
    """Frames smaller than the halo / the excluded border, and ragged tiles."""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(300, 5000, shape).astype(np.uint16)
    det = _engine().detect_batch(img)
    want = _want(img)
    assert det.total == len(want)
    assert np.array_equal(_cands(det), want)


@pytest.mark.parametrize("shape", [(96, 192), (96, 196), (96, 198), (104, 264), (72, 136)])
def test_interior_tile_staging_paths(shape):
    """Interior raw tiles are staged three ways by the packed pass-A kernel, chosen per tile: bulk-async row copies (frame
    base 16-byte aligned, W a multiple of 8), 64-bit loads (W a multiple of 4), reflected index-by-index loads (anything
    else and every border tile).  Widths 192 / 196 / 198 take one path each in frames that have interior tiles."""
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    imgs = rng.integers(300, 5000, (3,) + shape).astype(np.uint16)
    det = _engine().detect_batch(imgs)
    want = [_want(im) for im in imgs]
    got = _cands(det)
    off = 0
    for w in want:
        assert np.array_equal(got[off:off + len(w)], w)
        off += len(w)
    assert off == det.total


def test_flat_and_zero_frames():
    """std == 0: threshold == mean(cm) == 0, every interior pixel is kept (NOT '<', pflib.py:254)."""
    for v in (0, 400):
        img = np.full((40, 50), v, dtype=np.uint16)
        det = _engine().detect_batch(img)
        want = _want(img)
        assert len(want) == 36 * 46
        assert np.array_equal(_cands(det), want)
        assert float(det.thr[0].item()) == 0.0


def test_full_16bit_range_is_exact():
    rng = np.random.default_rng(8)
    img = rng.integers(0, 65536, (128, 128)).astype(np.uint16)
    img[40:45, 60:65] = 65535
    det = _engine().detect_batch(img)
    assert np.array_equal(_cands(det), _want(img))


def test_batch_equals_single_frames_and_capacity_protocol():
    from fluorosequencingimageanalysis_b200 import synth
    eng = _engine()
    frames = np.stack([synth.synth_frame(s, H=160, W=192, n_spots=60) for s in (1, 2, 3, 4, 5)])
    frames[3] = 400                                            # a flat frame inside the batch
    det = eng.detect_batch(frames, cap=64)                     # forces the re-launch path
    per = det.per_frame()
    counts = det.n_cand.cpu().numpy()
    assert counts[-1] == det.total == sum(len(p) for p in per)
    fr = det.cand_frame[:det.total].cpu().numpy()
    assert np.all(np.diff(fr) >= 0)                            # frames in order
    for f in range(len(frames)):
        want = _want(frames[f])
        assert np.array_equal(per[f], want)
        assert counts[f] == len(want)
        single = eng.detect_batch(frames[f])
        assert float(single.thr[0].item()) == float(det.thr[f].item())


def test_config4_frame_2048_high_density():
    """BASELINE.json configs[3] frame shape at full size against the numpy oracle."""
    from fluorosequencingimageanalysis_b200 import synth
    img = synth.synth_frame(77, H=2048, W=2048, n_spots=20000)
    det = _engine().detect_batch(img)
    want = _want(img)
    assert det.total == len(want) > 100000
    assert np.array_equal(_cands(det), want)


def test_pflib_psf_candidates_signature_and_layout():
    """Drop-in: list of (int, int) tuples, raster order (pflib.py:252-257); KAT-2 count."""
    from fluorosequencingimageanalysis_b200 import pflib, synth
    img = synth.synth_frame(0)
    out = pflib._psf_candidates(img)
    assert isinstance(out, list) and len(out) == 4988
    assert all(type(h) is int and type(w) is int for h, w in out[:10])
    assert out == sorted(out)
    assert out == [tuple(c) for c in golden("detect_c1_seed0.npz")["cands"].tolist()]
    with pytest.raises(ValueError, match="odd"):
        pflib._psf_candidates(img, correlation_matrix=np.ones((4, 4), dtype=int))
    out3 = pflib._psf_candidates(img, median_filter_size=3, c_std=3)
    assert out3 == po.psf_candidates(img, median_filter_size=3, c_std=3)
