"""
Drop-in mirror of the reference's ``pflib`` hot-path functions (same names, argument meaning,
return layouts and error behaviour), backed by the CUDA kernels of libfsq.so.

  _psf_candidates   pflib.py:217-258      find_peptides   pflib.py:284-520
  _fit_2d_gaussian  pflib.py:180-214      illumina_s_n    pflib.py:261-281
  image_batch / parallel_image_batch, save_psfs_*, read_image  pflib.py:523-746, 883-1111 -> psfio.py

``import fluorosequencingimageanalysis_b200.pflib as pflib`` (or the sys.modules alias shown
in INTEGRATION.md) lets flexlibrary / the basic_*_script drivers run unchanged.
"""
import math

import numpy as np

from . import engine
from .engine import DEFAULT_CORRELATION_MATRIX

default_correlation_matrix = DEFAULT_CORRELATION_MATRIX.copy()      # pflib.py:48-52

#: set to False to get clean-MINPACK fits instead of reproducing the reference's qrsolv
#: diagonal-view behaviour
FAITHFUL = True

#: which LM kernel the drop-in entry points run (fsq.h FSQ_SOLVER_*):
#:   "minpack" -- the reference's algorithm operation for operation (with FAITHFUL: its qrsolv defect included).
#:                Measured against reference-generated fits (tests/test_gpu_parity_table.py, four golden sets): equal
#:                to the reference on 0.94 of the fits whose reference trajectory is a clean one, 0.76 of the fits the
#:                reference accepts, 0.60 of all candidates, 0.88-0.90 of the final PSF keys -- the reference itself
#:                reproduces 0.85 / 0.69 / 0.54 of its own answers when exp() is perturbed by one ulp (the pflib call
#:                starts theta on its bound with equal widths, so the first step is rounding noise)
#:   "fast"    -- the production fitter (analytic Jacobian, normal equations; ~600x faster): 0.98 / 0.59 / 0.35 and
#:                0.79-0.84 of the final keys on the same sets; equal to the reference wherever the reference's answer
#:                is reproducible and its trajectory clean, the clean algorithm's constrained minimum elsewhere
SOLVER = "minpack"


def _py2_round(x):
    """Python-2 round(): half away from zero (pflib.py:515)."""
    x = float(x)
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def _psf_candidates(image, median_filter_size=5, correlation_matrix=default_correlation_matrix,
                    c_std=2, **kwargs):
    """pflib.py:217-258 -> [(h, w), ...] raster order."""
    image = np.asarray(image)
    if image.ndim != 2:
        raise ValueError("image must be two-dimensional")
    det = engine.detect_batch(image, median_filter_size, correlation_matrix, c_std)
    hw = det.cand_hw[:det.total].cpu().numpy()
    return [(int(h), int(w)) for h, w in hw]


def _pflib_limits(sub):
    """Start values / limits of pflib.py:199-213 for n windows [n,5,5] (host-side argument
    marshalling only; the kernel clamps like gaussfitter.py:202-204)."""
    n = sub.shape[0]
    flat = sub.reshape(n, 25)
    mx = flat.max(axis=1).astype(np.float64)
    p0 = np.tile(np.array([0, 0, 2.5, 2.5, 1, 1, 0], dtype=np.float64), (n, 1))
    p0[:, 0] = np.median(flat, axis=1)
    p0[:, 1] = mx
    lo = np.tile(np.array([0.0, 0.0, 2.0, 2.0, 0.75, 0.75, 0.0]), (n, 1))
    lo[:, 1] = (mx - flat.mean(axis=1)) / 3.0
    hi = np.tile(np.array([0.0, 0.0, 3.0, 3.0, 2.0, 2.0, 360.0]), (n, 1))
    lim_lo = np.ones((n, 7), dtype=np.uint8)
    lim_hi = np.tile(np.array([0, 0, 1, 1, 1, 1, 1], dtype=np.uint8), (n, 1))
    # gaussfitter.py:202-204 start clamp
    p0 = np.where((p0 > hi) & (lim_hi != 0), hi, p0)
    p0 = np.where((p0 < lo) & (lim_lo != 0), lo, p0)
    return p0, lo, hi, lim_lo, lim_hi


def _fit_2d_gaussian(subimage, implementation='agpy'):
    """pflib.py:180-214 -> (h_0, w_0, H, A, sigma_h, sigma_w, theta, fit_img)."""
    subimage = np.asarray(subimage)
    assert subimage.shape[0] == 5 and subimage.shape[1] == 5
    if implementation != 'agpy':
        raise NotImplementedError("Currently, only agpy is supported.")
    sub = subimage.astype(np.int64)[None] if subimage.dtype.kind in "iub" else subimage.astype(np.float64)[None]
    p0, lo, hi, lim_lo, lim_hi = _pflib_limits(sub)
    r = engine.gaussfit_batch(sub, p0, lo, hi, lim_lo, lim_hi, faithful=FAITHFUL, want_fit_img=True,
                              solver=SOLVER)
    H, A, h_0, w_0, sigma_h, sigma_w, theta = (float(v) for v in r.params[0].cpu().numpy())
    return (h_0, w_0, H, A, sigma_h, sigma_w, theta, r.fit_img[0].cpu().numpy())


def illumina_s_n(sub_img):
    """pflib.py:261-281."""
    sub_img = np.asarray(sub_img)
    if not (len(sub_img.shape) == 2 and sub_img.shape[0] == sub_img.shape[1]):
        raise ValueError("sub_img must be square, but has shape " + str(sub_img))
    if sub_img.shape[0] > 33:
        raise NotImplementedError("fsq_illumina_s_n handles windows up to 33x33 (numpy's un-split pairwise sum)")
    return float(engine.illumina_s_n_batch(sub_img[None])[0].item())


def consolidate_packed(cand_hw, fit, shape, r_2_threshold=0.7, consolidation_radius=4):
    """R^2 gate + rival consolidation + re-key of pflib.py:466-468, 479-519 on packed arrays of
    ONE frame.  cand_hw [n,2] raster order, fit [n,>=10] (h_0,w_0,...,r_2 at column 8).
    Returns (keys [m,2] int, idx [m] indices into the candidate arrays) in dict-insertion order.
    Runs on the device (fsq_consolidate); raises the reference's AssertionError (:518) on a key collision."""
    if consolidation_radius < 2:
        raise ValueError("consolidation_radius must be at least 2")          # pflib.py:431-432
    import torch
    engine.require_cuda()
    cand_hw = np.asarray(cand_hw)
    fit = np.asarray(fit, dtype=np.float64)
    n = cand_hw.shape[0]
    if n == 0:
        return np.zeros((0, 2), dtype=np.int64), np.zeros(0, dtype=np.int64)
    f12 = np.zeros((n, 12))
    f12[:, :min(fit.shape[1], 12)] = fit[:, :12]
    dev = torch.device("cuda", torch.cuda.current_device())
    c = engine.consolidate_batch(torch.from_numpy(np.ascontiguousarray(cand_hw.astype(np.int32))).to(dev),
                                 torch.zeros(n, dtype=torch.int32, device=dev),
                                 torch.from_numpy(np.ascontiguousarray(f12)).to(dev), n, 1,
                                 r_2_threshold, consolidation_radius)
    c.check()
    idx = engine.psf_dict_order(c.state.cpu().numpy()[:n])
    return c.key.cpu().numpy()[idx].astype(np.int64).reshape(-1, 2), idx.astype(np.int64)


def find_peptides(image, median_filter_size=5, correlation_matrix=default_correlation_matrix,
                  candidate_pixels=None, c_std=2, r_2_threshold=0.7, consolidation_radius=4,
                  fit_type='gauss', N_iter=10 ** 3):
    """pflib.py:284-520 -> {(h, w): (h_0, w_0, H, A, sigma_h, sigma_w, theta, sub_img, fit_img,
    rmse, r_2, s_n)} with sub_img 5x5 int64 and fit_img 5x5 float64."""
    if consolidation_radius < 2:
        raise ValueError("consolidation_radius must be at least 2")
    if fit_type != 'gauss':
        raise NotImplementedError("fit_type='monte_carlo' (pflib.py:117-177) is not part of the CUDA hot path")
    image = np.asarray(image)
    res = engine.find_peptides_batch(image, median_filter_size, correlation_matrix, c_std,
                                     faithful=FAITHFUL, want_fit_img=True, solver=SOLVER)
    return psfs_from_packed(image, res.cand_hw, res.fit, res.fit_img, r_2_threshold, consolidation_radius)


def psfs_from_packed(image, cand_hw, fit, fit_img, r_2_threshold=0.7, consolidation_radius=4):
    """Packed per-candidate arrays of ONE frame -> the reference's dictionary
    {(h, w): (h_0, w_0, H, A, sigma_h, sigma_w, theta, sub_img, fit_img, rmse, r_2, s_n)}
    (R^2 gate, consolidation, re-key: pflib.py:466-468, 479-519; tuple layout :475-477)."""
    keys, idx = consolidate_packed(cand_hw, fit, image.shape, r_2_threshold, consolidation_radius)
    out = {}
    for (kh, kw), i in zip(keys, idx):
        h, w = int(cand_hw[i, 0]), int(cand_hw[i, 1])
        f = fit[i]
        sub_img = image[h - 2:h + 3, w - 2:w + 3].astype(np.int64)            # pflib.py:443
        out[(int(kh), int(kw))] = (float(f[0]), float(f[1]), float(f[2]), float(f[3]), float(f[4]),
                                   float(f[5]), float(f[6]), sub_img, fit_img[i].reshape(5, 5).copy(),
                                   float(f[7]), float(f[8]), float(f[9]))
    return out


# file-level callers and PSF result files (pflib.py:523-746, 883-1111)
from .psfio import (_epoch_to_hash, _hash_to_epoch, _psfs_filename, save_psfs_pkl, save_psfs_csv,   # noqa: E402,F401
                    save_psfs_png, read_image, convert_image, image_batch, parallel_image_batch)
