"""
TEST INFRASTRUCTURE -- CPU oracle, not product code.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module.

Restatement of the reference's per-image spot path on the CPU:

  detection         pflib.py:217-258   (_psf_candidates)         -> psf_candidates / detect_maps
  Gaussian model    agpy/gaussfitter.py:63-140 (twodgaussian)    -> gauss2d
  moments start     agpy/gaussfitter.py:29-61                    -> moments
  fit wrapper       agpy/gaussfitter.py:142-255 (gaussfit)       -> gaussfit
  5x5 fit driver    pflib.py:180-214   (_fit_2d_gaussian)        -> fit_2d_gaussian
  S/N               pflib.py:261-281   (illumina_s_n)            -> illumina_s_n
  per-frame driver  pflib.py:284-520   (find_peptides)           -> find_peptides (+ packed form)
  photometry        flexlibrary.py:123-241, 264-284              -> photometry_*

The detection oracle is the closed form of SURVEY.md App. A written with numpy only (no
scipy), so it is independent of the scipy routines the reference calls; the pin test checks
it against the reference itself (oracle/_ref) and against KAT-2.

Parity pin: see oracle/lm_oracle.py header.
photometry_* are pinned against the reference's own Spot.photometry (oracle/_ref/flexlibrary_head.py, built from
flexlibrary.py:1-1319) by tests/test_oracle_pins.py.
"""
import math

import numpy as np

from . import lm_oracle

DEFAULT_CORRELATION_MATRIX = np.array(            # pflib.py:48-52
    [[-5935, -5935, -5935, -5935, -5935],
     [-5935, 8027, 8027, 8027, -5935],
     [-5935, 8027, 30742, 8027, -5935],
     [-5935, 8027, 8027, 8027, -5935],
     [-5935, -5935, -5935, -5935, -5935]], dtype=np.int64)


def py2_round(x):
    """Python-2 round(): half away from zero (pflib.py:515)."""
    x = float(x)
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


# ----------------------------------------------------------------------------- detection
def _windows(arr, size, pad_mode):
    """All size*size shifted views of arr (padded), stacked on axis 0.
    scipy.ndimage.median_filter(size=s, origin=0) centres the footprint so that an even
    size puts the extra sample on the low-index side: offsets -(s//2) .. s-1-(s//2)."""
    lo = size // 2
    hi = size - 1 - lo
    if pad_mode == "symmetric":
        p = np.pad(arr, ((lo, hi), (lo, hi)), mode="symmetric")
    else:
        p = np.pad(arr, ((lo, hi), (lo, hi)), mode="constant", constant_values=0)
    H, W = arr.shape
    return np.stack([p[i:i + H, j:j + W] for i in range(size) for j in range(size)])


def detect_maps(image, median_filter_size=5, correlation_matrix=DEFAULT_CORRELATION_MATRIX,
                c_std=2):
    """Returns (mf, cm, thr): background-removed image, clamped correlation map (int64) and
    the float64 threshold mean+c_std*std (pflib.py:241-250)."""
    K = np.asarray(correlation_matrix)
    if K.shape[0] != K.shape[1] or K.shape[0] % 2 == 0:           # pflib.py:236-239
        raise ValueError("correlation_matrix must be square, with an odd "
                         "number of rows and columns")
    im = np.asarray(image).astype(np.int64)                       # :241
    s = int(median_filter_size)
    win = _windows(im, s, "symmetric")                            # scipy mode='reflect'
    n = s * s
    med = np.partition(win, n // 2, axis=0)[n // 2]               # rank n//2 (0-based)
    mf = im - np.minimum(med, im)                                 # :243-245
    k = K.shape[0]
    cw = _windows(mf, k, "zero")                                  # correlate(mode='same'), zero pad
    cm = np.tensordot(K.astype(np.int64).ravel(), cw, axes=(0, 0))
    cm = np.maximum(cm, 0).astype(np.int64)                       # :247-248
    thr = np.mean(cm) + c_std * np.std(cm)                        # :250
    return mf, cm, thr


def psf_candidates(image, median_filter_size=5, correlation_matrix=DEFAULT_CORRELATION_MATRIX,
                   c_std=2, **kwargs):
    """pflib.py:217-258 -> list of (h, w) int tuples, raster order, 2-px border excluded."""
    _, cm, thr = detect_maps(image, median_filter_size, correlation_matrix, c_std)
    H, W = cm.shape
    if H <= 4 or W <= 4:
        return []
    inner = cm[2:H - 2, 2:W - 2]
    hh, ww = np.nonzero(~(inner < thr))                           # kept when NOT '<' (:254)
    return [(int(h) + 2, int(w) + 2) for h, w in zip(hh, ww)]


# ----------------------------------------------------------------------------- model
def gauss2d(p, shape):
    """agpy/gaussfitter.py:93-138 with circle=0, rotate=1, vheight=1.
    p = (height, amplitude, p2, p3, width_x, width_y, rota_deg); NOTE the reference pops the
    third parameter as center_y and the fourth as center_x (:100) and evaluates on
    numpy.indices (x = row index, y = col index) -- SURVEY.md section 0 fact 5."""
    height = float(p[0])
    amplitude = float(p[1])
    center_y = float(p[2])
    center_x = float(p[3])
    width_x = float(p[4])
    width_y = float(p[5])
    rota = np.pi / 180. * float(p[6])
    rcen_x = center_x * np.cos(rota) - center_y * np.sin(rota)
    rcen_y = center_x * np.sin(rota) + center_y * np.cos(rota)
    x, y = np.indices(shape)
    xp = x * np.cos(rota) - y * np.sin(rota)
    yp = x * np.sin(rota) + y * np.cos(rota)
    g = height + amplitude * np.exp(
        -(((rcen_x - xp) / width_x) ** 2 +
          ((rcen_y - yp) / width_y) ** 2) / 2.)
    return g


def moments(data):
    """agpy/gaussfitter.py:29-61 (circle=0, rotate=1, vheight=1, estimator=numpy.ma.median).
    Returns [height, amplitude, x, y, width_x, width_y, 0.]"""
    total = np.abs(data).sum()
    Y, X = np.indices(data.shape)
    y = np.argmax((X * np.abs(data)).sum(axis=1) / total)
    x = np.argmax((Y * np.abs(data)).sum(axis=0) / total)
    col = data[int(y), :]
    width_x = np.sqrt(np.abs((np.arange(col.size) - y) * col).sum() / np.abs(col).sum())
    row = data[:, int(x)]
    width_y = np.sqrt(np.abs((np.arange(row.size) - x) * row).sum() / np.abs(row).sum())
    height = np.ma.median(data.ravel())
    amplitude = data.max() - height
    if np.isnan(width_y) or np.isnan(width_x) or np.isnan(height) or np.isnan(amplitude):
        raise ValueError("something is nan")
    return [height, amplitude, x, y, width_x, width_y, 0.]


GAUSSFIT_DEFAULT_LIMITEDMIN = [False, False, False, False, True, True, True]     # gaussfitter.py:143
GAUSSFIT_DEFAULT_LIMITEDMAX = [False, False, False, False, False, False, True]   # :144
GAUSSFIT_DEFAULT_MINPARS = [0, 0, 0, 0, 0, 0, 0]                                 # :146
GAUSSFIT_DEFAULT_MAXPARS = [0, 0, 0, 0, 0, 0, 360]                               # :146


def gaussfit(data, params=(), limitedmin=GAUSSFIT_DEFAULT_LIMITEDMIN,
             limitedmax=GAUSSFIT_DEFAULT_LIMITEDMAX, minpars=GAUSSFIT_DEFAULT_MINPARS,
             maxpars=GAUSSFIT_DEFAULT_MAXPARS, faithful=True, trace=None):
    """agpy/gaussfitter.py:142-255 for the 7-parameter model, err=None.
    Returns (LMResult, fit_image)."""
    data = np.asarray(data)
    params = np.array(params, dtype='float')                      # :189
    if len(params) == 0:
        params = np.array(moments(data), dtype='float')           # :193-194 (a list there)
    for i in range(len(params)):                                  # :202-204
        if params[i] > maxpars[i] and limitedmax[i]:
            params[i] = maxpars[i]
        if params[i] < minpars[i] and limitedmin[i]:
            params[i] = minpars[i]

    def resid(p):                                                 # :214-215
        return np.ravel(data - gauss2d(p, data.shape))

    res = lm_oracle.lm_solve(resid, params, np.asarray(limitedmin, dtype=bool),
                             np.asarray(limitedmax, dtype=bool),
                             np.asarray(minpars, dtype=float), np.asarray(maxpars, dtype=float),
                             faithful=faithful, trace=trace)
    fitimage = gauss2d(res.params, data.shape)                    # :253
    return res, fitimage


def pflib_fit_args(subimage):
    """The start values and limits pflib hard-codes (pflib.py:199-213)."""
    params = (np.median(subimage), np.amax(subimage), 2.5, 2.5, 1, 1, 0)
    limitedmin = [True] * 7
    limitedmax = [False, False, True, True, True, True, True]
    minpars = np.array([0.00, (np.amax(subimage) - np.mean(subimage)) / 3.0,
                        2.00, 2.00, 0.75, 0.75, 0.00])
    maxpars = np.array([0.00, 0.00, 3.00, 3.00, 2.00, 2.00, 360.00])
    return params, limitedmin, limitedmax, minpars, maxpars


def fit_2d_gaussian(subimage, faithful=True, return_result=False, trace=None):
    """pflib.py:180-214 -> (h_0, w_0, H, A, sigma_h, sigma_w, theta, fit_img)."""
    assert subimage.shape[0] == 5 and subimage.shape[1] == 5       # :193
    params, lmin, lmax, mn, mx = pflib_fit_args(subimage)
    res, fit_img = gaussfit(subimage, params=params, limitedmin=lmin, limitedmax=lmax,
                            minpars=mn, maxpars=mx, faithful=faithful, trace=trace)
    H, A, h_0, w_0, sigma_h, sigma_w, theta = res.params
    out = (h_0, w_0, H, A, sigma_h, sigma_w, theta, fit_img)
    if return_result:
        return out, res
    return out


# ----------------------------------------------------------------------------- metrics
def illumina_s_n(sub_img):
    """pflib.py:261-281."""
    if not (len(sub_img.shape) == 2 and sub_img.shape[0] == sub_img.shape[1]):
        raise ValueError("sub_img must be square, but has shape " + str(sub_img))
    n = sub_img.shape[0]
    edge = ([sub_img[h, w] for h in [0, -1] for w in range(n)] +
            [sub_img[h, w] for h in range(1, n - 1) for w in [0, -1]])
    return (np.amax(sub_img) - np.mean(edge)) / np.std(edge)


def fit_metrics(sub_img, fit_img):
    """pflib.py:463-473 -> (r_2, rmse, s_n); Python sums, sequential order."""
    r_2 = (1.0 - sum(np.reshape((sub_img - fit_img) ** 2, -1)) /
           sum((np.reshape(sub_img, -1) - np.mean(sub_img)) ** 2))
    rmse = math.sqrt(sum([(sub_img[x, y] - fit_img[x, y]) ** 2
                          for x in range(5) for y in range(5)]) / 25.0)
    return r_2, rmse, illumina_s_n(sub_img)


# ----------------------------------------------------------------------------- frame driver
def consolidate(psfs, shape, consolidation_radius=4):
    """pflib.py:479-519 on a dict {(h,w): tuple with [0]=h_0,[1]=w_0,[10]=r_2}, iterated in
    insertion (= raster) order -- SURVEY.md App. C on the py2 dict-order caveat."""
    bins = dict(psfs)
    for (h, w), psf in list(bins.items()):
        if (h, w) not in bins:
            continue
        h_lo, h_hi = max(0, h - consolidation_radius - 2), min(h + consolidation_radius + 3, shape[0])
        w_lo, w_hi = max(0, w - consolidation_radius - 2), min(w + consolidation_radius + 3, shape[1])
        dead = False
        for h_d in range(h_lo, h_hi):
            for w_d in range(w_lo, w_hi):
                if h_d == h and w_d == w:
                    continue
                if (h_d, w_d) not in bins:
                    continue
                h_0, w_0 = bins[(h, w)][:2]
                h_0_d, w_0_d = bins[(h_d, w_d)][:2]
                if (h_0 - h_0_d) ** 2 + (w_0 - w_0_d) ** 2 > consolidation_radius ** 2:
                    continue
                if bins[(h, w)][10] > bins[(h_d, w_d)][10]:        # :508 strict '>'
                    del bins[(h_d, w_d)]
                else:
                    del bins[(h, w)]
                    dead = True
                    break
            if dead:
                break
    for (h, w), psf in list(bins.items()):                         # :514-519 re-key
        h_0_r, w_0_r = int(py2_round(psf[0])), int(py2_round(psf[1]))
        if h_0_r != h or w_0_r != w:
            del bins[(h, w)]
            assert (h_0_r, w_0_r) not in bins
            bins.setdefault((h_0_r, w_0_r), psf)
    return bins


def find_peptides(image, median_filter_size=5, correlation_matrix=DEFAULT_CORRELATION_MATRIX,
                  c_std=2, r_2_threshold=0.7, consolidation_radius=4, faithful=True,
                  candidate_subset=None, return_all_fits=False):
    """pflib.py:284-520 (fit_type='gauss').  ``candidate_subset`` (indices into the candidate
    list) lets tests bound the cost; ``return_all_fits`` also returns the per-candidate
    records before the R^2 gate."""
    if consolidation_radius < 2:
        raise ValueError("consolidation_radius must be at least 2")
    image = np.asarray(image)
    cands = psf_candidates(image, median_filter_size, correlation_matrix, c_std)
    if candidate_subset is not None:
        cands = [cands[i] for i in candidate_subset]
    bins = {}
    allfits = []
    for h, w in cands:
        sub_img = image[h - 2:h + 3, w - 2:w + 3].astype(np.int64)          # :443
        (h_0, w_0, H, A, sigma_h, sigma_w, theta, fit_img), res = \
            fit_2d_gaussian(sub_img, faithful=faithful, return_result=True)
        h_0, w_0 = h_0 + h - 2.5, w_0 + w - 2.5                             # :461
        r_2, rmse, s_n = fit_metrics(sub_img, fit_img)
        if return_all_fits:
            allfits.append(dict(h=h, w=w, params=res.params.copy(), status=res.status,
                                niter=res.niter, nfev=res.nfev, fnorm=res.fnorm,
                                n_qrsolv=res.n_qrsolv, r_2=r_2, rmse=rmse, s_n=s_n))
        if r_2 < r_2_threshold:
            continue
        bins.setdefault((h, w), (h_0, w_0, H, A, sigma_h, sigma_w, theta, sub_img, fit_img,
                                 rmse, r_2, s_n))
    out = consolidate(bins, image.shape, consolidation_radius)
    if return_all_fits:
        return out, allfits
    return out


# ----------------------------------------------------------------------------- photometry
def image_slice(image, h, w, radius):
    """flexlibrary.py:123-147 (border-truncated slice)."""
    return image[max(0, h - radius):min(image.shape[0], h + radius + 1),
                 max(0, w - radius):min(image.shape[1], w + radius + 1)]


def photometry_simple(image, h, w, size=5):
    """flexlibrary.py:160-170."""
    return np.sum(image_slice(image, h, w, (size - 1) // 2))


def photometry_mexican_hat(image, h, w, brim_size=6, radius=9):
    """flexlibrary.py:172-210: sum(crown) - len(crown)*median(brim), slice-local indices."""
    diameter = 2 * radius + 1
    sl = image_slice(image, h, w, radius)
    crown, brim = [], []
    for (hh, ww), p in np.ndenumerate(sl):
        if brim_size <= hh < diameter - brim_size and brim_size <= ww < diameter - brim_size:
            crown.append(p)
        else:
            brim.append(p)
    return sum(int(c) for c in crown) - len(crown) * np.median(brim)


def photometry_gaussian_volume(A, sigma_h, sigma_w, scaling=10 ** 6):
    """flexlibrary.py:212-230."""
    return float(scaling) * A * sigma_h * sigma_w


def photometry_maximum(image, h, w, radius=5, top=1):
    """flexlibrary.py:264-284 (background_adjust='none')."""
    r = np.sort(np.ravel(image_slice(image, h, w, radius)))
    return float(np.sum(r[-top:]))


def consolidate_packed(cand_hw, fit, shape, r_2_threshold=0.7, consolidation_radius=4):
    """R^2 gate + rival consolidation + re-key of pflib.py:466-468, 479-519 on packed arrays of
    ONE frame.  cand_hw [n,2] raster order, fit [n,>=10] (h_0,w_0,...,r_2 at column 8).
    Returns (keys [m,2] int, idx [m] indices into the candidate arrays) in dict-insertion order.
    numpy restatement of the dictionary logic in `consolidate` above (pinned to it by tests/test_host_logic.py);
    the checker of the device kernel fsq_consolidate."""
    if consolidation_radius < 2:
        raise ValueError("consolidation_radius must be at least 2")          # pflib.py:431-432
    H, W = shape
    n = cand_hw.shape[0]
    r2 = fit[:, 8]
    keep = ~(r2 < r_2_threshold)                                              # :466 discards only when '<'
    grid = -np.ones((H, W), dtype=np.int64)
    order = np.nonzero(keep)[0]
    grid[cand_hw[order, 0], cand_hw[order, 1]] = order
    alive = keep.copy()
    h0 = fit[:, 0]
    w0 = fit[:, 1]
    rad = consolidation_radius
    rr = rad ** 2
    for i in order:
        if not alive[i]:
            continue
        h, w = int(cand_hw[i, 0]), int(cand_hw[i, 1])
        sl = grid[max(0, h - rad - 2):min(h + rad + 3, H), max(0, w - rad - 2):min(w + rad + 3, W)]
        js = sl[sl >= 0]                                                      # raster order of (h_d, w_d)
        for j in js:
            if j == i or not alive[j]:
                continue
            if (h0[i] - h0[j]) ** 2 + (w0[i] - w0[j]) ** 2 > rr:
                continue
            if r2[i] > r2[j]:                                                 # :508
                alive[j] = False
                grid[cand_hw[j, 0], cand_hw[j, 1]] = -1
            else:
                alive[i] = False
                grid[h, w] = -1
                break
    idx = np.nonzero(alive)[0]
    # re-key (:514-519): delete + setdefault moves an entry to the END of the dict
    bins = {}
    for i in idx:
        bins[(int(cand_hw[i, 0]), int(cand_hw[i, 1]))] = int(i)
    for (h, w), i in list(bins.items()):
        k = (int(py2_round(h0[i])), int(py2_round(w0[i])))
        if k[0] != h or k[1] != w:
            del bins[(h, w)]
            assert k not in bins                                              # :518
            bins.setdefault(k, i)
    keys = np.array(list(bins.keys()), dtype=np.int64).reshape(-1, 2)
    return keys, np.array(list(bins.values()), dtype=np.int64)
