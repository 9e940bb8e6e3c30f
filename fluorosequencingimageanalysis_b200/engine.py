"""
Batched device API above the C-ABI: frames in, packed per-candidate arrays out.

These are the additions SURVEY.md section 8(b) describes (``find_peptides_batch``,
``gaussfit_batch``); the reference-signature single-call functions in ``pflib.py`` /
``gaussfitter.py`` are batch-of-one wrappers around them.  All tensors are torch CUDA tensors
(torch = device memory + streams only); every compute step is a kernel of libfsq.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib

DEFAULT_CORRELATION_MATRIX = np.array(            # pflib.py:48-52
    [[-5935, -5935, -5935, -5935, -5935],
     [-5935, 8027, 8027, 8027, -5935],
     [-5935, 8027, 30742, 8027, -5935],
     [-5935, 8027, 8027, 8027, -5935],
     [-5935, -5935, -5935, -5935, -5935]], dtype=np.int64)

_TORCH_DTYPE_CODE = {torch.uint8: _lib.FSQ_U8, torch.int16: _lib.FSQ_I16, torch.int32: _lib.FSQ_I32,
                     torch.float64: _lib.FSQ_F64, torch.int64: _lib.FSQ_I64}
if hasattr(torch, "uint16"):
    _TORCH_DTYPE_CODE[torch.uint16] = _lib.FSQ_U16

# out_fit columns of fit_candidates (fsq.h): image-coordinate PSF record
COL_H0, COL_W0, COL_H, COL_A, COL_SIGMA_H, COL_SIGMA_W, COL_THETA, COL_RMSE, COL_R2, COL_SN, COL_CHI2 = range(11)
ICOL_STATUS, ICOL_NITER, ICOL_NFEV, ICOL_NQRSOLV = range(4)

MAX_FRAMES_PER_CALL = 65535


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_i32(a, dev):
    """numpy / list / torch (any device) -> contiguous int32 CUDA tensor"""
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.int32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.int32))).to(dev)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("fluorosequencingimageanalysis_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    _lib.load()


def to_device_frames(frames, device=None):
    """Accepts numpy / torch, 2-D or 3-D, any integer dtype the reference would take
    (pflib.py:241 casts to int64); returns a contiguous CUDA tensor [F,H,W] of a supported
    pixel dtype."""
    require_cuda()
    device = device or torch.device("cuda", torch.cuda.current_device())
    if isinstance(frames, np.ndarray):
        a = frames
        if not a.flags.writeable:                      # e.g. arrays decoded by PIL: torch wants a writable buffer
            a = a.copy()
        if a.dtype == np.uint16:
            t = torch.from_numpy(a.view(np.int16)).view(torch.uint16)
        elif a.dtype in (np.uint8, np.int16, np.int32):
            t = torch.from_numpy(a)
        elif a.dtype.kind in "iu":
            if a.size and (a.min() < -2 ** 30 or a.max() >= 2 ** 30):
                raise OverflowError("pixel values beyond +-2^30 are outside the exact range of the kernels")
            t = torch.from_numpy(a.astype(np.int32))
        elif a.dtype.kind == "b":
            t = torch.from_numpy(a.astype(np.uint8))
        else:
            raise TypeError("frames must have an integer dtype (got %s)" % a.dtype)
        t = t.to(device, non_blocking=True)
    else:
        t = frames
        if t.dtype == torch.int64:
            t = t.to(torch.int32)
        if t.dtype not in _TORCH_DTYPE_CODE or t.dtype in (torch.float64,):
            raise TypeError("unsupported frame dtype %s" % t.dtype)
        t = t.to(device)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3:
        raise ValueError("frames must be [H,W] or [F,H,W]")
    return t.contiguous()


def _check_kernel(correlation_matrix):
    K = np.asarray(correlation_matrix)
    if K.ndim != 2 or K.shape[0] != K.shape[1] or K.shape[0] % 2 == 0:      # pflib.py:236-239
        raise ValueError("correlation_matrix must be square, with an odd "
                         "number of rows and columns")
    if K.dtype.kind not in "iub":
        if not np.all(K == np.rint(K)):
            raise ValueError("correlation_matrix must hold integers (the reference correlates in int64)")
    return np.ascontiguousarray(K.astype(np.int64))


class Detection(object):
    """Packed result of detect_batch: device tensors, raster order inside a frame."""
    __slots__ = ("cand_hw", "cand_frame", "n_cand", "thr", "total", "scratch", "shape")

    def per_frame(self):
        """-> list (per frame) of int32 numpy [n,2] arrays (one D2H copy)."""
        hw = self.cand_hw[:self.total].cpu().numpy()
        counts = self.n_cand[:-1].cpu().numpy()
        offs = np.concatenate([[0], np.cumsum(counts)])
        return [hw[offs[i]:offs[i + 1]] for i in range(len(counts))]


def detect_batch(frames, median_filter_size=5, correlation_matrix=DEFAULT_CORRELATION_MATRIX,
                 c_std=2, cap=None, keep_scratch=False):
    """pflib._psf_candidates (pflib.py:217-258) for a batch of frames on the GPU."""
    L = _lib.load()
    frames = to_device_frames(frames)
    F, H, W = frames.shape
    if F > MAX_FRAMES_PER_CALL:
        raise ValueError("at most %d frames per call" % MAX_FRAMES_PER_CALL)
    K = _check_kernel(correlation_matrix)
    dev = frames.device
    code = _TORCH_DTYPE_CODE[frames.dtype]
    sbytes = L.fsq_detect_scratch_bytes(F, H, W)
    scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
    n_cand = torch.empty(F + 1, dtype=torch.int64, device=dev)
    thr = torch.empty(F, dtype=torch.float64, device=dev)
    if cap is None:
        cap = max(4096, int(0.06 * F * H * W))
    Kp = K.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
    while True:
        cand_hw = torch.empty((cap, 2), dtype=torch.int32, device=dev)
        cand_frame = torch.empty(cap, dtype=torch.int32, device=dev)
        rc = L.fsq_detect(_ptr(frames), code, F, H, W, Kp, K.shape[0], int(median_filter_size),
                          float(c_std), _ptr(cand_hw), _ptr(cand_frame), _ptr(n_cand), _ptr(thr),
                          cap, _ptr(scratch), sbytes, _stream())
        _lib.check(rc)
        _lib.check(L.fsq_detect_flags(_ptr(scratch), F, H, W, _stream()))     # synchronises
        total = int(n_cand[F].item())
        if total <= cap:
            break
        cap = total                                                           # FSQ_E_CAPACITY protocol
    d = Detection()
    d.cand_hw, d.cand_frame, d.n_cand, d.thr, d.total = cand_hw, cand_frame, n_cand, thr, total
    d.scratch = scratch if keep_scratch else None
    d.shape = (F, H, W)
    return d


def detect_cm32(det):
    """Clamped correlation map (saturated to uint32) of the last detect_batch(keep_scratch=True)."""
    L = _lib.load()
    F, H, W = det.shape
    out = torch.empty((F, H, W), dtype=torch.int32, device=det.cand_hw.device)
    _lib.check(L.fsq_detect_copy_cm32(_ptr(det.scratch), F, H, W, _ptr(out), _stream()))
    return out.cpu().numpy().view(np.uint32)


def fit_candidates(frames, cand_hw, cand_frame, n, faithful=True, want_fit_img=False, n_dev=None,
                   opts=None, solver="minpack"):
    """pflib.find_peptides' per-candidate loop (pflib.py:441-477) for n candidates.
    Returns (out_fit [n,12] f64, out_int [n,4] i32, fit_img [n,25] f64 | None) on the device."""
    L = _lib.load()
    frames = to_device_frames(frames)
    F, H, W = frames.shape
    dev = frames.device
    o = opts or _lib.default_opts(faithful=faithful, solver=solver)
    out_fit = torch.empty((n, 12), dtype=torch.float64, device=dev)
    out_int = torch.empty((n, 4), dtype=torch.int32, device=dev)
    fit_img = torch.empty((n, 25), dtype=torch.float64, device=dev) if want_fit_img else None
    if n == 0:
        return out_fit, out_int, fit_img
    sbytes = L.fsq_fit_scratch_bytes(n)
    scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
    rc = L.fsq_fit_candidates(_ptr(frames), _TORCH_DTYPE_CODE[frames.dtype], F, H, W, _ptr(cand_hw),
                              _ptr(cand_frame), n, _ptr(n_dev), ctypes.byref(o), _ptr(out_fit),
                              _ptr(out_int), _ptr(fit_img), _ptr(scratch), sbytes, _stream())
    _lib.check(rc)
    return out_fit, out_int, fit_img


class FitBatch(object):
    __slots__ = ("params", "perror", "covar", "status", "niter", "nfev", "chi2", "n_qrsolv", "fit_img")


def gaussfit_batch(windows, p0, lo, hi, lim_lo, lim_hi, faithful=True, want_perror=False,
                   want_fit_img=False, opts=None, solver="minpack", rescue=None, fixed=None, err=None,
                   circle=False, want_covar=False, **mpfit_kw):
    """gaussfitter.gaussfit -> mpfit (agpy/gaussfitter.py:142-255) for n windows [n,win,win].

    ``fixed`` [n,7] (parinfo 'fixed'), ``err`` [n,win,win] (residual weights), ``circle`` (the circular model) and
    ``want_covar`` (mpfit .covar -> FitBatch.covar [n,7,7]) are the rest of gaussfit's surface
    (fsq_gaussfit_batch_ex, MINPACK solver only); parameters keep the 7-slot layout.

    ``rescue`` (default: on for solver="fast"): the FAST solver keeps its normal equations in FP32 and reports
    status -16 when they leave the finite range (measured: 0.03 % of 11x11 windows cut from a dense field, fits
    that chase a neighbour's tail 10 px outside the window); those fits are re-run by the reference-faithful
    FP64 MINPACK kernel -- on the device, like everything else -- and their rows replaced."""
    L = _lib.load()
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())

    def dv(a, dt):
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=dt).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)

    w = _device_windows(windows, dev)
    n, win, win2 = w.shape
    p0 = dv(p0, torch.float64).reshape(n, 7)
    lo = dv(lo, torch.float64).reshape(n, 7)
    hi = dv(hi, torch.float64).reshape(n, 7)
    lim_lo = dv(lim_lo, torch.uint8).reshape(n, 7)
    lim_hi = dv(lim_hi, torch.uint8).reshape(n, 7)
    o = opts or _lib.default_opts(faithful=faithful, want_perror=want_perror, solver=solver, **mpfit_kw)
    r = FitBatch()
    r.params = torch.empty((n, 7), dtype=torch.float64, device=dev)
    r.perror = torch.empty((n, 7), dtype=torch.float64, device=dev) if want_perror else None
    r.status = torch.empty(n, dtype=torch.int32, device=dev)
    r.niter = torch.empty(n, dtype=torch.int32, device=dev)
    r.nfev = torch.empty(n, dtype=torch.int32, device=dev)
    r.chi2 = torch.empty(n, dtype=torch.float64, device=dev)
    r.n_qrsolv = torch.empty(n, dtype=torch.int32, device=dev)
    r.fit_img = torch.empty((n, win, win), dtype=torch.float64, device=dev) if want_fit_img else None
    r.covar = torch.empty((n, 7, 7), dtype=torch.float64, device=dev) if want_covar else None
    fixed_d = dv(fixed, torch.uint8).reshape(n, 7) if fixed is not None else None
    err_d = dv(err, torch.float64).reshape(n, win, win) if err is not None else None
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    rc = L.fsq_gaussfit_batch_ex(_ptr(w), _TORCH_DTYPE_CODE[w.dtype], n, win, _ptr(p0), _ptr(lo), _ptr(hi),
                                 _ptr(lim_lo), _ptr(lim_hi), _ptr(fixed_d), _ptr(err_d), 1 if circle else 0,
                                 ctypes.byref(o), _ptr(r.params), _ptr(r.perror), _ptr(r.covar),
                                 _ptr(r.status), _ptr(r.niter), _ptr(r.nfev), _ptr(r.chi2), _ptr(r.n_qrsolv),
                                 _ptr(r.fit_img), _ptr(counter), _stream())
    _lib.check(rc)
    if rescue is None:
        rescue = (solver == "fast" and opts is None)
    if rescue and solver == "fast" and n:
        bad = torch.nonzero(r.status == -16).reshape(-1)
        if bad.numel():
            rr = gaussfit_batch(w[bad], p0[bad], lo[bad], hi[bad], lim_lo[bad], lim_hi[bad], faithful=faithful,
                                want_fit_img=want_fit_img, solver="minpack", rescue=False, **mpfit_kw)
            for name in ("params", "status", "niter", "nfev", "chi2", "n_qrsolv", "fit_img"):
                dst, src = getattr(r, name), getattr(rr, name)
                if dst is not None and src is not None:
                    dst[bad] = src
    return r


def gaussfit_batch_trace(windows, p0, lo, hi, lim_lo, lim_hi, faithful=True, trace_steps=256):
    """Debug/test variant of gaussfit_batch: returns (FitBatch, trace [n, trace_steps, 20] numpy)."""
    L = _lib.load()
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    w = torch.from_numpy(np.ascontiguousarray(np.asarray(windows).astype(np.float64))).to(dev)
    n, win, _ = w.shape
    td = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev).reshape(n, 7)
    p0, lo, hi = td(p0, torch.float64), td(lo, torch.float64), td(hi, torch.float64)
    lim_lo, lim_hi = td(lim_lo, torch.uint8), td(lim_hi, torch.uint8)
    o = _lib.default_opts(faithful=faithful)
    r = FitBatch()
    r.params = torch.empty((n, 7), dtype=torch.float64, device=dev)
    r.status = torch.empty(n, dtype=torch.int32, device=dev)
    r.niter = torch.empty(n, dtype=torch.int32, device=dev)
    r.nfev = torch.empty(n, dtype=torch.int32, device=dev)
    r.chi2 = torch.empty(n, dtype=torch.float64, device=dev)
    r.n_qrsolv = torch.empty(n, dtype=torch.int32, device=dev)
    r.perror = r.fit_img = None
    trace = torch.zeros((n, trace_steps, 20), dtype=torch.float64, device=dev)
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    rc = L.fsq_gaussfit_batch_trace(_ptr(w), _lib.FSQ_F64, n, win, _ptr(p0), _ptr(lo), _ptr(hi), _ptr(lim_lo),
                                    _ptr(lim_hi), ctypes.byref(o), _ptr(r.params), _ptr(r.status), _ptr(r.niter),
                                    _ptr(r.nfev), _ptr(r.chi2), _ptr(r.n_qrsolv), _ptr(trace), trace_steps, n,
                                    _ptr(counter), _stream())
    _lib.check(rc)
    return r, trace.cpu().numpy()


_NP_WINDOW_DTYPES = (np.uint8, np.uint16, np.int16, np.int32, np.int64, np.float64)


def _device_windows(windows, dev):
    """numpy / torch windows -> contiguous CUDA tensor [n,win,win]: camera integers (uint8 / uint16 / int16 / int32 /
    int64) and float64 keep their type -- the kernels read all six, and the 11x11 FAST kernels hold 8- / 16-bit pixels
    in shared memory as FP32 (exact), which doubles their occupancy -- anything else becomes int64 or float64"""
    if isinstance(windows, np.ndarray):
        a = windows
        if a.dtype not in _NP_WINDOW_DTYPES or (a.dtype == np.uint16 and not hasattr(torch, "uint16")):
            a = a.astype(np.int64 if a.dtype.kind in "iub" else np.float64)
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint16:
            w = torch.from_numpy(a.view(np.int16)).view(torch.uint16).to(dev)
        else:
            w = torch.from_numpy(a).to(dev)
    else:
        w = windows.to(dev).contiguous()
        if w.dtype not in _TORCH_DTYPE_CODE:
            w = w.to(torch.float64 if w.dtype.is_floating_point else torch.int64)
    if w.dim() == 2:
        w = w.unsqueeze(0)
    if w.dim() != 3 or w.shape[1] != w.shape[2]:
        raise ValueError("windows must be square")
    return w


def moments_batch(windows, lo=None, hi=None, lim_lo=None, lim_hi=None):
    """gaussfitter.moments (agpy/gaussfitter.py:29-61; circle=0, rotate=1, vheight=1, median estimator) for
    n windows [n,win,win] -> device tensor [n,7] (height, amplitude, x, y, width_x, width_y, 0); with the
    four limit vectors [7] the values are clipped into the limits like gaussfitter.py:202-204."""
    L = _lib.load()
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    w = _device_windows(windows, dev)
    n, win, _ = w.shape
    out = torch.empty((n, 7), dtype=torch.float64, device=dev)
    lims = [None] * 4
    if lo is not None:
        lims = [torch.as_tensor(np.asarray(v), dtype=dt).to(dev).reshape(7).contiguous()
                for v, dt in ((lo, torch.float64), (hi, torch.float64), (lim_lo, torch.uint8), (lim_hi, torch.uint8))]
    _lib.check(L.fsq_moments(_ptr(w), _TORCH_DTYPE_CODE[w.dtype], n, win, _ptr(lims[0]), _ptr(lims[1]),
                             _ptr(lims[2]), _ptr(lims[3]), _ptr(out), _stream()))
    return out


GAUSSFIT_DEFAULT_LIMITS = (np.zeros(7), np.array([0, 0, 0, 0, 0, 0, 360.]),                  # gaussfitter.py:143-146
                           np.array([0, 0, 0, 0, 1, 1, 1], dtype=np.uint8), np.array([0, 0, 0, 0, 0, 0, 1], dtype=np.uint8))


def gaussfit_default_batch(windows, solver="fast", faithful=True, **kw):
    """``gaussfit(data)`` with its default arguments (moments start, widths and angle bounded below, angle
    <= 360; agpy/gaussfitter.py:142-148, 188-204) for n windows, start values and fits all on the device --
    the BASELINE configs[0] / configs[3] "11x11 subimages" workload.  -> (FitBatch, p0 [n,7] device)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    require_cuda()
    w = _device_windows(windows, dev)
    n = w.shape[0]
    lo, hi, lmin, lmax = GAUSSFIT_DEFAULT_LIMITS
    p0 = moments_batch(w, lo, hi, lmin, lmax)
    ex = lambda v, dt: torch.as_tensor(v, dtype=dt).to(dev).expand(n, 7).contiguous()
    r = gaussfit_batch(w, p0, ex(lo, torch.float64), ex(hi, torch.float64), ex(lmin, torch.uint8), ex(lmax, torch.uint8),
                       faithful=faithful, solver=solver, **kw)
    return r, p0


PSF_GATED, PSF_DELETED, PSF_FINAL, PSF_REKEYED = 0, 1, 2, 3          # psf_state codes of fsq_consolidate (fsq.h)


class Consolidation(object):
    """Device result of consolidate_batch: per-candidate state / final key, final PSFs per frame."""
    __slots__ = ("state", "key", "n_psf", "flags", "scratch")

    def check(self):
        """Raises the reference's AssertionError (pflib.py:518) if a re-keyed PSF landed on an occupied key."""
        check_consolidation_flags(int(self.flags.item()))


def check_consolidation_flags(v):
    if v & 1:
        raise AssertionError("re-keyed PSF collides with an existing key (pflib.py:518)")


def consolidate_batch(cand_hw, cand_frame, fit, n, n_frames, r_2_threshold=0.7, consolidation_radius=4,
                      n_dev=None, out=None):
    """R^2 gate + rival consolidation + re-key (pflib.py:466-468, 479-519) for the packed candidates of a batch
    of frames, on the device (fsq_consolidate).  cand_hw [>=n,2] i32, cand_frame [>=n] i32, fit [>=n,12] f64
    device tensors in the order fsq_detect emits; ``n_dev`` = device-resident candidate count (no host sync)."""
    L = _lib.load()
    require_cuda()
    if consolidation_radius < 2:
        raise ValueError("consolidation_radius must be at least 2")          # pflib.py:431-432
    dev = fit.device
    n = int(n)
    r = out or Consolidation()
    if out is None:
        r.state = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
        r.key = torch.empty((max(n, 1), 2), dtype=torch.int32, device=dev)
        r.n_psf = torch.empty(max(int(n_frames), 1), dtype=torch.int64, device=dev)
        r.flags = torch.empty(1, dtype=torch.int32, device=dev)
        r.scratch = torch.empty(max(int(L.fsq_consolidate_scratch_bytes(n, int(n_frames))), 1), dtype=torch.uint8, device=dev)
    _lib.check(L.fsq_consolidate(_ptr(cand_hw), _ptr(cand_frame), _ptr(fit), n, _ptr(n_dev), int(n_frames),
                                 float(r_2_threshold), int(consolidation_radius), _ptr(r.state), _ptr(r.key),
                                 _ptr(r.n_psf), _ptr(r.flags), _ptr(r.scratch), r.scratch.numel(), _stream()))
    return r


class PackedPsfs(object):
    """Final PSFs of a batch in the reference's dictionary order (fsq_pack_psfs): ``fit`` [m,12] (the
    find_peptides 12-tuple minus the two 5x5 images, plus chi2 / fnorm), ``ints`` [m,4] = (frame, key_h, key_w,
    candidate index), ``base`` [F+1] offsets of each frame's PSFs."""
    __slots__ = ("fit", "ints", "base", "scratch")

    def frame(self, f):
        a, b = int(self.base[f]), int(self.base[f + 1])
        return self.ints[a:b], self.fit[a:b]


def pack_psfs_batch(cons, cand_frame, fit, n, n_frames, n_dev=None, cap_psf=None):
    """fsq_pack_psfs: the survivors of ``consolidate_batch`` compacted on the device -> PackedPsfs (device tensors)."""
    L = _lib.load()
    dev = fit.device
    n, F = int(n), int(n_frames)
    cap = int(cap_psf if cap_psf is not None else max(n, 1))
    r = PackedPsfs()
    r.fit = torch.empty((cap, 12), dtype=torch.float64, device=dev)
    r.ints = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    r.base = torch.empty(F + 1, dtype=torch.int64, device=dev)
    r.scratch = torch.empty(max(int(L.fsq_pack_psfs_scratch_bytes(n, F)), 1), dtype=torch.uint8, device=dev)
    _lib.check(L.fsq_pack_psfs(_ptr(cons.state), _ptr(cons.key), _ptr(cand_frame), _ptr(fit), n, _ptr(n_dev), F,
                               _ptr(r.fit), _ptr(r.ints), _ptr(r.base), cap, _ptr(r.scratch), r.scratch.numel(), _stream()))
    return r


def psf_dict_order(state, cand_frame=None):
    """Indices of the final PSFs in the reference's dictionary order (numpy, host): survivors that kept their
    candidate pixel as key in raster order, then the re-keyed ones (`del` + `setdefault` moves an entry to the
    end of the dict, pflib.py:514-519).  With ``cand_frame`` the order is per frame, frames in sequence."""
    state = np.asarray(state)
    keep = np.nonzero(state >= PSF_FINAL)[0]
    moved = state[keep] == PSF_REKEYED
    if cand_frame is None:
        return np.concatenate([keep[~moved], keep[moved]])
    fr = np.asarray(cand_frame)[keep]
    order = np.lexsort((keep, moved, fr))                     # frame, then unmoved before moved, then raster
    return keep[order]


def metrics_batch(sub, fit):
    """pflib.py:463-473 for n (sub_img, fit_img) pairs -> [n,3] (r_2, rmse, s_n)."""
    L = _lib.load()
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    s = torch.as_tensor(np.ascontiguousarray(np.asarray(sub, dtype=np.int64))).to(dev).reshape(-1, 25)
    f = torch.as_tensor(np.ascontiguousarray(np.asarray(fit, dtype=np.float64))).to(dev).reshape(-1, 25)
    out = torch.empty((s.shape[0], 3), dtype=torch.float64, device=dev)
    _lib.check(L.fsq_metrics(_ptr(s), _ptr(f), s.shape[0], _ptr(out), _stream()))
    return out


def illumina_s_n_batch(sub):
    """pflib.illumina_s_n (pflib.py:261-281) for n square integer windows [n,size,size] of any side <= 33 -> [n] f64
    device tensor, bit-identical to the reference (numpy's summation order)."""
    L = _lib.load()
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    a = np.asarray(sub)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3 or a.shape[1] != a.shape[2]:
        raise ValueError("sub_img must be square, but has shape " + str(a.shape[1:]))
    if a.dtype.kind not in "iub":
        raise TypeError("illumina_s_n_batch takes integer pixels (the reference's sub-images are int64 copies)")
    s = torch.from_numpy(np.ascontiguousarray(a.astype(np.int64))).to(dev)
    out = torch.empty(s.shape[0], dtype=torch.float64, device=dev)
    _lib.check(L.fsq_illumina_s_n(_ptr(s), s.shape[0], s.shape[1], _ptr(out), _stream()))
    return out


PHOT_METHODS = {"simple": 0, "mexican_hat": 1, "maximum": 2}


def photometry_batch(frames, spots_hw, spot_frame, method="mexican_hat", radius=None, brim_size=6,
                     size=5):
    """flexlibrary.Spot.photometry (flexlibrary.py:160-210, 264-284) for n spots -> [n] f64."""
    L = _lib.load()
    frames = to_device_frames(frames)
    F, H, W = frames.shape
    dev = frames.device
    if method not in PHOT_METHODS:
        raise ValueError("Uknown method specified.")                         # flexlibrary.py:315
    if radius is None:
        radius = {"simple": (size - 1) // 2, "mexican_hat": 9, "maximum": 5}[method]
    hw = _dev_i32(spots_hw, dev).reshape(-1, 2)
    fr = _dev_i32(spot_frame, dev).reshape(-1)
    out = torch.empty(hw.shape[0], dtype=torch.float64, device=dev)
    _lib.check(L.fsq_photometry(_ptr(frames), _TORCH_DTYPE_CODE[frames.dtype], F, H, W, _ptr(hw), _ptr(fr),
                                hw.shape[0], PHOT_METHODS[method], int(radius), int(brim_size), _ptr(out),
                                _stream()))
    return out


def photometry_from_fit(fit, method="gaussian_volume", scaling=10 ** 6):
    """The two Spot.photometry methods that only read the Gaussian fit (flexlibrary.py:212-241), for packed PSF
    records [n,12] (torch or numpy; same multiplication order as the reference):
    'gaussian_volume' = scaling * A * sigma_h * sigma_w, 'sigmas' = scaling * sigma_h * sigma_w."""
    A, sh, sw = fit[:, COL_A], fit[:, COL_SIGMA_H], fit[:, COL_SIGMA_W]
    if method == "gaussian_volume":
        return float(scaling) * A * sh * sw
    if method == "sigmas":
        return float(scaling) * sh * sw
    raise ValueError("Uknown method specified.")                             # flexlibrary.py:315


TRACK_NONE, TRACK_CENTROID, TRACK_STAYED, TRACK_INITIAL = 0, 1, 2, 3        # track_state codes of fsq_track_centroid


def track_centroid_batch(frames, spots_hw, spot_field=None, offsets=None, size=5, search_radius=3, s_n_cutoff=3.0):
    """Experiment.luminosity_centroid_particle_tracking (flexlibrary.py:1262-1317) for n spots of frame 0.
    frames [n_frames,H,W] (one field) or [n_fields,n_frames,H,W]; spots_hw [n,2]; spot_field [n] for several
    fields; offsets [n_frames,2] integer (delta_h, delta_w) or None.
    -> (track_hw [n,n_frames,2] i32, track_state [n,n_frames] u8, track_sn [n,n_frames] f64) device tensors."""
    L = _lib.load()
    require_cuda()
    if size % 2 == 0:
        raise AttributeError("Spot.size must be odd.")                       # flexlibrary.py:98-99
    if isinstance(frames, np.ndarray) and frames.ndim == 4:
        nf = frames.shape[0]
        fr = to_device_frames(frames.reshape((-1,) + frames.shape[2:]))
        fr = fr.reshape((nf, -1) + tuple(fr.shape[1:]))
    elif isinstance(frames, torch.Tensor) and frames.dim() == 4:
        fr = frames.contiguous()
        if fr.dtype not in _TORCH_DTYPE_CODE or fr.dtype in (torch.float64, torch.int64):
            raise TypeError("unsupported frame dtype %s" % fr.dtype)
    else:
        fr = to_device_frames(frames).unsqueeze(0)
    n_fields, n_frames, H, W = fr.shape
    dev = fr.device
    hw = _dev_i32(spots_hw, dev).reshape(-1, 2)
    n = hw.shape[0]
    sf = None
    if spot_field is not None:
        sf = _dev_i32(spot_field, dev).reshape(-1)
        if sf.numel() != n or (n and (int(sf.min()) < 0 or int(sf.max()) >= n_fields)):
            raise ValueError("spot_field must hold one field index in [0, n_fields) per spot")
    elif n_fields != 1:
        raise ValueError("spot_field is required when frames holds several fields")
    of = None
    if offsets is not None:
        o = np.asarray(offsets)
        if o.shape != (n_frames, 2) or not np.all(o == np.rint(o)):
            raise ValueError("offsets must be integer (delta_h, delta_w) pairs, one per frame")
        of = torch.as_tensor(np.ascontiguousarray(o.astype(np.int32))).to(dev)
    track_hw = torch.empty((n, n_frames, 2), dtype=torch.int32, device=dev)
    track_state = torch.empty((n, n_frames), dtype=torch.uint8, device=dev)
    track_sn = torch.empty((n, n_frames), dtype=torch.float64, device=dev)
    _lib.check(L.fsq_track_centroid(_ptr(fr), _TORCH_DTYPE_CODE[fr.dtype], n_fields, n_frames, H, W, _ptr(hw), _ptr(sf),
                                    _ptr(of), n, int(size), int(search_radius), float(s_n_cutoff), _ptr(track_hw),
                                    _ptr(track_state), _ptr(track_sn), _stream()))
    return track_hw, track_state, track_sn


def accumulate_offsets(offsets):
    """Experiment.accumulate_offsets (flexlibrary.py:566-593): Python sums, in sequence."""
    if tuple(offsets[0]) != (0, 0):
        raise ValueError("The first image's offset must be (0, 0) by definiton.")
    return [(sum([o[0] for o in offsets[:f + 1]]), sum([o[1] for o in offsets[:f + 1]])) for f in range(len(offsets))]


def track_greedy_batch(field_frame_spots, frame_shape, candidate_radius=2, offsets=None, spot_radius=0):
    """Experiment.greedy_particle_tracking (flexlibrary.py:680-1027) for several fields in one launch.
    field_frame_spots: per field a list (frames) of [k,2] arrays / lists of (h, w) -- or one such list of frames for
    a single field; offsets: per field a list of per-frame (delta_h, delta_w), or one list for all, or None.
    -> per field (traces [n_traces, n_frames] int64: the spot's index in its frame's list or -1 for None, in the
    reference's order of traces; total_discarded) -- a bare tuple when one field was given."""
    L = _lib.load()
    require_cuda()
    def depth(x):            # nesting depth down to the first number found (empty containers are skipped): 3 = one field
        if isinstance(x, np.ndarray):
            return x.ndim if x.size else None
        if isinstance(x, (list, tuple)):
            for e in x:
                d = depth(e)
                if d is not None:
                    return d + 1
            return None
        return 0
    d = depth(field_frame_spots)
    single = (d is None) or d <= 3
    fields = [field_frame_spots] if single else list(field_frame_spots)
    n_fields = len(fields)
    F = len(fields[0])
    if any(len(fr) != F for fr in fields):
        raise ValueError("every field must have the same number of frames")
    H, W = int(frame_shape[0]), int(frame_shape[1])
    dev = torch.device("cuda", torch.cuda.current_device())
    if offsets is None:
        cum = None
    else:
        per_field = offsets if (n_fields > 1 and np.ndim(offsets) == 3) else [offsets] * n_fields
        cum = np.array([accumulate_offsets([tuple(o) for o in off]) for off in per_field], dtype=np.float64).reshape(n_fields, F, 2)
    seg, pts = [0], []
    for fr in fields:
        for sp in fr:
            a = np.asarray(sp, dtype=np.float64).reshape(-1, 2)
            pts.append(a)
            seg.append(seg[-1] + len(a))
    n = seg[-1]
    hw = torch.from_numpy(np.ascontiguousarray(np.concatenate(pts) if n else np.zeros((0, 2)))).to(dev)
    seg_t = torch.tensor(seg, dtype=torch.int32, device=dev)
    cum_t = torch.from_numpy(cum).to(dev) if cum is not None else None
    anc = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    desc = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    bins = torch.empty((max(n, 1), 2), dtype=torch.int32, device=dev)
    disc = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    flags = torch.empty(n_fields, dtype=torch.int32, device=dev)
    sbytes = int(L.fsq_track_greedy_scratch_bytes(n_fields, H, W, n))
    scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
    _lib.check(L.fsq_track_greedy(_ptr(hw), _ptr(seg_t), _ptr(cum_t), n_fields, F, H, W, n, int(candidate_radius),
                                  float(spot_radius), _ptr(anc), _ptr(desc), _ptr(bins), _ptr(disc), _ptr(flags),
                                  _ptr(scratch), sbytes, _stream()))
    if int((flags & 1).any()):
        raise AssertionError("two spots of one frame round to the same pixel (flexlibrary.py:853-858)")
    anc, desc, bins, disc = anc.cpu().numpy()[:n], desc.cpu().numpy()[:n], bins.cpu().numpy()[:n], disc.cpu().numpy()[:n]
    seg = np.array(seg)
    frame_of = np.repeat(np.arange(n_fields * F), np.diff(seg))
    out = []
    for b in range(n_fields):
        lo, hi = seg[b * F], seg[(b + 1) * F]
        idx = np.arange(lo, hi)
        local_frame = frame_of[lo:hi] - b * F
        heads = idx[(anc[lo:hi] < 0) & (disc[lo:hi] == 0)]
        # the reference collects the heads frame by frame, each frame in raster order of the rounded pixels (:986-993)
        order = np.lexsort((bins[heads, 1], bins[heads, 0], frame_of[heads]))
        traces = -np.ones((len(heads), F), dtype=np.int64)
        for t, h in enumerate(heads[order]):
            cur = h
            while cur >= 0:
                traces[t, frame_of[cur] - b * F] = cur - seg[frame_of[cur]]
                cur = desc[cur]
        out.append((traces, int(disc[lo:hi].sum())))
    return out[0] if single else out


def timetrace_batch(frames, faithful=False, solver="fast", r_2_threshold=0.7, consolidation_radius=4,
                    search_radius=3, s_n_cutoff=3.0, photometry_method="mexican_hat", **photometry_kw):
    """The device part of basic_timetrace_script.py for one field's movie [n_frames,H,W]: frame 0
    is peak-fitted (find_peptides), its final PSFs are followed through the frames by the luminosity centroid
    (lc_create_traces), and Spot.photometry is evaluated for every (spot, frame) (stepfit_tracks' input).
    -> dict(psf_int [m,4], psf_fit [m,12], track_hw [m,F,2], track_state [m,F], track_sn [m,F], photometry [m,F])
    as numpy arrays; photometry is NaN where the spot is None."""
    frames = to_device_frames(frames)
    F = frames.shape[0]
    res = find_peptides_batch(frames[:1], faithful=faithful, solver=solver, to_host=False)
    n = int(res.cand_hw.shape[0])
    cons = consolidate_batch(res.cand_hw, res.cand_frame, res.fit, n, 1, r_2_threshold, consolidation_radius)
    cons.check()
    pk = pack_psfs_batch(cons, res.cand_frame, res.fit, n, 1)
    m = int(pk.base[1].item())
    psf_int, psf_fit = pk.ints[:m], pk.fit[:m]
    spots = psf_int[:, 1:3].contiguous()
    thw, tst, tsn = track_centroid_batch(frames, spots, search_radius=search_radius, s_n_cutoff=s_n_cutoff)
    live = tst.reshape(-1) != TRACK_NONE
    hw_flat = thw.reshape(-1, 2)
    fr_idx = torch.arange(F, dtype=torch.int32, device=frames.device).repeat(m)
    phot = torch.full((m * F,), float("nan"), dtype=torch.float64, device=frames.device)
    if int(live.sum()):
        phot[live] = photometry_batch(frames, hw_flat[live], fr_idx[live], method=photometry_method, **photometry_kw)
    return {"psf_int": psf_int.cpu().numpy(), "psf_fit": psf_fit.cpu().numpy(), "track_hw": thw.cpu().numpy(),
            "track_state": tst.cpu().numpy(), "track_sn": tsn.cpu().numpy(), "photometry": phot.reshape(m, F).cpu().numpy()}


class FieldResults(object):
    """Packed per-candidate results of find_peptides_batch (host numpy after .cpu())."""
    __slots__ = ("cand_hw", "cand_frame", "n_cand", "thr", "fit", "ints", "fit_img", "shape")


def find_peptides_batch(frames, median_filter_size=5, correlation_matrix=DEFAULT_CORRELATION_MATRIX,
                        c_std=2, faithful=True, want_fit_img=False, to_host=True, solver="minpack"):
    """Detection + per-candidate fit + metrics for a batch of frames (the GPU-relevant part of
    pflib.find_peptides, pflib.py:434-477).  The R^2 gate / consolidation / re-key
    (pflib.py:466-468, 479-519) operate on these packed arrays -- see pflib.consolidate_packed."""
    frames = to_device_frames(frames)
    det = detect_batch(frames, median_filter_size, correlation_matrix, c_std)
    fit, ints, fit_img = fit_candidates(frames, det.cand_hw, det.cand_frame, det.total,
                                        faithful=faithful, want_fit_img=want_fit_img, solver=solver)
    r = FieldResults()
    r.shape = tuple(frames.shape)
    if to_host:
        r.cand_hw = det.cand_hw[:det.total].cpu().numpy()
        r.cand_frame = det.cand_frame[:det.total].cpu().numpy()
        r.n_cand = det.n_cand.cpu().numpy()
        r.thr = det.thr.cpu().numpy()
        r.fit = fit.cpu().numpy()
        r.ints = ints.cpu().numpy()
        r.fit_img = fit_img.cpu().numpy() if fit_img is not None else None
    else:
        r.cand_hw, r.cand_frame, r.n_cand, r.thr = det.cand_hw[:det.total], det.cand_frame[:det.total], det.n_cand, det.thr
        r.fit, r.ints, r.fit_img = fit, ints, fit_img
    return r


class FieldPipeline(object):
    """Pre-allocated, sync-free detection -> fit -> metrics pipeline for batches of a fixed
    shape (the production path bench.py measures).  ``run`` only enqueues kernels on the
    current stream: the candidate count stays on the device (fsq_fit_candidates reads it through
    ``n_dev``); ``fetch`` does the one device->host copy of the packed results."""

    def __init__(self, n_frames, H, W, dtype=None, cap_per_frame=None, faithful=True,
                 median_filter_size=5, correlation_matrix=DEFAULT_CORRELATION_MATRIX, c_std=2,
                 device=None, solver="fast", park_after=None, warps_per_sm=None,
                 consolidate=False, r_2_threshold=0.7, consolidation_radius=4, cap_psf_per_frame=None):
        require_cuda()
        self.L = _lib.load()
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.F, self.H, self.W = int(n_frames), int(H), int(W)
        if consolidation_radius < 2:
            raise ValueError("consolidation_radius must be at least 2")      # pflib.py:431-432
        if self.F > MAX_FRAMES_PER_CALL:
            raise ValueError("at most %d frames per pipeline batch" % MAX_FRAMES_PER_CALL)
        self.dtype = dtype if dtype is not None else torch.uint16
        self.code = _TORCH_DTYPE_CODE[self.dtype]
        self.K = _check_kernel(correlation_matrix)
        self.mf = int(median_filter_size)
        self.c_std = float(c_std)
        if park_after is None and solver == "fast":
            park_after = -8        # drain parking (fsq.h): stragglers leave the launch once the queue is empty; scheduling only
        self.opts = _lib.default_opts(faithful=faithful, solver=solver, park_after=park_after, warps_per_sm=warps_per_sm)
        self.solver = solver
        if cap_per_frame is None:
            cap_per_frame = max(1024, int(0.06 * H * W))
        self.cap = int(cap_per_frame) * self.F
        self.sbytes = self.L.fsq_detect_scratch_bytes(self.F, self.H, self.W)
        d = self.dev
        self.scratch = torch.empty(self.sbytes, dtype=torch.uint8, device=d)
        self.cand_hw = torch.empty((self.cap, 2), dtype=torch.int32, device=d)
        self.cand_frame = torch.empty(self.cap, dtype=torch.int32, device=d)
        self.n_cand = torch.zeros(self.F + 1, dtype=torch.int64, device=d)
        self.thr = torch.empty(self.F, dtype=torch.float64, device=d)
        self.out_fit = torch.empty((self.cap, 12), dtype=torch.float64, device=d)
        self.out_int = torch.empty((self.cap, 4), dtype=torch.int32, device=d)
        self.fit_sbytes = self.L.fsq_fit_scratch_bytes(self.cap)
        self.fit_scratch = torch.empty(self.fit_sbytes, dtype=torch.uint8, device=d)
        self.F_run = self.F
        self.consolidate = False
        # detection: (cm, thr, rowmask) per chunk of frames whose correlation map fits the L2 (fsq_detect), then rowscan,
        # framescan, emit; + fit launches (FAST: prep, phase 1, phase 2 when parking, finish)
        self.kernels_per_run = self.launches_per_run()
        # optional tail of find_peptides on the device: R^2 gate + consolidation + re-key (7 launches), packed
        # final PSFs in dictionary order (3 launches) -- what then leaves the device is ~1 record per spot
        self.consolidate = bool(consolidate)
        if self.consolidate:
            self.r2_thr, self.radius = float(r_2_threshold), int(consolidation_radius)
            if cap_psf_per_frame is None:
                cap_psf_per_frame = max(256, int(0.015 * H * W))
            self.cap_psf = int(cap_psf_per_frame) * self.F
            self.psf_state = torch.empty(self.cap, dtype=torch.uint8, device=d)
            self.psf_key = torch.empty((self.cap, 2), dtype=torch.int32, device=d)
            self.n_psf = torch.empty(self.F, dtype=torch.int64, device=d)
            self.cons_flags = torch.empty(1, dtype=torch.int32, device=d)
            self.cons_sbytes = max(self.L.fsq_consolidate_scratch_bytes(self.cap, self.F),
                                   self.L.fsq_pack_psfs_scratch_bytes(self.cap, self.F))
            self.cons_scratch = torch.empty(self.cons_sbytes, dtype=torch.uint8, device=d)
            self.psf_fit = torch.empty((self.cap_psf, 12), dtype=torch.float64, device=d)
            self.psf_int = torch.empty((self.cap_psf, 4), dtype=torch.int32, device=d)
            self.psf_base = torch.zeros(self.F + 1, dtype=torch.int64, device=d)
            self.kernels_per_run = self.launches_per_run()

    def launches_per_run(self, n_frames=None):
        """Kernel launches one run() enqueues for ``n_frames`` frames (bench.py's gpu_launches)."""
        F = self.F if n_frames is None else int(n_frames)
        det_chunk = max(1, (64 << 20) // (self.H * self.W * 6))
        k = 3 * (-(-F // det_chunk)) + 3 + \
            ((3 + (1 if self.opts.park_after != 0 else 0)) if _lib.SOLVERS[self.solver] == 2 else 1)
        return k + (10 if self.consolidate else 0)

    def run(self, frames_dev, fit=True, n_frames=None):
        """Enqueue one pass over frames_dev [F,H,W] (device tensor of self.dtype).  ``n_frames`` <= the F the
        pipeline was sized for runs a partial batch through the same buffers (the last chunk of a field block)."""
        L = self.L
        F = self.F if n_frames is None else int(n_frames)
        if not (1 <= F <= self.F) or frames_dev.shape[0] < F:
            raise ValueError("n_frames must be in 1..%d and covered by frames_dev" % self.F)
        self.F_run = F
        Kp = self.K.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
        st = _stream()
        _lib.check(L.fsq_detect(_ptr(frames_dev), self.code, F, self.H, self.W, Kp, self.K.shape[0],
                                self.mf, self.c_std, _ptr(self.cand_hw), _ptr(self.cand_frame),
                                _ptr(self.n_cand), _ptr(self.thr), self.cap, _ptr(self.scratch),
                                self.sbytes, st))
        if fit:
            n_dev = ctypes.c_void_p(self.n_cand.data_ptr() + 8 * F)
            _lib.check(L.fsq_fit_candidates(_ptr(frames_dev), self.code, F, self.H, self.W,
                                            _ptr(self.cand_hw), _ptr(self.cand_frame), self.cap, n_dev,
                                            ctypes.byref(self.opts), _ptr(self.out_fit), _ptr(self.out_int),
                                            None, _ptr(self.fit_scratch), self.fit_sbytes, st))
            if self.consolidate:
                _lib.check(L.fsq_consolidate(_ptr(self.cand_hw), _ptr(self.cand_frame), _ptr(self.out_fit), self.cap, n_dev,
                                             F, self.r2_thr, self.radius, _ptr(self.psf_state), _ptr(self.psf_key),
                                             _ptr(self.n_psf), _ptr(self.cons_flags), _ptr(self.cons_scratch),
                                             self.cons_sbytes, st))
                _lib.check(L.fsq_pack_psfs(_ptr(self.psf_state), _ptr(self.psf_key), _ptr(self.cand_frame), _ptr(self.out_fit),
                                           self.cap, n_dev, F, _ptr(self.psf_fit), _ptr(self.psf_int), _ptr(self.psf_base),
                                           self.cap_psf, _ptr(self.cons_scratch), self.cons_sbytes, st))

    def total_psfs(self):
        """Synchronising read of the number of final PSFs; raises if the PSF capacity was exceeded or the
        reference's re-key assert (pflib.py:518) would have fired."""
        m = int(self.psf_base[self.F_run].item())
        if m > self.cap_psf:
            raise _lib.FsqError("capacity: %d PSFs > cap %d; enlarge cap_psf_per_frame" % (m, self.cap_psf))
        check_consolidation_flags(int(self.cons_flags.item()))
        return m

    def fetch_psfs(self):
        """-> (m, psf_int [m,4] (frame, key_h, key_w, candidate index), psf_fit [m,12], psf_base [F+1]) on the host"""
        self.total()
        m = self.total_psfs()
        return m, self.psf_int[:m].cpu(), self.psf_fit[:m].cpu(), self.psf_base[:self.F_run + 1].cpu()

    def run_detect_only(self, frames_dev):
        self.run(frames_dev, fit=False)

    def run_fit_only(self, frames_dev):
        """Re-fit the candidates of the last detection (used to time the fit kernel alone)."""
        n_dev = ctypes.c_void_p(self.n_cand.data_ptr() + 8 * self.F_run)
        _lib.check(self.L.fsq_fit_candidates(_ptr(frames_dev), self.code, self.F_run, self.H, self.W,
                                             _ptr(self.cand_hw), _ptr(self.cand_frame), self.cap, n_dev,
                                             ctypes.byref(self.opts), _ptr(self.out_fit), _ptr(self.out_int),
                                             None, _ptr(self.fit_scratch), self.fit_sbytes, _stream()))

    def total(self):
        """Synchronising read of the candidate total; raises if the capacity was exceeded."""
        n = int(self.n_cand[self.F_run].item())
        if n > self.cap:
            raise _lib.FsqError("capacity: %d candidates > cap %d; enlarge cap_per_frame" % (n, self.cap))
        return n

    def fetch(self, pinned=None):
        """One D2H copy of the packed per-candidate results -> (n, cand_hw, cand_frame, fit, ints, n_cand)."""
        n = self.total()
        if pinned is not None:
            pinned["fit"][:n].copy_(self.out_fit[:n], non_blocking=True)
            pinned["ints"][:n].copy_(self.out_int[:n], non_blocking=True)
            pinned["hw"][:n].copy_(self.cand_hw[:n], non_blocking=True)
            pinned["frame"][:n].copy_(self.cand_frame[:n], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return n, pinned["hw"][:n], pinned["frame"][:n], pinned["fit"][:n], pinned["ints"][:n]
        return (n, self.cand_hw[:n].cpu(), self.cand_frame[:n].cpu(), self.out_fit[:n].cpu(),
                self.out_int[:n].cpu())

    def pinned_buffers(self):
        return {"fit": torch.empty((self.cap, 12), dtype=torch.float64).pin_memory(),
                "ints": torch.empty((self.cap, 4), dtype=torch.int32).pin_memory(),
                "hw": torch.empty((self.cap, 2), dtype=torch.int32).pin_memory(),
                "frame": torch.empty(self.cap, dtype=torch.int32).pin_memory()}

    def d2h_bytes(self, n):
        return n * (12 * 8 + 4 * 4 + 2 * 4 + 4)

    def d2h_bytes_psfs(self, m):
        return m * (12 * 8 + 4 * 4) + (self.F + 1) * 8 + 16


class FieldStream(object):
    """Software-pipelined production path: ``depth`` FieldPipeline slots, each on its own CUDA
    stream with its own device and pinned host buffers, so that the host->device copy of batch
    k+1, the kernels of batch k and the device->host copy of batch k-1 overlap -- and so that the
    latency-bound tail of one batch's long fits runs underneath the next batch's bulk.

        t = fs.submit(frames)          # pinned host tensor [F,H,W] (copied in) or a device tensor
        fs.begin_fetch(t)              # waits for the candidate count, queues the D2H of exactly n rows
        n, hw, frame, fit, ints = fs.end_fetch(t)     # views into the slot's pinned buffers

    Call order for full overlap: submit(k); begin_fetch(k - depth + 2); end_fetch(k - depth + 1) -- the
    host then never waits for a batch it has just queued.  A slot is reused after ``depth`` submits; its
    previous results must have been fetched (or abandoned) by then."""

    def __init__(self, n_frames, H, W, dtype=None, depth=3, host_io=True, fetch="candidates", **kw):
        require_cuda()
        self.depth = int(depth)
        self.slots = []
        self.host_io = host_io
        if fetch not in ("candidates", "psfs"):
            raise ValueError("fetch must be 'candidates' or 'psfs'")
        self.fetch = fetch          # "psfs": consolidation on the device, only the final PSF records go to the host
        if fetch == "psfs":
            kw["consolidate"] = True
        for _ in range(self.depth):
            p = FieldPipeline(n_frames, H, W, dtype=dtype, **kw)
            sl = {"pipe": p, "stream": torch.cuda.Stream(device=p.dev),
                  "ev_count": torch.cuda.Event(), "ev_done": torch.cuda.Event(), "n": None}
            if host_io:
                sl["frames_dev"] = torch.empty((p.F, p.H, p.W), dtype=p.dtype, device=p.dev)
                sl["count_host"] = torch.zeros(4, dtype=torch.int64).pin_memory()
                if fetch == "psfs":
                    sl["pinned"] = {"psf_fit": torch.empty((p.cap_psf, 12), dtype=torch.float64).pin_memory(),
                                    "psf_int": torch.empty((p.cap_psf, 4), dtype=torch.int32).pin_memory(),
                                    "psf_base": torch.empty(p.F + 1, dtype=torch.int64).pin_memory(),
                                    "flags": torch.zeros(1, dtype=torch.int32).pin_memory()}
                else:
                    sl["pinned"] = p.pinned_buffers()
            self.slots.append(sl)
        self.k = 0
        self.kernels_per_run = self.slots[0]["pipe"].kernels_per_run
        # host -> device frame copies go through ONE dedicated stream, in submission order: a copy queued on a slot's own
        # stream shares a hardware work queue with other slots' kernels (CUDA maps streams onto a few queues) and holds
        # them back -- measured on B200: the e2e region ran 12-16 % below the resident one with per-slot copies
        self.copy_stream = torch.cuda.Stream(device=self.slots[0]["pipe"].dev) if host_io else None

    def submit(self, frames):
        """frames: [F', H, W] with F' <= the batch size the stream was built for (a shorter last chunk is allowed)"""
        sl = self.slots[self.k % self.depth]
        self.k += 1
        p = sl["pipe"]
        F = int(frames.shape[0])
        sl["stream"].wait_stream(torch.cuda.current_stream())
        if frames.device.type == "cpu":
            self.copy_stream.wait_stream(sl["stream"])          # the slot's previous batch has finished with frames_dev
            with torch.cuda.stream(self.copy_stream):
                sl["frames_dev"][:F].copy_(frames, non_blocking=True)
            sl["stream"].wait_stream(self.copy_stream)
            frames = sl["frames_dev"]
        with torch.cuda.stream(sl["stream"]):
            p.run(frames, n_frames=F)
            if self.host_io:
                sl["count_host"][0:1].copy_(p.n_cand[F:F + 1], non_blocking=True)
                if self.fetch == "psfs":
                    sl["count_host"][1:2].copy_(p.psf_base[F:F + 1], non_blocking=True)
                    sl["pinned"]["flags"].copy_(p.cons_flags, non_blocking=True)
                sl["ev_count"].record(sl["stream"])
        sl["n"] = None
        sl["F"] = F
        return sl

    def begin_fetch(self, sl):
        p = sl["pipe"]
        sl["ev_count"].synchronize()
        n = int(sl["count_host"][0])
        if n > p.cap:
            raise _lib.FsqError("capacity: %d candidates > cap %d; enlarge cap_per_frame" % (n, p.cap))
        pin = sl["pinned"]
        if self.fetch == "psfs":
            m = int(sl["count_host"][1])
            if m > p.cap_psf:
                raise _lib.FsqError("capacity: %d PSFs > cap %d; enlarge cap_psf_per_frame" % (m, p.cap_psf))
            check_consolidation_flags(int(pin["flags"][0]))
            with torch.cuda.stream(sl["stream"]):
                pin["psf_fit"][:m].copy_(p.psf_fit[:m], non_blocking=True)
                pin["psf_int"][:m].copy_(p.psf_int[:m], non_blocking=True)
                pin["psf_base"][:sl["F"] + 1].copy_(p.psf_base[:sl["F"] + 1], non_blocking=True)
                sl["ev_done"].record(sl["stream"])
            sl["n"], sl["m"] = n, m
            return n
        with torch.cuda.stream(sl["stream"]):
            pin["fit"][:n].copy_(p.out_fit[:n], non_blocking=True)
            pin["ints"][:n].copy_(p.out_int[:n], non_blocking=True)
            pin["hw"][:n].copy_(p.cand_hw[:n], non_blocking=True)
            pin["frame"][:n].copy_(p.cand_frame[:n], non_blocking=True)
            sl["ev_done"].record(sl["stream"])
        sl["n"] = n
        return n

    def end_fetch(self, sl):
        if sl["n"] is None:
            self.begin_fetch(sl)
        sl["ev_done"].synchronize()
        n, pin = sl["n"], sl["pinned"]
        if self.fetch == "psfs":        # (candidates fitted, final PSFs, psf_int [m,4], psf_fit [m,12], psf_base [F+1])
            m = sl["m"]
            return n, m, pin["psf_int"][:m], pin["psf_fit"][:m], pin["psf_base"][:sl["F"] + 1]
        return n, pin["hw"][:n], pin["frame"][:n], pin["fit"][:n], pin["ints"][:n]

    def synchronize(self):
        for sl in self.slots:
            sl["stream"].synchronize()
