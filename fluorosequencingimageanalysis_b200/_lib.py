"""
ctypes binding of libfsq.so (the C-ABI declared in include/fsq.h).

There is NO CPU fallback: if the CUDA library cannot be loaded (or built in-tree with nvcc)
importing the compute entry points raises.  PyTorch is used only for device memory and
streams.
"""
import ctypes
import os

from . import build as _build

_c = ctypes
_LIB = None

FSQ_U8, FSQ_U16, FSQ_I16, FSQ_I32, FSQ_F64, FSQ_I64 = 0, 1, 2, 3, 4, 5
FSQ_VERSION = 200          # include/fsq.h FSQ_VERSION this binding was written against
FSQ_OK, FSQ_E_ARG, FSQ_E_CAPACITY, FSQ_E_CUDA, FSQ_E_RANGE = 0, -1, -2, -3, -4

EXPORTED = ["fsq_version", "fsq_last_error", "fsq_detect_scratch_bytes", "fsq_detect",
            "fsq_detect_flags", "fsq_detect_copy_cm32", "fsq_lm_default_opts",
            "fsq_gaussfit_batch", "fsq_gaussfit_batch_ex", "fsq_gaussfit_batch_trace", "fsq_fit_candidates", "fsq_fit_scratch_bytes",
            "fsq_metrics", "fsq_illumina_s_n", "fsq_photometry", "fsq_moments", "fsq_consolidate", "fsq_consolidate_scratch_bytes", "fsq_pack_psfs", "fsq_pack_psfs_scratch_bytes", "fsq_track_centroid", "fsq_track_greedy", "fsq_track_greedy_scratch_bytes", "fsq_phase_correlate", "fsq_phase_correlate_scratch_bytes",
            "fsq_fma_peak"]


class LmOpts(_c.Structure):
    """struct fsq_lm_opts (mpfit keyword defaults, agpy/mpfit/mpfit.py:600-605)."""
    _fields_ = [("ftol", _c.c_double), ("xtol", _c.c_double), ("gtol", _c.c_double),
                ("factor", _c.c_double), ("maxiter", _c.c_int32), ("faithful", _c.c_int32),
                ("want_perror", _c.c_int32), ("solver", _c.c_int32),
                ("park_after", _c.c_int32), ("warps_per_sm", _c.c_int32)]


class FsqError(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load():
    """Load (building in-tree if necessary) libfsq.so; raises if that is impossible."""
    global _LIB
    if _LIB is not None:
        return _LIB
    # build() is a sha256 comparison of the sources with the stamp of the .so when nothing changed; it rebuilds a
    # stale library and raises when that is impossible (no nvcc), so a parity test or a benchmark can never run a
    # binary that does not match the sources / the ABI of include/fsq.h
    path = _build.build()
    L = _c.CDLL(path)
    vp, i32, i64, dbl = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_double
    L.fsq_version.restype = i32
    L.fsq_version.argtypes = []
    L.fsq_last_error.restype = _c.c_char_p
    L.fsq_last_error.argtypes = []
    L.fsq_detect_scratch_bytes.restype = i64
    L.fsq_detect_scratch_bytes.argtypes = [i32, i32, i32]
    L.fsq_detect.restype = i32
    L.fsq_detect.argtypes = [vp, i32, i32, i32, i32, _c.POINTER(i64), i32, i32, dbl,
                             vp, vp, vp, vp, i64, vp, i64, vp]
    L.fsq_detect_flags.restype = i32
    L.fsq_detect_flags.argtypes = [vp, i32, i32, i32, vp]
    L.fsq_detect_copy_cm32.restype = i32
    L.fsq_detect_copy_cm32.argtypes = [vp, i32, i32, i32, vp, vp]
    L.fsq_lm_default_opts.restype = None
    L.fsq_lm_default_opts.argtypes = [_c.POINTER(LmOpts)]
    L.fsq_gaussfit_batch.restype = i32
    L.fsq_gaussfit_batch.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, vp, _c.POINTER(LmOpts),
                                     vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.fsq_gaussfit_batch_ex.restype = i32
    L.fsq_gaussfit_batch_ex.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, i32, _c.POINTER(LmOpts),
                                        vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.fsq_gaussfit_batch_trace.restype = i32
    L.fsq_gaussfit_batch_trace.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, vp, _c.POINTER(LmOpts),
                                           vp, vp, vp, vp, vp, vp, vp, i32, i64, vp, vp]
    L.fsq_fit_candidates.restype = i32
    L.fsq_fit_candidates.argtypes = [vp, i32, i32, i32, i32, vp, vp, i64, vp, _c.POINTER(LmOpts),
                                     vp, vp, vp, vp, i64, vp]
    L.fsq_fit_scratch_bytes.restype = i64
    L.fsq_fit_scratch_bytes.argtypes = [i64]
    L.fsq_metrics.restype = i32
    L.fsq_metrics.argtypes = [vp, vp, i64, vp, vp]
    L.fsq_illumina_s_n.restype = i32
    L.fsq_illumina_s_n.argtypes = [vp, i64, i32, vp, vp]
    L.fsq_photometry.restype = i32
    L.fsq_photometry.argtypes = [vp, i32, i32, i32, i32, vp, vp, i64, i32, i32, i32, vp, vp]
    L.fsq_moments.restype = i32
    L.fsq_moments.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, vp, vp]
    L.fsq_consolidate_scratch_bytes.restype = i64
    L.fsq_consolidate_scratch_bytes.argtypes = [i64, i32]
    L.fsq_consolidate.restype = i32
    L.fsq_consolidate.argtypes = [vp, vp, vp, i64, vp, i32, dbl, i32, vp, vp, vp, vp, vp, i64, vp]
    L.fsq_pack_psfs_scratch_bytes.restype = i64
    L.fsq_pack_psfs_scratch_bytes.argtypes = [i64, i32]
    L.fsq_pack_psfs.restype = i32
    L.fsq_pack_psfs.argtypes = [vp, vp, vp, vp, i64, vp, i32, vp, vp, vp, i64, vp, i64, vp]
    L.fsq_track_centroid.restype = i32
    L.fsq_track_centroid.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, i64, i32, i32, dbl, vp, vp, vp, vp]
    L.fsq_track_greedy_scratch_bytes.restype = i64
    L.fsq_track_greedy_scratch_bytes.argtypes = [i32, i32, i32, i64]
    L.fsq_track_greedy.restype = i32
    L.fsq_track_greedy.argtypes = [vp, vp, vp, i32, i32, i32, i32, i64, i32, dbl, vp, vp, vp, vp, vp, vp, i64, vp]
    L.fsq_phase_correlate_scratch_bytes.restype = i64
    L.fsq_phase_correlate_scratch_bytes.argtypes = [i32, i32, i32, i32]
    L.fsq_phase_correlate.restype = i32
    L.fsq_phase_correlate.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, i64, vp]
    L.fsq_fma_peak.restype = i32
    L.fsq_fma_peak.argtypes = [i32, _c.POINTER(dbl), vp]
    if L.fsq_version() != FSQ_VERSION:
        raise FsqError("libfsq.so reports version %d, this package binds version %d" % (L.fsq_version(), FSQ_VERSION))
    _LIB = L
    return L


def last_error():
    return load().fsq_last_error().decode("utf-8", "replace")


def check(rc):
    """Map the C return convention onto Python exceptions (error table)."""
    if rc == FSQ_OK:
        return
    msg = last_error()
    if rc == FSQ_E_ARG:
        raise ValueError(msg)
    if rc == FSQ_E_CAPACITY:
        raise FsqError("capacity: " + msg)
    if rc == FSQ_E_RANGE:
        raise OverflowError(msg)
    raise FsqError("libfsq error %d: %s" % (rc, msg))


SOLVERS = {"minpack": 0, "fast64": 1, "fast": 2}


def default_opts(faithful=True, want_perror=False, solver="minpack", **kw):
    o = LmOpts()
    load().fsq_lm_default_opts(_c.byref(o))
    o.faithful = 1 if faithful else 0
    o.want_perror = 1 if want_perror else 0
    if solver not in SOLVERS:
        raise ValueError("solver must be one of %s" % sorted(SOLVERS))
    o.solver = SOLVERS[solver]
    for k, v in kw.items():
        if v is not None:
            setattr(o, k, v)
    return o
