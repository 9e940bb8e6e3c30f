"""CPU: the C-ABI shared library builds in-tree for sm_100a, loads, and exports every symbol
include/fsq.h declares.  No compute call is made here (there is no GPU); only the argument
validation that returns before any CUDA work is exercised."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from fluorosequencingimageanalysis_b200 import _lib, build

HEADER = os.path.join(ROOT, "include", "fsq.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsq_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    L = _lib.load()
    assert L.fsq_version() == _lib.FSQ_VERSION == 200


def test_every_declared_symbol_is_exported():
    L = _lib.load()
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), "include/fsq.h declares %s but libfsq.so does not export it" % n
    # and the python binding table lists the same set
    assert sorted(_lib.EXPORTED) == names


def test_library_holds_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.lib_path()], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_header_cites_reference_lines():
    src = open(HEADER).read()
    for needle in ("pflib.py:217-258", "agpy/gaussfitter.py:142-255", "agpy/mpfit/mpfit.py:600-1388",
                   "pflib.py:441-477", "pflib.py:261-281", "flexlibrary.py:160-210"):
        assert needle in src


def test_argument_validation_without_gpu():
    """Errors that are detected before any CUDA call: return code + fsq_last_error text."""
    L = _lib.load()
    o = _lib.default_opts()
    assert (o.ftol, o.xtol, o.gtol, o.factor, o.maxiter) == (1e-10, 1e-10, 1e-10, 100.0, 200)   # mpfit.py:600-605
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(8)          # never dereferenced: validation fails first
    K = (ctypes.c_int64 * 25)(*([1] * 25))
    rc = L.fsq_detect(null, _lib.FSQ_U16, 1, 8, 8, K, 5, 5, 2.0, null, null, null, null, 0, null, 0, null)
    assert rc == _lib.FSQ_E_ARG and "NULL" in _lib.last_error()
    rc = L.fsq_detect(one, _lib.FSQ_U16, 1, 8, 8, K, 4, 5, 2.0, one, one, one, one, 0, one, 0, null)
    assert rc == _lib.FSQ_E_ARG and "odd" in _lib.last_error()          # pflib.py:236-239
    rc = L.fsq_detect(one, _lib.FSQ_U16, 1, 8, 8, K, 5, 11, 2.0, one, one, one, one, 0, one, 0, null)
    assert rc == _lib.FSQ_E_ARG and "median_filter_size" in _lib.last_error()
    rc = L.fsq_detect(one, _lib.FSQ_U16, 1, 8, 8, K, 5, 5, 2.0, one, one, one, one, 0, one, 16, null)
    assert rc == _lib.FSQ_E_CAPACITY
    with pytest.raises(ValueError):
        _lib.check(_lib.FSQ_E_ARG)
    with pytest.raises(OverflowError):
        _lib.check(_lib.FSQ_E_RANGE)
    bad = _lib.default_opts()
    bad.ftol = 0.0                                                       # mpfit.py:986-989
    rc = L.fsq_gaussfit_batch(one, _lib.FSQ_F64, 1, 5, one, one, one, one, one, ctypes.byref(bad),
                              one, null, one, one, one, one, null, null, one, null)
    assert rc == _lib.FSQ_E_ARG and "inconsistent" in _lib.last_error()
    rc = L.fsq_gaussfit_batch(one, _lib.FSQ_F64, 1, 13, one, one, one, one, one, ctypes.byref(o),
                              one, null, one, one, one, one, null, null, one, null)
    assert rc == _lib.FSQ_E_ARG and "window side" in _lib.last_error()
    assert L.fsq_gaussfit_batch(one, _lib.FSQ_F64, 0, 5, one, one, one, one, one, ctypes.byref(o),
                                one, null, one, one, one, one, null, null, one, null) == 0   # empty batch
    assert L.fsq_fit_candidates(one, _lib.FSQ_U16, 1, 8, 8, one, one, 0, null, ctypes.byref(o), one, one,
                                null, one, 64, null) == 0
    # scratch smaller than fsq_fit_scratch_bytes(n) is refused before anything is launched
    assert L.fsq_fit_scratch_bytes(0) == 64 and L.fsq_fit_scratch_bytes(1000) == 64 + 256 * 1000
    rc = L.fsq_fit_candidates(one, _lib.FSQ_U16, 1, 8, 8, one, one, 10, null, ctypes.byref(o), one, one,
                              null, one, 64, null)
    assert rc == _lib.FSQ_E_CAPACITY and "scratch" in _lib.last_error()
    assert o.park_after == 0 and ctypes.sizeof(_lib.LmOpts) == 56
    rc = L.fsq_photometry(one, _lib.FSQ_U16, 1, 8, 8, one, one, 1, 7, 9, 6, one, null)
    assert rc == _lib.FSQ_E_ARG and "method" in _lib.last_error()         # flexlibrary.py:315
    assert L.fsq_detect_scratch_bytes(0, 8, 8) == 0
    assert L.fsq_detect_scratch_bytes(2, 512, 512) >= 2 * 512 * 512 * 4


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from fluorosequencingimageanalysis_b200 import pflib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pflib._psf_candidates(np.zeros((16, 16), dtype=np.uint16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pflib.find_peptides(np.zeros((16, 16), dtype=np.uint16))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fluorosequencingimageanalysis_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            assert "/root/reference" not in src, fn
