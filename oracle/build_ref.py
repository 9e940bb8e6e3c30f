#!/usr/bin/env python
"""
TEST INFRASTRUCTURE -- not product code.

Build recipe for ``oracle/_ref/``: a *mechanical* Python-2 -> Python-3
transliteration of the three reference files that hold the hot path

    /root/reference/agpy/mpfit/mpfit.py      (MINPACK-1 LM solver, class mpfit)
    /root/reference/agpy/gaussfitter.py      (lines 1-255: moments/twodgaussian/gaussfit)
    /root/reference/pflib.py                 (_psf_candidates, _fit_2d_gaussian, find_peptides ...)
    /root/reference/phase_correlate.py       (frame registration; one numpy-2 API rewrite)
    /root/reference/flexlibrary.py           (lines 1-1319: Spot / Image / Experiment -- photometry, the two trackers)

The reference is Python 2 and does not parse under the Python 3.12 of this
image.  This script plays the role a
Makefile plays for a C reference: it reads the sources *where they lie* under
/root/reference, applies the token-level rewrites listed in ``REWRITES`` below
(print statements, ``has_key``, ``types.ListType``, removed scipy/numpy names,
py2 ``round`` and dict-iteration semantics) and writes the result ONLY into
``oracle/_ref/`` which is git-ignored (so no reference source enters history)
but not gpurun-ignored (so it travels to the GPU box like a built .so).

No numerical statement is touched: every arithmetic line of the output is
byte-identical to the reference line it came from.  The rewrites are listed
one by one so a reviewer can audit them; ``--check`` prints a unified diff.

Usage:  python oracle/build_ref.py [--ref /root/reference] [--check]
"""
import argparse
import difflib
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

_PRINT_RE = re.compile(r"^(\s*)((?:if [^:]+:\s*)?)print\s+(?!\()(.*?)\s*$")
_PRINT_PAREN_TUPLE_RE = re.compile(r"^(\s*)((?:if [^:]+:\s*)?)print\s+(\(.*\).*,.*)$")
_HAS_KEY_RE = re.compile(r"([A-Za-z_][\w\.]*(?:\[[^\]]+\])?)\.has_key\(([^)]+)\)")


def _fix_print(line):
    """``print a, b`` -> ``print(a, b)`` (single physical line only)."""
    m = _PRINT_RE.match(line)
    if m:
        body = m.group(3)
        if body.endswith(","):          # py2 "no newline" form
            body = body[:-1] + ", end=' '"
        return "%s%sprint(%s)\n" % (m.group(1), m.group(2), body)
    m = _PRINT_PAREN_TUPLE_RE.match(line)
    if m and not line.lstrip().startswith("print("):
        return "%s%sprint(%s)\n" % (m.group(1), m.group(2), m.group(3).rstrip())
    return line


def convert_mpfit(src):
    out = []
    for line in src.splitlines(True):
        s = line
        if s.lstrip().startswith("#"):
            out.append(s)
            continue
        s = _fix_print(s)
        s = s.replace("types.ListType", "list")
        s = s.replace("types.DictionaryType", "dict")
        s = s.replace("types.IntType", "int")
        s = s.replace("types.FloatType", "float")
        s = _HAS_KEY_RE.sub(r"(\2 in \1)", s)
        s = s.replace("numpy.float)", "float)")       # mpfit.py:902
        s = s.replace("numpy.rank(", "numpy.ndim(")    # mpfit.py:2278
        out.append(s)
    return "".join(out)


def convert_gaussfitter(src):
    # only the 2-D part (gaussfitter.py:1-255) is on the hot path; the 1-D
    # fitters below line 256 contain multi-line py2 print statements and are
    # never called by pflib/flexlibrary.
    lines = src.splitlines(True)[:255]
    out = []
    for s in lines:
        s = s.replace("xrange(", "range(")                               # :202
        s = s.replace("elif params == [] or len(params)==0:",             # :193
                      "elif len(params)==0:")
        out.append(s)
    return "".join(out)


_PFLIB_HEADER = '''\
# ---- prologue injected by oracle/build_ref.py (py2 semantics shims) ----
import math as _math
def round(x, _floor=_math.floor):
    """Python-2 round(): half away from zero (pflib.py:515 relies on it)."""
    x = float(x)
    return _floor(x + 0.5) if x >= 0 else -_floor(-x + 0.5)
# ---- end prologue ----
'''


def convert_pflib(src):
    src = src.expandtabs(8)                                  # pflib.py:444-456 tab block
    out = []
    for s in src.splitlines(True):
        if s.startswith("from scipy.ndimage.filters import median_filter"):
            s = "from scipy.ndimage import median_filter\n"
        elif s.startswith("from scipy.misc import imread"):
            s = "imread = None  # scipy.misc.imread removed from scipy; file I/O is off the hot path\n"
        elif s.startswith("import cPickle"):
            s = "import pickle as cPickle\n"
        elif s.startswith("from skimage import exposure"):
            s = "exposure = None  # skimage absent; only the PNG helpers use it\n"
        elif s.startswith("from PIL import"):
            s = "try:\n    " + s + "except Exception:\n    pass\n"
        elif s.startswith("sys.path.insert(0, './agpy')"):
            s = "sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))\n"
        s = s.replace(".iteritems()", ".items()")
        # py2 dict.items() returned a list copy, so deleting while iterating was legal
        s = s.replace("in pixel_bins.items():", "in list(pixel_bins.items()):")
        out.append(s)
    text = "".join(out)
    # put the py2 round() shim right after the module docstring / imports
    marker = "import scipy\n"
    idx = text.find(marker)
    assert idx > 0
    idx += len(marker)
    return text[:idx] + "\n" + _PFLIB_HEADER + text[idx:]


def convert_phase_correlate(src):
    # phase_correlate.py is plain numpy and runs unchanged on Python 3 but for one numpy-2 API change:
    # numpy.array(..., copy=False) now raises when a copy is needed (uint16 -> float64 always needs one)
    return src.replace("np.array(ref_image, dtype=np.float64, copy=False)", "np.asarray(ref_image, dtype=np.float64)") \
              .replace("np.array(reg_image, dtype=np.float64, copy=False)", "np.asarray(reg_image, dtype=np.float64)")


_FLEX_HEADER = '''\
# ---- prologue injected by oracle/build_ref.py (py2 semantics shims, absent third-party modules) ----
import math as _math
def round(x, _floor=_math.floor):
    """Python-2 round(): half away from zero (flexlibrary.py:605, 1233-1234 rely on it)."""
    x = float(x)
    return _floor(x + 0.5) if x >= 0 else -_floor(-x + 0.5)
# ---- end prologue ----
'''


def convert_flexlibrary(src):
    """flexlibrary.py:1-1319 -- Spot, Image, Experiment (offset helpers, discard_dropouts, greedy_particle_tracking,
    next_frame_spot_by_luminosity_centroid, luminosity_centroid_particle_tracking).  The classes below line 1320 are
    consumers of the path (traces, step fitting, plotting), hold Python-2-only syntax (:1843) and are not needed by
    any parity test.  Rewrites: imports of absent / renamed modules; `(size - 1) / 2`, which is an INTEGER division
    in Python 2 (the value is used as a slice bound); numpy.object (removed from numpy); dict.iteritems."""
    lines = src.splitlines(True)[:1319]
    out = []
    for s in lines:
        if s.startswith("sys.path.insert(0, '/home/proteanseq/pflib')"):
            s = "sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))\n"
        elif s.startswith("from scipy.misc import imread"):
            s = "imread = None  # scipy.misc.imread removed from scipy; file I/O is off the path\n"
        elif s.startswith("import cPickle"):
            s = "import pickle as cPickle\n"
        elif s.startswith("from scipy.ndimage.measurements import center_of_mass"):
            s = "from scipy.ndimage import center_of_mass\n"
        elif s.startswith("import photutils"):
            s = "photutils = None  # absent; only the sextractor photometry uses it\n"
        elif s.startswith("import stepfitting_library"):
            s = "stepfitting_library = None  # Python-2 module; only the Trace classes below line 1320 use it\n"
        s = s.replace("(size - 1) / 2", "(size - 1) // 2").replace("(self.size - 1) / 2", "(self.size - 1) // 2")
        s = s.replace("dtype=np.object)", "dtype=object)")
        s = s.replace(".iteritems()", ".items()")
        out.append(s)
    text = "".join(out)
    marker = "import stepfitting_library"
    idx = text.find("stepfitting_library = None")
    idx = text.find("\n", idx) + 1
    return text[:idx] + "\n" + _FLEX_HEADER + text[idx:]


def build(ref_root="/root/reference", check=False, quiet=False):
    paths = {
        "flexlibrary": os.path.join(ref_root, "flexlibrary.py"),
        "phase_correlate": os.path.join(ref_root, "phase_correlate.py"),
        "mpfit": os.path.join(ref_root, "agpy", "mpfit", "mpfit.py"),
        "gaussfitter": os.path.join(ref_root, "agpy", "gaussfitter.py"),
        "pflib": os.path.join(ref_root, "pflib.py"),
    }
    for p in paths.values():
        if not os.path.isfile(p):
            raise FileNotFoundError(p)
    srcs = {k: open(p, encoding="utf-8", errors="replace").read() for k, p in paths.items()}
    conv = {
        "mpfit": convert_mpfit(srcs["mpfit"]),
        "gaussfitter": convert_gaussfitter(srcs["gaussfitter"]),
        "pflib": convert_pflib(srcs["pflib"]),
        "phase_correlate": convert_phase_correlate(srcs["phase_correlate"]),
        "flexlibrary": convert_flexlibrary(srcs["flexlibrary"]),
    }
    if check:
        for k in conv:
            a = srcs[k].expandtabs(8).splitlines(True) if k == "pflib" else srcs[k].splitlines(True)
            sys.stdout.writelines(difflib.unified_diff(a, conv[k].splitlines(True),
                                                       "ref/" + k, "_ref/" + k, n=0))
        return OUT
    os.makedirs(os.path.join(OUT, "agpy", "mpfit"), exist_ok=True)
    with open(os.path.join(OUT, "agpy", "__init__.py"), "w") as f:
        f.write("# minimal package shell: the reference's agpy/__init__.py imports modules\n"
                "# that are not in the tree (SURVEY.md section 1)\n")
    with open(os.path.join(OUT, "agpy", "mpfit", "__init__.py"), "w") as f:
        f.write("from .mpfit import mpfit\n")
    with open(os.path.join(OUT, "agpy", "mpfit", "mpfit.py"), "w") as f:
        f.write(conv["mpfit"])
    with open(os.path.join(OUT, "gaussfitter.py"), "w") as f:
        f.write(conv["gaussfitter"])
    with open(os.path.join(OUT, "pflib.py"), "w") as f:
        f.write(conv["pflib"])
    with open(os.path.join(OUT, "phase_correlate.py"), "w") as f:
        f.write(conv["phase_correlate"])
    with open(os.path.join(OUT, "flexlibrary_head.py"), "w") as f:
        f.write(conv["flexlibrary"])
    # compile check: every output must parse
    for rel in ("agpy/mpfit/mpfit.py", "gaussfitter.py", "pflib.py", "phase_correlate.py", "flexlibrary_head.py"):
        p = os.path.join(OUT, rel)
        compile(open(p).read(), p, "exec")
    if not quiet:
        print("oracle/_ref built from", ref_root)
    return OUT


def load():
    """Import the built reference modules; returns (pflib, gaussfitter, mpfit_module) or None."""
    if not os.path.isfile(os.path.join(OUT, "pflib.py")):
        return None
    import importlib
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    saved = {k: sys.modules.get(k) for k in ("pflib", "gaussfitter", "agpy", "agpy.mpfit")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        g = importlib.import_module("gaussfitter")
        p = importlib.import_module("pflib")
        m = importlib.import_module("agpy.mpfit.mpfit")
        return p, g, m
    finally:
        # do not leave top-level names 'pflib'/'gaussfitter' pointing at the reference
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
        if OUT in sys.path:
            sys.path.remove(OUT)


def load_flexlibrary():
    """The head of the reference's flexlibrary (Spot, Image, Experiment: flexlibrary.py:1-1319) from oracle/_ref, or
    None.  It imports the reference's own pflib / phase_correlate from the same directory."""
    path = os.path.join(OUT, "flexlibrary_head.py")
    if not os.path.isfile(path):
        return None
    import importlib.util
    saved = {k: sys.modules.get(k) for k in ("pflib", "gaussfitter", "agpy", "agpy.mpfit", "phase_correlate")}
    sys.path.insert(0, OUT)
    try:
        for k in saved:
            sys.modules.pop(k, None)
        spec = importlib.util.spec_from_file_location("_ref_flexlibrary_head", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
        if OUT in sys.path:
            sys.path.remove(OUT)


def load_phase_correlate():
    """The reference's phase_correlate module (phase_correlate.py:11-196) from oracle/_ref, or None."""
    path = os.path.join(OUT, "phase_correlate.py")
    if not os.path.isfile(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_phase_correlate", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--check", action="store_true", help="print the rewrite diff instead of building")
    a = ap.parse_args()
    build(a.ref, a.check)
