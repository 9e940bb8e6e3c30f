"""
In-tree build of libfsq.so (hand-written CUDA for sm_100a + the C-ABI of include/fsq.h).

    python -m fluorosequencingimageanalysis_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the gpurun snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libfsq.so")
STAMP = os.path.join(PKG, "csrc", ".build_stamp")
SOURCES = ["fsq_api.cu", "fsq_detect.cu", "fsq_lmfit.cu", "fsq_lmfast.cu", "fsq_lmwarp.cu", "fsq_consolidate.cu", "fsq_track.cu", "fsq_register.cu"]
HEADERS = ["fsq_common.cuh", "fsq_median.cuh", "fsq_median_pair.cuh", "fsq_chol7.cuh", os.path.join("..", "..", "include", "fsq.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# per-file extras.  The reference-faithful MINPACK kernel is compiled without implicit FMA
# contraction: every fused operation in it is an explicit fma(), so its arithmetic is fixed by
# the source (two template instantiations give bit-identical fits) -- DESIGN.md "Parity".
# The FAST LM kernel's FP32 half (Jacobian, normal equations, Cholesky, lmpar: quantities that only shape the step) is
# compiled with flush-to-zero and the approximate square root: the IEEE square-root fix-up and the denormal paths of
# rsqrt / reciprocal sit in the serial chains of every tick.  Measured on B200 (profiles/r02d_lm_ab.txt): 4.43 ->
# 4.01 ms per 200-frame launch with three in flight, same iteration counts, same parity figures.  Its FP64 half (residual,
# chi^2, parameters, bounds) is not touched by these flags.
EXTRA_FLAGS = {"fsq_lmfit.cu": ["-fmad=false"], "fsq_lmwarp.cu": ["-ftz=true", "-prec-sqrt=false"]}
if os.environ.get("FSQ_WMINB"):          # developer knob: occupancy target of the FAST LM kernel
    EXTRA_FLAGS["fsq_lmwarp.cu"] = EXTRA_FLAGS["fsq_lmwarp.cu"] + ["-DWMINB=" + os.environ["FSQ_WMINB"]]
if os.environ.get("FSQ_WDEFS"):          # developer knob: extra -D switches for the FAST LM kernel
    EXTRA_FLAGS.setdefault("fsq_lmwarp.cu", []).extend(os.environ["FSQ_WDEFS"].split())


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; libfsq.so cannot be built")


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update((" ".join(NVCC_FLAGS) + repr(sorted(EXTRA_FLAGS.items()))).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + EXTRA_FLAGS.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (src, out))
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
    # cuFFT (the two 2-D FFTs of phase_correlate, fsq_register.cu) is the one library libfsq.so links
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-L" + cuda_lib, "-lcufft",
                                                 "-Xlinker", "-rpath=" + cuda_lib]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError("link failed:\n" + out.stdout)
    with open(os.path.join(CSRC, "ptxas_info.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(dig)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
