"""
Drop-in mirror of the reference's ``phase_correlate`` module (phase_correlate.py:11-196, a port of Guizar-Sicairos'
efficient subpixel registration; caller: SequenceExperiment.offsets_from_frames, flexlibrary.py:1717-1741) over the
C entry point ``fsq_phase_correlate`` of libfsq.so: the three 2-D FFTs are cuFFT Z2Z transforms (complex128 like
numpy.fft -- the library calls SURVEY.md 8(f) allows for this row), the spectrum product, numpy's complex argmax, the
twiddle tables and the two matrix products of the upsampled DFT, error and phase are kernels of the library
(csrc/fsq_register.cu).  All pairs of a stack go through one launch sequence.  There is no CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import engine, _lib


def _device_frames(img, dev):
    """numpy / torch image stack -> contiguous CUDA tensor of a pixel type the kernels read (u8/u16/i16/i32/i64/f64)"""
    if isinstance(img, torch.Tensor):
        t = img.to(dev)
        if t.dtype not in engine._TORCH_DTYPE_CODE:
            t = t.to(torch.float64)
        return t.contiguous()
    a = np.asarray(img)
    if a.dtype == np.uint16:
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(torch.uint16).to(dev)
    if a.dtype in (np.uint8, np.int16, np.int32, np.int64, np.float64):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return torch.from_numpy(np.ascontiguousarray(a.astype(np.float64))).to(dev)


def _run(ref, reg_ptr, n_pairs, rows, cols, upsample_factor, dtype):
    L = _lib.load()
    dev = ref.device
    usf = int(upsample_factor)
    if usf != upsample_factor or usf < 1:
        raise ValueError("upsample_factor must be a positive integer")
    out = torch.empty((n_pairs, 4), dtype=torch.float64, device=dev)
    sbytes = int(L.fsq_phase_correlate_scratch_bytes(n_pairs, rows, cols, usf))
    scratch = torch.empty(max(sbytes, 1), dtype=torch.uint8, device=dev)
    _lib.check(L.fsq_phase_correlate(ctypes.c_void_p(ref.data_ptr()), ctypes.c_void_p(reg_ptr), engine._TORCH_DTYPE_CODE[dtype],
                                     n_pairs, rows, cols, usf, ctypes.c_void_p(out.data_ptr()),
                                     ctypes.c_void_p(scratch.data_ptr()), sbytes,
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def phase_correlate_batch(ref_images, reg_images, upsample_factor=1):
    """phase_correlate for B image pairs [B,rows,cols] at once -> (row_shift, col_shift, error, diffphase), each a
    float64 CUDA tensor [B]."""
    engine.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    ref = _device_frames(ref_images, dev)
    reg = _device_frames(reg_images, dev)
    if ref.shape != reg.shape:
        raise ValueError("Error: images must be same size for phase_correlate")       # phase_correlate.py:57-58
    if ref.dim() != 3:
        raise ValueError("Error: phase_correlate only supports 2D images")            # :61-62
    if reg.dtype != ref.dtype:
        ref, reg = ref.to(torch.float64), reg.to(torch.float64)
    B, rows, cols = ref.shape
    out = _run(ref, reg.data_ptr(), B, rows, cols, upsample_factor, ref.dtype)
    return out[:, 0], out[:, 1], out[:, 2], out[:, 3]


def phase_correlate(ref_image, reg_image, upsample_factor=1):
    """phase_correlate.py:11-134 -> (row_shift, col_shift, error, diffphase) as Python floats."""
    ref_image, reg_image = np.asarray(ref_image), np.asarray(reg_image)
    if ref_image.shape != reg_image.shape:
        raise ValueError("Error: images must be same size for phase_correlate")
    if ref_image.ndim != 2:
        raise ValueError("Error: phase_correlate only supports 2D images")
    r, c, e, d = phase_correlate_batch(ref_image[None], reg_image[None], upsample_factor)
    return float(r[0]), float(c[0]), float(e[0]), float(d[0])


def offsets_from_frames(frames, upsample_factor=20):
    """SequenceExperiment.offsets_from_frames (flexlibrary.py:1717-1741): offsets[0] = (0, 0), offsets[f + 1] =
    phase_correlate(frames[f], frames[f + 1]) -- every consecutive pair of the stack in one call (the reference and the
    registered image of pair f are frames f and f + 1 of the same device buffer).
    -> list of (d_h, d_w) float tuples."""
    engine.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    fr = _device_frames(frames, dev)
    if fr.dim() != 3:
        raise ValueError("frames must be [n_frames, rows, cols]")
    out = [(0, 0)]
    F, rows, cols = fr.shape
    if F > 1:
        res = _run(fr, fr.data_ptr() + rows * cols * fr.element_size(), F - 1, rows, cols, upsample_factor, fr.dtype)
        r = res.cpu().numpy()
        out += [(float(a), float(b)) for a, b in r[:, :2]]
    return out
