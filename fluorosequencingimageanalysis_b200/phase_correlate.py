"""
Drop-in mirror of the reference's ``phase_correlate`` module (phase_correlate.py:11-196, a port of Guizar-Sicairos'
efficient subpixel registration; caller: SequenceExperiment.offsets_from_frames, flexlibrary.py:1717-1741) on the
GPU: the three 2-D FFTs go through cuFFT (``torch.fft`` on CUDA tensors, complex128 like numpy), the upsampled
matrix-multiply DFT around the whole-pixel peak through cuBLAS.  These are LIBRARY calls -- this module is a
"next" row of SURVEY.md 8(f) (rank 2), not part of the hand-written hot path -- batched over all consecutive
frame pairs of a stack so that one field's registration is a handful of launches.  There is no CPU fallback.
"""
import math

import numpy as np
import torch

from . import engine


def _device_f64(img, dev):
    if isinstance(img, torch.Tensor):
        return img.to(device=dev, dtype=torch.float64)
    a = np.asarray(img)
    if a.dtype == np.uint16:
        return torch.from_numpy(a.astype(np.int32)).to(dev).to(torch.float64)
    return torch.from_numpy(np.ascontiguousarray(a.astype(np.float64))).to(dev)


def _centered_freq(n, dev):
    """ifftshift(arange(n)) - floor(n / 2)  (phase_correlate.py:184-186, 190-192)"""
    return torch.fft.ifftshift(torch.arange(n, device=dev, dtype=torch.float64)) - math.floor(n / 2)


def _dftups(data, up_rows, up_cols, usf, row_offset, col_offset):
    """Upsampled DFT of a batch by matrix products (phase_correlate.py:136-196).  data [B,rows,cols] complex128;
    row_offset / col_offset [B] float64 -> [B,up_rows,up_cols]."""
    B, rows, cols = data.shape
    dev = data.device
    kc = torch.arange(up_cols, device=dev, dtype=torch.float64)[None, None, :] - col_offset[:, None, None]      # [B,1,U]
    col_arg = (-2.0 * math.pi / (cols * usf)) * (_centered_freq(cols, dev)[None, :, None] * kc)                # [B,cols,U]
    kr = torch.arange(up_rows, device=dev, dtype=torch.float64)[None, :, None] - row_offset[:, None, None]      # [B,U,1]
    row_arg = (-2.0 * math.pi / (rows * usf)) * (kr * _centered_freq(rows, dev)[None, None, :])                # [B,U,rows]
    col_kernel = torch.polar(torch.ones_like(col_arg), col_arg)
    row_kernel = torch.polar(torch.ones_like(row_arg), row_arg)
    return torch.bmm(torch.bmm(row_kernel, data), col_kernel)


def phase_correlate_batch(ref_images, reg_images, upsample_factor=1):
    """phase_correlate for B image pairs [B,rows,cols] at once -> (row_shift, col_shift, error, diffphase), each a
    float64 CUDA tensor [B]."""
    engine.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    ref = _device_f64(ref_images, dev)
    reg = _device_f64(reg_images, dev)
    if ref.shape != reg.shape:
        raise ValueError("Error: images must be same size for phase_correlate")       # phase_correlate.py:57-58
    if ref.dim() != 3:
        raise ValueError("Error: phase_correlate only supports 2D images")            # :61-62
    B, rows, cols = ref.shape
    F_ref = torch.fft.fft2(ref.to(torch.complex128))
    F_reg = torch.fft.fft2(reg.to(torch.complex128))
    cc = torch.fft.ifft2(F_ref * F_reg.conj())                                        # :72
    # numpy's argmax orders complex numbers by real part first; the imaginary part of this array is rounding noise
    flat = cc.real.reshape(B, -1).argmax(dim=1)
    row_max = (flat // cols).to(torch.float64)
    col_max = (flat % cols).to(torch.float64)
    mid_row, mid_col = float(np.fix(rows / 2)), float(np.fix(cols / 2))               # :76-77
    row_shift = torch.where(row_max > mid_row, row_max - rows, row_max)               # :78-85
    col_shift = torch.where(col_max > mid_col, col_max - cols, col_max)
    if upsample_factor == 1:
        rfzero = (F_ref.abs() ** 2).sum(dim=(1, 2)) / (rows * cols)                   # :87-92
        rgzero = (F_reg.abs() ** 2).sum(dim=(1, 2)) / (rows * cols)
        ccmax = cc.reshape(B, -1).gather(1, flat[:, None])[:, 0]
        error = torch.sqrt(torch.abs(1.0 - (ccmax * ccmax.conj()).real / (rgzero * rfzero)))
        return row_shift, col_shift, error, torch.atan2(ccmax.imag, ccmax.real)
    usf = float(upsample_factor)
    # (torch divides a CUDA tensor by a Python scalar as a multiplication by its reciprocal -- 1 ulp off numpy's
    #  true division, e.g. -7 / 20; a tensor divisor keeps the IEEE division)
    usf_t = torch.full_like(row_shift, usf)
    row_shift = torch.round(row_shift * usf) / usf_t                                  # :97-98
    col_shift = torch.round(col_shift * usf) / usf_t
    up = int(np.ceil(usf * 1.5))                                                      # :99
    dftshift = float(np.fix(up / 2))                                                  # :101
    norm = mid_row * mid_col * usf ** 2
    cu = _dftups(F_reg * F_ref.conj(), up, up, usf, dftshift - row_shift * usf, dftshift - col_shift * usf).conj() / norm
    flat = cu.real.reshape(B, -1).argmax(dim=1)                                       # :111-112
    r_max = (flat // up).to(torch.float64) - dftshift
    c_max = (flat % up).to(torch.float64) - dftshift
    row_shift = row_shift + r_max / usf_t                                             # :115-116
    col_shift = col_shift + c_max / usf_t
    ccmax = cu.reshape(B, -1).gather(1, flat[:, None])[:, 0]
    rg00 = (F_ref * F_ref.conj()).sum(dim=(1, 2)) / norm                              # _dftups(..., 1, 1, usf): all-ones kernels
    rf00 = (F_reg * F_reg.conj()).sum(dim=(1, 2)) / norm
    error = torch.sqrt(torch.abs(1.0 - ccmax * ccmax.conj() / (rg00 * rf00)))         # :122-123
    diffphase = torch.atan2(ccmax.imag, ccmax.real)
    if mid_row == 1:                                                                  # :128-131
        row_shift = torch.zeros_like(row_shift)
    if mid_col == 1:
        col_shift = torch.zeros_like(col_shift)
    return row_shift, col_shift, error, diffphase


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
    phase_correlate(frames[f], frames[f + 1]) -- every consecutive pair of the stack in one batch.
    -> list of (d_h, d_w) float tuples."""
    engine.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    fr = _device_f64(frames, dev)
    if fr.dim() != 3:
        raise ValueError("frames must be [n_frames, rows, cols]")
    out = [(0, 0)]
    if fr.shape[0] > 1:
        r, c, _, _ = phase_correlate_batch(fr[:-1], fr[1:], upsample_factor)
        out += list(zip(r.cpu().tolist(), c.cpu().tolist()))
    return out
