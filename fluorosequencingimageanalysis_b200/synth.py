"""
Synthetic TIRF frame generator used by the tests and by bench.py.

This is the recipe of SURVEY.md section 8(d) ("Synthetic frame generator"), reproduced
exactly so the known-answer numbers of SURVEY.md App. D (KAT-2: seed 0 -> 4988
candidates, threshold 12744327.49987743) hold:

    rng = default_rng(seed); flat background 400; centres cr~U(8,H-8), cc~U(8,W-8)
    (all cr first, then all cc, then amplitudes A~U(800,4000)); isotropic Gaussian
    sigma=1.5 added on the 15x15 patch around the rounded centre; Poisson shot noise
    plus N(0,10) read noise; rint, clip to [0,65535], uint16.

Host-side numpy only; nothing here is on the measured path.
"""
import numpy as np


def spot_layout(seed, H=512, W=512, n_spots=500, amp_lo=800.0, amp_hi=4000.0):
    """Draw (rows, cols, amplitudes) in the order section 8(d) prescribes. Returns rng too."""
    rng = np.random.default_rng(seed)
    cr = rng.uniform(8, H - 8, n_spots)
    cc = rng.uniform(8, W - 8, n_spots)
    amp = rng.uniform(amp_lo, amp_hi, n_spots)
    return rng, cr, cc, amp


def render_clean(cr, cc, amp, H=512, W=512, sigma=1.5, bg=400.0):
    """Noise-free expectation image: background + Gaussians on 15x15 patches."""
    img = np.full((H, W), bg, dtype=np.float64)
    for r0, c0, a in zip(cr, cc, amp):
        ri, ci = int(round(r0)), int(round(c0))
        rlo, rhi = max(ri - 7, 0), min(ri + 8, H)
        clo, chi = max(ci - 7, 0), min(ci + 8, W)
        rr, cc_ = np.mgrid[rlo:rhi, clo:chi]
        img[rlo:rhi, clo:chi] += a * np.exp(-((rr - r0) ** 2 + (cc_ - c0) ** 2) / (2 * sigma ** 2))
    return img


def add_noise(clean, rng, read_sigma=10.0):
    img = rng.poisson(clean) + rng.normal(0, read_sigma, clean.shape)
    return np.clip(np.rint(img), 0, 65535).astype(np.uint16)


def synth_frame(seed, H=512, W=512, n_spots=500, sigma=1.5, bg=400.0):
    """One frame, config 1 of BASELINE.json (seed 0 -> KAT-2)."""
    rng, cr, cc, amp = spot_layout(seed, H, W, n_spots)
    return add_noise(render_clean(cr, cc, amp, H, W, sigma, bg), rng)


def synth_frame_with_truth(seed, H=512, W=512, n_spots=500, sigma=1.5, bg=400.0):
    rng, cr, cc, amp = spot_layout(seed, H, W, n_spots)
    return add_noise(render_clean(cr, cc, amp, H, W, sigma, bg), rng), cr, cc, amp


def synth_timetrace(seed, n_frames=40, H=512, W=512, n_spots=500, sigma=1.5, bg=400.0):
    """Config 2: one field, same centres in every frame, fresh noise per frame
    (noise seed = 1000 + frame)."""
    _, cr, cc, amp = spot_layout(seed, H, W, n_spots)
    clean = render_clean(cr, cc, amp, H, W, sigma, bg)
    out = np.empty((n_frames, H, W), dtype=np.uint16)
    for f in range(n_frames):
        out[f] = add_noise(clean, np.random.default_rng(1000 + f))
    return out


def synth_experiment(seed, n_fields=100, n_cycles=10, H=512, W=512, n_spots=1000,
                     sigma=1.5, bg=400.0, p_drop=0.1):
    """Config 3/5: (field, cycle) frames; each cycle drops each surviving spot with
    probability p_drop (Edman-like ON->OFF tracks).  Returns uint16 [n_fields, n_cycles, H, W]."""
    out = np.empty((n_fields, n_cycles, H, W), dtype=np.uint16)
    for fld in range(n_fields):
        rng, cr, cc, amp = spot_layout(seed * 100003 + fld, H, W, n_spots)
        alive = np.ones(n_spots, dtype=bool)
        for cyc in range(n_cycles):
            if cyc:
                alive &= rng.uniform(size=n_spots) >= p_drop
            clean = render_clean(cr[alive], cc[alive], amp[alive], H, W, sigma, bg)
            out[fld, cyc] = add_noise(clean, rng)
    return out


def dihedral_variants(stack):
    """The 8 symmetries of the square applied to the last two axes of ``stack`` [..., H, H] -> list of 8 arrays.
    A flipped / transposed synthetic frame is again a synthetic frame of the same recipe (isotropic spots, i.i.d.
    noise), so a few generated fields give 8x as many different ones for the cost of a copy (bench.py's frame pool)."""
    if stack.shape[-1] != stack.shape[-2]:
        raise ValueError("dihedral variants need square frames")
    out = []
    for t in (stack, np.swapaxes(stack, -1, -2)):
        out += [t, t[..., ::-1, :], t[..., :, ::-1], t[..., ::-1, ::-1]]
    return [np.ascontiguousarray(v) for v in out]


def experiment_field_pool(seed, n_fields, n_cycles=10, H=512, W=512, n_spots=1000, **kw):
    """[n_fields, n_cycles, H, W] uint16: ceil(n_fields / 8) fields by synth_experiment (config 3 / 5 recipe), the
    rest their dihedral variants."""
    n_base = -(-n_fields // 8)
    base = synth_experiment(seed, n_fields=n_base, n_cycles=n_cycles, H=H, W=W, n_spots=n_spots, **kw)
    pool = np.concatenate(dihedral_variants(base), axis=0) if H == W else np.concatenate([base] * 8, axis=0)
    return pool[:n_fields]


def cut_windows(frame, cr, cc, win=11):
    """Pre-cut win x win float64 windows centred on the rounded true centres
    (the direct-gaussfit 11x11 variant of config 1 / config 4)."""
    h = win // 2
    H, W = frame.shape
    out = []
    for r0, c0 in zip(cr, cc):
        ri, ci = int(round(r0)), int(round(c0))
        if ri - h < 0 or ci - h < 0 or ri + h + 1 > H or ci + h + 1 > W:
            continue
        out.append(frame[ri - h:ri + h + 1, ci - h:ci + h + 1].astype(np.float64))
    return np.stack(out)


def shifted_pair(seed, H, W, dy, dx, n_spots, sigma=1.5, bg=400.0):
    """Two frames of one spot layout, the second drifted by a sub-pixel (dy, dx), independent noise
    (registration test input: phase_correlate)."""
    _, cr, cc, amp = spot_layout(seed, H, W, n_spots)
    a = add_noise(render_clean(cr, cc, amp, H, W, sigma, bg), np.random.default_rng(100 + seed))
    cr2, cc2 = np.clip(cr + dy, 8, H - 8), np.clip(cc + dx, 8, W - 8)
    b = add_noise(render_clean(cr2, cc2, amp, H, W, sigma, bg), np.random.default_rng(200 + seed))
    return a, b
