"""TEST INFRASTRUCTURE -- CPU restatement of the reference's luminosity-centroid tracker.

Follows flexlibrary.py:1173-1317 (Experiment.next_frame_spot_by_luminosity_centroid,
Experiment.luminosity_centroid_particle_tracking), Spot.__init__ (flexlibrary.py:97-119), Spot.image_slice
(:123-147), Experiment.unapply_offset (:614-617) and pflib.illumina_s_n (pflib.py:261-281), with
scipy.ndimage.center_of_mass as the reference imports it (:61).

PARITY UNPINNED: flexlibrary is Python 2 and imports photutils / skimage (absent here), so this file was restated
by reading, not validated against a run of the reference.  Only tests/ may import it.
"""
import math

import numpy as np
from scipy.ndimage import center_of_mass

from .pflib_oracle import illumina_s_n, py2_round


class SpotError(AttributeError):
    pass


def make_spot(shape, h, w, size):
    """Spot.__init__ with gaussian_fit=None (flexlibrary.py:97-119): returns (h, w) or raises."""
    if size % 2 == 0:
        raise SpotError("Spot.size must be odd.")
    half = (size - 1) // 2
    if not (0 <= h - half and h + half < shape[0] and 0 <= w - half and w + half < shape[1]):
        raise SpotError("Spot area does not fit into parent_Image.image.shape")
    return (h, w)


def spot_slice(image, h, w, size):
    half = (size - 1) // 2                                                   # flexlibrary.py:140-147
    return image[max(0, h - half):min(image.shape[0], h + half + 1), max(0, w - half):min(image.shape[1], w + half + 1)]


def next_frame_spot(spot, size, next_image, offset=(0, 0), search_radius=3, s_n_cutoff=3.0):
    """flexlibrary.py:1173-1259 -> ((h, w) or None, state, s_n): state 1 centroid, 2 stayed at the prior position."""
    o_h, o_w = spot[0] - offset[0], spot[1] - offset[1]                      # unapply_offset
    lo_h, lo_w = o_h - search_radius, o_w - search_radius
    if lo_h < 0 or lo_w < 0:                                                 # a negative slice start wraps around in numpy:
        image_slice = next_image[0:0, 0:0]                                   # the reference's slice is then empty or mis-shaped
    else:
        image_slice = next_image[lo_h:o_h + search_radius + 1, lo_w:o_w + search_radius + 1]
    if image_slice.shape != (1 + 2 * search_radius, 1 + 2 * search_radius):
        return None, 0, float("nan")
    c_h, c_w = center_of_mass(image_slice)
    r_c_h = int(py2_round(c_h + o_h - search_radius))
    r_c_w = int(py2_round(c_w + o_w - search_radius))
    try:
        nxt = make_spot(next_image.shape, r_c_h, r_c_w, size)
    except AttributeError:
        return None, 0, float("nan")
    with np.errstate(divide="ignore", invalid="ignore"):
        sn = float(illumina_s_n(spot_slice(next_image, nxt[0], nxt[1], size)))
    if sn < s_n_cutoff:
        try:
            return make_spot(next_image.shape, int(py2_round(spot[0])), int(py2_round(spot[1])), size), 2, sn
        except AttributeError:
            return None, 0, sn
    return nxt, 1, sn


def track(frames, initial_spots, size=5, search_radius=3, s_n_cutoff=3.0, offsets=None):
    """flexlibrary.py:1262-1317 -> (track_hw [n,F,2] with -1 for None, state [n,F], s_n [n,F])."""
    n, F = len(initial_spots), len(frames)
    hw = -np.ones((n, F, 2), dtype=np.int64)
    state = np.zeros((n, F), dtype=np.uint8)
    sn = np.full((n, F), np.nan)
    for i, spot in enumerate(initial_spots):
        prior = (int(spot[0]), int(spot[1]))
        hw[i, 0] = prior
        state[i, 0] = 3
        for f in range(1, F):
            offset = offsets[f] if offsets is not None else (0, 0)
            nxt, st, s = next_frame_spot(prior, size, frames[f], offset, search_radius, s_n_cutoff)
            state[i, f], sn[i, f] = st, s
            if nxt is not None:
                hw[i, f] = nxt
                prior = nxt
    return hw, state, sn
