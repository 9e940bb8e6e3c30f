"""TEST INFRASTRUCTURE -- CPU restatement of the reference's luminosity-centroid tracker.

Follows flexlibrary.py:1173-1317 (Experiment.next_frame_spot_by_luminosity_centroid,
Experiment.luminosity_centroid_particle_tracking), Spot.__init__ (flexlibrary.py:97-119), Spot.image_slice
(:123-147), Experiment.unapply_offset (:614-617) and pflib.illumina_s_n (pflib.py:261-281), with
scipy.ndimage.center_of_mass as the reference imports it (:61).

PINNED: the head of the reference's flexlibrary (lines 1-1319: Spot, Image, Experiment) is built into
oracle/_ref/flexlibrary_head.py by oracle/build_ref.py (import stubs for absent modules, `(size - 1) / 2` as the
integer division it is in Python 2) and tests/test_oracle_pins.py checks `track` and `greedy_particle_tracking` below
against the reference's own Experiment.luminosity_centroid_particle_tracking / greedy_particle_tracking on seeded
movies and spot lists (identical positions, traces, trace order, drop-out counts).  Only tests/ may import it.
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


# ----------------------------------------------------------------------------- greedy cross-frame tracking
def accumulate_offsets(offsets):
    """flexlibrary.py:566-593 (Python sums, in sequence)."""
    if tuple(offsets[0]) != (0, 0):
        raise ValueError("The first image's offset must be (0, 0) by definiton.")
    return [(sum([o[0] for o in offsets[:f + 1]]), sum([o[1] for o in offsets[:f + 1]])) for f in range(len(offsets))]


def discard_dropouts(spots, spot_offset, frame_cumulative_offsets, image_shape, spot_radius=0):
    """flexlibrary.py:626-677 for the spots [(h, w)] of one frame -> (kept indices, number discarded)."""
    kept, number_discarded = [], 0
    for i, (h, w) in enumerate(spots):
        oh, ow = h + spot_offset[0], w + spot_offset[1]
        for off in frame_cumulative_offsets:
            gh, gw = oh - off[0], ow - off[1]
            if not (spot_radius <= gh < image_shape[0] - 0.5 - spot_radius and
                    spot_radius <= gw < image_shape[1] - 0.5 - spot_radius):
                number_discarded += 1
                break
        else:
            kept.append(i)
    return kept, number_discarded


def greedy_particle_tracking(frame_spots, frame_shape, candidate_radius=2, offsets=None, spot_radius=0):
    """flexlibrary.py:680-1027.  frame_spots: per frame a list of (h, w).  Returns (traces, total_discarded) with
    every trace a list of length n_frames holding the spot's index in its frame's list, or None.
    The reference's object arrays of dictionaries are restated as dictionaries keyed by the rounded pixel; every
    loop that walks an array with numpy.ndenumerate walks the keys in raster order here."""
    from scipy.spatial.distance import euclidean
    F = len(frame_spots)
    if offsets is None:
        offsets = [(0, 0) for _ in range(F)]
    cum = accumulate_offsets(offsets)
    H, W = frame_shape
    total_discarded = 0
    kept = []
    for f in range(F):
        k, nd = discard_dropouts(frame_spots[f], cum[f], cum, frame_shape, spot_radius)
        kept.append(k)
        total_discarded += nd
    bins = [dict() for _ in range(F)]                       # (rh, rw) -> {'spt': index, 'a_L': ..., 'd_L': ...}
    for f in range(F):
        for i in kept[f]:
            h, w = frame_spots[f][i][0] + cum[f][0], frame_spots[f][i][1] + cum[f][1]
            rh, rw = int(py2_round(h)), int(py2_round(w))
            assert (rh, rw) not in bins[f], "%s is already filled in frame_bins[%d]" % ((rh, rw), f)
            bins[f][(rh, rw)] = {'spt': i, 's_L': (f, rh, rw), 'a_L': None, 'd_L': None}
    cache = {}
    pos = lambda f, i: (frame_spots[f][i][0] + cum[f][0], frame_spots[f][i][1] + cum[f][1])
    for f in range(1, F):
        frame = bins[f]
        for (rh, rw) in sorted(bins[f - 1]):
            cache[(rh, rw)] = {'spt': bins[f - 1][(rh, rw)]['spt'], 's_L': (f - 1, rh, rw)}
        pairs = []
        for (ah, aw) in sorted(cache):
            a = cache[(ah, aw)]
            aaf = a['s_L'][0]
            for dh in range(max(ah - candidate_radius - 2, 0), min(ah + candidate_radius + 3, H)):
                for dw in range(max(aw - candidate_radius - 2, 0), min(aw + candidate_radius + 3, W)):
                    if (dh, dw) not in frame:
                        continue
                    d = frame[(dh, dw)]
                    distance = euclidean(pos(aaf, a['spt']), pos(f, d['spt']))
                    if distance < candidate_radius:
                        pairs.append((aaf, ah, aw, dh, dw, distance))
        pairs = sorted(pairs, key=lambda x: x[5])
        for (aaf, ah, aw, dh, dw, distance) in pairs:
            if (ah, aw) not in cache:
                continue                                      # ancestor has been paired
            elif frame[(dh, dw)]['a_L'] is not None:
                continue                                      # descendant has been paired
            frame[(dh, dw)]['a_L'] = (aaf, ah, aw)
            assert bins[aaf][(ah, aw)]['d_L'] is None
            bins[aaf][(ah, aw)]['d_L'] = (f, dh, dw)
            del cache[(ah, aw)]
    traces = []
    for f in range(F):
        for key in sorted(bins[f]):
            b = bins[f][key]
            if b['a_L'] is not None:
                continue
            trace = [None] * F
            trace[f] = b['spt']
            cur = b
            while cur['d_L'] is not None:
                df, dh, dw = cur['d_L']
                cur = bins[df][(dh, dw)]
                trace[df] = cur['spt']
            traces.append(trace)
    return traces, total_discarded
