"""GPU parity: the phase_correlate mirror (cuFFT / cuBLAS through torch) against outputs of the reference's own
phase_correlate.py (tests/golden/phase_correlate.npz).  Shifts exact (they are multiples of 1/upsample_factor);
error to 1e-6 absolute (1 - |CC|^2/(rg rf) cancels to ~1e-8 for aligned frames), diffphase to 1e-9."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def test_phase_correlate_matches_reference_outputs():
    from fluorosequencingimageanalysis_b200 import phase_correlate as pc, synth
    g = golden("phase_correlate.npz")
    for k, (seed, H, W, dy, dx, n) in enumerate(g["cases"]):
        a, b = synth.shifted_pair(int(seed), int(H), int(W), dy, dx, int(n))
        for usf, key in ((1, "out1"), (20, "out20")):
            r, c, e, d = pc.phase_correlate(a, b, usf)
            want = g[key][k]
            assert (r, c) == (want[0], want[1]), (k, usf, r, c, want)
            assert abs(e - want[2]) < 1e-6 and abs(d - want[3]) < 1e-9
    with pytest.raises(ValueError):
        pc.phase_correlate(a, b[:-1], 1)                                          # phase_correlate.py:57-58
    with pytest.raises(ValueError):
        pc.phase_correlate(a[0], b[0], 1)                                         # :61-62


def test_offsets_from_frames_batches_consecutive_pairs():
    from fluorosequencingimageanalysis_b200 import phase_correlate as pc, synth
    a, b = synth.shifted_pair(3, 256, 256, 0.35, -1.6, 150)
    c, _ = synth.shifted_pair(2, 256, 256, 0.0, 0.0, 150)
    stack = np.stack([a, b, a, c])
    off = pc.offsets_from_frames(stack, upsample_factor=20)
    assert off[0] == (0, 0) and len(off) == 4
    for f in range(3):
        assert off[f + 1] == pc.phase_correlate(stack[f], stack[f + 1], 20)[:2]
    assert off[1] == (-0.35, 1.6) and off[2] == (0.35, -1.6)
