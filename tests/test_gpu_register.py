"""GPU parity: the phase_correlate mirror (fsq_phase_correlate: cuFFT for the FFTs, own kernels for the rest) against outputs
of the reference's own phase_correlate.py (tests/golden/phase_correlate.npz).  Shifts exact (they are multiples of
1/upsample_factor); error to 1e-6 absolute (1 - |CC|^2/(rg rf) cancels to ~1e-8 for aligned frames), diffphase to 1e-9."""
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


def test_phase_correlate_batch_shapes_dtypes_and_odd_sizes():
    """The C entry point directly: float64 and uint16 inputs give the same answer, odd and rectangular frames, upsample
    factors 1 / 7 / 20, a batch equals its pairs one by one."""
    from fluorosequencingimageanalysis_b200 import phase_correlate as pc, synth
    pairs = [synth.shifted_pair(11 + k, 97, 130, 0.4 * k - 1.0, 1.7 - 0.6 * k, 40) for k in range(4)]
    ref = np.stack([p[0] for p in pairs])
    reg = np.stack([p[1] for p in pairs])
    for usf in (1, 7, 20):
        r, c, e, d = (v.cpu().numpy() for v in pc.phase_correlate_batch(ref, reg, usf))
        r64, c64, e64, d64 = (v.cpu().numpy() for v in pc.phase_correlate_batch(ref.astype(np.float64), reg.astype(np.float64), usf))
        assert np.array_equal(r, r64) and np.array_equal(c, c64) and np.array_equal(e, e64) and np.array_equal(d, d64)
        for k in range(4):
            one = pc.phase_correlate(ref[k], reg[k], usf)
            assert one == (r[k], c[k], e[k], d[k])
        if usf == 20:                                    # the drift that was put in comes back to 1/20 px
            for k in range(4):
                assert abs(-r[k] - (0.4 * k - 1.0)) <= 0.05 + 1e-12 and abs(-c[k] - (1.7 - 0.6 * k)) <= 0.05 + 1e-12
