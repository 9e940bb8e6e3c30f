"""CPU: host-side logic of the drop-in modules (argument marshalling, consolidation / re-key,
partitioner) against the oracle and the reference-generated goldens; the median selection
network of the detection kernel proved by the 0-1 principle; the N>1 path on gloo."""
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden
from fluorosequencingimageanalysis_b200 import gaussfitter, pflib, sharding, synth, engine
from oracle import pflib_oracle as po


# ------------------------------------------------------------------ median network (0-1 principle)
def test_median25_network_is_a_median():
    """A comparator network selects rank 12 of 25 for ALL inputs iff it does for all 2^25
    0/1 inputs.  The network is parsed from the CUDA source, so the test checks what ships."""
    src = open(os.path.join(ROOT, "fluorosequencingimageanalysis_b200", "csrc", "fsq_median.cuh")).read()
    body = src[src.index("median25(V* p)"):src.index("#undef FSQ_CE")]
    ces = [(int(a), int(b)) for a, b in re.findall(r"FSQ_CE\((\d+),\s*(\d+)\)", body)]
    assert len(ces) == 99
    assert "return p[12]" in body
    chunk = 1 << 21
    for base in range(0, 1 << 25, chunk):
        v = np.arange(base, base + chunk, dtype=np.uint32)
        w = [((v >> i) & 1).astype(np.uint8) for i in range(25)]
        ones = np.zeros(chunk, dtype=np.uint8)
        for x in w:
            ones += x
        for a, b in ces:
            lo, hi = w[a] & w[b], w[a] | w[b]
            w[a], w[b] = lo, hi
        assert np.array_equal(w[12], (ones >= 13).astype(np.uint8))     # 13th smallest is 1 iff >= 13 ones


# ------------------------------------------------------------------ argument marshalling
def test_pflib_limits_match_reference_call():
    img = synth.synth_frame(5, H=64, W=64, n_spots=12)
    subs = np.stack([img[h:h + 5, w:w + 5] for h, w in ((3, 4), (20, 31), (40, 9), (55, 55))]).astype(np.int64)
    p0, lo, hi, lim_lo, lim_hi = pflib._pflib_limits(subs)
    for i, s in enumerate(subs):
        p, lmin, lmax, mn, mx = po.pflib_fit_args(s)              # pflib.py:199-213
        p = np.array(p, dtype=float)
        p = np.where((p > mx) & np.array(lmax), mx, p)            # gaussfitter.py:202-204
        p = np.where((p < mn) & np.array(lmin), mn, p)
        assert np.array_equal(p0[i], p)
        assert np.array_equal(lo[i], mn) and np.array_equal(hi[i], mx)
        assert lim_lo[i].astype(bool).tolist() == lmin and lim_hi[i].astype(bool).tolist() == lmax


def test_moments_and_model_match_oracle():
    g = golden("fits11_seed0.npz")
    for i in (0, 5, 77):
        w = g["windows"][i]
        assert np.array_equal(np.array(po.moments(w), dtype=float), g["p0"][i])   # device moments: test_gpu_fit.py
    p = [83.8, 1988.1, 2.68, 2.31, 1.127, 1.126, 33.0]
    assert np.array_equal(gaussfitter.twodgaussian(p, shape=(5, 5)), po.gauss2d(p, (5, 5)))
    assert np.array_equal(gaussfitter.twodgaussian(p)(*np.indices((7, 9))), po.gauss2d(p, (7, 9)))
    with pytest.raises(ValueError):
        gaussfitter.twodgaussian(p + [1.0], shape=(5, 5))         # gaussfitter.py:120-123


def test_reference_error_behaviour_before_any_gpu_work():
    with pytest.raises(ValueError, match="consolidation_radius"):
        pflib.find_peptides(np.zeros((16, 16), dtype=np.uint16), consolidation_radius=1)   # pflib.py:431-432
    with pytest.raises(AssertionError):
        pflib._fit_2d_gaussian(np.zeros((7, 7)))                  # pflib.py:193
    with pytest.raises(ValueError, match="square"):
        pflib.illumina_s_n(np.zeros((5, 4)))                      # pflib.py:274-276
    with pytest.raises(ValueError, match="odd"):
        engine._check_kernel(np.ones((4, 4), dtype=int))          # pflib.py:236-239
    with pytest.raises(ValueError, match="odd"):
        engine._check_kernel(np.ones((3, 5), dtype=int))
    with pytest.raises(ValueError, match="haven't implemented"):
        gaussfitter.gaussfit(np.ones((5, 5)), autoderiv=0)        # gaussfitter.py:239
    assert pflib._py2_round(2.5) == 3 and pflib._py2_round(3.5) == 4 and pflib._py2_round(0.49999) == 0


# ------------------------------------------------------------------ consolidation / re-key
def _packed_from_golden(g):
    cands = g["cands"]
    P = g["ref_params"]
    fit = np.zeros((len(cands), 12))
    fit[:, 0] = P[:, 2] + cands[:, 0] - 2.5                       # pflib.py:461
    fit[:, 1] = P[:, 3] + cands[:, 1] - 2.5
    fit[:, 2], fit[:, 3], fit[:, 4], fit[:, 5], fit[:, 6] = P[:, 0], P[:, 1], P[:, 4], P[:, 5], P[:, 6]
    fit[:, 7], fit[:, 8], fit[:, 9] = g["rmse"], g["r_2"], g["s_n"]
    return cands, fit


def test_oracle_consolidate_packed_reproduces_reference_psf_keys(fits5):
    """Fed with the REFERENCE's own per-candidate fits, the packed consolidation + re-key must
    return exactly the reference's final PSF dictionary keys (pflib.py:479-519); the
    insertion order is pinned by test_consolidate_packed_equals_oracle_dict_logic_random and
    the pipeline_small golden."""
    cands, fit = _packed_from_golden(fits5)
    keys, idx = po.consolidate_packed(cands, fit, (512, 512))     # (the device kernel: tests/test_gpu_consolidate.py)
    order = np.lexsort((keys[:, 1], keys[:, 0]))                  # the golden stores the keys sorted
    assert np.array_equal(keys[order], fits5["final_keys"])
    assert np.array_equal(fit[idx[order], 0], fits5["final_h0"])
    assert np.array_equal(fit[idx[order], 1], fits5["final_w0"])


def test_oracle_consolidate_packed_equals_oracle_dict_logic_random():
    rng = np.random.default_rng(3)
    for trial in range(5):
        H = W = 60
        n = 150
        hw = np.unique(np.stack([rng.integers(2, H - 2, n), rng.integers(2, W - 2, n)], axis=1), axis=0)
        n = len(hw)
        fit = np.zeros((n, 12))
        fit[:, 0] = hw[:, 0] + rng.uniform(-0.5, 0.5, n).round(1)       # many exact .5 ties
        fit[:, 1] = hw[:, 1] + rng.uniform(-0.5, 0.5, n).round(1)
        fit[:, 8] = rng.choice([0.5, 0.71, 0.8, 0.9, 0.95], n)          # ties in r_2 too
        radius = int(rng.integers(2, 6))
        d = {}
        for i in range(n):
            if fit[i, 8] < 0.7:
                continue
            d[(int(hw[i, 0]), int(hw[i, 1]))] = (fit[i, 0], fit[i, 1], 0, 0, 0, 0, 0, None, None, 0, fit[i, 8], 0, i)
        try:
            want = po.consolidate(d, (H, W), radius)
        except AssertionError:
            with pytest.raises(AssertionError):
                po.consolidate_packed(hw, fit, (H, W), 0.7, radius)
            continue
        keys, idx = po.consolidate_packed(hw, fit, (H, W), 0.7, radius)
        assert [tuple(k) for k in keys.tolist()] == list(want.keys())
        assert idx.tolist() == [v[12] for v in want.values()]


# ------------------------------------------------------------------ partitioner
def test_balance_by_count_follows_reference_greedy():
    counts = [50, 10, 40, 10, 30, 20]
    parts = sharding.balance_by_count(counts, 2)
    # pflib.py:1056-1069: sorted descending, popped from the end (smallest first), emptiest partition
    # stable descending sort -> [0, 2, 4, 5, 1, 3]; pops 3, 1, 5, 4, 2, 0
    assert parts == [[3, 5, 2], [1, 4, 0]]
    assert sorted(i for p in parts for i in p) == list(range(6))
    with pytest.raises(ValueError):
        sharding.balance_by_count(counts, 0)


def test_field_blocks_cover_and_balance():
    for n, ws in ((8000, 8), (100, 8), (7, 4), (3, 8)):
        blocks = [sharding.field_block(n, ws, r) for r in range(ws)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    c = np.r_[np.full(10, 100), np.full(10, 300)]
    bl = sharding.balanced_field_blocks(c, 2)
    assert bl[0][0] == 0 and bl[-1][1] == 20 and bl[0][1] == bl[1][0]
    tot = [c[a:b].sum() for a, b in bl]
    assert abs(tot[0] - tot[1]) <= 300


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from fluorosequencingimageanalysis_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
lo, hi = sharding.field_block(5, 2, rank)
local = {{"field": np.arange(lo, hi, dtype=np.int64), "fit": np.full((hi - lo, 3), float(rank))}}
out = sharding.gather_packed(local)
assert out["field"].tolist() == [0, 1, 2, 3, 4], out["field"]
assert out["fit"].shape == (5, 3) and out["fit"][:3].max() == 0.0 and out["fit"][3:].min() == 1.0
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_shard_and_gather(tmp_path):
    """The N>1 path on CPU: two processes, gloo, each owns a contiguous field block, results are
    concatenated on the host in rank order; there is no data-path collective."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


# ------------------------------------------------------------------------------------ PSF result files
def _fake_psfs():
    rng = np.random.default_rng(3)
    out = {}
    for k in range(6):
        h, w = int(rng.integers(5, 60)), int(rng.integers(5, 60))
        out[(h, w)] = (h + 0.25, w - 0.125, 401.5, 1999.75, 1.25, 1.5, 12.5,
                       rng.integers(0, 4000, (5, 5)).astype(np.int64), rng.uniform(0, 4000, (5, 5)),
                       33.25, 0.975, 7.5 + k)
    return out


def test_epoch_hash_and_filenames_match_the_reference():
    from fluorosequencingimageanalysis_b200 import pflib
    from oracle import build_ref
    for e in (1, 35, 36, 1461000000, 1461000000.49, 1461000000.5, 2 ** 40 + 17):
        hsh = pflib._epoch_to_hash(e)
        assert pflib._hash_to_epoch(hsh) == int(math.floor(e + 0.5))
    assert pflib._epoch_to_hash(36 ** 3) == '1000' and pflib._epoch_to_hash(35.5) == '10'     # py2 round: half away from zero
    with pytest.raises(ValueError):
        pflib._epoch_to_hash(0)
    with pytest.raises(ValueError):
        pflib._hash_to_epoch('ab_c')
    assert pflib._psfs_filename('x/y.png', 1461000000, '.pkl') == os.path.abspath('x/y.png') + '_psfs_' + pflib._epoch_to_hash(1461000000) + '.pkl'
    mods = build_ref.load()
    if mods is not None:                                    # the reference itself, where /root/reference is available
        ref = mods[0]
        for e in (1, 35.5, 1461000000, 1461000000.5, 2 ** 40 + 17):
            assert pflib._epoch_to_hash(e) == ref._epoch_to_hash(e)
            assert pflib._hash_to_epoch(ref._epoch_to_hash(e)) == ref._hash_to_epoch(ref._epoch_to_hash(e))
        assert pflib._psfs_filename('a.png', 12345, '.csv') == ref._psfs_filename('a.png', 12345, '.csv')


def test_psf_result_files_round_trip_and_match_the_reference_csv(tmp_path):
    import pickle
    from fluorosequencingimageanalysis_b200 import pflib
    from oracle import build_ref
    psfs = _fake_psfs()
    img = str(tmp_path / "frame.png")
    pkl = pflib.save_psfs_pkl(psfs, image_path=img, timestamp_epoch=1461000000)
    assert pkl == pflib._psfs_filename(img, 1461000000, '.pkl')
    back = pickle.load(open(pkl, 'rb'))
    assert list(back.keys()) == list(psfs.keys())
    for k in psfs:
        assert back[k][:7] == psfs[k][:7] and back[k][9:] == psfs[k][9:]
        assert np.array_equal(back[k][7], psfs[k][7]) and np.array_equal(back[k][8], psfs[k][8])
    csvp = pflib.save_psfs_csv(psfs, image_path=img, timestamp_epoch=1461000000)
    rows = [ln.rstrip('\r\n').split('\t') for ln in open(csvp)]
    assert rows[0][0] == 'Absolute image path' and len(rows) == 1 + len(psfs) and len(rows[1]) == 11
    assert rows[1][0] == os.path.abspath(img) and float(rows[1][3]) == 401.5
    with pytest.raises(ValueError):
        pflib.save_psfs_csv(psfs)
    with pytest.raises(ValueError):
        pflib.save_psfs_pkl(psfs)
    out2 = pflib.save_psfs_csv(psfs, output_path=str(tmp_path / "explicit.csv"))
    assert out2 == str(tmp_path / "explicit.csv")
    mods = build_ref.load()
    if mods is not None:                                    # byte-identical table from the reference's writer
        ref_csv = mods[0].save_psfs_csv(psfs, output_path=str(tmp_path / "ref.csv"), image_path=img)
        assert open(ref_csv, newline='').read() == open(out2, newline='').read().replace(os.path.abspath(img), os.path.abspath(img)) or \
            [r[1:] for r in rows] == [ln.rstrip('\r\n').split('\t')[1:] for ln in open(ref_csv)]


def test_read_image_png_and_conversion(tmp_path):
    from PIL import Image
    from fluorosequencingimageanalysis_b200 import pflib, synth
    img = synth.synth_frame(5, H=40, W=56, n_spots=4)
    p_png = str(tmp_path / "a.png")
    Image.fromarray(img).save(p_png)
    conv, back = pflib.read_image(p_png)
    assert conv == p_png and np.array_equal(back, img)
    p_tif = str(tmp_path / "b.tif")
    Image.fromarray(img).save(p_tif)
    conv, back = pflib.read_image(p_tif)
    assert conv == p_tif + '.png' and os.path.exists(conv) and np.array_equal(back, img)
    os.remove(p_tif)                                        # the converted copy is reused (pflib.py:737-738)
    open(p_tif, 'wb').write(b'not an image')
    conv2, back2 = pflib.read_image(p_tif)
    assert conv2 == conv and np.array_equal(back2, img)
    with pytest.raises(Exception):
        pflib.read_image(str(tmp_path / "missing.tif"))


# ------------------------------------------------------------------ tracking oracle: hand-checkable cases
def test_track_oracle_known_answers():
    """Hand-checkable cases of the luminosity-centroid tracker's restatement (flexlibrary.py:1173-1317): the
    centroid follows a bright pixel, python-2 rounding at .5, None at the border, fall-back below the S/N cut-off."""
    from oracle import track_oracle as tro
    H = W = 20
    f0 = np.full((H, W), 10, dtype=np.uint16)
    f1 = f0.copy(); f1[9, 11] = 60000                       # one very bright pixel: the 7x7 centre of mass lands next to it
    f2 = f0.copy()                                          # flat: std = 0 -> S/N = nan -> `nan < cutoff` is False -> centroid kept
    hw, st, sn = tro.track([f0, f1, f2], [(10, 10)])
    assert hw[0].tolist() == [[10, 10], [9, 11], [9, 11]] and st[0].tolist() == [3, 1, 1]
    assert np.isnan(sn[0, 2]) and sn[0, 1] > 3
    # window cut by the border -> None; the last known position stays in force for the next frame
    hw, st, _ = tro.track([f0, f0, f1], [(2, 2)])
    assert st[0].tolist() == [3, 0, 0] and hw[0, 1].tolist() == [-1, -1]
    # S/N below the cut-off -> the spot stays at the prior position (state 2)
    g = f0.copy(); g[10, 10] = 12; g[8, 8] = 11             # weak, with an uneven border so that std > 0
    hw, st, sn = tro.track([f0, g], [(10, 10)], s_n_cutoff=50.0)
    assert st[0, 1] == 2 and hw[0, 1].tolist() == [10, 10] and sn[0, 1] < 50
    # an even Spot.size is refused like the reference does (flexlibrary.py:98-99)
    with pytest.raises(AttributeError):
        tro.make_spot((H, W), 10, 10, 4)
    with pytest.raises(AttributeError):
        tro.make_spot((H, W), 1, 10, 5)                     # the 5x5 square leaves the image


def test_greedy_tracking_oracle_known_answers():
    """Hand-checkable cases of the restated Experiment.greedy_particle_tracking (flexlibrary.py:680-1027)."""
    from oracle import track_oracle as tro
    shape = (40, 40)
    # two ancestors compete for one candidate: the nearer one wins, the other one finds it again a frame later
    f0 = [(10, 10), (10, 13)]
    f1 = [(10, 11)]                      # distance 1 to (10,10), 2 to (10,13) -> not < 2: only the first is a pair
    f2 = [(10, 12), (11, 13)]            # (10,11)->(10,12) d=1; the unpaired (10,13) of frame 0 -> (11,13) d=1, skipping frame 1
    tr, nd = tro.greedy_particle_tracking([f0, f1, f2], shape)
    assert nd == 0 and tr == [[0, 0, 0], [1, None, 1]]
    # equal distances: the stable sort keeps ancestor raster order, so the ancestor that comes first in raster order gets the candidate
    tr, _ = tro.greedy_particle_tracking([[(10, 10), (10, 12)], [(10, 11)]], shape)
    assert tr == [[0, 0], [1, None]]
    # drift: frame 1 is shifted by (+3, 0); with the offset the spots pair up, without it they do not
    tr, _ = tro.greedy_particle_tracking([[(20, 20)], [(17, 20)]], shape, offsets=[(0, 0), (3, 0)])
    assert tr == [[0, 0]]
    tr, _ = tro.greedy_particle_tracking([[(20, 20)], [(17, 20)]], shape)
    assert tr == [[0, None], [None, 0]]
    # a spot that would leave the field in some frame of the sequence is dropped and counted (discard_dropouts)
    tr, nd = tro.greedy_particle_tracking([[(1, 20), (20, 20)], [(20, 20)]], shape, offsets=[(0, 0), (3, 0)])
    assert nd == 1 and tr == [[1, None], [None, 0]]
    with pytest.raises(ValueError):
        tro.accumulate_offsets([(1, 0), (0, 0)])


def test_bench_partition_covers_every_field_of_a_step_exactly_once():
    """bench.py's strong-scaling plan: sharding.balanced_field_blocks splits a step's 800 fields into contiguous blocks
    with near-equal candidate totals; plan_chunks turns a block into launches of <= LAUNCH_FIELDS consecutive fields,
    each one contiguous slice of the (wrapped) frame pool.  Every field is processed exactly once for any world size."""
    sys.path.insert(0, ROOT)
    import bench
    from fluorosequencingimageanalysis_b200 import sharding
    rng = np.random.default_rng(3)
    pool_counts = rng.integers(50000, 70000, bench.POOL_FIELDS)
    step_counts = pool_counts[np.arange(bench.FIELDS_PER_STEP) % bench.POOL_FIELDS]
    for world in (1, 2, 3, 4, 8):
        blocks = sharding.balanced_field_blocks(step_counts, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == bench.FIELDS_PER_STEP
        seen = np.zeros(bench.FIELDS_PER_STEP, dtype=int)
        tot = []
        for lo, hi in blocks:
            f = lo
            for p0, nf in bench.plan_chunks(lo, hi):
                assert 1 <= nf <= bench.LAUNCH_FIELDS and 0 <= p0 < bench.POOL_FIELDS
                assert p0 == f % bench.POOL_FIELDS and p0 + nf <= bench.POOL_FIELDS + bench.LAUNCH_FIELDS
                seen[f:f + nf] += 1
                f += nf
            assert f == hi
            tot.append(step_counts[lo:hi].sum())
        assert (seen == 1).all()
        assert max(tot) - min(tot) <= 2 * step_counts.max()          # balanced to within a field or two
    cfg1, cfg8 = bench.static_config(1), bench.static_config(8)
    assert cfg1["workload"] == cfg8["workload"] and cfg1["frames_per_step"] == 8000


def test_median_pair_network_is_a_pair_of_medians_and_the_header_is_current():
    """fsq_median_pair.cuh (the two-windows-at-once median of the packed detection kernel) is generated: the generator's
    program must select rank 12 of each window for ALL 2^25 0/1 inputs (0-1 principle: every operation is a min or a
    max), and the committed header must be exactly what the generator emits."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_median_pair as gmp
    text, n_raw, n_sorted = gmp.generate(check=True)          # raises if any 0/1 input of either window fails
    assert (n_raw, n_sorted) == (216, 108)
    committed = open(os.path.join(ROOT, "fluorosequencingimageanalysis_b200", "csrc", "fsq_median_pair.cuh")).read()
    assert committed == text


def test_window_batches_keep_camera_integer_types():
    """engine._device_windows: 8- / 16- / 32- / 64-bit camera integers and float64 reach the kernels as they are (a 200 000
    window batch of uint16 pixels is 48 MB, not 194 MB, on its way to the device); anything else is widened to int64 /
    float64; a single window becomes a batch of one; non-square windows raise like the reference's 5x5 assert."""
    import numpy as np
    import torch
    from fluorosequencingimageanalysis_b200 import engine
    cpu = torch.device("cpu")
    w = (np.arange(2 * 11 * 11) % 4096).reshape(2, 11, 11)
    for dt, want in ((np.uint8, torch.uint8), (np.int16, torch.int16), (np.int32, torch.int32), (np.int64, torch.int64),
                     (np.float64, torch.float64), (np.float32, torch.float64), (np.uint32, torch.int64), (np.bool_, torch.int64)):
        t = engine._device_windows((w % 200).astype(dt) if dt in (np.uint8, np.bool_) else w.astype(dt), cpu)
        assert t.dtype == want and tuple(t.shape) == (2, 11, 11) and t.is_contiguous(), (dt, t.dtype)
    if hasattr(torch, "uint16"):
        t = engine._device_windows(w.astype(np.uint16), cpu)
        assert t.dtype == torch.uint16 and np.array_equal(t.view(torch.int16).numpy().view(np.uint16), w.astype(np.uint16))
    assert tuple(engine._device_windows(w[0].astype(np.float64), cpu).shape) == (1, 11, 11)
    t = engine._device_windows(torch.from_numpy(w.astype(np.float32)), cpu)
    assert t.dtype == torch.float64
    try:
        engine._device_windows(np.zeros((3, 5, 7)), cpu)
    except ValueError:
        pass
    else:
        raise AssertionError("non-square windows must raise")
