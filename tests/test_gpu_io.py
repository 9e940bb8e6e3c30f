"""GPU: the file-level callers of the hot path (pflib.image_batch / parallel_image_batch,
pflib.py:883-1111) -- same result files as calling find_peptides image by image."""
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_frames(tmp_path, n=5):
    from PIL import Image
    from fluorosequencingimageanalysis_b200 import synth
    paths, imgs = [], []
    for i in range(n):
        img = synth.synth_frame(70 + i, H=96, W=128, n_spots=12 + 3 * i)
        p = str(tmp_path / ("field_%02d.png" % i))
        Image.fromarray(img).save(p)
        paths.append(p)
        imgs.append(img)
    odd = synth.synth_frame(99, H=64, W=80, n_spots=6)           # a second shape in the same call
    p = str(tmp_path / "odd.tif")
    Image.fromarray(odd).save(p)
    return paths + [p], imgs + [odd]


def _same_psfs(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k][:7] == b[k][:7] and a[k][9:] == b[k][9:]
        assert np.array_equal(a[k][7], b[k][7]) and np.array_equal(a[k][8], b[k][8])


@pytest.mark.parametrize("solver", ["minpack", "fast"])
def test_image_batch_writes_what_find_peptides_returns(tmp_path, solver):
    from fluorosequencingimageanalysis_b200 import pflib
    paths, imgs = _write_frames(tmp_path)
    old = pflib.SOLVER
    pflib.SOLVER = solver
    try:
        out = pflib.image_batch(paths + paths[:2], timestamp_epoch=1461000000)        # duplicates are ignored
        assert sorted(out.keys()) == sorted(os.path.abspath(p) for p in paths)
        for p, img in zip(paths, imgs):
            conv, pkl, csvp, png = out[os.path.abspath(p)]
            assert conv == (p if p.endswith('.png') else p + '.png')
            assert pkl == pflib._psfs_filename(conv, 1461000000, '.pkl') and os.path.exists(csvp) and os.path.exists(png)
            _same_psfs(pickle.load(open(pkl, 'rb')), pflib.find_peptides(img))
            n_rows = sum(1 for _ in open(csvp)) - 1
            assert n_rows == len(pflib.find_peptides(img))
        # non-default parameters are forwarded (pflib.py:284-287)
        out2 = pflib.image_batch(paths[:1], find_peptides_parameters={'c_std': 3, 'r_2_threshold': 0.8},
                                 timestamp_epoch=1461000001)
        _same_psfs(pickle.load(open(out2[os.path.abspath(paths[0])][1], 'rb')),
                   pflib.find_peptides(imgs[0], c_std=3, r_2_threshold=0.8))
        # an unreadable file is logged and skipped, the others are processed (pflib.py:960-964)
        bad = str(tmp_path / "broken.png")
        open(bad, 'wb').write(b'nope')
        out3 = pflib.image_batch([bad, paths[1]], timestamp_epoch=1461000002)
        assert list(out3.keys()) == [os.path.abspath(paths[1])]
    finally:
        pflib.SOLVER = old


def test_parallel_image_batch_equals_image_batch(tmp_path):
    """Partitions balanced by candidate count like pflib.py:1056-1069, one partition per device
    (all on device 0 here if the box has one GPU); same files as the serial call."""
    from fluorosequencingimageanalysis_b200 import pflib
    paths, imgs = _write_frames(tmp_path)
    ser = pflib.image_batch(paths, timestamp_epoch=1461000000)
    # host threads (the default) and worker processes (the reference's multiprocessing.Pool): same files either way
    for epoch, workers in ((1461000010, "process"), (1461000030, "thread")):
        par = pflib.parallel_image_batch(paths, timestamp_epoch=epoch, num_processes=3, workers=workers)
        assert sorted(par.keys()) == sorted(ser.keys()), workers
        for k in ser:
            _same_psfs(pickle.load(open(ser[k][1], 'rb')), pickle.load(open(par[k][1], 'rb')))
    with pytest.raises(ValueError):
        pflib.parallel_image_batch(paths, num_processes=2, workers="fibre")
    with pytest.raises(ValueError):
        pflib.parallel_image_batch(paths, num_processes=2.5)
    assert pflib.parallel_image_batch(paths[:1], timestamp_epoch=1461000020, num_processes=4).keys() == \
        pflib.image_batch(paths[:1], timestamp_epoch=1461000020).keys()


def test_image_batch_goes_through_bounded_chunks_and_loses_only_the_failing_image(tmp_path, monkeypatch):
    """image_batch reads lazily and sends bounded same-shape chunks to the device (psfio.BATCH_MAX_FRAMES); when a chunk
    fails as a whole every image of it is retried on its own, so only the failing image is skipped -- the reference's
    granularity (pflib.py:957-996)."""
    from fluorosequencingimageanalysis_b200 import pflib, psfio, engine
    paths, imgs = _write_frames(tmp_path, n=5)
    monkeypatch.setattr(psfio, "BATCH_MAX_FRAMES", 2)
    calls = []
    real = engine.find_peptides_batch

    def spy(frames, *a, **kw):
        f = np.asarray(frames)
        calls.append(f.shape[0] if f.ndim == 3 else 1)
        return real(frames, *a, **kw)
    monkeypatch.setattr(engine, "find_peptides_batch", spy)
    out = pflib.image_batch(paths, timestamp_epoch=1461000010)
    assert sorted(out.keys()) == sorted(os.path.abspath(p) for p in paths)
    assert max(calls) <= 2 and sum(calls) == len(paths)              # 5 same-shape frames -> 2 + 2 + 1, the odd one alone
    for p, img in zip(paths, imgs):
        _same_psfs(pickle.load(open(out[os.path.abspath(p)][1], 'rb')), pflib.find_peptides(img))

    # a batch that fails as a whole: the images are retried one by one and only the bad one is lost
    bad_marker = imgs[1]

    def flaky(frames, *a, **kw):
        f = np.asarray(frames)
        f3 = f if f.ndim == 3 else f[None]
        if any(np.array_equal(x, bad_marker) for x in f3):
            raise RuntimeError("simulated device failure")
        return real(frames, *a, **kw)
    monkeypatch.setattr(engine, "find_peptides_batch", flaky)
    out2 = pflib.image_batch(paths, timestamp_epoch=1461000011)
    assert sorted(out2.keys()) == sorted(os.path.abspath(p) for i, p in enumerate(paths) if i != 1)
