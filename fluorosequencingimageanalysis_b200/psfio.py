"""
File-level callers of the hot path and the PSF result files: the
interchange between the spot finder and ``flexlibrary`` / the ``basic_*_script`` drivers.

  _epoch_to_hash / _hash_to_epoch / _psfs_filename   pflib.py:523-591
  save_psfs_pkl / save_psfs_csv                      pflib.py:594-711
  read_image                                         pflib.py:714-746
  image_batch / parallel_image_batch                 pflib.py:883-1111

Same names, arguments, return layouts and error behaviour as the reference.  What is different
underneath: ``image_batch`` does not loop ``find_peptides`` image by image -- it stacks all
images of one shape and sends them through detection + fitting as ONE device batch; and
``parallel_image_batch`` fans the partitions out over GPUs (one host thread and CUDA stream set
per device) instead of over ``multiprocessing`` workers, with the reference's own
candidate-count balancing (``sharding.balance_by_count``).  Images are decoded with PIL (the
reference shells out to ImageMagick ``convert`` and ``scipy.misc.imread``, neither of which
touches the arithmetic of the path).
"""
import csv
import logging
import os
import pickle
import threading
import time

import numpy as np

_HASHCHARS = '0123456789abcdefghijklmnopqrstuvwxyz'

CSV_HEADER = ['Absolute image path', 'PSF center (h) coordinate', 'PSF center (w) coordinate',
              'PSF base (H)eight', 'PSF (A)mplitude', 'PSF width (sigma_h)', 'PSF width (sigma_w)',
              'PSF (theta)', 'PSF (rmse)', 'PSF (r_2)', 'PSF (s_n)']                      # pflib.py:688-698


def _py2_round(x):
    """Python-2 ``round``: half away from zero, returns a float."""
    x = float(x)
    return float(np.floor(x + 0.5)) if x >= 0 else -float(np.floor(-x + 0.5))


def _epoch_to_hash(epoch):
    """Base-36 text of a Unix epoch rounded to the second (pflib.py:523-543)."""
    if epoch <= 0:
        raise ValueError("epoch must be positive.")
    n = int(_py2_round(epoch))
    digits = []
    while n > 0:
        n, d = divmod(n, len(_HASHCHARS))
        digits.append(_HASHCHARS[d])
    return ''.join(reversed(digits))


def _hash_to_epoch(epoch_hash):
    """Inverse of _epoch_to_hash (pflib.py:546-566)."""
    epoch = 0
    for c in epoch_hash:
        k = _HASHCHARS.find(c)
        if k < 0:
            raise ValueError("epoch_hash contains unrecognized character(s).")
        epoch = epoch * len(_HASHCHARS) + k
    return epoch


def _psfs_filename(image_path, timestamp_epoch, format_suffix):
    """abspath(image_path) + '_psfs_' + hash + suffix (pflib.py:569-591)."""
    if timestamp_epoch is None:
        timestamp_epoch = _py2_round(time.time())
    return os.path.abspath(image_path) + '_psfs_' + _epoch_to_hash(timestamp_epoch) + format_suffix


def _resolve_output(image_path, timestamp_epoch, output_path, suffix):
    if image_path is None and output_path is None:
        raise ValueError("Either image_path or output_path must be provided.")
    if image_path is not None:
        image_path = os.path.abspath(image_path)
    if output_path is None:
        if timestamp_epoch is None:
            timestamp_epoch = _py2_round(time.time())
        output_path = _psfs_filename(image_path, timestamp_epoch, suffix)
    return image_path, output_path


def save_psfs_pkl(psfs, image_path=None, timestamp_epoch=None, output_path=None):
    """Pickle of the find_peptides dictionary (pflib.py:594-636); protocol 2 so that a Python-2
    flexlibrary can still unpickle the container structure."""
    _, output_path = _resolve_output(image_path, timestamp_epoch, output_path, '.pkl')
    with open(output_path, 'wb') as fh:
        pickle.dump(psfs, fh, protocol=2)
    return output_path


def save_psfs_csv(psfs, image_path=None, timestamp_epoch=None, output_path=None):
    """Tab-delimited table, one row per PSF in dictionary order (pflib.py:639-711)."""
    image_path, output_path = _resolve_output(image_path, timestamp_epoch, output_path, '.csv')
    with open(output_path, 'w', newline='') as fh:
        w = csv.writer(fh, dialect='excel-tab')
        w.writerow(CSV_HEADER)
        for (h, wpix), v in psfs.items():
            h_0, w_0, H, A, sigma_h, sigma_w, theta, sub_img, fit_img, rmse, r_2, s_n = v
            w.writerow([image_path] + [str(x) for x in (h_0, w_0, H, A, sigma_h, sigma_w, theta, rmse, r_2, s_n)])
    return output_path


def _stretch_limits(a):
    """1st and 99.9th percentile of the frame for the 8-bit contrast stretch: order statistics from a histogram for
    integer camera data (0.7 ms on a 512 x 512 frame where two numpy.percentile calls need 18 ms), numpy.percentile
    otherwise."""
    if a.dtype.kind in 'iu' and a.size and int(a.min()) >= 0 and int(a.max()) < (1 << 20):
        cs = np.cumsum(np.bincount(a.ravel()))
        n = int(cs[-1])
        return [float(np.searchsorted(cs, q / 100.0 * (n - 1), side='right')) for q in (1.0, 99.9)]
    lo, hi = np.percentile(np.asarray(a, dtype=np.float64), [1.0, 99.9])
    return [float(lo), float(hi)]


def save_psfs_png(psfs, image_path=None, timestamp_epoch=None, output_path=None, image=None, square_size=9):
    """Sanity-check picture: the frame, contrast-stretched to 8 bits, with a square around every
    PSF (the role of pflib.py:749-880; rendering details are not part of any numeric contract).
    Written as a palette image (254 greys + the squares' red) at PNG compression level 1: the writer was 85 % of
    image_batch's host time as a level-6 RGB file (86 ms -> 10 ms per 512 x 512 frame)."""
    from PIL import Image, ImageDraw
    image_path, output_path = _resolve_output(image_path, timestamp_epoch, output_path, '.png')
    if image is None:
        image = np.asarray(Image.open(image_path))
    image = np.asarray(image)
    lo, hi = _stretch_limits(image)
    g = np.clip((image.astype(np.float64) - lo) / max(hi - lo, 1e-12), 0.0, 1.0)
    im = Image.fromarray((g * 253.0).astype(np.uint8), mode='P')          # grey levels 0 .. 253; 255 = the squares
    pal = [v for k in range(254) for v in (k * 255 // 253,) * 3] + [0, 0, 0] + [255, 64, 64]
    im.putpalette(pal)
    dr = ImageDraw.Draw(im)
    r = square_size // 2
    for (h, w) in psfs:
        dr.rectangle([w - r, h - r, w + r, h + r], outline=255)
    im.save(output_path, compress_level=1)
    return output_path


def convert_image(image_path):
    """Write a PNG copy next to a non-PNG image and return its path (the role of the reference's
    ImageMagick call; 16-bit depth is kept)."""
    from PIL import Image
    out = image_path + '.png'
    a = np.asarray(Image.open(image_path))
    if a.dtype.kind in 'iu' and a.dtype.itemsize > 1:
        Image.fromarray(a.astype(np.uint16)).save(out)
    else:
        Image.fromarray(a).save(out)
    return out


def read_image(image_path):
    """-> (converted_path, image array); a non-PNG is converted once, `path + '.png'` is reused
    when it exists (pflib.py:714-746)."""
    from PIL import Image
    logger = logging.getLogger()
    converted_path = image_path = os.path.abspath(image_path)
    if image_path[-4:] != '.png':
        if os.path.exists(image_path + '.png'):
            converted_path += '.png'
        else:
            try:
                converted_path = convert_image(image_path)
            except Exception as e:
                logger.exception(e, exc_info=True)
                raise
    image = np.asarray(Image.open(converted_path))
    if image.ndim == 3:                                  # grey image stored with colour channels
        image = image[:, :, 0]
    return converted_path, image


def _unique_abs(paths):
    seen, out = set(), []
    for p in paths:
        p = os.path.abspath(p)
        if p not in seen:
            seen.add(p)
            out.append(p)
    return out


#: one device batch of image_batch holds at most this many frames / this many pixel bytes: detection scratch costs
#: ~10 B per pixel per frame and want_fit_img 200 B per candidate, so a directory of hundreds of 2048x2048 frames
#: must not become one launch (and a failure must not take more than one chunk's images with it)
BATCH_MAX_FRAMES = 32
BATCH_MAX_BYTES = 256 << 20


def _find_peptides_many(images, params):
    """find_peptides for a list of same-shape images through ONE device batch.  -> list with, per image, either the
    PSF dictionary or the exception that image raised (the reference loses only the failing image, pflib.py:957-996)."""
    from . import pflib, engine
    kw = dict(params)
    kw.pop('candidate_pixels', None)                     # "Not yet implemented" in the reference (pflib.py:374)
    r2t = kw.pop('r_2_threshold', 0.7)
    rad = kw.pop('consolidation_radius', 4)
    fit_type = kw.pop('fit_type', 'gauss')
    kw.pop('N_iter', None)
    if rad < 2:
        raise ValueError("consolidation_radius must be at least 2")
    if fit_type != 'gauss':
        raise NotImplementedError("fit_type='monte_carlo' (pflib.py:117-177) is not part of the CUDA hot path")
    import torch
    stack = np.stack(images)
    F = len(images)
    res = engine.find_peptides_batch(stack, faithful=pflib.FAITHFUL, want_fit_img=True, solver=pflib.SOLVER, to_host=False, **kw)
    n = int(res.fit.shape[0])
    if n == 0:
        return [{} for _ in images]
    # R^2 gate / consolidation / re-key for the whole batch on the device, then only the FINAL PSFs -- in the
    # reference's dictionary order (fsq_pack_psfs) -- come to the host: ~0.5 k records per frame instead of ~5 k
    # candidates with their model images
    cons = engine.consolidate_batch(res.cand_hw, res.cand_frame, res.fit, n, F, r2t, rad)
    if int(cons.flags.item()) & 1:
        # a re-keyed PSF collides with an existing key somewhere in the batch: the reference raises for THAT image
        # (pflib.py:518), so the images are consolidated one by one
        offs = np.concatenate([[0], np.cumsum(res.n_cand.cpu().numpy()[:-1])]).astype(np.int64)
        cand_hw, fit, fit_img = res.cand_hw.cpu().numpy(), res.fit.cpu().numpy(), res.fit_img.cpu().numpy()
        out = []
        for f, image in enumerate(images):
            sl = slice(offs[f], offs[f + 1])
            try:
                out.append(pflib.psfs_from_packed(image, cand_hw[sl], fit[sl], fit_img[sl], r2t, rad))
            except Exception as e:
                out.append(e)
        return out
    packed = engine.pack_psfs_batch(cons, res.cand_frame, res.fit, n, F)
    base = packed.base.cpu().numpy()
    m = int(base[-1])
    ints = packed.ints[:m].cpu().numpy()                  # (frame, key_h, key_w, candidate index)
    pfit = packed.fit[:m].cpu().numpy()
    cidx = packed.ints[:m, 3].long()
    pimg = res.fit_img[cidx].cpu().numpy().reshape(m, 5, 5)
    phw = res.cand_hw[cidx].cpu().numpy()
    out = []
    for f, image in enumerate(images):
        d = {}
        for i in range(int(base[f]), int(base[f + 1])):
            h, w = int(phw[i, 0]), int(phw[i, 1])
            v = pfit[i]
            d[(int(ints[i, 1]), int(ints[i, 2]))] = (float(v[0]), float(v[1]), float(v[2]), float(v[3]), float(v[4]), float(v[5]), float(v[6]),
                                                     image[h - 2:h + 3, w - 2:w + 3].astype(np.int64), pimg[i].copy(),
                                                     float(v[7]), float(v[8]), float(v[9]))
        out.append(d)
    return out


def _process_chunk(chunk, params, timestamp_epoch, processed, logger):
    """One bounded device batch of same-shape images -> result files.  If the batch as a whole fails, every image of
    it is retried on its own, so that only the failing image is logged and skipped (the reference's granularity)."""
    try:
        all_psfs = _find_peptides_many([g[2] for g in chunk], params)
    except Exception as e:
        if len(chunk) == 1:
            logger.exception(e, exc_info=True)
            return
        for item in chunk:
            _process_chunk([item], params, timestamp_epoch, processed, logger)
        return
    for (orig, converted, image), psfs in zip(chunk, all_psfs):
        try:
            if isinstance(psfs, Exception):
                raise psfs
            pkl = save_psfs_pkl(psfs, image_path=converted, timestamp_epoch=timestamp_epoch)
            csvp = save_psfs_csv(psfs, image_path=converted, timestamp_epoch=timestamp_epoch)
            png = save_psfs_png(psfs, image_path=converted, timestamp_epoch=timestamp_epoch, image=image)
        except Exception as e:
            logger.exception(e, exc_info=True)
            continue
        processed.setdefault(orig, (converted, pkl, csvp, png))


def image_batch(image_paths, find_peptides_parameters=None, timestamp_epoch=None):
    """pflib.py:883-997 -> {original_path: (converted_path, pkl_path, csv_path, png_path)}.
    Per-image failures are logged and skipped, as in the reference.  Images are read lazily and go to the device in
    bounded chunks of one shape (BATCH_MAX_FRAMES / BATCH_MAX_BYTES), so host and device memory do not grow with the
    length of the list."""
    logger = logging.getLogger()
    if timestamp_epoch is None:
        timestamp_epoch = _py2_round(time.time())
    image_paths = _unique_abs(image_paths)
    params = dict(find_peptides_parameters or {})
    processed = {}
    pending = {}                                          # (shape, dtype) -> images read but not yet processed
    for p in image_paths:
        try:
            converted, image = read_image(p)
        except Exception as e:
            logger.exception(e, exc_info=True)
            continue
        key = (image.shape, image.dtype.str)
        chunk = pending.setdefault(key, [])
        chunk.append((p, converted, image))
        if len(chunk) >= BATCH_MAX_FRAMES or len(chunk) * image.nbytes >= BATCH_MAX_BYTES:
            _process_chunk(pending.pop(key), params, timestamp_epoch, processed, logger)
    for chunk in pending.values():
        _process_chunk(chunk, params, timestamp_epoch, processed, logger)
    return processed


#: how parallel_image_batch runs its partitions: "thread" = host threads of this process, one per partition, each bound
#: to its GPU; "process" = one worker process per partition (multiprocessing, spawn context), the reference's arrangement
#: (multiprocessing.Pool, pflib.py:1082).  Measured on a B200 box (16 host cores, one GPU, tools/gpu_io_bench.py, 384
#: frames of 512 x 512 with every result file written): image_batch 60 images/s, 8 threads 72 images/s, 8 processes
#: 8.8 images/s -- what is left per image is Python (CSV rows, dictionaries, pickle), so threads gain little over one,
#: while every worker process pays ~15 s of start-up (interpreter, torch import, CUDA context) and processes that share
#: a GPU time-slice it.  Processes are for one worker per GPU on lists long enough to amortise the start-up.
PARALLEL_WORKERS = "thread"
#: seconds after which a worker process that has not answered is given up (logged; its images are reported missing)
PARALLEL_TIMEOUT_S = 3600.0


def _partition_worker(job):
    """One partition of parallel_image_batch in a worker process: bind to the partition's GPU, carry the parent's
    solver switches over (module globals are not inherited by a spawned process) and run image_batch."""
    dev, paths, params, timestamp_epoch, solver, faithful = job
    import torch
    from . import pflib
    torch.cuda.set_device(dev)
    pflib.SOLVER, pflib.FAITHFUL = solver, faithful
    return image_batch(paths, find_peptides_parameters=params, timestamp_epoch=timestamp_epoch)


def parallel_image_batch(image_paths, find_peptides_parameters=None, timestamp_epoch=None, num_processes=None, workers=None):
    """pflib.py:1000-1111 with one GPU per worker: images are balanced over ``num_processes`` partitions by candidate
    count exactly like the reference (detection runs once on the first device for the count), each partition runs
    ``image_batch`` on its own GPU (partition k -> device k mod device_count) -- in its own host thread by default, or in
    its own worker process as the reference does (``workers="process"`` / psfio.PARALLEL_WORKERS).  No collective; the
    parent merges the workers' {path: result files} dictionaries (pflib.py:1085-1108)."""
    import torch
    from . import engine, sharding, pflib
    logger = logging.getLogger()
    if num_processes == 1 or len(image_paths) == 1:
        return image_batch(image_paths, find_peptides_parameters=find_peptides_parameters, timestamp_epoch=timestamp_epoch)
    if num_processes is None:
        num_processes = max(1, torch.cuda.device_count())
    if timestamp_epoch is None:
        timestamp_epoch = _py2_round(time.time())
    if num_processes < 1 or round(num_processes) != num_processes:
        raise ValueError("Number of processes must be an integer >= 1")                  # pflib.py:1059-1060
    workers = workers or PARALLEL_WORKERS
    if workers not in ("process", "thread"):
        raise ValueError("workers must be 'process' or 'thread'")
    image_paths = _unique_abs(image_paths)
    params = dict(find_peptides_parameters or {})
    det_kw = {k: params[k] for k in ('median_filter_size', 'correlation_matrix', 'c_std') if k in params}
    # candidate counts for the balance (pflib.py:1056-1069): detection only, same-shape images in bounded device batches
    paths, counts, pending = [], [], {}

    def count_group(group):
        try:
            nc = engine.detect_batch(np.stack([g[1] for g in group]), **det_kw).n_cand.cpu().numpy()
            got = [(g[0], int(nc[i])) for i, g in enumerate(group)]
        except Exception:
            got = []
            for g in group:                                  # one by one: only the failing image is lost
                try:
                    got.append((g[0], int(engine.detect_batch(g[1], **det_kw).total)))
                except Exception as e:
                    logger.exception(e, exc_info=True)
        for p, c in got:
            paths.append(p)
            counts.append(c)

    for p in image_paths:
        try:
            _, image = read_image(p)
        except Exception as e:
            logger.exception(e, exc_info=True)
            continue
        key = (image.shape, image.dtype.str)
        group = pending.setdefault(key, [])
        group.append((p, image))
        if len(group) >= BATCH_MAX_FRAMES or len(group) * image.nbytes >= BATCH_MAX_BYTES:
            count_group(pending.pop(key))
    for group in pending.values():
        count_group(group)
    order = {p: i for i, p in enumerate(image_paths)}        # keep the caller's order (the balance depends on it)
    paths, counts = (list(t) for t in zip(*sorted(zip(paths, counts), key=lambda pc: order[pc[0]]))) if paths else ([], [])
    parts = sharding.balance_by_count(counts, int(num_processes))
    ndev = max(1, torch.cuda.device_count())
    live = [k for k in range(len(parts)) if parts[k]]
    results = {}

    if workers == "process" and len(live) > 1:
        import multiprocessing
        ctx = multiprocessing.get_context("spawn")          # the parent holds a CUDA context: fork is not an option
        jobs = {k: (k % ndev, [paths[i] for i in parts[k]], params, timestamp_epoch, pflib.SOLVER, pflib.FAITHFUL) for k in live}
        pool = ctx.Pool(processes=len(live))
        try:
            pending = {k: pool.apply_async(_partition_worker, (jobs[k],)) for k in live}
            for k, fut in pending.items():
                try:
                    results[k] = fut.get(timeout=PARALLEL_TIMEOUT_S)
                except Exception as e:                       # a failed partition is logged, the others survive
                    logger.exception(e, exc_info=True)
        finally:
            pool.terminate()
            pool.join()
    else:
        def work(k):
            try:
                with torch.cuda.device(k % ndev):
                    results[k] = image_batch([paths[i] for i in parts[k]], find_peptides_parameters=params,
                                             timestamp_epoch=timestamp_epoch)
            except Exception as e:
                logger.exception(e, exc_info=True)

        threads = [threading.Thread(target=work, args=(k,)) for k in live]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    processed = {}
    for k in live:
        for key, v in (results.get(k) or {}).items():
            processed.setdefault(key, v)
    return processed
