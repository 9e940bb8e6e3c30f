"""Developer diagnostic (GPU): the file-level callers on a directory of synthetic frames -- image_batch against
parallel_image_batch with worker processes and with host threads (pflib.py:883-1111).  Everything a caller of the
reference gets is inside the timings: image decoding, detection + fits, the PSF dictionaries, pickle / CSV / PNG files.
    python tools/gpu_io_bench.py [n_images] [num_processes]"""
import os, sys, tempfile, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np, torch
    from PIL import Image
    from fluorosequencingimageanalysis_b200 import pflib, synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    nproc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    pflib.SOLVER, pflib.FAITHFUL = "fast", False
    tmp = tempfile.mkdtemp(prefix="fsq_io_")
    try:
        paths = []
        for i in range(n):
            p = os.path.join(tmp, "field_%03d.png" % i)
            Image.fromarray(synth.synth_frame(200 + i)).save(p)           # 512 x 512 uint16, 500 spots
            paths.append(p)
        pflib.image_batch(paths[:2], timestamp_epoch=1461000000)           # warm-up: CUDA context, library load
        torch.cuda.synchronize()
        rows = []
        t0 = time.time(); out = pflib.image_batch(paths, timestamp_epoch=1461000100); rows.append(("image_batch (one process)", time.time() - t0, len(out)))
        for k, workers in enumerate(("thread", "process")):
            t0 = time.time()
            out = pflib.parallel_image_batch(paths, timestamp_epoch=1461000200 + 100 * k, num_processes=nproc, workers=workers)
            rows.append(("parallel_image_batch, %d %s workers" % (nproc, workers), time.time() - t0, len(out)))
        if os.environ.get("FSQ_IO_PROFILE"):
            import cProfile, pstats
            pr = cProfile.Profile(); pr.enable()
            pflib.image_batch(paths[:32], timestamp_epoch=1461000900)
            pr.disable()
            pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
        for name, dt, m in rows:
            print("%-46s %3d images %7.2f s  %6.2f images/s" % (name, m, dt, m / dt), flush=True)
        print("(%d GPU(s) visible; %d host cores; worker processes pay their start-up -- interpreter, torch import, CUDA context -- inside the timing)"
              % (torch.cuda.device_count(), os.cpu_count()))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
