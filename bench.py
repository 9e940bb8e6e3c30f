"""
bench.py -- headline benchmark of the spot hot path (BASELINE.json metric: 2-D Gaussian PSF
candidate fits/s and frames/s) on N B200s of one node, with the reference CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[4] -- the scaling job of 8000 fields x 10 Edman cycles of
512x512 frames at ~1000 spots/field (the configs[2] experiment recipe), end-to-end detection + per-candidate
5x5 LM fit + metrics + R^2 gate / consolidation / re-key.  One "step" = one tenth of that job: a block of
800 fields x 10 cycles = 8000 frames (~4.6e7 candidate fits), so that the driver's --steps 20 is a timed region of
seconds, not of one pipeline wave.  N > 1 is STRONG scaling: the step's 800 fields are split into contiguous blocks
with near-equal candidate totals by sharding.balanced_field_blocks (all cycles of a field on one GPU,
pflib.py:1056-1069 is the reference's balancing), every rank runs its block, no data-path collective; in the e2e
region every rank's final PSF records land in its own pinned host memory and the per-frame result table is gathered
with sharding.gather_packed at the end (what pflib.parallel_image_batch's parent collects, pflib.py:1085-1108).
Frames come from a pool of 80 different fields (800 frames, 420 MB > 126 MB L2) cycled through the job.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every key.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# 32 hardware work queues instead of 8, before the CUDA context exists (see fluorosequencingimageanalysis_b200/__init__.py)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

H, W, N_SPOTS, CYCLES = 512, 512, 1000, 10
JOB_FIELDS = 8000                 # configs[4]
FIELDS_PER_STEP = 800             # one step = 1/10 of the job
POOL_FIELDS = 80                  # different fields cycled through the job (field i of a step -> pool field i % 80)
LAUNCH_FIELDS = 20                # fields (x 10 cycles = 200 frames) one pass of the kernels processes
POOL_SEED = 5
# SURVEY.md 8(d) FLOP convention, `minpack` mode, per executed LM iteration, P = 25, n = 7:
# 8 evaluations (176 P) + Householder QR (2 P n^2 - 2/3 n^3 = 2221) + Q^T f (4 P n = 700) + lmpar (~10 x 350)
FLOP_PER_LM_ITER_5x5 = 176 * 25 + 2221 + 700 + 3500
METRIC = "2-D Gaussian PSF candidate fits/s (detection + 5x5 LM fit + metrics), frames/s alongside"
WORKLOAD = ("configs[4]: 8000 fields x 10 cycles of 512x512 u16 frames, ~1000 spots/field (configs[2] recipe), field-sharded; "
            "one step = 800 fields x 10 cycles = 8000 frames (1/10 of the job); every frame: detection + 5x5 LM fit of "
            "every candidate + metrics + R^2 gate / consolidation / re-key")


def static_config(world):
    """The workload description both arms print (identical dicts: the driver compares them)."""
    return {"workload": WORKLOAD, "fields_per_step": FIELDS_PER_STEP, "cycles": CYCLES, "frames_per_step": FIELDS_PER_STEP * CYCLES,
            "frame": "%dx%d uint16" % (H, W), "spots_per_field": N_SPOTS,
            "l2": "pool of %d different fields (%d frames, %d MB > 126 MB L2) cycled: inputs larger than L2"
                  % (POOL_FIELDS, POOL_FIELDS * CYCLES, POOL_FIELDS * CYCLES * H * W * 2 >> 20),
            "parallelism": "field-sharded x%d (contiguous field blocks balanced by candidate count), no collective" % world}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_pool(n_fields=POOL_FIELDS):
    from fluorosequencingimageanalysis_b200 import synth
    return synth.experiment_field_pool(POOL_SEED, n_fields, n_cycles=CYCLES, H=H, W=W, n_spots=N_SPOTS)


# --------------------------------------------------------------------------- CPU reference arm
def _ref_worker_init():
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    global _REF_FIT
    _REF_FIT = None


def _ref_fit_one(sub):
    """One candidate through the reference's pflib._fit_2d_gaussian + metrics (oracle/_ref when
    built, else the restated oracle port)."""
    global _REF_FIT
    if _REF_FIT is None:
        from oracle import build_ref, pflib_oracle as po
        mods = build_ref.load()
        if mods is not None:
            pf = mods[0]

            def fit(s):
                out = pf._fit_2d_gaussian(s)
                fit_img = out[7]
                r_2 = 1.0 - sum(np.reshape((s - fit_img) ** 2, -1)) / sum((np.reshape(s, -1) - np.mean(s)) ** 2)
                return r_2, pf.illumina_s_n(s)
            _REF_FIT = ("reference", fit)
        else:
            def fit(s):
                out = po.fit_2d_gaussian(s, faithful=True)
                return po.fit_metrics(s, out[7])[0], po.illumina_s_n(s)
            _REF_FIT = ("port", fit)
    return _REF_FIT[1](sub)


def _ref_kind():
    from oracle import build_ref
    return "reference" if build_ref.load() is not None else "port"


def _ref_detect(img):
    from oracle import build_ref, pflib_oracle as po
    mods = build_ref.load()
    if mods is not None:
        return mods[0]._psf_candidates(img)
    return po.psf_candidates(img)


def cpu_reference_sample(frames, n_fits, cores, pool=None):
    """Times the reference CPU path on a bounded sample of the workload: detection on full frames
    (2 frames) and `n_fits` candidate fits fanned out over `cores` processes exactly like
    pflib.parallel_image_batch's Pool (pflib.py:1082).  Returns dict(fits_per_s, frames_per_s, ...)."""
    import multiprocessing
    t0 = time.perf_counter()
    cands = _ref_detect(frames[0])
    t_det = time.perf_counter() - t0
    rng = np.random.default_rng(12345)
    pick = rng.choice(len(cands), size=min(n_fits, len(cands)), replace=False)
    subs = [frames[0][cands[i][0] - 2:cands[i][0] + 3, cands[i][1] - 2:cands[i][1] + 3].astype(np.int64) for i in pick]
    own = pool is None
    if own:
        pool = multiprocessing.Pool(cores, initializer=_ref_worker_init)
        pool.map(_ref_fit_one, subs[:cores])           # import / warm the workers
    t0 = time.perf_counter()
    pool.map(_ref_fit_one, subs, chunksize=max(1, len(subs) // (cores * 4)))
    t_fit = time.perf_counter() - t0
    if own:
        pool.close()
        pool.join()
    fits_per_s = len(subs) / t_fit
    frames_per_s = 1.0 / (t_det / 1.0 / cores + len(cands) / fits_per_s)   # detection also fans out over files
    return dict(fits_per_s=fits_per_s, frames_per_s=frames_per_s, n_fits=len(subs), t_fit=t_fit,
                t_detect_per_frame=t_det, cands_per_frame=len(cands))


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on this box's host cores."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    import multiprocessing
    cores = os.cpu_count() or 1
    frames = make_pool(1)[0, :1]                      # field 0, cycle 0 of the workload's pool
    kind = _ref_kind()
    per_step = max(cores * 24, 64)                    # ~2.5 s of CPU work per step at ~0.1 s/fit/core
    pool = multiprocessing.Pool(cores, initializer=_ref_worker_init)
    cands = _ref_detect(frames[0])
    rng = np.random.default_rng(999)
    steps = args.warmup + args.steps
    times, nf = [], []
    for s in range(steps):
        pick = rng.choice(len(cands), size=min(per_step, len(cands)), replace=False)
        subs = [frames[0][cands[i][0] - 2:cands[i][0] + 3, cands[i][1] - 2:cands[i][1] + 3].astype(np.int64) for i in pick]
        t0 = time.perf_counter()
        pool.map(_ref_fit_one, subs, chunksize=max(1, len(subs) // (cores * 4)))
        times.append(time.perf_counter() - t0)
        nf.append(len(subs))
    pool.close()
    pool.join()
    t = sum(times[args.warmup:])
    n = sum(nf[args.warmup:])
    value = n / t
    t0 = time.perf_counter()
    _ref_detect(frames[0])
    t_det = time.perf_counter() - t0
    sample = "%d candidate fits per step (random candidates of one 512x512 frame of the workload) on %d processes; detection timed on 1 full frame" % (per_step, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "fits/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": static_config(max(args.gpus, 1)),
        "frames_per_s": 1.0 / (t_det / cores + len(cands) / value),
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler(object):
    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, windows=()):
        """windows: (t0, t1) perf_counter intervals of the timed regions; samples inside them are
        preferred, all samples of the run are the fall-back when the regions were shorter than the
        sampling period."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (t, ln) in self.lines if any(a <= t <= b + 0.15 for a, b in windows)]
        scope = "timed regions"
        if len(inside) < 3:
            inside = [ln for (_, ln) in self.lines]
            scope = "whole GPU phase of the run (timed regions shorter than the sampling period)"
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


# --------------------------------------------------------------------------- our arm
SOLVERS = {"fast": ("fast", False), "fast64": ("fast64", False),
           "minpack-faithful": ("minpack", True), "minpack-clean": ("minpack", False)}
# SURVEY.md 8(d) FLOP convention per executed LM iteration, P = 25 pixels, n = 7 parameters
#   fast   : 1 evaluation (22 P) + analytic Jacobian (40 P) + J^T J, J^T r (70 P) + 7x7 damped solve (~300)
#   minpack: 8 evaluations (176 P) + Householder QR (2 P n^2 - 2/3 n^3 = 2221) + Q^T f (4 P n) + lmpar (~10 x 350)
FLOP_PER_LM_ITER = {"fast": 132 * 25 + 300, "fast64": 132 * 25 + 300, "minpack": FLOP_PER_LM_ITER_5x5}


def plan_chunks(lo, hi, launch_fields=LAUNCH_FIELDS):
    """A rank's field block [lo, hi) of one step -> [(pool_field_start, n_fields)] launches of <= launch_fields
    consecutive fields; field i of the step is pool field i % POOL_FIELDS and the pool holds launch_fields extra wrapped
    fields, so every launch is one contiguous slice of it."""
    out = []
    f = lo
    while f < hi:
        n = min(launch_fields, hi - f)
        out.append((f % POOL_FIELDS, n))
        f += n
    return out


def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from fluorosequencingimageanalysis_b200 import engine, _lib, sharding

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L = _lib.load()
    solver, faithful = SOLVERS[args.solver]
    warmup = max(args.warmup, 3)
    LF = args.launch_fields
    FL = LF * CYCLES                              # frames per full launch
    # ---- synthetic inputs: the pool of different fields (+ LF wrapped ones), pinned host copy and device copy
    pool = make_pool()                                                        # [POOL_FIELDS, CYCLES, H, W] u16
    pool = np.concatenate([pool, pool[:LF]], axis=0).reshape(-1, H, W)
    pool_host = torch.from_numpy(pool.view(np.int16)).view(torch.uint16).pin_memory()
    pool_dev = pool_host.to(dev)
    kw = dict(dtype=torch.uint16, faithful=faithful, solver=solver)
    if args.park is not None:
        kw["park_after"] = args.park
    kw["warps_per_sm"] = args.warps_per_sm
    if args.fetch == "psfs":
        kw["consolidate"] = True          # the tail of find_peptides runs on the device in both timed regions
    cur = torch.cuda.current_stream()

    # ---- candidate count per pool field (one detection pass over the pool), then the partition of a step's fields:
    #      contiguous blocks with near-equal candidate totals (sharding.balanced_field_blocks), identical on all ranks
    fs = engine.FieldStream(FL, H, W, depth=args.depth, host_io=False, **kw)
    pipe = fs.slots[0]["pipe"]
    pool_counts = np.zeros(POOL_FIELDS, dtype=np.int64)
    for p0 in range(0, POOL_FIELDS, LF):
        nf = min(LF, POOL_FIELDS - p0)
        pipe.run(pool_dev[p0 * CYCLES:(p0 + nf) * CYCLES], fit=False, n_frames=nf * CYCLES)
        torch.cuda.synchronize()
        pool_counts[p0:p0 + nf] = pipe.n_cand[:nf * CYCLES].cpu().numpy().reshape(nf, CYCLES).sum(axis=1)
    step_counts = pool_counts[np.arange(FIELDS_PER_STEP) % POOL_FIELDS]
    if args.scaling == "strong":
        blocks = sharding.balanced_field_blocks(step_counts, world)
    else:
        blocks = [(0, FIELDS_PER_STEP)] * world                  # weak: every rank runs a whole step
    lo, hi = blocks[rank]
    chunks = plan_chunks(lo, hi, LF)
    my_fits_per_step = int(step_counts[lo:hi].sum())
    n_launch = args.steps * len(chunks)

    # ---- timed region A: inputs resident in HBM; K steps software-pipelined over `depth` streams
    def resident_steps(n_steps):
        for _ in range(n_steps):
            for (p0, nf) in chunks:
                fs.submit(pool_dev[p0 * CYCLES:(p0 + nf) * CYCLES])
        for sl in fs.slots:
            cur.wait_stream(sl["stream"])

    sampler = ClockSampler(local_rank)
    sampler.start()
    resident_steps(warmup)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    win_a0 = time.perf_counter()
    ev[0].record()
    resident_steps(args.steps)
    ev[1].record()
    host_enqueue_ms = (time.perf_counter() - win_a0) * 1e3 / args.steps      # host time to queue one step
    barrier()
    win_a1 = time.perf_counter()
    ms_total = ev[0].elapsed_time(ev[1])
    fits_total = my_fits_per_step * args.steps
    for sl in fs.slots:                                                       # every launch stayed inside its buffers
        if int(sl["pipe"].n_cand[sl["pipe"].F_run].item()) > sl["pipe"].cap:
            raise RuntimeError("candidate capacity exceeded")
    gpu_launches = args.steps * sum(pipe.launches_per_run(nf * CYCLES) for (_, nf) in chunks)

    # ---- LM iterations per fit over the whole pool (the timed regions cycle through exactly these frames), and the
    #      fit / detection launches of one full 200-frame launch timed alone
    sum_niter = sum_nfev = n_pool = 0
    for p0 in range(0, POOL_FIELDS, LF):
        nf = min(LF, POOL_FIELDS - p0)
        pipe.run(pool_dev[p0 * CYCLES:(p0 + nf) * CYCLES], n_frames=nf * CYCLES)
        n_k = pipe.total()
        sum_niter += int(pipe.out_int[:n_k, engine.ICOL_NITER].sum().item())
        sum_nfev += int(pipe.out_int[:n_k, engine.ICOL_NFEV].sum().item())
        n_pool += n_k
    frames_k = pool_dev[:FL]
    pipe.run(frames_k)
    fit_ms = []
    for k in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.run_fit_only(frames_k)
        e1.record()
        e1.synchronize()
        fit_ms.append(e0.elapsed_time(e1))
    n_last = pipe.total()
    niter_last = int(pipe.out_int[:n_last, engine.ICOL_NITER].sum().item())
    fit_ms_avg = float(np.mean(fit_ms[1:]))
    det_ms = []
    for k in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.run_detect_only(pool_dev[(k % 3) * FL:(k % 3 + 1) * FL])
        e1.record()
        e1.synchronize()
        det_ms.append(e0.elapsed_time(e1))
    det_ms_avg = float(np.mean(det_ms[1:]))
    n_det = pipe.total()
    # un-pipelined launches (one stream, one batch at a time): what a single isolated call costs
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(3):
        pipe.run(pool_dev[k * FL:(k + 1) * FL])
    e1.record()
    e1.synchronize()
    serial_ms_per_launch = e0.elapsed_time(e1) / 3.0

    # ---- host <-> device link, measured (all ranks at once: the e2e region feeds N GPUs from one host's memory)
    pcie = {}
    buf_d = torch.empty(FL * H * W, dtype=torch.int16, device=dev)
    buf_h = torch.empty(FL * H * W, dtype=torch.int16).pin_memory()
    for name, dst, src in (("h2d_gbs", buf_d, buf_h), ("d2h_gbs", buf_h, buf_d)):
        dst.copy_(src, non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        e1.record()
        e1.synchronize()
        pcie[name] = 4 * buf_h.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del buf_d, buf_h

    # ---- timed region B (e2e): pinned host frames in, final PSF records back in pinned host memory, every launch;
    #      submit(k) / begin_fetch(k-lag) / end_fetch(k-lag-1) keeps H2D, kernels and D2H of neighbouring launches in
    #      flight; the region ends with the gather of the per-frame result table (sharding.gather_packed)
    #      (the host learns a launch's PSF count only when the launch has ended, so `e2e_depth` slots keep the same number
    #      of launches queued ahead of the GPU as region A's `depth` streams do)
    e2e_depth = args.e2e_depth if args.e2e_depth else args.depth + 2
    fe = engine.FieldStream(FL, H, W, depth=e2e_depth, host_io=True, fetch=args.fetch, **kw)
    epipe = fe.slots[0]["pipe"]

    def e2e_steps(n_steps, gather):
        fits = d2h = h2d = 0
        tick, meta = [], []
        table = {"field": [], "n_psf": []}
        lag_b, lag_e = max(e2e_depth - 2, 1), max(e2e_depth - 1, 2)

        def finish(k):
            nonlocal fits, d2h
            r = fe.end_fetch(tick[k])
            fits += r[0]
            if args.fetch == "psfs":
                d2h += epipe.d2h_bytes_psfs(r[1])
                f0, nf = meta[k]
                table["field"].append(np.arange(f0, f0 + nf, dtype=np.int32))
                table["n_psf"].append(np.diff(r[4].numpy()).reshape(nf, CYCLES).astype(np.int32))
            else:
                d2h += epipe.d2h_bytes(r[0])

        k = 0
        for s_ in range(n_steps):
            f = lo
            for (p0, nf) in chunks:
                src = (pool_dev if args.diag_no_h2d else pool_host)[p0 * CYCLES:(p0 + nf) * CYCLES]
                tick.append(fe.submit(src))
                meta.append((s_ * FIELDS_PER_STEP + f, nf))
                h2d += src.numel() * 2
                f += nf
                if k >= lag_b:
                    fe.begin_fetch(tick[k - lag_b])
                if k >= lag_e:
                    finish(k - lag_e)
                k += 1
        for j in range(max(0, k - lag_e), k):
            finish(j)
        gathered = None
        if gather and args.fetch == "psfs":
            local = {"field": np.concatenate(table["field"]), "n_psf": np.concatenate(table["n_psf"])}
            gathered = sharding.gather_packed(local)
        return fits, d2h, h2d, gathered

    e2e_steps(min(warmup, 3), False)
    barrier()
    t0 = time.perf_counter()
    e2e_fits, d2h, h2d, gathered = e2e_steps(args.steps, True)
    barrier()
    e2e_s = time.perf_counter() - t0
    win_b = (t0, t0 + e2e_s)
    if gathered is not None and args.scaling == "strong":
        want = FIELDS_PER_STEP * args.steps
        if gathered["field"].shape[0] != want or len(np.unique(gathered["field"])) != want:
            raise RuntimeError("gathered result table covers %d of %d (field, step) rows" % (len(np.unique(gathered["field"])), want))

    # ---- weak-scaling companion (N > 1 only): every rank runs whole 800-field steps, resident inputs
    weak = None
    if world > 1 and args.scaling == "strong":
        wchunks = plan_chunks(0, FIELDS_PER_STEP, LF)
        wsteps = max(2, args.steps // 5)

        def weak_steps(n):
            for _ in range(n):
                for (p0, nf) in wchunks:
                    fs.submit(pool_dev[p0 * CYCLES:(p0 + nf) * CYCLES])
            for sl in fs.slots:
                cur.wait_stream(sl["stream"])
        weak_steps(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        weak_steps(wsteps)
        e1.record()
        barrier()
        weak = [e0.elapsed_time(e1), float(step_counts.sum()) * wsteps, wsteps]
    clocks = sampler.stop([(win_a0, win_a1), win_b])

    # ---- the parity instrument (reference-faithful MINPACK solver) on 40 frames of the same pool
    parity = None
    if rank == 0 and solver != "minpack" and not args.no_parity_solver:
        pp = engine.FieldPipeline(40, H, W, dtype=torch.uint16, faithful=True, solver="minpack")
        pp.run(pool_dev[:40])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pp.run(pool_dev[:40])
        e1.record()
        e1.synchronize()
        parity = {"solver": "minpack-faithful (reference behaviour incl. qrsolv diagonal view, FD Jacobian, QR)",
                  "value": pp.total() / (e0.elapsed_time(e1) * 1e-3), "unit": "fits/s", "ms_per_40_frames": e0.elapsed_time(e1)}
        del pp

    # ---- BASELINE configs[3] beside the headline (N = 1 only): 2048x2048 frames at ~20 000 spots through the same
    #      pipeline, and the fitter-bound batch of 11x11 windows cut from such a frame (gaussfit default arguments)
    other = None
    if world == 1 and not args.no_other_configs:
        from fluorosequencingimageanalysis_b200 import synth
        big, cr, cc, _ = synth.synth_frame_with_truth(4, H=2048, W=2048, n_spots=20000)
        frames4 = np.stack(synth.dihedral_variants(big)[:4])
        f4 = torch.from_numpy(frames4.view(np.int16)).view(torch.uint16).to(dev)
        p4 = engine.FieldPipeline(4, 2048, 2048, dtype=torch.uint16, faithful=faithful, solver=solver, consolidate=True,
                                  warps_per_sm=0)
        p4.run(f4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            p4.run(f4)
        e1.record()
        e1.synchronize()
        ms4 = e0.elapsed_time(e1) / 5.0
        n4, m4 = p4.total(), p4.total_psfs()
        del p4
        win = synth.cut_windows(big, cr, cc, 11)
        wd = torch.from_numpy(np.concatenate([win] * (-(-200000 // len(win))))[:200000]).to(dev)
        res11 = {}
        for sv in ("fast", "minpack"):
            wsub = wd if sv == "fast" else wd[:20000]
            engine.gaussfit_default_batch(wsub, solver=sv, faithful=(sv == "minpack"))
            torch.cuda.synchronize()
            ms11 = []
            for _ in range(3 if sv == "fast" else 1):            # a single 200 000-window call is a 10 ms measurement: best of three
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r11, _ = engine.gaussfit_default_batch(wsub, solver=sv, faithful=(sv == "minpack"))
                e1.record()
                e1.synchronize()
                ms11.append(e0.elapsed_time(e1))
            res11[sv] = {"windows": int(wsub.shape[0]), "ms": min(ms11), "fits_per_s": wsub.shape[0] / (min(ms11) * 1e-3), "calls_timed": len(ms11),
                         "mean_niter": float(r11.niter.double().mean().item()), "status_gt0": float((r11.status > 0).double().mean().item())}
        other = {"configs[3] frames": {"what": "4 frames 2048x2048 x ~20 000 spots, detection + 5x5 fits + metrics + consolidation, one stream",
                                       "candidates": n4, "final_psfs": m4, "ms": ms4, "fits_per_s": n4 / (ms4 * 1e-3), "frames_per_s": 4 / (ms4 * 1e-3)},
                 "configs[3] 11x11 windows": {"what": "windows cut around the spots of that frame (neighbours inside most windows), gaussfit default "
                                                      "arguments, moments start values on the device (fsq_moments) inside the timed call",
                                              "fast": res11["fast"], "minpack_faithful": res11["minpack"]}}
        del wd, f4

    # ---- max over ranks, sum of work
    if world > 1:
        tt = torch.tensor([ms_total, e2e_s * 1e3, weak[0] if weak else 0.0, -pcie["h2d_gbs"], -pcie["d2h_gbs"]],
                          dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms = float(tt[0].item()), float(tt[1].item())
        if weak:
            weak[0] = float(tt[2].item())
        pcie_min = {"h2d_gbs": -float(tt[3].item()), "d2h_gbs": -float(tt[4].item())}
        cc = torch.tensor([fits_total, e2e_fits, h2d, d2h, gpu_launches], dtype=torch.int64, device=dev)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        fits_all, e2e_fits_all, h2d_all, d2h_all, launches_all = (int(v) for v in cc.tolist())
    else:
        e2e_ms = e2e_s * 1e3
        fits_all, e2e_fits_all, h2d_all, d2h_all, launches_all = fits_total, e2e_fits, h2d, d2h, gpu_launches
        pcie_min = dict(pcie)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    frames_per_step_all = FIELDS_PER_STEP * CYCLES * (world if args.scaling == "weak" else 1)
    value = fits_all / (ms_total * 1e-3)
    frames_per_s = args.steps * frames_per_step_all / (ms_total * 1e-3)
    # ---- roofline of the dominant kernel (the LM fitter; FP pipes, never tensor cores)
    peak = {}
    for nm, flag in (("fp64", 1), ("fp32", 0)):
        v = ctypes.c_double(0.0)
        _lib.check(L.fsq_fma_peak(flag, ctypes.byref(v), None))
        peak[nm] = v.value
    fl_iter = FLOP_PER_LM_ITER[solver]
    iters_per_fit = sum_niter / max(n_pool, 1)
    achieved = niter_last * fl_iter / (fit_ms_avg * 1e-3) / 1e12
    pk = "fp64" if solver in ("minpack", "fast64") else "fp32"
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_file):
        try:
            hbm_peak = float(json.load(open(peaks_file))["hbm_gbs"])
            hbm_src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    # the ncu captures are per 40-frame stack of the configs[1] workload: scaled by frames (detection) / fits (LM)
    lm_traffic = traffic.get("lmwarp_kernel", {}).get("bytes") if solver == "fast" else None
    lm_traffic = lm_traffic * n_last / traffic["lmwarp_kernel"].get("fits", 202564) if lm_traffic else None
    det_traffic = traffic.get("detect_cm_packed_kernel", {}).get("bytes")
    det_traffic = det_traffic * FL / 40 if det_traffic else None
    det_bytes = FL * H * W * 2 + 8 * n_det
    det_gbs = det_bytes / (det_ms_avg * 1e-3) / 1e9
    kname = {"fast": "lmwarp_kernel (+ fit_prep_kernel, fit_finish_kernel) behind fsq_fit_candidates",
             "fast64": "lmfast_kernel<double,true>", "minpack": "lmfit_kernel<8,true>"}[solver]
    pipe_tflops = iters_per_fit * fl_iter * (fits_all / world) / (ms_total * 1e-3) / 1e12     # per GPU
    roofline = {"bound": pk, "achieved": pipe_tflops, "peak": peak[pk] / 1e12, "unit": "TFLOP/s",
                "frac": pipe_tflops / (peak[pk] / 1e12), "traffic": lm_traffic,
                "achieved_lone_launch": achieved, "frac_lone_launch": achieved / (peak[pk] / 1e12),
                "achieved_in_pipeline": pipe_tflops, "frac_in_pipeline": pipe_tflops / (peak[pk] / 1e12),
                "kernel": kname, "ms_per_launch": fit_ms_avg,
                "fits_per_launch": n_last, "lm_iterations_per_launch": niter_last, "lm_iterations_per_fit": iters_per_fit,
                "passes_per_fit": sum_nfev / max(n_pool, 1), "flop_per_lm_iteration": fl_iter,
                "peak_source": "fsq_fma_peak %s FMA micro-benchmark, measured in this run by this repo's own kernel "
                               "(MEASURED_PEAKS.json holds no FP32 vector peak; nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s)" % pk.upper(),
                "fp32_fma_peak_tflops": peak["fp32"] / 1e12, "fp64_fma_peak_tflops": peak["fp64"] / 1e12,
                "share_of_serial_launch": fit_ms_avg / serial_ms_per_launch,
                "note": "FLOPs by the SURVEY 8(d) convention (3600 per executed LM iteration, from the device niter counters of the "
                        "frame pool the timed region cycles through). achieved / frac = those FLOPs per GPU over the timed region A "
                        "(CUDA events around all K steps), in which several LM launches share the GPU with each other and with the "
                        "detection / consolidation kernels: the kernel's sustained rate, a lower bound; *_lone_launch = the fit kernels "
                        "of one 200-frame launch timed alone by CUDA events. Residual / chi^2 are FP64, Jacobian / normal equations / "
                        "Cholesky FP32, so the FP32 peak is an upper bound the kernel cannot reach"}
    alu_pct, alu_src = None, "profiles/r02f_detect_kernel.txt"
    for cand_src in ("profiles/r02f_detect_kernel.txt", "profiles/r02_detect_kernel.txt"):
        try:
            for ln in open(os.path.join(ROOT, cand_src)):
                if ln.startswith("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"):
                    alu_pct, alu_src = float(ln.split()[-1]), cand_src
            if alu_pct is not None:
                break
        except Exception:
            pass
    roofline_detect = {"bound": "hbm", "alu_pipe_active_pct_ncu": alu_pct, "alu_pipe_source": alu_src, "achieved": det_gbs, "peak": hbm_peak, "unit": "GB/s",
                       "frac": det_gbs / hbm_peak, "traffic": det_traffic, "peak_source": hbm_src,
                       "kernels": "detect_cm + thr + rowmask + scans + emit", "ms_per_launch": det_ms_avg, "frames_per_launch": FL,
                       "algorithmic_bytes_per_launch": det_bytes,
                       "note": "ALU-pipe bound, not HBM bound: the exact 5x5 median (packed-u16 selection network) keeps the ALU "
                               "pipe ~80 % busy with math_pipe_throttle as the top stall (ncu, alu_pipe_source); at the HBM roofline the "
                               "budget would be ~6 ALU operations per pixel, below any exact 5x5 median"}

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_sample = max(cores * 160, 256)
        r = cpu_reference_sample(pool[:1], n_sample, cores)
        cpu = {"value": r["fits_per_s"], "unit": "fits/s", "cores": cores, "kind": _ref_kind(),
               "sample": "%d random candidates of one 512x512 frame of the workload fitted on %d processes "
                         "(%.1f s); detection on 1 full frame (%.3f s)" % (r["n_fits"], cores, r["t_fit"], r["t_detect_per_frame"]),
               "frames_per_s": r["frames_per_s"]}

    line = {
        "metric": METRIC, "value": value, "unit": "fits/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64+f32" if solver == "fast" else "f64", "data": "synthetic",
        "config": static_config(world),
        "run": {"solver": args.solver, "pipeline_depth": args.depth, "e2e_pipeline_depth": e2e_depth, "lm_warps_per_sm_per_batch": args.warps_per_sm,
                "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "fields_per_launch": LF, "frames_per_launch": FL, "launches_per_step_per_rank": len(chunks),
                "candidates_per_step": int(step_counts.sum()), "field_blocks": [list(map(int, b)) for b in blocks],
                "block_candidates_per_step": [int(step_counts[a:b].sum()) for a, b in blocks]},
        "frames_per_s": frames_per_s, "serial_ms_per_launch": serial_ms_per_launch,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "e2e": {"value": e2e_fits_all / (e2e_ms * 1e-3), "unit": "fits/s", "h2d_bytes_per_step": h2d_all // max(args.steps, 1),
                "d2h_bytes_per_step": d2h_all // max(args.steps, 1), "ms_per_step": e2e_ms / args.steps,
                "frames_per_s": args.steps * frames_per_step_all / (e2e_ms * 1e-3),
                "returns": ("final PSF records of find_peptides (R^2 gate, consolidation, re-key on the device) into each rank's pinned host "
                            "memory; per-frame result table gathered with sharding.gather_packed at the end of the region" if args.fetch == "psfs"
                            else "every candidate's fit record"),
                "api": "engine.FieldStream submit/begin_fetch/end_fetch over fsq_detect / fsq_fit_candidates"
                       + (" / fsq_consolidate / fsq_pack_psfs" if args.fetch == "psfs" else "") + " (pinned host frames in, packed results out)",
                "pcie_measured": {"h2d_gbs_per_gpu_min": pcie_min["h2d_gbs"], "d2h_gbs_per_gpu_min": pcie_min["d2h_gbs"],
                                  "how": "%d-MB pinned copies, all %d ranks at once, slowest rank" % (FL * H * W * 2 >> 20, world)}},
        "gpu_launches": launches_all,
        "clocks": clocks, "roofline": roofline, "roofline_detect": roofline_detect,
    }
    if weak:
        line["weak_scaling"] = {"value": world * weak[1] / (weak[0] * 1e-3), "unit": "fits/s", "steps": weak[2],
                                "note": "every rank runs whole 800-field steps (resident inputs), max over ranks"}
    if other is not None:
        line["other_configs"] = other
    if parity is not None:
        line["parity_solver"] = parity
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10, help="timed steps; one step = 800 fields x 10 cycles (default 10 = one configs[4] job)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--solver", default="fast", choices=["fast", "fast64", "minpack-faithful", "minpack-clean"])
    ap.add_argument("--depth", type=int, default=5, help="launches in flight (streams) in the pipelined regions")
    ap.add_argument("--e2e-depth", type=int, default=0, help="slots of the e2e region's FieldStream (default: depth + 2)")
    ap.add_argument("--no-parity-solver", action="store_true")
    ap.add_argument("--diag-no-h2d", action="store_true", help="diagnostic: the e2e region submits device-resident frames (no H2D copy)")
    ap.add_argument("--park", type=int, default=None, help="fsq_lm_opts.park_after override (scheduling only)")
    ap.add_argument("--warps-per-sm", type=int, default=4, choices=[0, 1, 2, 4, 8],
                    help="fsq_lm_opts.warps_per_sm: warps per SM of ONE batch's LM launch (scheduling only; 0 = fill the SM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the configs[3] side measurements (N = 1)")
    ap.add_argument("--launch-fields", type=int, default=LAUNCH_FIELDS,
                    help="fields (x 10 cycles) one pass of the kernels processes: a larger launch amortises the drain of each "
                         "LM launch's last long fits")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: split each step's 800 fields over the ranks (default), or give every rank whole steps")
    ap.add_argument("--fetch", default="psfs", choices=["psfs", "candidates"],
                    help="what a step returns to the host in the e2e region: the final PSF records of find_peptides "
                         "(R^2 gate + consolidation + re-key on the device; default) or every candidate's fit record")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
