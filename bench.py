#!/usr/bin/env python
"""
bench.py -- headline benchmark of the spot hot path (BASELINE.json metric: 2-D Gaussian PSF
candidate fits/s and frames/s) on N B200s of one node, with the reference CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1] -- a 40-frame stack of one 512x512 field,
~500 spots, every frame through detection + per-candidate 5x5 fit + metrics (what
basic_image_script would do on that directory).  One "step" = one pass over one 40-frame stack.
A pool of 8 different stacks (168 MB > 126 MB L2) is cycled so no step re-reads inputs that are
still L2-resident.  N > 1: every rank runs its own stacks (weak scaling, no collective).

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

import numpy as np  # noqa: E402

N_FRAMES, H, W, N_SPOTS = 40, 512, 512, 500
N_VARIANTS = 8
# SURVEY.md 8(d) FLOP convention, `minpack` mode, per executed LM iteration, P = 25, n = 7:
# 8 evaluations (176 P) + Householder QR (2 P n^2 - 2/3 n^3 = 2221) + Q^T f (4 P n = 700) + lmpar (~10 x 350)
FLOP_PER_LM_ITER_5x5 = 176 * 25 + 2221 + 700 + 3500
METRIC = "2-D Gaussian PSF candidate fits/s (detection + 5x5 LM fit + metrics), frames/s alongside"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_stack(seed):
    from fluorosequencingimageanalysis_b200 import synth
    return synth.synth_timetrace(seed, n_frames=N_FRAMES, H=H, W=W, n_spots=N_SPOTS)


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
    frames = make_stack(7000)[:1]
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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: 40-frame stack of one 512x512 field, ~500 spots, every frame fitted",
                   "reference_sample": sample},
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


def stacks_per_launch(steps, requested=None):
    """How many 40-frame stacks (steps) one pass of the kernels processes: the timed regions run EXACTLY `steps` steps, so
    the group size divides it -- 4 by default, else the nearest size that does; an explicit request is reduced to a divisor."""
    steps = max(int(steps), 1)
    if requested is None:
        return next(g for g in (4, 5, 6, 8, 7, 3, 2, 1) if steps % g == 0)
    g = max(1, min(int(requested), steps))
    while steps % g:
        g -= 1
    return g


def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from fluorosequencingimageanalysis_b200 import engine, _lib

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
    # ---- steps per launch: G stacks (steps) go through the kernels together; G divides the number of timed steps
    G = stacks_per_launch(args.steps, args.stacks_per_launch)
    n_launch = args.steps // G                    # launches in each timed region: n_launch * G = args.steps steps exactly
    w_launch = max(3, -(-warmup // G))
    FL = N_FRAMES * G                             # frames per launch
    # ---- synthetic inputs: N_VARIANTS different groups of G 40-frame stacks, host (pinned) and device copies
    base_stacks = [make_stack(1 + 100 * rank + v) for v in range(N_VARIANTS)]
    stacks_host = []
    for v in range(N_VARIANTS):
        st = np.concatenate([base_stacks[(v + j) % N_VARIANTS] for j in range(G)]) if G > 1 else base_stacks[v]
        t = torch.from_numpy(st.view(np.int16)).view(torch.uint16).pin_memory()
        stacks_host.append(t)
    stacks_dev = [t.to(dev) for t in stacks_host]
    in_bytes = FL * H * W * 2
    kw = dict(dtype=torch.uint16, faithful=faithful, solver=solver)
    if args.park is not None:
        kw["park_after"] = args.park
    kw["warps_per_sm"] = args.warps_per_sm
    if args.fetch == "psfs":
        kw["consolidate"] = True          # the tail of find_peptides runs on the device in both timed regions
    cur = torch.cuda.current_stream()

    # ---- timed region A: inputs resident in HBM; K steps software-pipelined over `depth` streams
    fs = engine.FieldStream(FL, H, W, depth=args.depth, host_io=False, **kw)
    totals = torch.zeros(max(n_launch, w_launch), dtype=torch.int64, device=dev)

    def resident_steps(n_steps, first):
        for k in range(n_steps):
            sl = fs.submit(stacks_dev[(first + k) % N_VARIANTS])
            with torch.cuda.stream(sl["stream"]):
                totals[k].copy_(sl["pipe"].n_cand[sl["pipe"].F])          # device-side bookkeeping, no sync
        for sl in fs.slots:
            cur.wait_stream(sl["stream"])

    sampler = ClockSampler(local_rank)
    sampler.start()
    resident_steps(w_launch, 0)
    torch.cuda.synchronize()
    n_probe = int(totals[0].item())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    win_a0 = time.perf_counter()
    ev[0].record()
    resident_steps(n_launch, 3)
    ev[1].record()
    host_enqueue_ms = (time.perf_counter() - win_a0) * 1e3 / args.steps      # host time to queue one step
    barrier()
    win_a1 = time.perf_counter()
    ms_total = ev[0].elapsed_time(ev[1])
    fits_total = int(totals[:n_launch].sum().item())
    pipe = fs.slots[0]["pipe"]
    if int(totals[:n_launch].max().item()) > pipe.cap:
        raise RuntimeError("candidate capacity exceeded")

    # ---- fit launches alone (roofline): re-fit the last detection of slot 0, events per launch
    frames_k = stacks_dev[0]
    pipe.run(frames_k)
    fit_ms = []
    for k in range(max(4, min(args.steps, 8))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.run_fit_only(frames_k)
        e1.record()
        e1.synchronize()
        fit_ms.append(e0.elapsed_time(e1))
    n_last = pipe.total()
    sum_niter = int(pipe.out_int[:n_last, engine.ICOL_NITER].sum().item())
    sum_nfev = int(pipe.out_int[:n_last, engine.ICOL_NFEV].sum().item())
    fit_ms_avg = float(np.mean(fit_ms[1:]))
    # detection alone
    det_ms = []
    for k in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.run_detect_only(stacks_dev[(k + 5) % N_VARIANTS])
        e1.record()
        e1.synchronize()
        det_ms.append(e0.elapsed_time(e1))
    det_ms_avg = float(np.mean(det_ms[1:]))
    # un-pipelined step (one stream, one batch at a time): what a single isolated call costs
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(4):
        pipe.run(stacks_dev[(k + 1) % N_VARIANTS])
    e1.record()
    e1.synchronize()
    serial_ms_per_step = e0.elapsed_time(e1) / 4.0

    # ---- timed region B (e2e): pinned host frames in, packed results back on the host, every step;
    #      submit(k) / begin_fetch(k-1) / end_fetch(k-2) keeps H2D, kernels and D2H of neighbouring steps in flight
    fe = engine.FieldStream(FL, H, W, depth=args.depth, host_io=True, fetch=args.fetch, **kw)

    def e2e_steps(n_steps, first):
        # submit(k), begin_fetch(k - depth + 2), end_fetch(k - depth + 1): depth - 1 batches stay in flight,
        # the host never waits for the batch it has just queued
        fits = d2h = 0
        tick = []
        lag_b, lag_e = max(args.depth - 2, 1), max(args.depth - 1, 2)
        for k in range(n_steps):
            tick.append(fe.submit(stacks_host[(first + k) % N_VARIANTS]))
            if k >= lag_b:
                fe.begin_fetch(tick[k - lag_b])
            if k >= lag_e:
                r = fe.end_fetch(tick[k - lag_e])
                fits += r[0]
                d2h += pipe.d2h_bytes_psfs(r[1]) if args.fetch == "psfs" else pipe.d2h_bytes(r[0])
        for t in tick[max(0, n_steps - lag_e):]:
            r = fe.end_fetch(t)
            fits += r[0]
            d2h += pipe.d2h_bytes_psfs(r[1]) if args.fetch == "psfs" else pipe.d2h_bytes(r[0])
        return fits, d2h

    e2e_steps(max(3, args.depth), 0)
    barrier()
    t0 = time.perf_counter()
    e2e_fits, d2h = e2e_steps(n_launch, 2)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop([(win_a0, win_a1), (t0, t0 + e2e_s)])

    # ---- the parity instrument (reference-faithful MINPACK solver) on the same batch, 2 launches
    parity = None
    if rank == 0 and solver != "minpack" and not args.no_parity_solver:
        pp = engine.FieldPipeline(N_FRAMES, H, W, dtype=torch.uint16, faithful=True, solver="minpack")
        pp.run(frames_k[:N_FRAMES])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pp.run(frames_k[:N_FRAMES])
        e1.record()
        e1.synchronize()
        parity = {"solver": "minpack-faithful (reference behaviour incl. qrsolv diagonal view, FD Jacobian, QR)",
                  "value": pp.total() / (e0.elapsed_time(e1) * 1e-3), "unit": "fits/s", "ms_per_step": e0.elapsed_time(e1)}
        del pp

    # ---- max over ranks, sum of work
    if world > 1:
        tt = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms = float(tt[0].item()), float(tt[1].item())
        cc = torch.tensor([fits_total, e2e_fits], dtype=torch.int64, device=dev)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        fits_all, e2e_fits_all = int(cc[0].item()), int(cc[1].item())
    else:
        e2e_ms = e2e_s * 1e3
        fits_all, e2e_fits_all = fits_total, e2e_fits

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = fits_all / (ms_total * 1e-3)
    frames_per_s = world * args.steps * N_FRAMES / (ms_total * 1e-3)
    # ---- roofline of the dominant kernel (the LM fitter; FP pipes, never tensor cores)
    peak = {}
    for nm, flag in (("fp64", 1), ("fp32", 0)):
        v = ctypes.c_double(0.0)
        _lib.check(L.fsq_fma_peak(flag, ctypes.byref(v), None))
        peak[nm] = v.value
    fl_iter = FLOP_PER_LM_ITER[solver]
    flops = sum_niter * fl_iter
    achieved = flops / (fit_ms_avg * 1e-3) / 1e12
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
    lm_traffic = traffic.get("lmwarp_kernel", {}).get("bytes") if solver == "fast" else None
    det_traffic = traffic.get("detect_cm_packed_kernel", {}).get("bytes")
    if G > 1:                               # the ncu captures are of one-stack launches: G stacks move G times the bytes
        lm_traffic = lm_traffic * G if lm_traffic else None
        det_traffic = det_traffic * G if det_traffic else None
    det_bytes = FL * H * W * 2 + 8 * n_last
    det_gbs = det_bytes / (det_ms_avg * 1e-3) / 1e9
    kname = {"fast": "lmwarp_kernel (+ fit_prep_kernel, fit_finish_kernel) behind fsq_fit_candidates",
             "fast64": "lmfast_kernel<double,true>", "minpack": "lmfit_kernel<8,true>"}[solver]
    pipe_tflops = flops / n_last * (fits_all / world) / (ms_total * 1e-3) / 1e12
    roofline = {"bound": pk, "achieved": pipe_tflops, "peak": peak[pk] / 1e12, "unit": "TFLOP/s",
                "frac": pipe_tflops / (peak[pk] / 1e12), "traffic": lm_traffic,
                "achieved_lone_launch": achieved, "frac_lone_launch": achieved / (peak[pk] / 1e12),
                "achieved_in_pipeline": pipe_tflops, "frac_in_pipeline": pipe_tflops / (peak[pk] / 1e12),
                "kernel": kname, "ms_per_launch": fit_ms_avg,
                "fits_per_launch": n_last, "lm_iterations_per_launch": sum_niter, "passes_per_launch": sum_nfev,
                "flop_per_lm_iteration": fl_iter,
                "peak_source": "fsq_fma_peak %s FMA micro-benchmark, measured in this run (of measured)" % pk.upper(),
                "fp32_fma_peak_tflops": peak["fp32"] / 1e12, "fp64_fma_peak_tflops": peak["fp64"] / 1e12,
                "share_of_serial_step": fit_ms_avg / serial_ms_per_step,
                "note": "FLOPs by the SURVEY 8(d) convention (3600 per executed LM iteration, from the device niter counters). "
                        "achieved / frac = those FLOPs over the timed region A (CUDA events around all K steps), in which "
                        "several LM launches share the GPU with each other and with the detection / consolidation kernels, so "
                        "it is the kernel's sustained rate and a lower bound; *_lone_launch = one launch's fit kernels timed "
                        "alone by CUDA events (one 4-warp block per SM, as launched in the pipeline: latency-bound on its own). "
                        "The kernel keeps residual / chi^2 in FP64 and the Jacobian / normal equations / Cholesky in FP32, so "
                        "the FP32 peak is an upper bound it cannot reach"}
    alu_pct, alu_src = None, "profiles/r01i_detect_kernel.txt"
    try:
        for ln in open(os.path.join(ROOT, alu_src)):
            if ln.startswith("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"):
                alu_pct = float(ln.split()[-1])
    except Exception:
        pass
    roofline_detect = {"bound": "hbm", "alu_pipe_active_pct_ncu": alu_pct, "alu_pipe_source": alu_src, "achieved": det_gbs, "peak": hbm_peak, "unit": "GB/s",
                       "frac": det_gbs / hbm_peak, "traffic": det_traffic, "peak_source": hbm_src,
                       "kernels": "detect_cm + thr + rowmask + scans + emit", "ms_per_launch": det_ms_avg,
                       "algorithmic_bytes_per_launch": det_bytes,
                       "note": "ALU-pipe bound, not HBM bound: the packed-u16 median network (99 compare-exchanges per pixel) keeps the ALU "
                               "pipe 80 % busy with math_pipe_throttle as the top stall (ncu, alu_pipe_source); at the HBM roofline the "
                               "budget would be ~6 ALU operations per pixel, below any exact 5x5 median"}

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_sample = max(cores * 160, 256)
        sample_frames = stacks_host[0].view(torch.int16).numpy().view(np.uint16)[:1]
        r = cpu_reference_sample(sample_frames, n_sample, cores)
        cpu = {"value": r["fits_per_s"], "unit": "fits/s", "cores": cores, "kind": _ref_kind(),
               "sample": "%d random candidates of one 512x512 frame of the workload fitted on %d processes "
                         "(%.1f s); detection on 1 full frame (%.3f s)" % (r["n_fits"], cores, r["t_fit"], r["t_detect_per_frame"]),
               "frames_per_s": r["frames_per_s"]}

    line = {
        "metric": METRIC, "value": value, "unit": "fits/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32" if solver == "fast" else "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: 40-frame stack of one 512x512 field, ~500 spots (sigma 1.5) = one step; "
                               + ("%d stacks go through the kernels per launch; " % G if G > 1 else "") + "every frame: "
                               "detection + 5x5 LM fit of every candidate + metrics",
                   "frames_per_step": N_FRAMES, "candidates_per_step": n_probe // G, "stacks_per_launch": G,
                   "solver": args.solver, "pipeline_depth": args.depth, "lm_warps_per_sm_per_batch": args.warps_per_sm,
                   "l2": "8 different stacks cycled (168 MB > 126 MB L2): inputs larger than L2", "parallelism": "field-sharded x%d, no collective" % world},
        "frames_per_s": frames_per_s, "serial_ms_per_step": serial_ms_per_step / G,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "e2e": {"value": e2e_fits_all / (e2e_ms * 1e-3), "unit": "fits/s", "h2d_bytes_per_step": in_bytes // G,
                "d2h_bytes_per_step": d2h // max(args.steps, 1), "frames_per_s": world * args.steps * N_FRAMES / (e2e_ms * 1e-3),
                "returns": ("final PSF records of find_peptides (R^2 gate, consolidation, re-key on the device)" if args.fetch == "psfs"
                            else "every candidate's fit record"),
                "api": "engine.FieldStream submit/begin_fetch/end_fetch over fsq_detect / fsq_fit_candidates"
                       + (" / fsq_consolidate / fsq_pack_psfs" if args.fetch == "psfs" else "") + " (pinned host frames in, packed results out)"},
        "gpu_launches": n_launch * fs.kernels_per_run,
        "clocks": clocks, "roofline": roofline, "roofline_detect": roofline_detect,
    }
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
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--solver", default="fast", choices=["fast", "fast64", "minpack-faithful", "minpack-clean"])
    ap.add_argument("--depth", type=int, default=5, help="launches in flight (streams) in the pipelined regions")
    ap.add_argument("--no-parity-solver", action="store_true")
    ap.add_argument("--park", type=int, default=None, help="fsq_lm_opts.park_after override (scheduling only)")
    ap.add_argument("--warps-per-sm", type=int, default=4, choices=[0, 1, 2, 4, 8],
                    help="fsq_lm_opts.warps_per_sm: warps per SM of ONE batch's LM launch (scheduling only; 0 = fill the SM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stacks-per-launch", type=int, default=None,
                    help="40-frame stacks (steps) processed by one pass of the kernels: a larger launch amortises the "
                         "drain of each LM launch's last long fits; default: 4, or the nearest size that divides --steps; "
                         "an explicit value is reduced to a divisor of --steps")
    ap.add_argument("--fetch", default="psfs", choices=["psfs", "candidates"],
                    help="what a step returns to the host in the e2e region: the final PSF records of find_peptides "
                         "(R^2 gate + consolidation + re-key on the device; default) or every candidate's fit record")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
