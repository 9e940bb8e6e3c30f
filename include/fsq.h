/*
 * fsq.h -- C-ABI of libfsq.so: the B200 (sm_100a) implementation of the per-field spot hot
 * path of marcottelab/FluorosequencingImageAnalysis.
 *
 * Every entry point replaces the inside of one reference function (cited per function as
 * file:line under /root/reference); the reference-side binding is the ctypes stub shown in
 * INTEGRATION.md.  Plain pointers and sizes only: all data pointers are DEVICE pointers
 * unless the name ends in _host; the caller (Python/torch) owns every buffer, inputs and
 * outputs, including scratch.  All launches are asynchronous on `stream` (a cudaStream_t
 * passed as void*; NULL = legacy default stream).  The library keeps no mutable global state
 * besides a thread-local error string and is safe to call from N threads/processes each
 * bound to one GPU.
 *
 * Return convention: 0 success; <0 failure (see FSQ_E_*); text via fsq_last_error().
 * Per-fit failures are NOT errors: they are `status` values with mpfit's meaning
 * (agpy/mpfit/mpfit.py:754-790): 0 bad input, 1..4 converged, 5 maxiter, 6..8 tolerance too
 * small, -16 non-finite.
 */
#ifndef FSQ_H_
#define FSQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSQ_VERSION 200            /* 0.2.0 */

#define FSQ_OK          0
#define FSQ_E_ARG      -1          /* bad argument (NULL pointer, even kernel size, unsupported dtype ...) */
#define FSQ_E_CAPACITY -2          /* output capacity too small; required count is written where documented */
#define FSQ_E_CUDA     -3          /* a CUDA runtime call failed */
#define FSQ_E_RANGE    -4          /* data outside the exact-arithmetic range of the kernels */

/* pixel dtype codes for frame / window inputs */
#define FSQ_U8   0
#define FSQ_U16  1
#define FSQ_I16  2
#define FSQ_I32  3
#define FSQ_F64  4                 /* windows only (gaussfit on float data) */
#define FSQ_I64  5                 /* windows only (pflib passes int64 sub-images) */

int         fsq_version(void);
const char* fsq_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Candidate detection -- replaces pflib._psf_candidates (pflib.py:217-258) for a batch of
 * frames:  int64 cast -> s x s median background removal (scipy 'reflect' borders, rank
 * s*s/2) -> k x k integer correlation (zero padded) -> clamp >= 0 -> per-frame threshold
 * mean + c_std*std over ALL pixels -> every pixel >= threshold with a 2-px border excluded,
 * raster order within a frame, frames in order.
 *
 *  frames      [n_frames, H, W] pixels of `dtype_code` (FSQ_U8/U16/I16/I32), contiguous
 *  K_host      HOST pointer, [ksize*ksize] int64 correlation template, ksize odd, <= 9
 *  mf_size     median filter side, 1..9
 *  c_std       threshold coefficient
 *  cand_hw     out [cap, 2] int32 (h, w)
 *  cand_frame  out [cap] int32 frame index of each candidate
 *  n_cand      out [n_frames + 1] int64: per-frame counts, then the total in the last slot
 *  thr         out [n_frames] float64 thresholds
 *  cap         capacity of cand_hw / cand_frame in candidates; candidates beyond cap are
 *              dropped (the caller compares n_cand[n_frames] with cap after synchronising and
 *              re-launches with a larger buffer -- that is the FSQ_E_CAPACITY protocol, made
 *              asynchronous)
 *  scratch     device scratch of at least fsq_detect_scratch_bytes(n_frames, H, W) bytes
 * ------------------------------------------------------------------------------------------ */
int64_t fsq_detect_scratch_bytes(int n_frames, int H, int W);

int fsq_detect(const void* frames, int dtype_code, int n_frames, int H, int W,
               const int64_t* K_host, int ksize, int mf_size, double c_std,
               int32_t* cand_hw, int32_t* cand_frame, int64_t* n_cand, double* thr,
               int64_t cap, void* scratch, int64_t scratch_bytes, void* stream);

/* Synchronises `stream` and reports FSQ_E_RANGE when the last fsq_detect on this scratch met
 * data outside the exact-arithmetic range (correlation value >= 2^41, or a threshold beyond
 * the saturated uint32 correlation scratch, i.e. > 2^32). */
int fsq_detect_flags(const void* scratch, int n_frames, int H, int W, void* stream);

/* Optional outputs of the detection intermediates (tests / debugging): the clamped
 * correlation map saturated to uint32 as stored in scratch, [n_frames, H, W]. */
int fsq_detect_copy_cm32(const void* scratch, int n_frames, int H, int W, uint32_t* cm32_out,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * Levenberg-Marquardt options -- the keyword arguments of class mpfit
 * (agpy/mpfit/mpfit.py:600-605); gaussfit leaves them all at their defaults.
 * ------------------------------------------------------------------------------------------ */
typedef struct fsq_lm_opts {
    double ftol;        /* 1e-10 */
    double xtol;        /* 1e-10 */
    double gtol;        /* 1e-10 */
    double factor;      /* 100   */
    int32_t maxiter;    /* 200   */
    int32_t faithful;   /* 1: reproduce the reference's qrsolv diagonal-view behaviour
                              (mpfit.py:1915,1956,1976-1977); 0: clean MINPACK            */
    int32_t want_perror;/* 1: compute covariance -> perror (mpfit.py:1361-1388); MINPACK solver only */
    int32_t solver;     /* FSQ_SOLVER_*                                                    */
    int32_t park_after; /* FAST solver scheduling only (results do not depend on it).
                           > 0: a fit that has not ended after this many passes over its window is parked and
                                finished by a second launch over the parked fits;
                           < 0: drain parking -- once the work queue is empty, whatever is still running
                                -park_after iterations later (the 100+-iteration stragglers, 0.1 % of the fits)
                                is parked, so the launch returns its SM slots instead of keeping hundreds of
                                warps alive for one lane each; a second launch finishes the parked fits packed
                                into 16 blocks.  Measured on B200 in the pipelined step: +4 % (four stacks per
                                launch) to +16 % (one stack per launch); engine.FieldPipeline uses -8.
                           0 = one launch (default).                                          */
    int32_t warps_per_sm;/* FAST solver scheduling only: warps per SM of the persistent LM launch: 0 (the
                           default: the launch fills the machine, 12 warps per SM), 4, 2 or 1.  A small value leaves
                           most of every SM to launches queued on other streams: with several batches in
                           flight (engine.FieldStream) each thread then works through many more fits, so
                           far fewer warp-ticks are spent on half-empty warps waiting for their last long
                           fit, and one batch's tail runs underneath the other batches' bulk.
                           fsq_gaussfit_batch, 11x11 windows: 0 = a thread-per-window launch whose
                           fits are parked after park_after passes (default 32) and finished by the kernel with 8 or
                           with 4 lanes per window (pass split over the lanes, shuffle reductions; which of the two
                           depends on the number of parked fits and is decided on the device); -1 = thread per window
                           only; -2 / -3 / -4 = 4 / 8 / 2 lanes per window for every fit; -5 / -6 = thread-per-window
                           launch + 4- / 8-lane finish.  Same algorithm, different summation order in the lane-group passes: fits
                           that end before they are parked are bit-identical in every arrangement.          */
} fsq_lm_opts;

/* Solvers behind the two fit entry points.
 *  MINPACK     the reference's algorithm operation for operation in FP64: forward-difference
 *              Jacobian (8 model evaluations per iteration), pivoted Householder QR, lmpar/qrsolv
 *              (with the reference's diagonal-view behaviour when faithful = 1).  One sub-warp per
 *              window.  This is the parity instrument.
 *  FAST        the production fitter of the frame path (fsq_fit_candidates): the same bounded
 *              trust-region LM (same pegging / alpha / snapping / termination rules) driven by the
 *              analytic Jacobian through column-scaled normal equations and a 7x7 Cholesky.  One
 *              thread per 5x5 window, the 32 fits of a warp in lock step, one pass over the
 *              window per LM iteration; residual / chi^2 in FP64, Jacobian and normal equations
 *              in FP32.  (fsq_gaussfit_batch runs FAST64 for this code.)
 *  FAST64      the same algorithm entirely in FP64, one thread per window, two passes per
 *              iteration; 5x5 windows.  Cross-check of FAST and generic-window entry point.
 * `faithful` and `want_perror` are ignored by the FAST solvers; n_qrsolv then counts damped
 * (par > 0) solves, so n_qrsolv == 0 still means "every step was a plain Gauss-Newton step". */
#define FSQ_SOLVER_MINPACK    0
#define FSQ_SOLVER_FAST64     1
#define FSQ_SOLVER_FAST       2

void fsq_lm_default_opts(fsq_lm_opts* o);

/* ------------------------------------------------------------------------------------------
 * Batched 2-D Gaussian fit -- replaces gaussfitter.gaussfit (agpy/gaussfitter.py:142-255)
 * -> class mpfit (agpy/mpfit/mpfit.py:600-1388) for the 7-parameter rotated elliptical
 * Gaussian (height, amplitude, p2, p3, width_x, width_y, rota_deg) with err=None.
 *
 *  windows     [n, win, win] of `dtype_code` (FSQ_F64, FSQ_I64, FSQ_U16, FSQ_I32), win <= 11
 *  p0          [n, 7] float64 start values (already clamped like gaussfitter.py:202-204)
 *  lo, hi      [n, 7] float64 limits;  lim_lo, lim_hi [n, 7] uint8 "limited" flags
 *  params      out [n, 7]   perror out [n, 7] (may be NULL unless want_perror)
 *  status/niter/nfev out [n] int32;  chi2 out [n] (= mpfit .fnorm)
 *  n_qrsolv    out [n] int32 number of qrsolv calls (0 <=> robust set); may be NULL
 *  fit_img     out [n, win, win] float64 model at the final parameters; may be NULL
 * ------------------------------------------------------------------------------------------ */
int fsq_gaussfit_batch(const void* windows, int dtype_code, int64_t n, int win,
                       const double* p0, const double* lo, const double* hi,
                       const uint8_t* lim_lo, const uint8_t* lim_hi, const fsq_lm_opts* opts,
                       double* params, double* perror, int32_t* status, int32_t* niter,
                       int32_t* nfev, double* chi2, int32_t* n_qrsolv, double* fit_img,
                       int64_t* work_counter, void* stream);

/* The rest of gaussfit's surface (agpy/gaussfitter.py:188-232) and mpfit's .covar (agpy/mpfit/mpfit.py:1361-1388,
 * 2274-2336); FSQ_SOLVER_MINPACK only.  Every extra argument may be NULL / 0, which gives fsq_gaussfit_batch.
 *  fixed   [n, 7] uint8   parinfo 'fixed' (mpfit.py:917-921): the parameter keeps its start value, the algorithm runs on
 *                         the free ones only (x = xall[ifree], mpfit.py:943-948).  gaussfit's vheight = 0 is "height
 *                         fixed at 0" (gaussfitter.py:195-198), rotate = 0 is "rota fixed at 0" (cos 0 = 1, sin 0 = 0
 *                         exactly, so the rotated model equals the unrotated one bit for bit)
 *  err     [n, win, win]  float64: residuals are (data - model) / err (gaussfitter.py:218)
 *  circle  1 = the circular model (gaussfitter.py:104-107): width_y := width_x, no rotation; parameters 5 and 6 are not
 *          fitted and come back as (width, 0); params / perror / covar keep the 7-slot layout
 *  covar   out [n, 7, 7]  mpfit .covar: (R^T R)^-1 un-pivoted with zero rows / columns for fixed parameters; NaN when
 *          the fit ended with status <= 0 (the reference leaves None there)
 * perror is zero for fixed parameters (mpfit.py:1382-1386). */
int fsq_gaussfit_batch_ex(const void* windows, int dtype_code, int64_t n, int win,
                          const double* p0, const double* lo, const double* hi,
                          const uint8_t* lim_lo, const uint8_t* lim_hi, const uint8_t* fixed,
                          const double* err, int circle, const fsq_lm_opts* opts,
                          double* params, double* perror, double* covar, int32_t* status, int32_t* niter,
                          int32_t* nfev, double* chi2, int32_t* n_qrsolv, double* fit_img,
                          int64_t* work_counter, void* stream);

/* Same fit with a per-trial-step trace for the first trace_n windows (tests / debugging).
 * trace is [trace_n, trace_steps, 20] float64, zero-initialised by the caller; record 0 of a
 * window holds the number of records written in [0]; record k >= 1 = (niter, accepted, status,
 * fnorm, fnorm1, delta, par, ratio, alpha, pnorm, xnew[0..6], n_qrsolv, actred, prered). */
int fsq_gaussfit_batch_trace(const void* windows, int dtype_code, int64_t n, int win,
                             const double* p0, const double* lo, const double* hi,
                             const uint8_t* lim_lo, const uint8_t* lim_hi, const fsq_lm_opts* opts,
                             double* params, int32_t* status, int32_t* niter, int32_t* nfev,
                             double* chi2, int32_t* n_qrsolv, double* trace, int trace_steps,
                             int64_t trace_n, int64_t* work_counter, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fit every candidate of a batch of frames the way pflib.find_peptides does
 * (pflib.py:441-477 -> _fit_2d_gaussian pflib.py:180-214): 5x5 window sliced from the RAW
 * frame around (h, w); start (median, max, 2.5, 2.5, 1, 1, 0); limits
 * [0,inf) x [(max-mean)/3, inf) x [2,3]^2 x [0.75,2]^2 x [0,360]; then the fit-quality
 * metrics of pflib.py:461-473 fused at the end.
 *
 *  out_fit     [n, 12] float64: h_0, w_0, H, A, sigma_h, sigma_w, theta (image coordinates,
 *              pflib.py:461), rmse, r_2, s_n, chi2, (reserved)
 *  out_int     [n, 4] int32: status, niter, nfev, n_qrsolv
 *  fit_img     [n, 25] float64 or NULL
 *  scratch     device scratch of at least fsq_fit_scratch_bytes(n) bytes (work-queue head, one
 *              128-byte start record per candidate, the states of parked fits); the call initialises it
 * ------------------------------------------------------------------------------------------ */
int fsq_fit_candidates(const void* frames, int dtype_code, int n_frames, int H, int W,
                       const int32_t* cand_hw, const int32_t* cand_frame, int64_t n,
                       const int64_t* n_dev /* optional device count (n_cand total); NULL = use n */,
                       const fsq_lm_opts* opts, double* out_fit, int32_t* out_int,
                       double* fit_img, void* scratch, int64_t scratch_bytes, void* stream);

int64_t fsq_fit_scratch_bytes(int64_t n);

/* ------------------------------------------------------------------------------------------
 * Start values from image moments -- replaces gaussfitter.moments (agpy/gaussfitter.py:29-61) with
 * circle = 0, rotate = 1, vheight = 1 and the default median estimator, for a batch of windows:
 *  p0_out [n, 7] = (height = median, amplitude = max - height, x, y, width_x, width_y, 0).
 * A window holding a NaN gives NaN height / amplitude (the Python mirror raises the reference's
 * ValueError("something is nan")).  windows [n, win, win] of any FSQ_* dtype, win <= 11.  Sums run
 * in numpy's add.reduce order, so float64 windows reproduce the reference's start values bit for bit.
 * lo / hi / lim_lo / lim_hi [7] (device, all four or all NULL): clip the start values into the
 * limits as gaussfit does before calling mpfit (gaussfitter.py:202-204).
 * ------------------------------------------------------------------------------------------ */
int fsq_moments(const void* windows, int dtype_code, int64_t n, int win, const double* lo, const double* hi,
                const uint8_t* lim_lo, const uint8_t* lim_hi, double* p0_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * R^2 gate + rival consolidation + re-key -- replaces the tail of pflib.find_peptides
 * (pflib.py:466-468 `if r_2 < r_2_threshold: continue`, :479-512 the consolidation loop, :514-519 the
 * re-key) for a whole batch of frames, on the packed records fsq_detect / fsq_fit_candidates wrote.
 *   in   cand_hw [n,2], cand_frame [n]   candidates sorted by (frame, h, w) -- the order fsq_detect emits
 *        out_fit [n,12]                  columns 0,1 = fitted centre h_0, w_0; column 8 = r_2
 *        n, n_dev                        capacity / device-resident count (as fsq_fit_candidates)
 *   out  psf_state [n] u8   0 = dropped by the R^2 gate, 1 = deleted by a rival, 2 = final PSF keyed by its
 *                           candidate pixel, 3 = final PSF re-keyed to (round(h_0), round(w_0)) -- in the
 *                           reference's dictionary such an entry moves to the END (delete + setdefault)
 *        psf_key [n,2] i32  the final key (python-2 round(): half away from zero) of states 2/3
 *        n_psf [n_frames] i64 (may be NULL)  final PSFs per frame
 *        flags [1] i32      bit 0: a re-keyed PSF landed on an occupied key -- the reference's assert at
 *                           pflib.py:518 would fire (impossible for pflib fits, whose centres stay within
 *                           0.5 px of the candidate pixel; detected within the rival reach)
 * The reference's result depends on its visiting order (raster) and on `>` at :508 (ties delete the visiting
 * PSF); both are reproduced exactly: rivals form tiny connected components, each replayed sequentially.
 * Returns FSQ_E_ARG for consolidation_radius < 2 (ValueError at pflib.py:431-432).
 * ------------------------------------------------------------------------------------------ */
int64_t fsq_consolidate_scratch_bytes(int64_t n, int n_frames);
int fsq_consolidate(const int32_t* cand_hw, const int32_t* cand_frame, const double* out_fit, int64_t n,
                    const int64_t* n_dev, int n_frames, double r_2_threshold, int consolidation_radius,
                    uint8_t* psf_state, int32_t* psf_key, int64_t* n_psf, int32_t* flags,
                    void* scratch, int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * The final PSFs of a batch, packed in the order of the dictionary pflib.find_peptides returns
 * (pflib.py:514-520): frames in sequence; inside a frame first the PSFs keyed by their candidate pixel
 * in raster order, then the re-keyed ones (the reference deletes and re-inserts those, which appends).
 *   in   psf_state, psf_key from fsq_consolidate; cand_frame, out_fit, n, n_dev as there
 *   out  psf_fit [cap_psf,12] f64   the out_fit rows of the final PSFs
 *        psf_int [cap_psf,4] i32    (frame, key_h, key_w, candidate index)
 *        psf_base [n_frames+1] i64  offset of each frame's PSFs; psf_base[n_frames] = total (rows beyond
 *                                   cap_psf are not written: compare the total with cap_psf)
 * This is what leaves the device in the production pipeline: ~1 record per spot instead of ~10 candidates.
 * ------------------------------------------------------------------------------------------ */
int64_t fsq_pack_psfs_scratch_bytes(int64_t n, int n_frames);
int fsq_pack_psfs(const uint8_t* psf_state, const int32_t* psf_key, const int32_t* cand_frame,
                  const double* out_fit, int64_t n, const int64_t* n_dev, int n_frames,
                  double* psf_fit, int32_t* psf_int, int64_t* psf_base, int64_t cap_psf,
                  void* scratch, int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Luminosity-centroid particle tracking -- replaces Experiment.luminosity_centroid_particle_tracking and
 * next_frame_spot_by_luminosity_centroid (flexlibrary.py:1173-1317), the tracker of the time-trace path
 * (basic_timetrace_script.py -> TimetraceExperiment.lc_create_traces, flexlibrary.py:3309-3382), for n
 * spots of frame 0 through the n_frames frames of their field.
 *   in   frames [n_fields, n_frames, H, W]; spots_hw [n,2] i32 positions in frame 0; spot_field [n] i32
 *        (NULL: one field); offsets [n_frames,2] i32 per-frame (delta_h, delta_w) (NULL: none; the
 *        reference's caller passes None); spot_size (Spot.size, 3 or 5), search_radius (3), s_n_cutoff (3.0)
 *   out  track_hw [n, n_frames, 2] i32   the spot's pixel in every frame, (-1, -1) where the reference has None
 *        track_state [n, n_frames] u8    0 None (window or spot square cut by the border), 1 centroid accepted,
 *                                        2 Illumina S/N below the cut-off: stays at the prior position, 3 frame 0
 *        track_sn [n, n_frames] f64      pflib.illumina_s_n of the square at the centroid position (NaN for None)
 * Integer pixel sums; centre of mass = one float64 division per axis (scipy center_of_mass); python-2 round().
 * A window whose pixels are all zero gives None (the reference divides by zero and raises there).
 * ------------------------------------------------------------------------------------------ */
int fsq_track_centroid(const void* frames, int dtype_code, int n_fields, int n_frames, int H, int W,
                       const int32_t* spots_hw, const int32_t* spot_field, const int32_t* offsets, int64_t n,
                       int spot_size, int search_radius, double s_n_cutoff,
                       int32_t* track_hw, uint8_t* track_state, double* track_sn, void* stream);

/* ------------------------------------------------------------------------------------------
 * Greedy cross-frame particle tracking -- replaces Experiment.greedy_particle_tracking
 * (flexlibrary.py:680-1027; drop-outs :626-677, offsets :566-617), the tracker of the experiment path (one
 * frame per Edman cycle), for the spots of n_fields fields at once.
 *   in   spot_hw [n,2] f64 (Spot.h, Spot.w), spots sorted by (field, frame); seg_start [n_fields*n_frames+1]
 *        i32 first spot of every (field, frame); cum_offsets [n_fields, n_frames, 2] f64 = accumulate_offsets
 *        of the per-frame offsets (NULL: no drift); candidate_radius (2), spot_radius (0)
 *   out  anc / desc [n] i32  the spot's ancestor / descendant (global spot index, -1 none): the doubly linked
 *                            lists a_L / d_L of the reference's frame_bins; a trace is a chain from a spot with
 *                            anc = -1, and may skip frames
 *        bin_hw [n,2] i32    rounded drift-corrected pixel (python-2 round), (-1,-1) for dropped spots
 *        discarded [n] u8    1: the spot would leave some frame of the sequence (discard_dropouts)
 *        flags [n_fields] i32  bit 0: two spots of one frame share a pixel (the reference asserts, :853-858)
 * The reference sorts all (ancestor, candidate) pairs of a frame by distance (stable sort over ancestors in
 * raster order, candidates in raster order) and links greedily; the kernel replays that walk exactly in
 * parallel rounds of mutually-best pairs.  Scratch is two H x W index grids per field.
 * ------------------------------------------------------------------------------------------ */
int64_t fsq_track_greedy_scratch_bytes(int n_fields, int H, int W, int64_t n);
int fsq_track_greedy(const double* spot_hw, const int32_t* seg_start, const double* cum_offsets,
                     int n_fields, int n_frames, int H, int W, int64_t n, int candidate_radius,
                     double spot_radius, int32_t* anc, int32_t* desc, int32_t* bin_hw, uint8_t* discarded,
                     int32_t* flags, void* scratch, int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Frame registration -- replaces phase_correlate.phase_correlate (phase_correlate.py:11-196; caller
 * SequenceExperiment.offsets_from_frames, flexlibrary.py:1717-1741) for n_pairs image pairs in one call.
 *  ref, reg   [n_pairs, rows, cols] of `dtype_code` (any FSQ_* code); the two may overlap (consecutive frames of one
 *             stack: reg = ref + one frame)
 *  out        [n_pairs, 4] float64: row_shift, col_shift, error, diffphase -- the reference's return tuple
 *  upsample_factor = 1: whole-pixel registration; > 1: the matrix-multiply DFT on a ceil(1.5 usf)^2 grid around the peak
 * The two forward and the inverse 2-D FFT are cuFFT Z2Z transforms (complex128 like numpy.fft); everything else --
 * spectrum product, numpy's complex argmax, twiddle tables, the two DFT products, error / phase -- are kernels of this
 * library.  cuFFT plans are cached per (device, rows, cols, n_pairs); scratch is caller-owned.
 * ------------------------------------------------------------------------------------------ */
int64_t fsq_phase_correlate_scratch_bytes(int n_pairs, int rows, int cols, int upsample_factor);
int fsq_phase_correlate(const void* ref, const void* reg, int dtype_code, int n_pairs, int rows, int cols,
                        int upsample_factor, double* out, void* scratch, int64_t scratch_bytes, void* stream);

/* Fit-quality metrics for arbitrary (sub_img, fit_img) pairs -- pflib.py:463-473 and
 * illumina_s_n pflib.py:261-281.  sub [n,25] int64, fit [n,25] float64 -> out [n,3] (r_2, rmse, s_n) */
int fsq_metrics(const int64_t* sub, const double* fit, int64_t n, double* out, void* stream);

/* pflib.illumina_s_n (pflib.py:261-281) for n square windows of any side 1..33 (Spot sizes are a parameter,
 * flexlibrary.py:319-320): sub [n,size,size] int64 -> out [n] float64, in numpy's own summation order, so the value
 * is bit-identical to the reference's */
int fsq_illumina_s_n(const int64_t* sub, int64_t n, int size, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Photometry on spots -- flexlibrary.Spot.photometry family (flexlibrary.py:160-210, 264-284)
 *  method 0: simple (sum of size x size), 1: mexican_hat (radius, brim), 2: maximum (radius, top=1)
 *  spots_hw [n,2] int32, spot_frame [n] int32, out [n] float64
 * ------------------------------------------------------------------------------------------ */
int fsq_photometry(const void* frames, int dtype_code, int n_frames, int H, int W,
                   const int32_t* spots_hw, const int32_t* spot_frame, int64_t n,
                   int method, int radius, int brim, double* out, void* stream);

/* FP64/FP32 FMA micro-benchmark used for the roofline denominator (bench.py); returns
 * achieved FLOP/s through *flops_out (host pointer). */
int fsq_fma_peak(int fp64, double* flops_out_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSQ_H_ */
