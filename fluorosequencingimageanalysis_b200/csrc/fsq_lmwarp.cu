// `fast` solver (FSQ_SOLVER_FAST): the production fitter of the frame path.
//
// Same bounded trust-region Levenberg-Marquardt as class mpfit (agpy/mpfit/mpfit.py:600-1388):
// pegging by exact equality and gradient sign (:1073-1091), More's lmpar (:2077-2190), step
// clipping / alpha scaling / snapping (:1184-1231), ratio / delta / par updates (:1253-1288),
// termination tests (:1301-1335), .fnorm (:1357-1359) -- driven by the ANALYTIC Jacobian of the
// rotated elliptical Gaussian (agpy/gaussfitter.py:63-140) through column-scaled normal
// equations and a 7x7 Cholesky, instead of 7 finite-difference model evaluations and a
// Householder QR per iteration.
//
// Execution model (what makes it fast on an SM):
//   * one THREAD per 5x5 window, but the 32 fits of a warp advance in lock step: every lane
//     runs the same "tick" = [one pass over the 25 pixels at its current trial point] ->
//     [accept / reject bookkeeping] -> [solve for the next trial point].  No lane is ever in
//     a different phase than its neighbours, so the warp does not serialise;
//   * ONE pass per LM iteration: residual, chi^2, Jacobian, J^T J and J^T f are all formed at
//     the trial point.  When the step is accepted (the common case) they are the next
//     iteration's normal equations; when it is rejected the previous ones are still in shared
//     memory.  (The reference: 7 + 1 evaluations per iteration.)
//   * precision split by what each quantity decides: residual and chi^2 -- which drive the
//     ftol = 1e-10 termination test and the gain ratio -- in FP64 (exp in FP64); the Jacobian,
//     J^T J, J^T f and the Cholesky solve -- which only shape the step -- in FP32 with column
//     scaling (cond(J) <= 40 after scaling); parameters, bounds and the
//     trust-region scalars in FP64;
//   * per-thread matrices live in shared memory ([entry][thread], conflict free), registers
//     hold the accumulators of the running pass only;
//   * a lane whose fit has ended takes the next window from an atomic queue (persistent grid),
//     so a 200-iteration fit never holds 31 finished neighbours.
//
// Per-window start values (median / max / mean of the 25 pixels, pflib.py:199-213) and the two
// fit-independent quality figures (total sum of squares for r_2, Illumina S/N) come from a
// separate fully-parallel kernel (fit_prep_kernel), so the refill inside the LM kernel is 25
// loads and a handful of flops.
#include "fsq_common.cuh"
#include "fsq_median.cuh"
#include "fsq_chol7.cuh"
#include <string.h>
#include <atomic>
#include <type_traits>

namespace fsq {

constexpr int WNP = 7;
constexpr int WNT = 28;
constexpr int WTHREADS = 128;
#ifndef WMINB
#define WMINB 3      // 3 CTAs/SM (168 registers, no spills since the forward-differenced pass freed the FP64 tables):
                     // +4..10 % in the pipelined step over 2 CTAs/SM (212 registers); 4 CTAs/SM (128 registers) spills and is slower
#endif
#define WQ_MACHEP 2.220446049250313e-16
#define WQ_DWARF 2.2250738585072014e-308
#define WQ_DEG2RAD 0.017453292519943295

__host__ __device__ constexpr int wtri(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

struct WarpArgs {
    const void* frames; int fdtype; int H; int W;
    const int32_t* cand_hw; const int32_t* cand_frame;
    // generic-window entry (fsq_gaussfit_batch): windows + per-fit start / limits in, mpfit fields out
    const void* windows; int wdtype;
    const double* p0; const double* lo; const double* hi; const uint8_t* lim_lo; const uint8_t* lim_hi;
    double* params; int32_t* status; int32_t* niter; int32_t* nfev; double* chi2; int32_t* n_damped;
    long long n; const long long* n_dev;
    fsq_lm_opts o;
    float ftol_f, xtol_f, gtol_f, factor_f;   // the tolerances of o as FP32 (converted once on the host: a conversion in the kernel is re-done every tick)
    double* out_fit; int32_t* out_int; double* fit_img;
    unsigned long long* work_counter;        // queue head of the running launch
    unsigned long long* strag_count;         // number of parked fits (phase 1 appends, phase 2 consumes)
    struct StragRec* strag;                  // parked fit states
    struct PrepRec* prep;                    // per-candidate start records
    int cap;                                 // phase 1: park a fit after this many passes (0 = never)
    int drain_grace;                         // phase 1: once the queue is empty, park what is still running after this many ticks (0 = never)
    int resume;                              // phase 2: the work list is strag[0 .. *strag_count)
    long long resume_lo, resume_hi;          // phase 2 runs only if resume_lo <= *strag_count <= resume_hi (else this launch is a no-op)
};

// Everything the LM kernel needs to start one candidate, written by fit_prep_kernel: a refill is eight
// 16-byte loads from one 128-byte line instead of two dependent round trips (candidate list, then 25
// scattered pixels) -- the refill was 10 % of the kernel's time with 2 lanes of every warp refilling per tick.
struct __align__(16) PrepRec {
    int px[25];                  // the 5x5 raw window, raster order
    int med, max;                // numpy.median / max of the window (pflib.py:199-200)
    unsigned hw;                 // candidate pixel: h << 16 | w
    double sst;                  // total sum of squares around the window mean (r_2, pflib.py:464)
    double lo1;                  // amplitude floor (max - mean) / 3 (pflib.py:205): two FP64 divisions that the fully
                                 // parallel start-record kernel does, not the two refilling lanes of an LM warp
};
static_assert(sizeof(PrepRec) == 128, "PrepRec must be 128 bytes");

// State of a fit parked by phase 1 (everything the LM iteration carries from one accepted point
// to the next; the normal equations are re-formed from x by the resuming launch, bit for bit).
struct StragRec {
    long long idx;
    double x[7];
    double ss0;
    float diag[7];
    float delta, par, xnorm;
    int niter, nfev, n_damped;
    int pad;
};
static_assert(sizeof(StragRec) == 128, "StragRec must be 128 bytes");

__device__ __forceinline__ int w_ld_int(const void* base, int dtype, size_t off) {
    switch (dtype) {
        case FSQ_U8:  return (int)((const uint8_t*)base)[off];
        case FSQ_U16: return (int)((const uint16_t*)base)[off];
        case FSQ_I16: return (int)((const int16_t*)base)[off];
        default:      return ((const int32_t*)base)[off];
    }
}

__device__ __forceinline__ double w_ld_dbl(const void* base, int dtype, size_t off) {
    switch (dtype) {
        case FSQ_U8:  return (double)((const uint8_t*)base)[off];
        case FSQ_U16: return (double)((const uint16_t*)base)[off];
        case FSQ_I16: return (double)((const int16_t*)base)[off];
        case FSQ_I32: return (double)((const int32_t*)base)[off];
        case FSQ_I64: return (double)((const long long*)base)[off];
        default:      return ((const double*)base)[off];
    }
}

// pflib limits (pflib.py:199-213): lower on all seven, upper on centres, widths, angle
__device__ __forceinline__ double pf_lo(int j, double lo1) {
    return j == 0 ? 0.0 : j == 1 ? lo1 : (j == 2 || j == 3) ? 2.0 : (j == 4 || j == 5) ? 0.75 : 0.0;
}
__device__ __forceinline__ double pf_hi(int j) { return (j == 2 || j == 3) ? 3.0 : (j == 4 || j == 5) ? 2.0 : 360.0; }
constexpr unsigned PF_QLL = 0x7fu, PF_QUL = 0x7cu;

// Box limits of one fit: compile-time constants on the pflib path (only the amplitude floor varies),
// per-fit arrays in global memory on the generic path (read where needed: twice per tick).
template <bool PFLIB>
struct Lim {
    double lo1; const double* lo; const double* hi; unsigned qll, qul;
    __device__ __forceinline__ bool has_lo(int j) const { return PFLIB ? true : ((qll >> j) & 1u); }
    __device__ __forceinline__ bool has_hi(int j) const { return PFLIB ? ((PF_QUL >> j) & 1u) : ((qul >> j) & 1u); }
    __device__ __forceinline__ double lower(int j) const { return PFLIB ? pf_lo(j, lo1) : lo[j]; }
    __device__ __forceinline__ double upper(int j) const { return PFLIB ? pf_hi(j) : hi[j]; }
};

// -------------------------------------------------------------------------------------------
// Start values + fit-independent metrics, one thread per candidate.
//   out_int[i] = (median, max, sum, 0)  -- consumed and overwritten by the LM kernel
//   out_fit[i][8] = total sum of squares (becomes r_2), out_fit[i][9] = s_n (final)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
fit_prep_kernel(const WarpArgs a) {
    long long n_total = a.n;
    if (a.n_dev) { const long long nd = *a.n_dev; n_total = nd < a.n ? nd : a.n; }
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const int ch = a.cand_hw[2 * i], cw = a.cand_hw[2 * i + 1];
    const size_t fbase = (size_t)a.cand_frame[i] * a.H * a.W + (size_t)(ch - 2) * a.W + (cw - 2);
    int v[25];
    long long isum = 0;
    int imax = -2147483647 - 1;
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int p = w_ld_int(a.frames, a.fdtype, fbase + (size_t)r * a.W + c);
            v[r * 5 + c] = p;
            isum += p; imax = max(imax, p);
        }
    // sum((sub - mean(sub))**2) as Python's sum() forms it: exact mean, unfused squares added in raster order (pflib.py:464)
    const double dmean = __ddiv_rn((double)isum, 25.0);
    double sst = 0.0;
#pragma unroll
    for (int q = 0; q < 25; ++q) {
        const double e = __dsub_rn((double)v[q], dmean);
        sst = __dadd_rn(sst, __dmul_rn(e, e));
    }
    const double s_n = illumina_sn<5>(5, [&](int r, int c) { return (long long)v[r * 5 + c]; });   // bit-identical to pflib.illumina_s_n
    // (the selection network permutes its input: the record keeps the window, the network gets a copy)
    PrepRec rec;
#pragma unroll
    for (int q = 0; q < 25; ++q) rec.px[q] = v[q];
    const int imed = median25<int>(v);                           // numpy.median of 25 (pflib.py:199)
    rec.med = imed; rec.max = imax;
    rec.hw = ((unsigned)ch << 16) | (unsigned)cw;
    rec.sst = sst;
    rec.lo1 = ((double)imax - (double)(int)isum / 25.0) / 3.0;                                // pflib.py:205
    uint4* dst = reinterpret_cast<uint4*>(a.prep + i);
    const uint4* src = reinterpret_cast<const uint4*>(&rec);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = src[q];
    a.out_fit[i * 12 + 9] = s_n;                                          // illumina_s_n, pflib.py:261-281 (final)
}

// double -> float, round to nearest.  The file is compiled with -ftz=true, under which a plain (float) cast becomes
// cvt.rn.ftz.f32.f64 -- emulated in SASS by an FP64 compare and a multiply around every conversion (56 of them in the tick,
// ten per pixel row of the pass).  The non-flushing form is one F2F; a denormal result is flushed by the FP32 operation that
// consumes it.
__device__ __forceinline__ float d2f(double v) {
    float r;
    asm("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(v));
    return r;
}

// -------------------------------------------------------------------------------------------
// Branch-free FP64 helpers of the pass.  exp: Cody-Waite reduction by ln2 + degree-12 Taylor
// polynomial (|r| <= ln2/2: 2 ulp, measured against numpy.exp), argument <= 0 and bounded by the
// pflib limits (widths >= 0.75, centres in [2,3] => |arg| < 64), so no range checks.
// -------------------------------------------------------------------------------------------
template <bool CLAMP>
__device__ __forceinline__ double w_exp_neg(double u) {
    if (CLAMP) u = fmax(u, -700.0);              // generic windows: widths / offsets are not bounded
    const double kf = fma(u, 1.4426950408889634, 6755399441055744.0);        // round(u / ln2) in the low word
    const int k = __double2loint(kf);
    const double kd = kf - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, u);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 2.08767569878680989792e-09;                                   // 1/12!
    p = fma(p, r, 2.50521083854417187751e-08);
    p = fma(p, r, 2.75573192239858906526e-07);
    p = fma(p, r, 2.75573192239858906526e-06);
    p = fma(p, r, 2.48015873015873015873e-05);
    p = fma(p, r, 1.98412698412698412698e-04);
    p = fma(p, r, 1.38888888888888888889e-03);
    p = fma(p, r, 8.33333333333333333333e-03);
    p = fma(p, r, 4.16666666666666666667e-02);
    p = fma(p, r, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// sin / cos of an angle given in DEGREES (gaussfitter.py:115 converts with pi/180), exact at
// multiples of 90: reduce by quarter turns in degrees (exact), then odd / even polynomials on
// |r| <= pi/4.
__device__ __forceinline__ void w_sincos_deg(double th, double* sn, double* cs) {
    const double qf = rint(th * (1.0 / 90.0));
    const int q = (int)qf;
    const double r = fma(qf, -90.0, th) * WQ_DEG2RAD;
    const double r2 = r * r;
    double ps = 1.58969099521155010221e-10;          // 1/13!
    ps = fma(ps, r2, -2.50521083854417187751e-08);
    ps = fma(ps, r2, 2.75573192239858906526e-06);
    ps = fma(ps, r2, -1.98412698412698412698e-04);
    ps = fma(ps, r2, 8.33333333333333333333e-03);
    ps = fma(ps, r2, -1.66666666666666666667e-01);
    const double s = fma(ps * r2, r, r);
    double pc = -1.14707455977297247139e-11;         // -1/14!
    pc = fma(pc, r2, 2.08767569878680989792e-09);
    pc = fma(pc, r2, -2.75573192239858906526e-07);
    pc = fma(pc, r2, 2.48015873015873015873e-05);
    pc = fma(pc, r2, -1.38888888888888888889e-03);
    pc = fma(pc, r2, 4.16666666666666666667e-02);
    pc = fma(pc, r2, -0.5);
    const double c = fma(pc, r2, 1.0);
    const bool swap = q & 1;
    const double ss = swap ? c : s, cc = swap ? s : c;
    *sn = (q & 2) ? -ss : ss;
    *cs = ((q + 1) & 2) ? -cc : cc;
}

#ifndef WPASS_NO_FFMA2
#define WPASS_FFMA2 1      // measured on B200: bit-identical fits, 4.555 -> 4.410 ms per 200-frame launch with three in flight
#endif
// Packed FP32 accumulation of the normal equations (FFMA2, new on sm_100: two FP32 multiply-adds per issue slot, one
// operand may be a broadcast scalar).  With e = (j2, j3, j4, j5, j1, j6, ff, -1) -- J's columns 1..6 of one pixel in the
// order their packed evaluation produces them, the residual, and the constant column 0 -- every sum the pass needs is an
// entry of e e^T: j_k j_l, j_k ff (J^T f), -j_k (J^T J's column 0).  Row a of e e^T is accumulated against the pairs
// (e1,e2) (e3,e4) (e5,e6) (e7,e8) from the pair that holds a onwards: 18 packed multiply-adds per pixel instead of 27
// multiply-adds and 6 additions; each packed lane is the same round-to-nearest FMA as before, so the sums are bit-identical.
struct OuterAcc {
    float2 R[18];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 18; ++i) R[i] = make_float2(0.0f, 0.0f);
    }
    __device__ __forceinline__ void add(const float2 P1, const float2 P2, const float2 P3, const float ff) {
        const float2 P4 = make_float2(ff, -1.0f);
        const float2 B1 = make_float2(P1.x, P1.x), B2 = make_float2(P1.y, P1.y), B3 = make_float2(P2.x, P2.x);
        const float2 B4 = make_float2(P2.y, P2.y), B5 = make_float2(P3.x, P3.x), B6 = make_float2(P3.y, P3.y);
        R[0] = __ffma2_rn(B1, P1, R[0]); R[1] = __ffma2_rn(B1, P2, R[1]); R[2] = __ffma2_rn(B1, P3, R[2]); R[3] = __ffma2_rn(B1, P4, R[3]);
        R[4] = __ffma2_rn(B2, P1, R[4]); R[5] = __ffma2_rn(B2, P2, R[5]); R[6] = __ffma2_rn(B2, P3, R[6]); R[7] = __ffma2_rn(B2, P4, R[7]);
        R[8] = __ffma2_rn(B3, P2, R[8]); R[9] = __ffma2_rn(B3, P3, R[9]); R[10] = __ffma2_rn(B3, P4, R[10]);
        R[11] = __ffma2_rn(B4, P2, R[11]); R[12] = __ffma2_rn(B4, P3, R[12]); R[13] = __ffma2_rn(B4, P4, R[13]);
        R[14] = __ffma2_rn(B5, P3, R[14]); R[15] = __ffma2_rn(B5, P4, R[15]);
        R[16] = __ffma2_rn(B6, P3, R[16]); R[17] = __ffma2_rn(B6, P4, R[17]);
    }
    // -> packed lower triangle A (column 0 = -sum j_k; A[0] is set by the caller) and g[1..6] (g[0] is summed by the caller)
    __device__ __forceinline__ void unpack(float (&A)[WNT], float (&g)[WNP]) const {
        constexpr int M[7] = {0, 5, 1, 2, 3, 4, 6};               // position of column k in e
        constexpr int rowbase[7] = {0, 0, 4, 8, 11, 14, 16};      // first accumulator of row a
        constexpr int rowp0[7] = {0, 1, 1, 2, 2, 3, 3};           // first pair of row a
#pragma unroll
        for (int k = 1; k < WNP; ++k) {
            const float2 last = R[rowbase[M[k]] + 4 - rowp0[M[k]]];
            g[k] = last.x;
            A[wtri(k, 0)] = last.y;
#pragma unroll
            for (int l = 1; l <= k; ++l) {
                const int lo = M[k] < M[l] ? M[k] : M[l], hi = M[k] < M[l] ? M[l] : M[k];
                const float2 v = R[rowbase[lo] + (hi + 1) / 2 - rowp0[lo]];
                A[wtri(k, l)] = (hi & 1) ? v.x : v.y;
            }
        }
    }
};

#if defined(WPASS_FFMA2) && !defined(WPASS_NO_BASEVEC)
#define WPASS_BASEVEC 1     // measured on B200: 3.37 -> 3.26 ms per 200-frame launch with three in flight, every parity figure
                            // unchanged to four digits; 11x11: 6.4 -> 5.1 ms on 200 000 isolated windows
#endif
// WPASS_BASEVEC: the pass accumulates the outer products of the BASE vector v = (E, AE a, AE b, AE a^2, AE b^2, AE a b)
// instead of the Jacobian row j = T v (j1 = -v1, (j2, j3) = R (v2, v3) with R = [[-sx, cyw], [cxw, sy]], j4 = -v4 / wx,
// j5 = -v5 / wy, j6 = -krot v6): T is the same for every pixel of a pass, so J^T J = T (sum v v^T) T^T, J^T f = T (sum v f)
// and column 0 follow from the accumulated sums by ~75 multiply-adds per pass, and the eight multiplications per pixel that
// formed j leave the pixel loop.  S / sg come in as sum v_k v_l, -sum v_k (column 0), sum v_k f and leave as J^T J, J^T f.
__device__ __forceinline__ void w_basevec_transform(float (&A)[WNT], float (&g)[WNP], const float sx, const float cxw, const float sy,
                                                    const float cyw, const float iwxf, const float iwyf, const float krot) {
    const float d4 = -iwxf, d5 = -iwyf, d6 = -krot;
    // rows 2, 3 of T applied to a pair (p2, p3): (-sx p2 + cyw p3, cxw p2 + sy p3)
#define W_R2(p2, p3) fmaf(cyw, (p3), -sx * (p2))
#define W_R3(p2, p3) fmaf(sy, (p3), cxw * (p2))
    // column 0 and J^T f: the same linear map
    {
        const float s10 = A[wtri(1, 0)], s20 = A[wtri(2, 0)], s30 = A[wtri(3, 0)];
        A[wtri(1, 0)] = -s10; A[wtri(2, 0)] = W_R2(s20, s30); A[wtri(3, 0)] = W_R3(s20, s30);
        A[wtri(4, 0)] *= d4; A[wtri(5, 0)] *= d5; A[wtri(6, 0)] *= d6;
        const float g1 = g[1], g2 = g[2], g3 = g[3];
        g[1] = -g1; g[2] = W_R2(g2, g3); g[3] = W_R3(g2, g3);
        g[4] *= d4; g[5] *= d5; g[6] *= d6;
    }
    const float s21 = A[wtri(2, 1)], s31 = A[wtri(3, 1)], s22 = A[wtri(2, 2)], s32 = A[wtri(3, 2)], s33 = A[wtri(3, 3)];
    // row / column 1 (d1 = -1)
    A[wtri(2, 1)] = -W_R2(s21, s31); A[wtri(3, 1)] = -W_R3(s21, s31);
    A[wtri(4, 1)] *= iwxf; A[wtri(5, 1)] *= iwyf; A[wtri(6, 1)] *= krot;          // d_k d_1 = +1/wx, +1/wy, +krot
    // the (2, 3) block: R S23 R^T
    {
        const float m22 = W_R2(s22, s32), m23 = W_R2(s32, s33);     // row 2 of R S23
        const float m32 = W_R3(s22, s32), m33 = W_R3(s32, s33);     // row 3 of R S23
        A[wtri(2, 2)] = W_R2(m22, m23);
        A[wtri(3, 2)] = W_R2(m32, m33);
        A[wtri(3, 3)] = W_R3(m32, m33);
    }
    // rows 4, 5, 6 against columns 2, 3 and among themselves
    {
        const float s42 = A[wtri(4, 2)], s43 = A[wtri(4, 3)], s52 = A[wtri(5, 2)], s53 = A[wtri(5, 3)], s62 = A[wtri(6, 2)], s63 = A[wtri(6, 3)];
        A[wtri(4, 2)] = d4 * W_R2(s42, s43); A[wtri(4, 3)] = d4 * W_R3(s42, s43);
        A[wtri(5, 2)] = d5 * W_R2(s52, s53); A[wtri(5, 3)] = d5 * W_R3(s52, s53);
        A[wtri(6, 2)] = d6 * W_R2(s62, s63); A[wtri(6, 3)] = d6 * W_R3(s62, s63);
        A[wtri(4, 4)] *= d4 * d4; A[wtri(5, 4)] *= d5 * d4; A[wtri(5, 5)] *= d5 * d5;
        A[wtri(6, 4)] *= d6 * d4; A[wtri(6, 5)] *= d6 * d5; A[wtri(6, 6)] *= d6 * d6;
    }
#undef W_R2
#undef W_R3
}

// One pass over the window at pt: chi^2 in FP64; J^T J (packed), J^T f in FP32 (J = d residual / dp).
//
// RECUR (the pflib frame path: 5x5 window, widths >= 0.75, centres in [2,3], so every exponent below is
// within +-64): the exponent u(r,c) = -(a^2 + b^2)/2 is a quadratic form in the pixel indices, so the 25
// values exp(u) follow from SIX exp evaluations and two multiplications per pixel by forward differencing:
//   E(r,c+1) = E(r,c) G(r,c),  G(r,c+1) = G(r,c) Kc;   E(r+1,0) = E(r,0) Gr(r),  Gr(r+1) = Gr(r) Kr,
//   G(r+1,0) = G(r,0) Kx,  with Kc, Kr, Kx = exp of the (constant) second differences.
// Measured against an 80-bit evaluation over the whole parameter box: relative error <= 9.3e-15 (the direct
// polynomial: 5.0e-15), four orders below what the ftol = 1e-10 test resolves.
template <int WIN, int TPB, bool CLAMP, bool RECUR, typename PXT>
__device__ __forceinline__ void w_pass(const double (&pt)[WNP], const PXT* __restrict__ sd,
                                       float (&A)[WNT], float (&g)[WNP], double& ss_out) {
    static_assert(!RECUR || (WIN <= 5 && !CLAMP), "forward differencing needs bounded exponents");
    const double Hh = pt[0], Aa = pt[1];
    double sn, cs;
    w_sincos_deg(pt[6], &sn, &cs);                                            // gaussfitter.py:115
    // A width sitting exactly on a lower limit of 0 (gaussfit's default limits, gaussfitter.py:143-146): the reference
    // divides the finite rotated offset by it, gets +-inf, and exp(-inf) = 0 leaves the flat model H.  Pre-multiplied
    // reciprocals would give inf - inf = NaN here, so the generic entry maps that case to E = 0 explicitly.
    const bool flat = CLAMP && ((pt[4] == 0.0) || (pt[5] == 0.0));
    const double ezero = flat ? 0.0 : 1.0;
    const double iwx = (CLAMP && pt[4] == 0.0) ? 0.0 : 1.0 / pt[4], iwy = (CLAMP && pt[5] == 0.0) ? 0.0 : 1.0 / pt[5];
    const double cxs = cs * iwx, sxs = sn * iwx, cys = cs * iwy, sys = sn * iwy;
    // per-column terms of the rotated offsets (numpy.indices: y = column index pairs with p[2]);
    // 5x5: tables in registers; larger windows: one multiply-add per pixel instead
#ifdef WPASS_ROLLED
    constexpr bool TABLES = false;
#else
    constexpr bool TABLES = (WIN <= 5);
#endif
    double ca[(TABLES && !RECUR) ? WIN : 1], cb[(TABLES && !RECUR) ? WIN : 1];
    float caf[TABLES ? WIN : 1], cbf[TABLES ? WIN : 1];
    if (TABLES) {
#pragma unroll
        for (int c = 0; c < (TABLES ? WIN : 1); ++c) {
            const double dy = pt[2] - (double)c;
            const double cav = dy * sxs, cbv = dy * cys;
            if (!RECUR) { ca[c] = cav; cb[c] = cbv; }
            caf[c] = d2f(cav); cbf[c] = d2f(cbv);
        }
    }
    double Er = 0.0, Grow = 0.0, Grr = 0.0, Kc = 0.0, Kr = 0.0, Kx = 0.0;
    if (RECUR) {
        const double a00 = fma(pt[3], cxs, -pt[2] * sxs), b00 = fma(pt[3], sys, pt[2] * cys);   // a, b at pixel (0, 0)
        // steps: column +1 -> (a, b) += (sxs, -cys); row +1 -> (a, b) += (-cxs, -sys)
        const double qc = fma(sxs, sxs, cys * cys), qr = fma(cxs, cxs, sys * sys), qx = fma(cxs, sxs, -sys * cys);
        Er = w_exp_neg<false>(-0.5 * fma(b00, b00, a00 * a00));
        Grow = w_exp_neg<false>(fma(b00, cys, -a00 * sxs) - 0.5 * qc);
        Grr = w_exp_neg<false>(fma(a00, cxs, b00 * sys) - 0.5 * qr);
        Kc = w_exp_neg<false>(-qc);
        Kr = w_exp_neg<false>(-qr);
        Kx = w_exp_neg<false>(qx);
    }
    const float Af = d2f(Aa), sx = d2f(sxs), cxw = d2f(cxs), sy = d2f(sys), cyw = d2f(cys);
    const float iwxf = d2f(iwx), iwyf = d2f(iwy);
    const float krot = d2f((pt[5] * iwx - pt[4] * iwy) * WQ_DEG2RAD);
    const float cyf = d2f(pt[2]);
#pragma unroll
    for (int i = 0; i < WNT; ++i) A[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < WNP; ++i) g[i] = 0.0f;
#ifdef WPASS_FFMA2
    OuterAcc R;
    R.clear();
#endif
    double ss = 0.0;
    double dx = pt[3];                                  // x = row index pairs with p[3]
#ifdef WPASS_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int r = 0; r < WIN; ++r) {
        const double ra = dx * cxs, rb = dx * sys;
        const float raf = d2f(ra), rbf = d2f(rb);
        const PXT* drow = sd + r * WIN * TPB;
        dx -= 1.0;
        double Ec = Er, Gc = Grow;
        if (RECUR) { Er *= Grr; Grr *= Kr; Grow *= Kx; }
#ifdef WPASS_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int c = 0; c < WIN; ++c) {
            double av, bv;
            float af, bf;
            if (RECUR) {
                av = 0.0; bv = 0.0;
                if (TABLES) { af = raf - caf[c]; bf = rbf + cbf[c]; }
                else { const float dyf = cyf - (float)c; af = fmaf(-dyf, sx, raf); bf = fmaf(dyf, cyw, rbf); }
            } else if (TABLES) {
                av = ra - ca[c]; bv = rb + cb[c];
                af = raf - caf[c]; bf = rbf + cbf[c];
            } else {
                const double dy = pt[2] - (double)c;
                const float dyf = cyf - (float)c;
                av = fma(-dy, sxs, ra); bv = fma(dy, cys, rb);
                af = fmaf(-dyf, sx, raf); bf = fmaf(dyf, cyw, rbf);
            }
            double E;
            if (RECUR) { E = Ec; Ec *= Gc; Gc *= Kc; }
            else E = CLAMP ? ezero * w_exp_neg<CLAMP>(-0.5 * fma(bv, bv, av * av)) : w_exp_neg<CLAMP>(-0.5 * fma(bv, bv, av * av));
            const double f = (double)drow[c * TPB] - fma(Aa, E, Hh);
            ss = fma(f, f, ss);
            const float Ef = d2f(E), ff = d2f(f);
            const float AE = Af * Ef;
            const float AEa = AE * af, AEb = AE * bf;
#if defined(WPASS_FFMA2) && defined(WPASS_BASEVEC)
            g[0] -= ff;
            R.add(make_float2(AEa, AEb), make_float2(AEa * af, AEb * bf), make_float2(Ef, AEa * bf), ff);
#else
            float j[WNP];
            j[1] = -Ef;
            j[2] = AEb * cyw - AEa * sx;                // d/d p[2] (centre along axis 1)
            j[3] = AEa * cxw + AEb * sy;                // d/d p[3] (centre along axis 0)
            j[4] = -AEa * af * iwxf;
            j[5] = -AEb * bf * iwyf;
            j[6] = -AEa * bf * krot;                    // degrees
            // column 0 of J is the constant -1
            g[0] -= ff;
#endif
#if defined(WPASS_FFMA2) && defined(WPASS_BASEVEC)
#elif defined(WPASS_FFMA2)
            R.add(make_float2(j[2], j[3]), make_float2(j[4], j[5]), make_float2(j[1], j[6]), ff);
#else
#pragma unroll
            for (int k = 1; k < WNP; ++k) {
                g[k] = fmaf(j[k], ff, g[k]);
                A[wtri(k, 0)] -= j[k];
#pragma unroll
                for (int l = 1; l <= k; ++l) A[wtri(k, l)] = fmaf(j[k], j[l], A[wtri(k, l)]);
            }
#endif
        }
    }
#ifdef WPASS_FFMA2
    R.unpack(A, g);
#ifdef WPASS_BASEVEC
    w_basevec_transform(A, g, sx, cxw, sy, cyw, iwxf, iwyf, krot);
#endif
#endif
    A[0] = (float)(WIN * WIN);
    ss_out = ss;
}

// -------------------------------------------------------------------------------------------
// Lane-group pass of the 11x11 kernel (GRP lanes per window, GRP = 2, 4 or 8): 121 pixels are too many for one thread
// when latency matters -- a thread-per-window tick is ~9 700 dependent-ish instructions, and a window that runs to
// maxiter = 200 (1-3 % of real 11x11 windows do, in the reference as well) keeps its warp alive for 200 of them -- and too
// few for a warp.  Here GRP adjacent lanes own ONE window for its whole life: lane gl takes pixels gl, gl + GRP, ... of
// the pass, the 27 + 7 FP32 sums and the FP64 chi^2 are reduced over the group by an xor butterfly (bitwise identical
// sums in every lane of the group), and everything else of the tick -- bookkeeping, Cholesky, lmpar, bounds -- is executed
// redundantly by all lanes of the group on identical numbers, so the group never diverges and needs no broadcast.
// Pixels sit in shared memory as [window][128 + GRP] (the lanes of a warp read conflict-free), in FP32 when the
// window data are narrow integers (exact), in FP64 otherwise.
// -------------------------------------------------------------------------------------------
template <int WIN, int G, bool CLAMP, typename PXT>
__device__ __forceinline__ void g_pass(const double (&pt)[WNP], const PXT* __restrict__ px, const int gl,
                                       float (&A)[WNT], float (&g)[WNP], double& ss_out) {
    constexpr int P = WIN * WIN;
    constexpr int SL = (P + G - 1) / G;             // pixel slots per lane
    static_assert(G < WIN, "one row wrap per slot");
    const double Hh = pt[0], Aa = pt[1];
    double sn, cs;
    w_sincos_deg(pt[6], &sn, &cs);
    const bool flat = CLAMP && ((pt[4] == 0.0) || (pt[5] == 0.0));     // see w_pass
    const double ezero = flat ? 0.0 : 1.0;
    const double iwx = (CLAMP && pt[4] == 0.0) ? 0.0 : 1.0 / pt[4], iwy = (CLAMP && pt[5] == 0.0) ? 0.0 : 1.0 / pt[5];
    const double cxs = cs * iwx, sxs = sn * iwx, cys = cs * iwy, sys = sn * iwy;
    // a(r, c) = (p3 - r) cxs - (p2 - c) sxs,  b(r, c) = (p3 - r) sys + (p2 - c) cys
    const double a00 = fma(pt[3], cxs, -pt[2] * sxs), b00 = fma(pt[3], sys, pt[2] * cys);
    const float Af = d2f(Aa), sx = d2f(sxs), cxw = d2f(cxs), sy = d2f(sys), cyw = d2f(cys);
    const float iwxf = d2f(iwx), iwyf = d2f(iwy);
    const float krot = d2f((pt[5] * iwx - pt[4] * iwy) * WQ_DEG2RAD);
#pragma unroll
    for (int i = 0; i < WNT; ++i) A[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < WNP; ++i) g[i] = 0.0f;
    double ss = 0.0;
#ifdef WPASS_FFMA2
    OuterAcc R;
    R.clear();
#endif
    int r = 0, c = gl;                                               // pixel gl + G s = (r, c), advanced by G per slot
#ifndef W11_UNROLL
#define W11_UNROLL 2
#endif
    constexpr int UNR = W11_UNROLL;
#pragma unroll UNR
    for (int sl = 0; sl < SL; ++sl) {
        const bool ok = (sl * G + gl) < P;
        const double rd_ = (double)r, cd = (double)c;
        const double av = fma(cd, sxs, fma(-rd_, cxs, a00));
        const double bv = fma(-cd, cys, fma(-rd_, sys, b00));
        const double E0 = w_exp_neg<CLAMP>(-0.5 * fma(bv, bv, av * av));
        const double E = ok ? (CLAMP ? ezero * E0 : E0) : 0.0;
        const double d = ok ? (double)px[sl * G] : Hh;              // beyond the window: zero residual, zero Jacobian
        const double f = d - fma(Aa, E, Hh);
        ss = fma(f, f, ss);
        const float af = d2f(av), bf = d2f(bv);
        const float Ef = d2f(E), ff = d2f(f);
        const float AE = Af * Ef;
        const float AEa = AE * af, AEb = AE * bf;
#if defined(WPASS_FFMA2) && defined(WPASS_BASEVEC)
        g[0] -= ff;
        R.add(make_float2(AEa, AEb), make_float2(AEa * af, AEb * bf), make_float2(Ef, AEa * bf), ff);
#else
        float j[WNP];
        j[1] = -Ef;
        j[2] = AEb * cyw - AEa * sx;
        j[3] = AEa * cxw + AEb * sy;
        j[4] = -AEa * af * iwxf;
        j[5] = -AEb * bf * iwyf;
        j[6] = -AEa * bf * krot;
        g[0] -= ff;
#endif
#if defined(WPASS_FFMA2) && defined(WPASS_BASEVEC)
#elif defined(WPASS_FFMA2)
        R.add(make_float2(j[2], j[3]), make_float2(j[4], j[5]), make_float2(j[1], j[6]), ff);
#else
#pragma unroll
        for (int k = 1; k < WNP; ++k) {
            g[k] = fmaf(j[k], ff, g[k]);
            A[wtri(k, 0)] -= j[k];
#pragma unroll
            for (int l = 1; l <= k; ++l) A[wtri(k, l)] = fmaf(j[k], j[l], A[wtri(k, l)]);
        }
#endif
        c += G;
        if (c >= WIN) { c -= WIN; ++r; }
    }
#ifdef WPASS_FFMA2
    R.unpack(A, g);
#ifdef WPASS_BASEVEC
    w_basevec_transform(A, g, sx, cxw, sy, cyw, iwxf, iwyf, krot);      // (linear: applied to this lane's partial sums)
#endif
#endif
    ss_out = ss;
}

enum { MODE_FIRST = 0, MODE_TRIAL = 1, MODE_RESUME = 2 };
#ifndef WRECUR
#define WRECUR 1           // forward-differenced exponentials on the pflib frame path (w_pass)
#endif
#define WQ_TINYF 1.0e-37f
#ifndef WLMPAR_MAX
#define WLMPAR_MAX 10      // lmpar iteration limit (mpfit.py:2148)
#endif

// GRP = lanes per window (1: one thread per window; 2 / 4 / 8: lane groups, generic-window entry only, see g_pass);
// PXT = type of the pixels in shared memory (double; float when the window data are narrow integers -- exact)
template <int WIN, int GRP>
struct WLayout {
    static constexpr int P = WIN * WIN;
    static constexpr int PSTR = ((P + 31) / 32) * 32 + GRP;       // elements per window row of the lane-group pixel block
    template <int TPB, typename PXT>
    __host__ __device__ static constexpr size_t pixel_bytes() {
        return GRP == 1 ? (size_t)P * TPB * sizeof(PXT) : (((size_t)(TPB / GRP) * PSTR * sizeof(PXT) + 15) / 16) * 16;
    }
    template <int TPB, typename PXT>
    __host__ __device__ static constexpr size_t smem_bytes() {
        return pixel_bytes<TPB, PXT>() + (size_t)(WNT + WNP) * TPB * sizeof(float) + (GRP > 1 ? (size_t)(TPB / GRP) * 14 * sizeof(double) : 0);
    }
};

template <int WIN, int TPB, int MINB, bool PFLIB, int GRP = 1, typename PXT = double>
__global__ void __launch_bounds__(TPB, MINB)
lmwarp_kernel(const WarpArgs a) {
    constexpr int P = WIN * WIN;
    static_assert(GRP == 1 || !PFLIB, "lane groups serve the generic-window entry");
    static_assert(GRP == 1 || GRP == 2 || GRP == 4 || GRP == 8, "lanes per window");
    using LY = WLayout<WIN, GRP>;
    constexpr int PSTR = LY::PSTR;
    extern __shared__ __align__(16) unsigned char w_smem[];
    // GRP == 1: [P][TPB] pixels | [28][TPB] | [7][TPB];   GRP > 1: [TPB / GRP][PSTR] pixels | [28][TPB] | [7][TPB] | [TPB / GRP][14] limits
    PXT* const s_d = reinterpret_cast<PXT*>(w_smem);
    float* const s_A = reinterpret_cast<float*>(w_smem + LY::template pixel_bytes<TPB, PXT>());   // column-scaled J^T J at the current point
    float* const s_g = s_A + WNT * TPB;                                    // [7][TPB]  column-scaled J^T f
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    const int gl = (int)(lane & (unsigned)(GRP - 1));                      // lane within its group
    const unsigned gbit = 1u << (lane & ~(unsigned)(GRP - 1));             // ballot bit of the group's first lane
    constexpr unsigned LEADERS = GRP == 1 ? 0xffffffffu : GRP == 2 ? 0x55555555u : GRP == 4 ? 0x11111111u : 0x01010101u;
    const PXT* const sd = GRP == 1 ? s_d + tid : s_d + (tid / GRP) * PSTR + gl;
    float* const sA = s_A + tid;
    float* const sg = s_g + tid;
    // lane groups: the window's box limits (lower[7] | upper[7]) -- the serial phase reads them twice per tick, and from
    // global memory that is fourteen dependent round trips in the chain that sets the length of a long fit's tick
    double* const s_lim = reinterpret_cast<double*>(s_g + WNP * TPB) + (GRP > 1 ? (tid / GRP) * 14 : 0);

    const float ftol = a.ftol_f, xtol = a.xtol_f, gtol = a.gtol_f, factor = a.factor_f;
    const float machep = (float)WQ_MACHEP;
    const int maxiter = a.o.maxiter;
    long long n_total = a.n;
    if (a.n_dev) { const long long nd = *a.n_dev; n_total = nd < a.n ? nd : a.n; }
    if (a.resume) {
        const long long ns = (long long)*a.strag_count;
        if (ns < a.resume_lo || ns > a.resume_hi) return;          // another arrangement finishes this batch's parked fits
        n_total = ns < n_total ? ns : n_total;
    }

    // ---- per-lane fit state
    bool active = false, exhausted = false;
    long long idx = 0;
    int cand_h = 0, cand_w = 0;
    int mode = MODE_FIRST, status = 0, niter = 1, nfev = 0, n_damped = 0;
    unsigned lpeg = 0, upeg = 0;
    bool nonfinite = false;
    double x[WNP], y[WNP];
    double ss0 = -1.0, ss1 = -1.0;                      // chi^2 at x, chi^2 of the last trial point
    Lim<PFLIB> lim;
    lim.lo1 = 0.0; lim.lo = nullptr; lim.hi = nullptr; lim.qll = PF_QLL; lim.qul = PF_QUL;
    float delta = 0.0f, par = 0.0f, xnorm = 0.0f, gnorm = 0.0f, pnorm = 0.0f, prered = 0.0f, dirder = 0.0f, rss0 = 0.0f;
    float diag[WNP], iS[WNP];
#pragma unroll
    for (int j = 0; j < WNP; ++j) { x[j] = 0.0; y[j] = 1.0; diag[j] = 1.0f; iS[j] = 1.0f; }

    // (Claiming one slot ahead per lane and prefetching its start record was measured on B200: it hides the two
    //  round trips of a refill but a lane inside a 200-iteration fit then sits on an unstarted candidate, and the
    //  40-frame launch got 10 % SLOWER (1.65 -> 1.82 ms).  A warp-level pool with guided claim sizes (up to 32
    //  candidates per atomic, records prefetched, claims shrinking to 1 towards the end of the queue) avoids that
    //  imbalance and measured no gain either (1.86 vs 1.78 ms; 160 frames and the pipelined step unchanged): the
    //  refill's round trips are hidden by the other resident warps.  The queue stays strictly on demand.)
    int drained_ticks = 0;                              // warp-uniform: ticks since this warp first found the queue empty
    for (;;) {
        // ------------------------------------------------------------------ refill idle lanes
        __syncwarp();
        if (a.drain_grace > 0 && __any_sync(0xffffffffu, exhausted)) ++drained_ticks;
        // (lane groups: the lanes of a group hold identical state, so the ballot has whole groups; one slot per group)
        const unsigned want = __ballot_sync(0xffffffffu, !active && !exhausted) & LEADERS;
        bool fresh = false;                                 // this lane was handed a new window in this tick
        if (want) {
            const int leader = __ffs(want) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(a.work_counter, (unsigned long long)__popc(want));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want & gbit) {
                const long long slot = (long long)(base + (unsigned long long)__popc(want & (gbit - 1u)));
                if (slot >= n_total) exhausted = true;
                else {
                    idx = a.resume ? a.strag[slot].idx : slot;
                    if (PFLIB) {
                        uint4 q[8];                                   // one 128-byte start record
                        {
                            const uint4* src = reinterpret_cast<const uint4*>(a.prep + idx);
#pragma unroll
                            for (int k = 0; k < 8; ++k) q[k] = src[k];
                        }
#define WREC(i) ((i) % 4 == 0 ? q[(i) / 4].x : (i) % 4 == 1 ? q[(i) / 4].y : (i) % 4 == 2 ? q[(i) / 4].z : q[(i) / 4].w)
#pragma unroll
                        for (int k = 0; k < P; ++k) s_d[k * TPB + tid] = (PXT)(int)WREC(k);
                        cand_h = (int)(WREC(27) >> 16); cand_w = (int)(WREC(27) & 0xffffu);
                        lim.lo1 = __hiloint2double((int)WREC(31), (int)WREC(30));          // pflib.py:205 (fit_prep_kernel)
                        x[0] = (double)(int)WREC(25); x[1] = (double)(int)WREC(26); x[2] = 2.5; x[3] = 2.5; x[4] = 1.0; x[5] = 1.0; x[6] = 0.0;
#undef WREC
#pragma unroll
                        for (int j = 0; j < WNP; ++j) {                                    // gaussfitter.py:202-204
                            if (lim.has_hi(j) && x[j] > lim.upper(j)) x[j] = lim.upper(j);
                            if (x[j] < lim.lower(j)) x[j] = lim.lower(j);
                            y[j] = x[j];
                        }
                    } else {
                        fresh = true;                                  // pixels: loaded by the whole warp below
                        if (GRP > 1) {
                            for (int k = gl; k < 14; k += GRP) s_lim[k] = (k < WNP) ? a.lo[idx * WNP + k] : a.hi[idx * WNP + k - WNP];
                            lim.lo = s_lim; lim.hi = s_lim + WNP;       // (visible to the group after the __syncwarp below)
                        } else { lim.lo = a.lo + idx * WNP; lim.hi = a.hi + idx * WNP; }
                        lim.qll = 0; lim.qul = 0;
#pragma unroll
                        for (int j = 0; j < WNP; ++j) {
                            x[j] = a.p0[idx * WNP + j]; y[j] = x[j];
                            lim.qll |= (a.lim_lo[idx * WNP + j] ? 1u : 0u) << j;
                            lim.qul |= (a.lim_hi[idx * WNP + j] ? 1u : 0u) << j;
                        }
                    }
                    active = true; mode = MODE_FIRST; status = 0; niter = 1; nfev = 0; n_damped = 0;
                    ss0 = -1.0; ss1 = -1.0; par = 0.0f; nonfinite = false;
                    if (!PFLIB) {
                        // mpfit.py:956-964: start outside the limits / inconsistent limits -> status 0, nothing runs
                        // (the limits come from global memory here: a lane group's shared copy is being written)
                        bool bad = false;
                        const double* glo = a.lo + idx * WNP;
                        const double* ghi = a.hi + idx * WNP;
#pragma unroll
                        for (int j = 0; j < WNP; ++j) {
                            const bool ql = lim.has_lo(j), qu = lim.has_hi(j);
                            bad |= (ql && x[j] < glo[j]) || (qu && x[j] > ghi[j]) || (ql && qu && glo[j] >= ghi[j]);
                        }
                        if (bad) {
                            if (gl == 0) {
#pragma unroll
                                for (int j = 0; j < WNP; ++j) a.params[idx * WNP + j] = x[j];
                                a.status[idx] = 0; a.niter[idx] = 0; a.nfev[idx] = 0; a.chi2[idx] = -1.0;
                                if (a.n_damped) a.n_damped[idx] = 0;
                            }
                            active = false;
                        }
                    }
                    if (a.resume) {
                        const StragRec* rec = a.strag + slot;
#pragma unroll
                        for (int j = 0; j < WNP; ++j) { x[j] = rec->x[j]; y[j] = x[j]; diag[j] = rec->diag[j]; }
                        ss0 = rec->ss0; delta = rec->delta; par = rec->par; xnorm = rec->xnorm;
                        niter = rec->niter; nfev = rec->nfev; n_damped = rec->n_damped;
                        mode = MODE_RESUME;
                    }
                }
            }
        }
        if (!PFLIB && want) {
            // The pixels of every newly claimed window are fetched by the WHOLE warp, coalesced (lane k takes elements
            // k, k + 32, ...): a lane copying its own 121 pixels one after the other kept the other 31 lanes waiting for
            // 121 dependent round trips per refill -- with ~2 refills per warp and tick that was most of the kernel's time
            unsigned needm = __ballot_sync(0xffffffffu, fresh) & LEADERS;
            while (needm) {
                const int w = __ffs(needm) - 1;
                needm &= needm - 1u;
                const long long iw = __shfl_sync(0xffffffffu, idx, w);
                double v[(P + 31) / 32];
#pragma unroll
                for (int k = 0; k < (P + 31) / 32; ++k) {
                    const int q = k * 32 + (int)lane;
                    v[k] = (q < P) ? w_ld_dbl(a.windows, a.wdtype, (size_t)iw * P + q) : 0.0;
                }
#pragma unroll
                for (int k = 0; k < (P + 31) / 32; ++k) {
                    const int q = k * 32 + (int)lane;
                    if (q < P) {
                        if (GRP == 1) s_d[q * TPB + (tid & ~31) + w] = (PXT)v[k];
                        else s_d[(((tid & ~31) + w) / GRP) * PSTR + q] = (PXT)v[k];
                    }
                }
            }
            __syncwarp();
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) {
            if (__ballot_sync(0xffffffffu, !exhausted) == 0u) break;      // queue drained and nothing running
            continue;                                                    // (a refill can end at once: status 0)
        }

        // -------------------------------------------------------------- pass at the trial point
        float An[WNT], gn[WNP];
        double ss = 0.0;
        if (GRP > 1) {
            // every lane sums its share of the window's pixels; the butterfly leaves identical totals in the whole group
            // (idle groups shuffle zeros along: the exchange is warp-wide)
#pragma unroll
            for (int i = 0; i < WNT; ++i) An[i] = 0.0f;
#pragma unroll
            for (int i = 0; i < WNP; ++i) gn[i] = 0.0f;
            if (active) g_pass<WIN, GRP, !PFLIB, PXT>(y, sd, gl, An, gn, ss);
#pragma unroll
            for (int m = GRP / 2; m >= 1; m >>= 1) {
                ss += __shfl_xor_sync(0xffffffffu, ss, m);
#pragma unroll
                for (int i = 1; i < WNT; ++i) An[i] += __shfl_xor_sync(0xffffffffu, An[i], m);
#pragma unroll
                for (int i = 0; i < WNP; ++i) gn[i] += __shfl_xor_sync(0xffffffffu, gn[i], m);
            }
            An[0] = (float)P;
        }
        if (active) {
            if (GRP == 1) w_pass<WIN, TPB, !PFLIB, PFLIB && WIN == 5 && WRECUR, PXT>(y, sd, An, gn, ss);      // y == x on the first tick of a fit
            if (mode != MODE_RESUME) ++nfev;

            bool have_new = false;
            if (mode == MODE_FIRST) {
                ss0 = ss;                                                                // mpfit.py:999, :1019
                have_new = true;
            } else if (mode == MODE_RESUME) {
                have_new = true;                            // same x, same arithmetic: ss == ss0 bit for bit
            } else {
                // ---------------------------------------------------------- trial bookkeeping (:1245-1335)
                ss1 = ss;
                float actred = -1.0f;
                if (0.01 * ss1 < ss0) actred = d2f(ss0 - ss1) * rss0;               // 1 - (fnorm1/fnorm)^2
                float ratio = 0.0f;
                if (prered != 0.0f) ratio = __fdividef(actred, prered);
                if (ratio <= 0.25f) {                                                    // :1276-1288
                    float temp;
                    if (actred >= 0.0f) temp = 0.5f;
                    else temp = __fdividef(0.5f * dirder, dirder + 0.5f * actred);
                    if ((0.01 * ss1 >= ss0) || (temp < 0.1f)) temp = 0.1f;
                    delta = temp * fminf(delta, pnorm * 10.0f);
                    par = __fdividef(par, temp);
                } else if ((par == 0.0f) || (ratio >= 0.75f)) {
                    delta = pnorm * 2.0f;
                    par = 0.5f * par;
                }
                const bool accepted = ratio >= 0.0001f;                                  // :1291-1298
                if (accepted) {
                    float s = 0.0f;
#pragma unroll
                    for (int j = 0; j < WNP; ++j) { x[j] = y[j]; const float t = diag[j] * d2f(x[j]); s = fmaf(t, t, s); }
                    xnorm = sqrtf(s);
                    ss0 = ss1;
                    ++niter;
                }
                const bool c1 = (fabsf(actred) <= ftol) && (prered <= ftol) && (0.5f * ratio <= 1.0f);   // :1301-1323
                if (c1) status = 1;
                if (delta <= xtol * xnorm) status = 2;
                if (c1 && status == 2) status = 3;
                if (status == 0) {
                    if (niter >= maxiter) status = 5;
                    if ((fabsf(actred) <= machep) && (prered <= machep) && (0.5f * ratio <= 1.0f)) status = 6;
                    if (delta <= machep * xnorm) status = 7;
                    if (gnorm <= machep) status = 8;
                }
                if (status == 0 && !accepted && (nonfinite || !isfinite(ratio))) status = -16;   // :1330-1335
                have_new = accepted;
                if (status == 0 && accepted &&
                    ((a.cap > 0 && nfev >= a.cap) || (a.drain_grace > 0 && drained_ticks > a.drain_grace))) {
                    // ------------------------------------------------------ park: a long fit leaves the lane
                    if (gl == 0) {
                        StragRec* rec = a.strag + atomicAdd(a.strag_count, 1ull);
                        rec->idx = idx;
#pragma unroll
                        for (int j = 0; j < WNP; ++j) { rec->x[j] = x[j]; rec->diag[j] = diag[j]; }
                        rec->ss0 = ss0; rec->delta = delta; rec->par = par; rec->xnorm = xnorm;
                        rec->niter = niter; rec->nfev = nfev; rec->n_damped = n_damped; rec->pad = 0;
                    }
                    active = false;
                }
            }

            if (!active) {
                // parked
            } else if (status == 0 && have_new) {
                // ---------------------------------------------------------- new linearisation at x
                rss0 = __fdividef(1.0f, d2f(ss0));
                // pegged parameters: zero the column when the gradient pushes outwards (:1073-1091)
                // (a zeroed column is not multiplied out of the 28 accumulators: its diagonal is read as 0 here and
                //  its scale factor as 0 in the store below, which leaves the same numbers in shared memory)
                lpeg = 0; upeg = 0;
                float acn[WNP], iSz[WNP];
                float gmax = 0.0f;
#pragma unroll
                for (int j = 0; j < WNP; ++j) {
                    // (bitwise & and |: no short-circuit branches in the serial chain -- every operand is cheap and side-effect free)
                    const bool lp = lim.has_lo(j) & (x[j] == lim.lower(j));
                    const bool up = lim.has_hi(j) & (x[j] == lim.upper(j));
                    lpeg |= (lp ? 1u : 0u) << j; upeg |= (up ? 1u : 0u) << j;
                    const bool zero = (lp & (gn[j] > 0.0f)) | (up & (gn[j] < 0.0f));
                    gn[j] = zero ? gn[j] * 0.0f : gn[j];
                    const float ajj = zero ? An[wtri(j, j)] * 0.0f : An[wtri(j, j)];
                    const float rs = ajj > 0.0f ? rsqrtf(ajj) : 0.0f;
                    acn[j] = ajj * rs;                                                   // column norms (:1758)
                    iS[j] = ajj > 0.0f ? rs : 1.0f;
                    iSz[j] = zero ? 0.0f : iS[j];
                    gn[j] *= iS[j];                                                      // scaled gradient
                    if (ajj > 0.0f) gmax = fmaxf(gmax, fabsf(gn[j]));                    // :1142-1148
                }
                gnorm = (ss0 != 0.0) ? gmax * sqrtf(rss0) : 0.0f;
                if (mode == MODE_FIRST) {                                                // :1099-1110
                    float s = 0.0f;
#pragma unroll
                    for (int j = 0; j < WNP; ++j) {
                        diag[j] = (acn[j] == 0.0f) ? 1.0f : acn[j];
                        const float t = diag[j] * d2f(x[j]);
                        s = fmaf(t, t, s);
                    }
                    xnorm = sqrtf(s);
                    delta = factor * xnorm;
                    if (delta == 0.0f) delta = factor;
                }
                if (gnorm <= gtol) status = 4;                                           // :1151
                else if (maxiter == 0) status = 5;
#pragma unroll
                for (int j = 0; j < WNP; ++j) diag[j] = fmaxf(diag[j], acn[j]);          // :1160
#pragma unroll
                for (int i = 0; i < WNP; ++i) {
#pragma unroll
                    for (int k = 0; k <= i; ++k) sA[wtri(i, k) * TPB] = An[wtri(i, k)] * (iSz[i] * iSz[k]);
                    sg[i * TPB] = gn[i];
                }
            }

            if (!active) {
                // parked
            } else if (status == 0) {
                // ---------------------------------------------------------- lmpar (:2077-2190), FP32
                Chol7 ch;
                float rhs[WNP], z[WNP], T[WNP], pf[WNP];
#pragma unroll
                for (int i = 0; i < WNP; ++i) { rhs[i] = -sg[i * TPB]; const float t = diag[i] * iS[i]; T[i] = t * t; }
                // (A "damped-first" search -- first factorisation at the par carried over from the previous iteration, the
                //  Gauss-Newton step only if the Newton iteration on par runs into zero -- was measured on B200: same
                //  parity figures, no gain (1.766 vs 1.775 ms per 40-frame launch, 2 % slower on 160 frames); More's
                //  order is kept.)
                float prr = 0.0f, par_used = 0.0f, fp = 0.0f, parl = 0.0f, paru = 0.0f, dxnorm = 0.0f;
                unsigned ok = 0;
#pragma unroll 1
                for (int it = 0; it <= WLMPAR_MAX; ++it) {
                    const unsigned okk = chol7_factor<TPB>(sA, T, prr, it == 0 ? 16.0f * 1.1920929e-07f : 0.0f, ch);
                    if (it == 0) ok = okk;
                    chol7_fwd(ch, rhs, z);
                    chol7_bwd(ch, z);
                    float dx2 = 0.0f;
#pragma unroll
                    for (int j = 0; j < WNP; ++j) { pf[j] = z[j] * iS[j]; const float t = diag[j] * pf[j]; dx2 = fmaf(t, t, dx2); }
                    dxnorm = sqrtf(dx2);
                    const float temp = fp;
                    fp = dxnorm - delta;
                    if (it == 0) {
                        if (fp <= 0.1f * delta) break;                // Gauss-Newton step inside the region (:2112)
                    } else {
                        ++n_damped;
                        par_used = prr;
                        if ((fabsf(fp) <= 0.1f * delta) || ((parl == 0.0f) && (fp <= temp) && (temp < 0.0f)) || it == WLMPAR_MAX) break;
                    }
                    float u[WNP], w[WNP];
                    const float idn = __fdividef(1.0f, dxnorm);
#pragma unroll
                    for (int j = 0; j < WNP; ++j) u[j] = T[j] * z[j] * idn;          // D^2 p / |D p| in scaled variables
                    chol7_fwd(ch, u, w);
                    float t2 = 0.0f;
#pragma unroll
                    for (int j = 0; j < WNP; ++j) t2 = fmaf(w[j], w[j], t2);
                    const float parc = __fdividef(__fdividef(fp, delta), t2);
                    if (it == 0) {
                        parl = (ok == 0x7fu && t2 > 0.0f) ? parc : 0.0f;
                        float gs2 = 0.0f;
#pragma unroll
                        for (int j = 0; j < WNP; ++j) { const float t = __fdividef(rhs[j], diag[j] * iS[j]); gs2 = fmaf(t, t, gs2); }
                        const float gsn = sqrtf(gs2);
                        paru = __fdividef(gsn, delta);
                        if (paru == 0.0f) paru = __fdividef(WQ_TINYF, fminf(delta, 0.1f));
                        prr = fminf(fmaxf(par, parl), paru);
                        if (prr == 0.0f) prr = __fdividef(gsn, dxnorm);
                    } else {
                        if (fp > 0.0f) parl = fmaxf(parl, prr);
                        if (fp < 0.0f) paru = fminf(paru, prr);
                        prr = fmaxf(parl, prr + parc);
                    }
                    if (prr == 0.0f) prr = fmaxf(WQ_TINYF, paru * 0.001f);
                }
                par = par_used;

                // ---------------------------------------------------------- bounds (:1184-1231)
                {
                    float mx = pf[0], mn = pf[0];
#pragma unroll
                    for (int j = 1; j < WNP; ++j) { mx = fmaxf(mx, pf[j]); mn = fminf(mn, pf[j]); }
#pragma unroll
                    for (int j = 0; j < WNP; ++j) {
                        if ((lpeg >> j) & 1u) pf[j] = fminf(fmaxf(pf[j], 0.0f), mx);
                        if ((upeg >> j) & 1u) pf[j] = fminf(fmaxf(pf[j], mn), 0.0f);
                    }
                }
                // step scale that lands on the first bound hit: the ratio is formed in FP32 and nudged
                // up by 3 ulp, so the limiting parameter overshoots its bound by a hair and is snapped
                // onto it below (the reference reaches the bound to within machep and snaps likewise)
                float alpha = 1.0f;
#pragma unroll
                for (int j = 0; j < WNP; ++j) {
                    const double xn = x[j] + (double)pf[j];
                    // branch-free: a step hits at most one of a parameter's two bounds, so one division per parameter serves
                    // both tests (twelve short branches with a MUFU each used to sit in this chain)
                    const bool big = fabsf(pf[j]) > machep;
                    const bool vlo = big & lim.has_lo(j) & (xn < lim.lower(j));
                    const bool vhi = big & lim.has_hi(j) & (xn > lim.upper(j));
                    {
                        const double bnd = vlo ? lim.lower(j) : lim.upper(j);
                        const float r = __fdividef(d2f(bnd - x[j]), pf[j]) * (1.0f + 4e-7f);
                        alpha = (vlo | vhi) ? fminf(alpha, r) : alpha;
                    }
                }
                float pn = 0.0f;
                nonfinite = false;
#pragma unroll
                for (int j = 0; j < WNP; ++j) {
                    pf[j] *= alpha;
                    double xn = x[j] + (double)pf[j];
                    if (PFLIB) {
                        const double ll = lim.lower(j);
                        const double llim1 = ll * (1.0 + WQ_MACHEP) + ((ll == 0.0) ? WQ_MACHEP : 0.0);   // ll >= 0 here
                        if (lim.has_hi(j)) {
                            const double ul = lim.upper(j);
                            if (xn >= ul * (1.0 - WQ_MACHEP)) xn = ul;
                        }
                        if (xn <= llim1) xn = ll;
                    } else {                                                             // :1220-1231, any sign
                        if (lim.has_hi(j)) {
                            const double ul = lim.upper(j);
                            const double sgnu = (ul >= 0.0) ? 1.0 : -1.0;
                            if (xn >= ul * (1.0 - sgnu * WQ_MACHEP) - ((ul == 0.0) ? WQ_MACHEP : 0.0)) xn = ul;
                        }
                        if (lim.has_lo(j)) {
                            const double ll = lim.lower(j);
                            const double sgnl = (ll >= 0.0) ? 1.0 : -1.0;
                            if (xn <= ll * (1.0 + sgnl * WQ_MACHEP) + ((ll == 0.0) ? WQ_MACHEP : 0.0)) xn = ll;
                        }
                    }
                    y[j] = xn;
                    const float t = diag[j] * pf[j];
                    pn = fmaf(t, t, pn);
                    nonfinite |= !(isfinite(pf[j]) && isfinite(xn));
                }
                pnorm = sqrtf(pn);
                if (niter == 1) delta = fminf(delta, pnorm);                             // :1237-1238
                float pAp = 0.0f;                                                        // |J p|^2 = zs^T As zs, zs = p / iS
                {
                    float zs[WNP];
#pragma unroll
                    for (int j = 0; j < WNP; ++j) zs[j] = __fdividef(pf[j], iS[j]);
#pragma unroll
                    for (int i = 0; i < WNP; ++i) {
                        float s = 0.0f;
#pragma unroll
                        for (int j = 0; j < WNP; ++j) s = fmaf(sA[((i >= j) ? wtri(i, j) : wtri(j, i)) * TPB], zs[j], s);
                        pAp = fmaf(s, zs[i], pAp);
                    }
                }
                // mpfit applies alpha to the (already scaled) step once more here (:1265)
                const float t1sq = alpha * alpha * fmaxf(pAp, 0.0f) * rss0;
                const float t2sq = alpha * par * pnorm * pnorm * rss0;
                prered = t1sq + 2.0f * t2sq;
                dirder = -(t1sq + t2sq);
                mode = MODE_TRIAL;
            } else {
                // ---------------------------------------------------------- results (pflib.py:461-477)
                if (status > 0) ++nfev;                                                  // :1351-1355
                if (PFLIB) {
                    double* o = a.out_fit + idx * 12;
                    o[0] = (x[2] + (double)cand_h) - 2.5;                                // pflib.py:461
                    o[1] = (x[3] + (double)cand_w) - 2.5;
                    o[2] = x[0]; o[3] = x[1]; o[4] = x[4]; o[5] = x[5]; o[6] = x[6];
                    o[7] = ss0;                 // residual sum of squares at the final parameters: fit_finish_kernel turns
                    o[10] = fmax(ss0, ss1);     // it into rmse / r_2 / sqrt (FP64 sqrt and divisions at full lanes there);
                                                // o[10] = mpfit .fnorm (:1357-1359)
                    *reinterpret_cast<int4*>(a.out_int + idx * 4) = make_int4(status, niter, nfev, n_damped);
                } else if (gl == 0) {
#pragma unroll
                    for (int j = 0; j < WNP; ++j) a.params[idx * WNP + j] = x[j];
                    a.status[idx] = status; a.niter[idx] = niter; a.nfev[idx] = nfev;
                    a.chi2[idx] = fmax(ss0, ss1);
                    if (a.n_damped) a.n_damped[idx] = n_damped;
                }
                active = false;
            }
        }
    }
}

// r_2, rmse (pflib.py:463-472) and sqrt(chi^2) from the residual sum of squares the LM kernel left in column 7
__global__ void __launch_bounds__(256)
fit_finish_kernel(const WarpArgs a) {
    long long n_total = a.n;
    if (a.n_dev) { const long long nd = *a.n_dev; n_total = nd < a.n ? nd : a.n; }
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    double* o = a.out_fit + i * 12;
    const double ss0 = o[7];
    o[7] = sqrt(ss0 / 25.0);
    o[8] = 1.0 - ss0 / a.prep[i].sst;
    o[11] = sqrt(ss0);
}

// model image at the fitted parameters (gaussfitter.py:252-254), one thread per fit
__global__ void __launch_bounds__(128)
fit_image_kernel(const WarpArgs a) {
    long long n_total = a.n;
    if (a.n_dev) { const long long nd = *a.n_dev; n_total = nd < a.n ? nd : a.n; }
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const double* o = a.out_fit + i * 12;
    const double cy = (o[0] - (double)a.cand_hw[2 * i]) + 2.5, cx = (o[1] - (double)a.cand_hw[2 * i + 1]) + 2.5;
    double sn, cs;
    sincos(WQ_DEG2RAD * o[6], &sn, &cs);
    const double iwx = 1.0 / o[4], iwy = 1.0 / o[5];
    for (int r = 0; r < 5; ++r) {
        const double dx = cx - (double)r;
        for (int c = 0; c < 5; ++c) {
            const double dy = cy - (double)c;
            const double aa = (dx * cs - dy * sn) * iwx;
            const double bb = (dx * sn + dy * cs) * iwy;
            a.fit_img[i * 25 + r * 5 + c] = o[2] + o[3] * exp(-0.5 * (aa * aa + bb * bb));
        }
    }
}

// same for the generic entry: params [n,7] -> fit_img [n,win,win]
__global__ void __launch_bounds__(128)
fit_image_generic_kernel(const double* __restrict__ params, const int32_t* __restrict__ status, long long n, int win,
                         double* __restrict__ fit_img) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* o = params + i * 7;
    double sn, cs;
    sincos(WQ_DEG2RAD * o[6], &sn, &cs);
    const double iwx = 1.0 / o[4], iwy = 1.0 / o[5];
    for (int r = 0; r < win; ++r) {
        const double dx = o[3] - (double)r;
        for (int c = 0; c < win; ++c) {
            const double dy = o[2] - (double)c;
            const double aa = (dx * cs - dy * sn) * iwx;
            const double bb = (dx * sn + dy * cs) * iwy;
            fit_img[(i * win + r) * win + c] = o[0] + o[1] * exp(-0.5 * (aa * aa + bb * bb));
        }
    }
}

// [64 B header | n start records | n parked-fit records]
long long warp_scratch_bytes(long long n) { return 64 + (long long)(sizeof(PrepRec) + sizeof(StragRec)) * (n > 0 ? n : 0); }

// One instantiation of the LM kernel: shared-memory size, resident blocks per SM (function attributes and occupancy
// are per device: cached per device ordinal -- one process may drive several GPUs from several host threads,
// psfio.parallel_image_batch does), launch.
template <int WIN, int TPB, int MINB, bool PFLIB, int GRP = 1, typename PXT = double>
struct WKernel {
    static constexpr int tpb = TPB, grp = GRP;
    static constexpr size_t smem = WLayout<WIN, GRP>::template smem_bytes<TPB, PXT>();
    static int per_sm(int* out) {
        static std::atomic<int> per_sm_dev[64];
        int dev = 0;
        FSQ_CUDA_CHECK(cudaGetDevice(&dev));
        int v = (dev >= 0 && dev < 64) ? per_sm_dev[dev].load(std::memory_order_acquire) : 0;
        if (v == 0) {
            FSQ_CUDA_CHECK(cudaFuncSetAttribute(lmwarp_kernel<WIN, TPB, MINB, PFLIB, GRP, PXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, lmwarp_kernel<WIN, TPB, MINB, PFLIB, GRP, PXT>, TPB, smem) != cudaSuccess || v < 1) v = 1;
            if (dev >= 0 && dev < 64) per_sm_dev[dev].store(v, std::memory_order_release);
        }
        *out = v;
        return FSQ_OK;
    }
    static int launch(const WarpArgs& a, long long blocks, cudaStream_t st) {
        lmwarp_kernel<WIN, TPB, MINB, PFLIB, GRP, PXT><<<(unsigned)blocks, TPB, smem, st>>>(a);
        FSQ_LAUNCH_CHECK();
        return FSQ_OK;
    }
};

// Persistent launch(es): phase 1 (kernel K1) over every fit -- parked after `park_after` passes when that is set --
// and phase 2 (kernel K2, by default the same) over the parked fits.
// (K3, when given: a second finishing kernel.  K2 takes the parked fits when there are at most `k2_max` of them, K3
//  otherwise -- both are launched, each reads the parked count on the device and the one whose turn it is not returns
//  at once, so the choice costs no host synchronisation.)
template <typename K1, typename K2 = K1, typename K3 = void>
static int launch_warp(WarpArgs& a, int ctas_per_sm, unsigned long long* head, cudaStream_t st, long long k2_max = 0) {
    int per_sm = 1;
    { const int rc = K1::per_sm(&per_sm); if (rc != FSQ_OK) return rc; }
    const int use_per_sm = (ctas_per_sm > 0 && ctas_per_sm < per_sm) ? ctas_per_sm : per_sm;
    long long blocks = (long long)sm_count() * use_per_sm;
    const long long need = (a.n * K1::grp + K1::tpb - 1) / K1::tpb;
    if (need < blocks) blocks = need < 1 ? 1 : need;
    // park_after > 0: park every fit after that many passes; park_after < 0: drain parking -- once the queue is
    // empty, the fits still running -park_after ticks later (the 100+-iteration stragglers) are parked, so that the
    // launch gives its SM slots back instead of keeping hundreds of warps alive for one lane each; the second launch
    // finishes them packed into a few blocks
    const int park = a.o.park_after;
    a.cap = park > 0 ? park : 0; a.drain_grace = park < 0 ? -park : 0; a.resume = 0;
    { const int rc = K1::launch(a, blocks, st); if (rc != FSQ_OK) return rc; }
    if (park != 0) {
        FSQ_CUDA_CHECK(cudaMemsetAsync(head, 0, sizeof(unsigned long long), st));
        a.cap = 0; a.drain_grace = 0; a.resume = 1;
        int per_sm2 = 1;
        { const int rc = K2::per_sm(&per_sm2); if (rc != FSQ_OK) return rc; }
        // drain parking: the parked fits are few and latency bound -- a few blocks leave the rest of the machine to
        // whatever the caller has queued on other streams (the next batch's detection and phase 1); parking by pass
        // count (the 11x11 entry): every parked fit gets its lane group at once
        long long blocks2 = park < 0 ? 16 : (long long)sm_count() * per_sm2;
        const long long need2 = (a.n * K2::grp + K2::tpb - 1) / K2::tpb;
        if (need2 < blocks2) blocks2 = need2 < 1 ? 1 : need2;
        a.resume_lo = 0; a.resume_hi = 0x7fffffffffffffffLL;
        if constexpr (!std::is_void<K3>::value) a.resume_hi = k2_max;
        { const int rc = K2::launch(a, blocks2, st); if (rc != FSQ_OK) return rc; }
        if constexpr (!std::is_void<K3>::value) {
            FSQ_CUDA_CHECK(cudaMemsetAsync(head, 0, sizeof(unsigned long long), st));
            int per_sm3 = 1;
            { const int rc = K3::per_sm(&per_sm3); if (rc != FSQ_OK) return rc; }
            long long blocks3 = (long long)sm_count() * per_sm3;
            const long long need3 = (a.n * K3::grp + K3::tpb - 1) / K3::tpb;
            if (need3 < blocks3) blocks3 = need3 < 1 ? 1 : need3;
            a.resume_lo = k2_max + 1; a.resume_hi = 0x7fffffffffffffffLL;
            const int rc = K3::launch(a, blocks3, st);
            if (rc != FSQ_OK) return rc;
        }
    }
    return FSQ_OK;
}

// Persistent launch of the frame-path kernel with `warps_per_sm` warps per SM (fsq.h): block size and
// blocks per SM are chosen so that the register budget per thread stays the same (168 registers).
static int launch_frame_path(WarpArgs& a, unsigned long long* head, cudaStream_t st) {
    switch (a.o.warps_per_sm) {
        case 1:  return launch_warp<WKernel<5, 32, 8, true>>(a, 1, head, st);
        case 2:  return launch_warp<WKernel<5, 64, 4, true>>(a, 1, head, st);
        case 4:  return launch_warp<WKernel<5, WTHREADS, WMINB, true>>(a, 1, head, st);
        default: return launch_warp<WKernel<5, WTHREADS, WMINB, true>>(a, 0, head, st);
    }
}

int warp_fit_candidates(const void* frames, int dtype_code, int H, int W, const int32_t* cand_hw,
                        const int32_t* cand_frame, long long n, const long long* n_dev, const fsq_lm_opts* opts,
                        double* out_fit, int32_t* out_int, double* fit_img, void* scratch, cudaStream_t st) {
    WarpArgs a;
    memset(&a, 0, sizeof(a));
    a.frames = frames; a.fdtype = dtype_code; a.H = H; a.W = W; a.cand_hw = cand_hw; a.cand_frame = cand_frame;
    a.n = n; a.n_dev = n_dev; a.o = *opts; a.ftol_f = (float)opts->ftol; a.xtol_f = (float)opts->xtol; a.gtol_f = (float)opts->gtol; a.factor_f = (float)opts->factor; a.out_fit = out_fit; a.out_int = out_int; a.fit_img = fit_img;
    unsigned long long* head = (unsigned long long*)scratch;          // [0] queue head, [1] parked count
    a.work_counter = head; a.strag_count = head + 1;
    a.prep = (PrepRec*)((char*)scratch + 64);
    a.strag = (StragRec*)((char*)scratch + 64 + sizeof(PrepRec) * (size_t)n);
    const unsigned flat = (unsigned)((n + 127) / 128);
    FSQ_CUDA_CHECK(cudaMemsetAsync(head, 0, 64, st));
    fit_prep_kernel<<<flat, 128, 0, st>>>(a);
    FSQ_LAUNCH_CHECK();
    const int rc = launch_frame_path(a, head, st);
    if (rc != FSQ_OK) return rc;
    fit_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a);
    FSQ_LAUNCH_CHECK();
    if (fit_img) {
        fit_image_kernel<<<flat, 128, 0, st>>>(a);
        FSQ_LAUNCH_CHECK();
    }
    return FSQ_OK;
}

// 11x11 windows.  Two kernels share the work: the bulk runs one thread per window (the fewest
// instructions per fit), and every fit still running after W11_PARK passes -- the 5 % that take long, among them the
// 1-3 % that run to maxiter = 200 (250-330 passes) and set the launch's duration -- is parked and finished by the
// lane-group kernel, whose ticks are 3x shorter (measured on B200, a warp alone: 26 us per tick with one thread per
// window, 11.7 us with 4 lanes, 8.3 us with 8 lanes per window).  Which fits are parked depends on their own pass count
// only, so results do not depend on scheduling.  opts.warps_per_sm selects the other arrangements (cross-checks):
// -1 thread per window only, -2 / -3 / -4: 4 / 8 / 2 lanes per window for every fit, -5: bulk + 4-lane finish.
// (Measured and not kept: FP32 pixels in shared memory for 8- / 16-bit window data -- 10 instead of 6 warps per SM for
//  the thread-per-window kernel at 168 instead of 254 registers: 13.1 instead of 8.9 ms, the tail's ticks get longer;
//  cutting the batch in two halves on two streams -- persistent grids do not overlap: 10 ms.)
#ifndef W11_PARK
#define W11_PARK 32
#endif
static int launch_win11(WarpArgs& a, unsigned long long* head, cudaStream_t st) {
    typedef WKernel<11, 64, 3, false, 1> KT;        // thread per window
    typedef WKernel<11, 128, 3, false, 4> KG4;
    typedef WKernel<11, 128, 3, false, 8> KG8;
    typedef WKernel<11, 128, 3, false, 2> KG2;
    switch (a.o.warps_per_sm) {
        case -1: return launch_warp<KT>(a, 0, head, st);
        case -2: return launch_warp<KG4>(a, 0, head, st);
        case -3: return launch_warp<KG8>(a, 0, head, st);
        case -4: return launch_warp<KG2>(a, 0, head, st);
        case -5: if (a.o.park_after == 0) a.o.park_after = W11_PARK; return launch_warp<KT, KG4>(a, 0, head, st);
        case -6: if (a.o.park_after == 0) a.o.park_after = W11_PARK; return launch_warp<KT, KG8>(a, 0, head, st);
        default: {
            // 8 lanes per window have the shorter ticks but half the windows per wave: beyond ~1.75 waves of 8-lane
            // groups (isolated spots park 5 % of 200 000 windows, a dense field 8 %) the 4-lane kernel finishes sooner
            if (a.o.park_after == 0) a.o.park_after = W11_PARK;
            int per_sm8 = 1;
            { const int rc = KG8::per_sm(&per_sm8); if (rc != FSQ_OK) return rc; }
            const long long wave8 = (long long)sm_count() * per_sm8 * (KG8::tpb / KG8::grp);
            return launch_warp<KT, KG8, KG4>(a, 0, head, st, wave8 * 7 / 4);
        }
    }
}

// generic windows (fsq_gaussfit_batch with FSQ_SOLVER_FAST): win = 5 or 11, per-fit start / limits
int warp_gaussfit_batch(const void* windows, int dtype_code, long long n, int win, const double* p0, const double* lo,
                        const double* hi, const uint8_t* lim_lo, const uint8_t* lim_hi, const fsq_lm_opts* opts,
                        double* params, int32_t* status, int32_t* niter, int32_t* nfev, double* chi2,
                        int32_t* n_damped, double* fit_img, cudaStream_t st) {
    WarpArgs a;
    memset(&a, 0, sizeof(a));
    a.windows = windows; a.wdtype = dtype_code; a.p0 = p0; a.lo = lo; a.hi = hi; a.lim_lo = lim_lo; a.lim_hi = lim_hi;
    a.n = n; a.o = *opts; a.ftol_f = (float)opts->ftol; a.xtol_f = (float)opts->xtol; a.gtol_f = (float)opts->gtol; a.factor_f = (float)opts->factor; a.params = params; a.status = status; a.niter = niter; a.nfev = nfev; a.chi2 = chi2;
    a.n_damped = n_damped;
    // this entry has no caller-provided scratch: a stream-ordered allocation holds the queue head(s) and the parked states
    fsq_lm_opts o = *opts;
    a.o = o;
    const bool need_strag = (o.park_after != 0) || (win == 11);
    // The parked-state scratch (128 B per fit: 26 MB for 200 000 windows) comes from the device's stream-ordered pool.
    // With the pool's default release threshold of 0 the memory goes back to the driver at every synchronisation and
    // each call pays for a fresh allocation (measured: a 7 ms call became 35 ms inside bench.py); the threshold is
    // raised once per device so that the pool keeps what it has been given.
    {
        static std::atomic<unsigned long long> pool_done{0};
        int dev = 0;
        FSQ_CUDA_CHECK(cudaGetDevice(&dev));
        const unsigned long long bit = 1ull << (dev & 63);
        if (!(pool_done.load(std::memory_order_acquire) & bit)) {
            cudaMemPool_t pool = nullptr;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
            pool_done.fetch_or(bit, std::memory_order_acq_rel);
        }
    }
    void* scratch = nullptr;
    const size_t bytes = 64 + (need_strag ? sizeof(StragRec) * (size_t)n : 0);
    FSQ_CUDA_CHECK(cudaMallocAsync(&scratch, bytes, st));
    unsigned long long* head = (unsigned long long*)scratch;
    a.work_counter = head; a.strag_count = head + 1; a.strag = (StragRec*)((char*)scratch + 64);
    FSQ_CUDA_CHECK(cudaMemsetAsync(head, 0, 64, st));
    int rc;
    if (win == 5) rc = launch_warp<WKernel<5, 128, 2, false>>(a, 0, head, st);
    else if (win == 11) rc = launch_win11(a, head, st);
    else { set_error("the FAST solver takes 5x5 or 11x11 windows (got %d)", win); rc = FSQ_E_ARG; }
    cudaFreeAsync(scratch, st);
    if (rc != FSQ_OK) return rc;
    if (fit_img) {
        fit_image_generic_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(params, status, n, win, fit_img);
        FSQ_LAUNCH_CHECK();
    }
    return FSQ_OK;
}

}  // namespace fsq
