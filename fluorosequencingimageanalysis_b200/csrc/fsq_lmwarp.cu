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
#include <string.h>

namespace fsq {

constexpr int WNP = 7;
constexpr int WNT = 28;
constexpr int WTHREADS = 128;
#define WQ_MACHEP 2.220446049250313e-16
#define WQ_DWARF 2.2250738585072014e-308
#define WQ_DEG2RAD 0.017453292519943295

__host__ __device__ constexpr int wtri(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

struct WarpArgs {
    const void* frames; int fdtype; int H; int W;
    const int32_t* cand_hw; const int32_t* cand_frame;
    long long n; const long long* n_dev;
    fsq_lm_opts o;
    double* out_fit; int32_t* out_int; double* fit_img;
    unsigned long long* work_counter;
};

__device__ __forceinline__ int w_ld_int(const void* base, int dtype, size_t off) {
    switch (dtype) {
        case FSQ_U8:  return (int)((const uint8_t*)base)[off];
        case FSQ_U16: return (int)((const uint16_t*)base)[off];
        case FSQ_I16: return (int)((const int16_t*)base)[off];
        default:      return ((const int32_t*)base)[off];
    }
}

// pflib limits (pflib.py:199-213): lower on all seven, upper on centres, widths, angle
__device__ __forceinline__ double pf_lo(int j, double lo1) {
    return j == 0 ? 0.0 : j == 1 ? lo1 : (j == 2 || j == 3) ? 2.0 : (j == 4 || j == 5) ? 0.75 : 0.0;
}
__device__ __forceinline__ double pf_hi(int j) { return (j == 2 || j == 3) ? 3.0 : (j == 4 || j == 5) ? 2.0 : 360.0; }
constexpr unsigned PF_QLL = 0x7fu, PF_QUL = 0x7cu;

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
    long long isum = 0, esum = 0;
    int imax = -2147483647 - 1;
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int p = w_ld_int(a.frames, a.fdtype, fbase + (size_t)r * a.W + c);
            v[r * 5 + c] = p;
            isum += p; imax = max(imax, p);
            if (r == 0 || r == 4 || c == 0 || c == 4) esum += p;
        }
    const double dmean = (double)isum / 25.0, emean = (double)esum / 16.0;
    double sst = 0.0, evar = 0.0;
#pragma unroll
    for (int q = 0; q < 25; ++q) {                               // raster order (pflib.py:464)
        const double e = (double)v[q] - dmean;
        sst += e * e;
        const int r = q / 5, c = q % 5;
        if (r == 0 || r == 4 || c == 0 || c == 4) { const double e3 = (double)v[q] - emean; evar += e3 * e3; }
    }
    const int imed = median25<int>(v);                           // numpy.median of 25 (pflib.py:199)
    int* oi = a.out_int + i * 4;
    oi[0] = imed; oi[1] = imax; oi[2] = (int)isum; oi[3] = 0;
    double* o = a.out_fit + i * 12;
    o[8] = sst;
    o[9] = ((double)imax - emean) / sqrt(evar / 16.0);           // illumina_s_n, pflib.py:261-281
}

// -------------------------------------------------------------------------------------------
// 7x7 Cholesky of the column-scaled, damped normal matrix
//     M = S^-1 A S^-1 + par (D/S)^2        (unit diagonal at par = 0)
// A read from shared memory ([entry][thread]); singular pivots are skipped (Li = 0).
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned w_chol7(const float* __restrict__ sA, const float (&iS)[WNP], const float (&diag)[WNP],
                                            float par, float (&L)[WNT], float (&Li)[WNP], float eps) {
    unsigned ok = 0;
#pragma unroll
    for (int j = 0; j < WNP; ++j) {
        const float dsj = diag[j] * iS[j];
        float dj = fmaf(par * dsj, dsj, sA[wtri(j, j) * WTHREADS] * (iS[j] * iS[j]));
#pragma unroll
        for (int k = 0; k < j; ++k) dj = fmaf(-L[wtri(j, k)], L[wtri(j, k)], dj);
        const bool good = dj > eps;
        const float inv = good ? rsqrtf(dj) : 0.0f;
        Li[j] = inv;
        L[wtri(j, j)] = dj * inv;
        ok |= (good ? 1u : 0u) << j;
#pragma unroll
        for (int i = j + 1; i < WNP; ++i) {
            float sacc = sA[wtri(i, j) * WTHREADS] * (iS[i] * iS[j]);
#pragma unroll
            for (int k = 0; k < j; ++k) sacc = fmaf(-L[wtri(i, k)], L[wtri(j, k)], sacc);
            L[wtri(i, j)] = sacc * inv;
        }
    }
    return ok;
}

__device__ __forceinline__ void w_fwd7(const float (&L)[WNT], const float (&Li)[WNP], const float (&rhs)[WNP], float (&z)[WNP]) {
#pragma unroll
    for (int j = 0; j < WNP; ++j) {
        float s = rhs[j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fmaf(-L[wtri(j, k)], z[k], s);
        z[j] = s * Li[j];
    }
}

__device__ __forceinline__ void w_bwd7(const float (&L)[WNT], const float (&Li)[WNP], float (&z)[WNP]) {
#pragma unroll
    for (int j = WNP - 1; j >= 0; --j) {
        float s = z[j];
#pragma unroll
        for (int i = j + 1; i < WNP; ++i) s = fmaf(-L[wtri(i, j)], z[i], s);
        z[j] = s * Li[j];
    }
}

// One pass over the window at pt: chi^2 in FP64; J^T J (packed), J^T f in FP32 (J = d residual / dp).
__device__ __forceinline__ void w_pass(const double (&pt)[WNP], const double* __restrict__ sd,
                                       float (&A)[WNT], float (&g)[WNP], double& ss_out) {
    const double Hh = pt[0], Aa = pt[1];
    double sn, cs;
    sincos(WQ_DEG2RAD * pt[6], &sn, &cs);                                     // gaussfitter.py:115
    const double iwx = 1.0 / pt[4], iwy = 1.0 / pt[5];
    const double cxs = cs * iwx, sxs = sn * iwx, cys = cs * iwy, sys = sn * iwy;
    double ca[5], cb[5];
    float caf[5], cbf[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const double dy = pt[2] - (double)c;            // numpy.indices: y = column index pairs with p[2]
        ca[c] = dy * sxs; cb[c] = dy * cys;
        caf[c] = (float)ca[c]; cbf[c] = (float)cb[c];
    }
    const float Af = (float)Aa, sx = (float)sxs, cxw = (float)cxs, sy = (float)sys, cyw = (float)cys;
    const float iwxf = (float)iwx, iwyf = (float)iwy;
    const float krot = (float)((pt[5] * iwx - pt[4] * iwy) * WQ_DEG2RAD);
#pragma unroll
    for (int i = 0; i < WNT; ++i) A[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < WNP; ++i) g[i] = 0.0f;
    double ss = 0.0;
#pragma unroll 1
    for (int r = 0; r < 5; ++r) {
        const double dx = pt[3] - (double)r;            // x = row index pairs with p[3]
        const double ra = dx * cxs, rb = dx * sys;
        const float raf = (float)ra, rbf = (float)rb;
        const double* drow = sd + r * 5 * WTHREADS;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const double av = ra - ca[c], bv = rb + cb[c];
            const double E = exp(-0.5 * fma(bv, bv, av * av));
            const double f = drow[c * WTHREADS] - fma(Aa, E, Hh);
            ss = fma(f, f, ss);
            const float af = raf - caf[c], bf = rbf + cbf[c];
            const float Ef = (float)E, ff = (float)f;
            const float AE = Af * Ef;
            const float AEa = AE * af, AEb = AE * bf;
            float j[WNP];
            j[1] = -Ef;
            j[2] = -(AEa * sx - AEb * cyw);             // d/d p[2] (centre along axis 1)
            j[3] = AEa * cxw + AEb * sy;                // d/d p[3] (centre along axis 0)
            j[4] = -AEa * af * iwxf;
            j[5] = -AEb * bf * iwyf;
            j[6] = -AEa * bf * krot;                    // degrees
            // column 0 of J is the constant -1
            g[0] -= ff;
#pragma unroll
            for (int k = 1; k < WNP; ++k) {
                g[k] = fmaf(j[k], ff, g[k]);
                A[wtri(k, 0)] -= j[k];
#pragma unroll
                for (int l = 1; l <= k; ++l) A[wtri(k, l)] = fmaf(j[k], j[l], A[wtri(k, l)]);
            }
        }
    }
    A[0] = 25.0f;
    ss_out = ss;
}

enum { MODE_FIRST = 0, MODE_TRIAL = 1 };

__global__ void __launch_bounds__(WTHREADS, 3)
lmwarp_kernel(const WarpArgs a) {
    __shared__ double s_d[25 * WTHREADS];
    __shared__ float s_A[WNT * WTHREADS];
    __shared__ float s_g[WNP * WTHREADS];
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    double* const sd = s_d + tid;
    float* const sA = s_A + tid;
    float* const sg = s_g + tid;

    const double ftol = a.o.ftol, xtol = a.o.xtol, gtol = a.o.gtol, factor = a.o.factor;
    const int maxiter = a.o.maxiter;
    long long n_total = a.n;
    if (a.n_dev) { const long long nd = *a.n_dev; n_total = nd < a.n ? nd : a.n; }

    // ---- per-lane fit state
    bool active = false, exhausted = false;
    long long idx = 0;
    int cand_h = 0, cand_w = 0;
    int mode = MODE_FIRST, status = 0, niter = 1, nfev = 0, n_damped = 0;
    unsigned lpeg = 0, upeg = 0;
    bool nonfinite = false;
    double x[WNP], y[WNP];
    double lo1 = 0.0, fnorm = -1.0, fnorm1 = -1.0, delta = 0.0, par = 0.0, xnorm = 0.0, gnorm = 0.0;
    double pnorm = 0.0, prered = 0.0, dirder = 0.0;
    float diag[WNP], iS[WNP];
#pragma unroll
    for (int j = 0; j < WNP; ++j) { x[j] = 0.0; y[j] = 1.0; diag[j] = 1.0f; iS[j] = 1.0f; }

    for (;;) {
        // ------------------------------------------------------------------ refill idle lanes
        __syncwarp();
        const unsigned want = __ballot_sync(0xffffffffu, !active && !exhausted);
        if (want) {
            const int leader = __ffs(want) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(a.work_counter, (unsigned long long)__popc(want));
            base = __shfl_sync(0xffffffffu, base, leader);
            if ((want >> lane) & 1u) {
                idx = (long long)(base + (unsigned long long)__popc(want & ((1u << lane) - 1u)));
                if (idx >= n_total) exhausted = true;
                else {
                    cand_h = a.cand_hw[2 * idx]; cand_w = a.cand_hw[2 * idx + 1];
                    const size_t fbase = (size_t)a.cand_frame[idx] * a.H * a.W + (size_t)(cand_h - 2) * a.W + (cand_w - 2);
#pragma unroll
                    for (int r = 0; r < 5; ++r)
#pragma unroll
                        for (int c = 0; c < 5; ++c)
                            sd[(r * 5 + c) * WTHREADS] = (double)w_ld_int(a.frames, a.fdtype, fbase + (size_t)r * a.W + c);
                    const int4 pre = *reinterpret_cast<const int4*>(a.out_int + idx * 4);
                    const double dmax = (double)pre.y, dmean = (double)pre.z / 25.0;
                    lo1 = (dmax - dmean) / 3.0;                                        // pflib.py:205
                    x[0] = (double)pre.x; x[1] = dmax; x[2] = 2.5; x[3] = 2.5; x[4] = 1.0; x[5] = 1.0; x[6] = 0.0;
#pragma unroll
                    for (int j = 0; j < WNP; ++j) {                                    // gaussfitter.py:202-204
                        if (((PF_QUL >> j) & 1u) && x[j] > pf_hi(j)) x[j] = pf_hi(j);
                        if (x[j] < pf_lo(j, lo1)) x[j] = pf_lo(j, lo1);
                        y[j] = x[j];
                    }
                    active = true; mode = MODE_FIRST; status = 0; niter = 1; nfev = 0; n_damped = 0;
                    fnorm = -1.0; fnorm1 = -1.0; par = 0.0; nonfinite = false;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) break;

        if (active) {
            // -------------------------------------------------------------- pass at the trial point
            float An[WNT], gn[WNP];
            double ss;
            w_pass(y, sd, An, gn, ss);                      // y == x on the first tick of a fit
            ++nfev;

            bool have_new = false;
            if (mode == MODE_FIRST) {
                fnorm = sqrt(ss);                                                        // mpfit.py:999, :1019
                have_new = true;
            } else {
                // ---------------------------------------------------------- trial bookkeeping (:1245-1335)
                fnorm1 = sqrt(ss);
                double actred = -1.0;
                if (0.1 * fnorm1 < fnorm) { const double r = fnorm1 / fnorm; actred = 1.0 - r * r; }
                double ratio = 0.0;
                if (prered != 0.0) ratio = actred / prered;
                if (ratio <= 0.25) {                                                     // :1276-1288
                    double temp;
                    if (actred >= 0.0) temp = 0.5;
                    else temp = 0.5 * dirder / (dirder + 0.5 * actred);
                    if ((0.1 * fnorm1 >= fnorm) || (temp < 0.1)) temp = 0.1;
                    delta = temp * fmin(delta, pnorm / 0.1);
                    par = par / temp;
                } else if ((par == 0.0) || (ratio >= 0.75)) {
                    delta = pnorm / 0.5;
                    par = 0.5 * par;
                }
                const bool accepted = ratio >= 0.0001;                                   // :1291-1298
                if (accepted) {
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < WNP; ++j) { x[j] = y[j]; const double t = (double)diag[j] * x[j]; s += t * t; }
                    xnorm = sqrt(s);
                    fnorm = fnorm1;
                    ++niter;
                }
                const bool c1 = (fabs(actred) <= ftol) && (prered <= ftol) && (0.5 * ratio <= 1.0);   // :1301-1323
                if (c1) status = 1;
                if (delta <= xtol * xnorm) status = 2;
                if (c1 && status == 2) status = 3;
                if (status == 0) {
                    if (niter >= maxiter) status = 5;
                    if ((fabs(actred) <= WQ_MACHEP) && (prered <= WQ_MACHEP) && (0.5 * ratio <= 1.0)) status = 6;
                    if (delta <= WQ_MACHEP * xnorm) status = 7;
                    if (gnorm <= WQ_MACHEP) status = 8;
                }
                if (status == 0 && !accepted && (nonfinite || !isfinite(ratio))) status = -16;   // :1330-1335
                have_new = accepted;
            }

            if (status == 0 && have_new) {
                // ---------------------------------------------------------- new linearisation at x
                // pegged parameters: zero the column when the gradient pushes outwards (:1073-1091)
                lpeg = 0; upeg = 0;
#pragma unroll
                for (int j = 0; j < WNP; ++j) {
                    const bool lp = (x[j] == pf_lo(j, lo1));
                    const bool up = ((PF_QUL >> j) & 1u) && (x[j] == pf_hi(j));
                    lpeg |= (lp ? 1u : 0u) << j; upeg |= (up ? 1u : 0u) << j;
                    const bool zero = (lp && gn[j] > 0.0f) || (up && gn[j] < 0.0f);
                    if (zero) {
                        gn[j] = 0.0f;
#pragma unroll
                        for (int k = 0; k < WNP; ++k) An[(k >= j) ? wtri(k, j) : wtri(j, k)] = 0.0f;
                    }
                }
                float acn[WNP];
                float gmax = 0.0f;
#pragma unroll
                for (int j = 0; j < WNP; ++j) {
                    const float ajj = An[wtri(j, j)];
                    const float rs = ajj > 0.0f ? rsqrtf(ajj) : 0.0f;
                    acn[j] = ajj * rs;                                                   // column norms (:1758)
                    iS[j] = ajj > 0.0f ? rs : 1.0f;
                    if (ajj > 0.0f) gmax = fmaxf(gmax, fabsf(gn[j] * rs));               // :1142-1148
                }
                gnorm = (fnorm != 0.0) ? (double)gmax / fnorm : 0.0;
                if (mode == MODE_FIRST) {                                                // :1099-1110
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < WNP; ++j) {
                        diag[j] = (acn[j] == 0.0f) ? 1.0f : acn[j];
                        const double t = (double)diag[j] * x[j];
                        s += t * t;
                    }
                    xnorm = sqrt(s);
                    delta = factor * xnorm;
                    if (delta == 0.0) delta = factor;
                }
                if (gnorm <= gtol) status = 4;                                           // :1151
                else if (maxiter == 0) status = 5;
#pragma unroll
                for (int j = 0; j < WNP; ++j) diag[j] = fmaxf(diag[j], acn[j]);          // :1160
#pragma unroll
                for (int i = 0; i < WNT; ++i) sA[i * WTHREADS] = An[i];
#pragma unroll
                for (int i = 0; i < WNP; ++i) sg[i * WTHREADS] = gn[i];
            }

            if (status == 0) {
                // ---------------------------------------------------------- lmpar (:2077-2190)
                float L[WNT], Li[WNP], rhs[WNP], z[WNP];
#pragma unroll
                for (int i = 0; i < WNP; ++i) rhs[i] = -sg[i * WTHREADS] * iS[i];
                const unsigned ok = w_chol7(sA, iS, diag, 0.0f, L, Li, 16.0f * 1.1920929e-07f);
                w_fwd7(L, Li, rhs, z);
                w_bwd7(L, Li, z);
                float pf[WNP];
                float dx2f = 0.0f;
#pragma unroll
                for (int j = 0; j < WNP; ++j) { pf[j] = z[j] * iS[j]; const float t = diag[j] * pf[j]; dx2f = fmaf(t, t, dx2f); }
                double dxnorm = sqrt((double)dx2f);
                double fp = dxnorm - delta;
                double par_used = 0.0;
                if (fp > 0.1 * delta) {                                   // Gauss-Newton step too long (:2112)
                    double parl = 0.0;
                    if (ok == 0x7fu) {
                        float u[WNP], w[WNP];
                        const float idx_ = (float)(1.0 / dxnorm);
#pragma unroll
                        for (int j = 0; j < WNP; ++j) u[j] = diag[j] * diag[j] * iS[j] * pf[j] * idx_;
                        w_fwd7(L, Li, u, w);
                        float t2 = 0.0f;
#pragma unroll
                        for (int j = 0; j < WNP; ++j) t2 = fmaf(w[j], w[j], t2);
                        if (t2 > 0.0f) parl = (fp / delta) / (double)t2;
                    }
                    float gs2 = 0.0f;
#pragma unroll
                    for (int j = 0; j < WNP; ++j) { const float t = sg[j * WTHREADS] / diag[j]; gs2 = fmaf(t, t, gs2); }
                    const double gsn = sqrt((double)gs2);
                    double paru = gsn / delta;
                    if (paru == 0.0) paru = WQ_DWARF / fmin(delta, 0.1);
                    double prr = fmin(fmax(par, parl), paru);
                    if (prr == 0.0) prr = gsn / dxnorm;
#pragma unroll 1
                    for (int it = 0; it < 10; ++it) {
                        if (prr == 0.0) prr = fmax(WQ_DWARF, paru * 0.001);
                        w_chol7(sA, iS, diag, fmaxf((float)prr, 1e-30f), L, Li, 0.0f);
                        w_fwd7(L, Li, rhs, z);
                        w_bwd7(L, Li, z);
                        ++n_damped;
                        dx2f = 0.0f;
#pragma unroll
                        for (int j = 0; j < WNP; ++j) { pf[j] = z[j] * iS[j]; const float t = diag[j] * pf[j]; dx2f = fmaf(t, t, dx2f); }
                        dxnorm = sqrt((double)dx2f);
                        const double temp = fp;
                        fp = dxnorm - delta;
                        par_used = prr;
                        if ((fabs(fp) <= 0.1 * delta) || ((parl == 0.0) && (fp <= temp) && (temp < 0.0)) || it == 9) break;
                        float u[WNP], w[WNP];
                        const float idx_ = (float)(1.0 / dxnorm);
#pragma unroll
                        for (int j = 0; j < WNP; ++j) u[j] = diag[j] * diag[j] * iS[j] * pf[j] * idx_;
                        w_fwd7(L, Li, u, w);
                        float t2 = 0.0f;
#pragma unroll
                        for (int j = 0; j < WNP; ++j) t2 = fmaf(w[j], w[j], t2);
                        const double parc = (fp / delta) / (double)t2;
                        if (fp > 0.0) parl = fmax(parl, prr);
                        if (fp < 0.0) paru = fmin(paru, prr);
                        prr = fmax(parl, prr + parc);
                    }
                }
                par = par_used;

                // ---------------------------------------------------------- bounds (:1184-1231)
                double p[WNP];
#pragma unroll
                for (int j = 0; j < WNP; ++j) p[j] = (double)pf[j];
                double alpha = 1.0;
                {
                    double mx = p[0], mn = p[0];
#pragma unroll
                    for (int j = 1; j < WNP; ++j) { mx = fmax(mx, p[j]); mn = fmin(mn, p[j]); }
#pragma unroll
                    for (int j = 0; j < WNP; ++j) {
                        if ((lpeg >> j) & 1u) p[j] = fmin(fmax(p[j], 0.0), mx);
                        if ((upeg >> j) & 1u) p[j] = fmin(fmax(p[j], mn), 0.0);
                    }
#pragma unroll
                    for (int j = 0; j < WNP; ++j) {
                        if (fabs(p[j]) > WQ_MACHEP) {
                            if (x[j] + p[j] < pf_lo(j, lo1)) alpha = fmin(alpha, (pf_lo(j, lo1) - x[j]) / p[j]);
                            if (((PF_QUL >> j) & 1u) && (x[j] + p[j] > pf_hi(j))) alpha = fmin(alpha, (pf_hi(j) - x[j]) / p[j]);
                        }
                    }
                }
                double pn = 0.0;
                nonfinite = false;
#pragma unroll
                for (int j = 0; j < WNP; ++j) {
                    p[j] *= alpha;
                    double xn = x[j] + p[j];
                    const double ll = pf_lo(j, lo1);
                    const double llim1 = ll * (1.0 + WQ_MACHEP) + ((ll == 0.0) ? WQ_MACHEP : 0.0);   // ll >= 0 here
                    if ((PF_QUL >> j) & 1u) {
                        const double ul = pf_hi(j);
                        if (xn >= ul * (1.0 - WQ_MACHEP)) xn = ul;
                    }
                    if (xn <= llim1) xn = ll;
                    y[j] = xn;
                    pf[j] = (float)p[j];
                    const double t = (double)diag[j] * p[j];
                    pn += t * t;
                    nonfinite |= !(isfinite(p[j]) && isfinite(xn));
                }
                pnorm = sqrt(pn);
                if (niter == 1) delta = fmin(delta, pnorm);                              // :1237-1238
                float pAp = 0.0f;                                                        // |J p|^2
#pragma unroll
                for (int i = 0; i < WNP; ++i) {
                    float s = 0.0f;
#pragma unroll
                    for (int j = 0; j < WNP; ++j) s = fmaf(sA[((i >= j) ? wtri(i, j) : wtri(j, i)) * WTHREADS], pf[j], s);
                    pAp = fmaf(s, pf[i], pAp);
                }
                // mpfit applies alpha to the (already scaled) step once more here (:1265)
                const double t1sq = alpha * alpha * fmax((double)pAp, 0.0) / (fnorm * fnorm);
                const double t2sq = alpha * par * pnorm * pnorm / (fnorm * fnorm);
                prered = t1sq + t2sq / 0.5;
                dirder = -(t1sq + t2sq);
                mode = MODE_TRIAL;
            } else {
                // ---------------------------------------------------------- results (pflib.py:461-477)
                if (status > 0) ++nfev;                                                  // :1351-1355
                const double fn = fmax(fnorm, fnorm1);
                double* o = a.out_fit + idx * 12;
                const double sst = o[8];
                const double ssr = fnorm * fnorm;                  // residual sum of squares at the final parameters
                o[0] = (x[2] + (double)cand_h) - 2.5;                                    // pflib.py:461
                o[1] = (x[3] + (double)cand_w) - 2.5;
                o[2] = x[0]; o[3] = x[1]; o[4] = x[4]; o[5] = x[5]; o[6] = x[6];
                o[7] = sqrt(ssr / 25.0); o[8] = 1.0 - ssr / sst;
                o[10] = fn * fn;                                                         // mpfit .fnorm (:1357-1359)
                o[11] = fnorm;
                *reinterpret_cast<int4*>(a.out_int + idx * 4) = make_int4(status, niter, nfev, n_damped);
                active = false;
            }
        }
    }
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

int warp_fit_candidates(const void* frames, int dtype_code, int H, int W, const int32_t* cand_hw,
                        const int32_t* cand_frame, long long n, const long long* n_dev, const fsq_lm_opts* opts,
                        double* out_fit, int32_t* out_int, double* fit_img, unsigned long long* work_counter,
                        cudaStream_t st) {
    WarpArgs a;
    memset(&a, 0, sizeof(a));
    a.frames = frames; a.fdtype = dtype_code; a.H = H; a.W = W; a.cand_hw = cand_hw; a.cand_frame = cand_frame;
    a.n = n; a.n_dev = n_dev; a.o = *opts; a.out_fit = out_fit; a.out_int = out_int; a.fit_img = fit_img;
    a.work_counter = work_counter;
    const unsigned flat = (unsigned)((n + 127) / 128);
    fit_prep_kernel<<<flat, 128, 0, st>>>(a);
    FSQ_LAUNCH_CHECK();
    static int per_sm = 0;
    if (per_sm == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, lmwarp_kernel, WTHREADS, 0) != cudaSuccess || v < 1) v = 2;
        per_sm = v;
    }
    long long blocks = (long long)sm_count() * per_sm;
    const long long need = (n + WTHREADS - 1) / WTHREADS;
    if (need < blocks) blocks = need < 1 ? 1 : need;
    FSQ_CUDA_CHECK(cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), st));
    lmwarp_kernel<<<(unsigned)blocks, WTHREADS, 0, st>>>(a);
    FSQ_LAUNCH_CHECK();
    if (fit_img) {
        fit_image_kernel<<<flat, 128, 0, st>>>(a);
        FSQ_LAUNCH_CHECK();
    }
    return FSQ_OK;
}

}  // namespace fsq
