// `fast` solver: one THREAD per 5x5 window.  The same bounded trust-region Levenberg-Marquardt
// as class mpfit (agpy/mpfit/mpfit.py:600-1388) -- pegging by exact equality and gradient sign
// (:1073-1091), More's lmpar (:2077-2190), step clipping / alpha scaling / snapping (:1184-1231),
// ratio / delta / par updates (:1253-1288), termination tests (:1301-1335), .fnorm (:1357-1359)
// -- but driven by the ANALYTIC Jacobian of the rotated elliptical Gaussian
// (agpy/gaussfitter.py:63-140) through column-scaled normal equations and a 7x7 Cholesky held in
// registers, instead of 7 finite-difference model evaluations and a Householder QR per
// iteration.  One model+Jacobian pass and one model pass per LM iteration (the reference: 8 + 1).
//
// Everything a fit needs lives in the registers of its thread: the 25 pixels, the 7 parameters,
// the packed 28-entry J^T J, J^T f, the Cholesky factor.  No shared memory, no shuffles, no
// divergence between the lanes of a warp inside an iteration; a lane whose fit has ended pulls
// the next window from an atomic work queue (persistent grid), so long fits never hold a warp.
//
// All arithmetic in FP64 (the T = float instantiation of the pixel math is not built).
//
// What is NOT reproduced on purpose: the rounding noise of the reference's finite-difference
// rotation column at the pflib start point and the qrsolv diagonal-view behaviour -- those are
// what the MINPACK kernel (fsq_lmfit.cu, opts.solver = FSQ_SOLVER_MINPACK) exists for.
#include "fsq_common.cuh"
#include "fsq_median.cuh"
#include <string.h>
#include <type_traits>

namespace fsq {

constexpr int FNP = 7;
constexpr int FNT = 28;                  // packed lower triangle
constexpr int FAST_THREADS = 128;
#define FQ_MACHEP 2.220446049250313e-16
#define FQ_DWARF 2.2250738585072014e-308
#define FQ_DEG2RAD 0.017453292519943295

__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

struct FastArgs {
    const void* windows; int wdtype;
    const double* p0; const double* lo; const double* hi;
    const uint8_t* lim_lo; const uint8_t* lim_hi;
    const void* frames; int fdtype; int H; int W;
    const int32_t* cand_hw; const int32_t* cand_frame;
    long long n; const long long* n_dev;
    fsq_lm_opts o;
    double* params; int32_t* status; int32_t* niter; int32_t* nfev; double* chi2; int32_t* n_damped;
    double* fit_img; double* out_fit; int32_t* out_int;
    unsigned long long* work_counter;
    int resume;          // 1: start from the parameters already in params / out_fit (polish launch)
    int write_metrics;   // 0: coarse phase, only parameters + counters are written
};

template <typename T> struct Mth;
template <> struct Mth<double> {
    static __device__ __forceinline__ double ex(double x) { return exp(x); }
    static __device__ __forceinline__ double rsq(double x) { return 1.0 / sqrt(x); }
    static __device__ __forceinline__ void sc(double x, double* s, double* c) { sincos(x, s, c); }
    static __device__ __forceinline__ double rank_eps() { return 16.0 * FQ_MACHEP; }
};
template <> struct Mth<float> {
    static __device__ __forceinline__ float ex(float x) { return __expf(x); }
    static __device__ __forceinline__ float rsq(float x) { return rsqrtf(x); }
    static __device__ __forceinline__ void sc(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float rank_eps() { return 16.0f * 1.1920929e-07f; }
};

// ---- rotated-Gaussian geometry shared by the three passes -------------------------------------
template <typename T>
struct Geo {
    T H, A, cy, cx, iwx, iwy, cs, sn, wx, wy;
    __device__ __forceinline__ void set(const double* x) {
        H = (T)x[0]; A = (T)x[1]; cy = (T)x[2]; cx = (T)x[3]; wx = (T)x[4]; wy = (T)x[5];
        iwx = (T)1 / wx; iwy = (T)1 / wy;
        Mth<T>::sc((T)(FQ_DEG2RAD * x[6]), &sn, &cs);          // gaussfitter.py:115
    }
};

// residual sum of squares at x (model pass)                       gaussfitter.py:133-135, :214
template <typename T, typename DS>
__device__ __forceinline__ double chi_pass(const double* x, const DS (&d)[25]) {
    Geo<T> q; q.set(x);
    T ss = 0;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const T dx = q.cx - (T)r;                  // numpy.indices: x = row index pairs with p[3]
        const T ur = dx * q.cs, vr = dx * q.sn;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const T dy = q.cy - (T)c;
            const T a = (ur - dy * q.sn) * q.iwx;
            const T b = (vr + dy * q.cs) * q.iwy;
            const T E = Mth<T>::ex((T)-0.5 * (a * a + b * b));
            const T f = (T)d[r * 5 + c] - (q.H + q.A * E);
            ss = fma(f, f, ss);
        }
    }
    return (double)ss;
}

// model + analytic Jacobian pass: A = J^T J (packed), g = J^T f, ss = f.f with J = d(residual)/dp
template <typename T, typename DS>
__device__ __forceinline__ void normal_pass(const double* x, const DS (&d)[25], T (&A)[FNT], T (&g)[FNP], T& ss) {
    Geo<T> q; q.set(x);
#pragma unroll
    for (int i = 0; i < FNT; ++i) A[i] = 0;
#pragma unroll
    for (int i = 0; i < FNP; ++i) g[i] = 0;
    ss = 0;
    const T krot = (q.wy * q.iwx - q.wx * q.iwy) * (T)FQ_DEG2RAD;
    const T sx = q.sn * q.iwx, cxw = q.cs * q.iwx, sy = q.sn * q.iwy, cyw = q.cs * q.iwy;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const T dx = q.cx - (T)r;
        const T ur = dx * q.cs, vr = dx * q.sn;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const T dy = q.cy - (T)c;
            const T a = (ur - dy * q.sn) * q.iwx;
            const T b = (vr + dy * q.cs) * q.iwy;
            const T E = Mth<T>::ex((T)-0.5 * (a * a + b * b));
            const T AE = q.A * E;
            const T f = (T)d[r * 5 + c] - (q.H + AE);
            T j[FNP];
            j[0] = (T)-1;
            j[1] = -E;
            j[2] = -AE * (a * sx - b * cyw);               // d/d p[2] (centre along axis 1)
            j[3] = AE * (a * cxw + b * sy);                // d/d p[3] (centre along axis 0)
            j[4] = -AE * a * a * q.iwx;
            j[5] = -AE * b * b * q.iwy;
            j[6] = -AE * a * b * krot;                     // degrees
            ss = fma(f, f, ss);
#pragma unroll
            for (int k = 0; k < FNP; ++k) {
                g[k] = fma(j[k], f, g[k]);
#pragma unroll
                for (int l = 0; l <= k; ++l) A[tri(k, l)] = fma(j[k], j[l], A[tri(k, l)]);
            }
        }
    }
}

// Cholesky (FP32, also in the FP64 flavour: the step only has to be a descent step accurate to
// ~1e-4, the fixed point g = 0 is set by the FP64 gradient) of the column-scaled, damped matrix
//     M = S^-1 A S^-1 + par * (D/S)^2,   unit diagonal at par = 0,
// read from the packed A on the fly, with singular-pivot skipping.  Li = 1 / L_jj (0 when skipped).
template <typename TA>
__device__ __forceinline__ unsigned chol7(const TA (&A)[FNT], const double (&iS)[FNP], const double (&diag)[FNP],
                                          double par, float (&L)[FNT], float (&Li)[FNP], float eps) {
    unsigned ok = 0;
#pragma unroll
    for (int j = 0; j < FNP; ++j) {
        const double dsj = diag[j] * iS[j];
        float dj = (float)((double)A[tri(j, j)] * (iS[j] * iS[j]) + par * dsj * dsj);
#pragma unroll
        for (int k = 0; k < j; ++k) dj = fmaf(-L[tri(j, k)], L[tri(j, k)], dj);
        const bool good = dj > eps;
        const float inv = good ? rsqrtf(dj) : 0.0f;
        Li[j] = inv;
        L[tri(j, j)] = dj * inv;
        ok |= (good ? 1u : 0u) << j;
#pragma unroll
        for (int i = j + 1; i < FNP; ++i) {
            float sacc = (float)((double)A[tri(i, j)] * (iS[i] * iS[j]));
#pragma unroll
            for (int k = 0; k < j; ++k) sacc = fmaf(-L[tri(i, k)], L[tri(j, k)], sacc);
            L[tri(i, j)] = sacc * inv;
        }
    }
    return ok;
}

__device__ __forceinline__ void fwd7(const float (&L)[FNT], const float (&Li)[FNP], const float (&rhs)[FNP], float (&z)[FNP]) {
#pragma unroll
    for (int j = 0; j < FNP; ++j) {
        float s = rhs[j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fmaf(-L[tri(j, k)], z[k], s);
        z[j] = s * Li[j];
    }
}

__device__ __forceinline__ void bwd7(const float (&L)[FNT], const float (&Li)[FNP], float (&z)[FNP]) {   // in place
#pragma unroll
    for (int j = FNP - 1; j >= 0; --j) {
        float s = z[j];
#pragma unroll
        for (int i = j + 1; i < FNP; ++i) s = fmaf(-L[tri(i, j)], z[i], s);
        z[j] = s * Li[j];
    }
}

__device__ __forceinline__ int ld_int(const void* base, int dtype, size_t off) {
    switch (dtype) {
        case FSQ_U8:  return (int)((const uint8_t*)base)[off];
        case FSQ_U16: return (int)((const uint16_t*)base)[off];
        case FSQ_I16: return (int)((const int16_t*)base)[off];
        default:      return ((const int32_t*)base)[off];
    }
}

__device__ __forceinline__ double ld_dbl(const void* base, int dtype, size_t off) {
    switch (dtype) {
        case FSQ_U8:  return (double)((const uint8_t*)base)[off];
        case FSQ_U16: return (double)((const uint16_t*)base)[off];
        case FSQ_I16: return (double)((const int16_t*)base)[off];
        case FSQ_I32: return (double)((const int32_t*)base)[off];
        case FSQ_I64: return (double)((const long long*)base)[off];
        default:      return ((const double*)base)[off];
    }
}

// ==========================================================================================
template <typename T, bool PFLIB>
__global__ void __launch_bounds__(FAST_THREADS)
lmfast_kernel(const FastArgs a) {
    const double ftol = a.o.ftol, xtol = a.o.xtol, gtol = a.o.gtol, factor = a.o.factor;
    const int maxiter = a.o.maxiter;
    long long n_total = a.n;
    if (a.n_dev) { const long long nd = *a.n_dev; n_total = nd < a.n ? nd : a.n; }

    for (;;) {
        const long long idx = (long long)atomicAdd(a.work_counter, 1ull);
        if (idx >= n_total) return;

        // ------------------------------------------------------------ window, start, limits
        // pixels are integers < 2^24 on the frame path: FP32 storage is exact and halves the registers
        typedef typename std::conditional<PFLIB, float, T>::type DS;
        DS d[25];
        double x[FNP], lo[FNP], hi[FNP];
        unsigned qll = 0, qul = 0;
        int cand_h = 0, cand_w = 0;
        double dmax = 0.0, dmean = 0.0, esum = 0.0;          // pflib metrics inputs
        if (PFLIB) {
            cand_h = a.cand_hw[2 * idx]; cand_w = a.cand_hw[2 * idx + 1];
            const size_t fbase = (size_t)a.cand_frame[idx] * a.H * a.W + (size_t)(cand_h - 2) * a.W + (cand_w - 2);
            int v[25];
            long long isum = 0; int imax = -2147483647 - 1;
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const int p = ld_int(a.frames, a.fdtype, fbase + (size_t)r * a.W + c);
                    v[r * 5 + c] = p; d[r * 5 + c] = (DS)p;
                    isum += p; imax = max(imax, p);
                    if (r == 0 || r == 4 || c == 0 || c == 4) esum += (double)p;
                }
            const int imed = median25<int>(v);                 // numpy.median of 25 (pflib.py:199)
            dmax = (double)imax; dmean = (double)isum / 25.0;
            // pflib.py:199-213
            x[0] = (double)imed; x[1] = dmax; x[2] = 2.5; x[3] = 2.5; x[4] = 1.0; x[5] = 1.0; x[6] = 0.0;
            lo[0] = 0.0; lo[1] = (dmax - dmean) / 3.0; lo[2] = 2.0; lo[3] = 2.0; lo[4] = 0.75; lo[5] = 0.75; lo[6] = 0.0;
            hi[0] = 0.0; hi[1] = 0.0; hi[2] = 3.0; hi[3] = 3.0; hi[4] = 2.0; hi[5] = 2.0; hi[6] = 360.0;
            qll = 0x7fu; qul = 0x7cu;
        } else {
#pragma unroll
            for (int i = 0; i < 25; ++i) d[i] = (DS)ld_dbl(a.windows, a.wdtype, (size_t)idx * 25 + i);
#pragma unroll
            for (int j = 0; j < FNP; ++j) {
                x[j] = a.p0[idx * FNP + j]; lo[j] = a.lo[idx * FNP + j]; hi[j] = a.hi[idx * FNP + j];
                qll |= (a.lim_lo[idx * FNP + j] ? 1u : 0u) << j;
                qul |= (a.lim_hi[idx * FNP + j] ? 1u : 0u) << j;
            }
        }
        int base_niter = 0, base_nfev = 0, base_damped = 0;
        if (a.resume) {
            // polish launch: continue from the coarse phase's end point (window coordinates)
            if (PFLIB) {
                const double* o = a.out_fit + idx * 12;
                x[0] = o[2]; x[1] = o[3]; x[2] = (o[0] - (double)cand_h) + 2.5; x[3] = (o[1] - (double)cand_w) + 2.5;
                x[4] = o[4]; x[5] = o[5]; x[6] = o[6];
                base_niter = a.out_int[idx * 4 + 1] - 1; base_nfev = a.out_int[idx * 4 + 2]; base_damped = a.out_int[idx * 4 + 3];
            } else {
#pragma unroll
                for (int j = 0; j < FNP; ++j) x[j] = a.params[idx * FNP + j];
                base_niter = a.niter[idx] - 1; base_nfev = a.nfev[idx];
                base_damped = a.n_damped ? a.n_damped[idx] : 0;
            }
        }
        // gaussfitter.py:202-204 start clamp (pflib path; the generic path is clamped by the caller)
        bool bad = false;
#pragma unroll
        for (int j = 0; j < FNP; ++j) {
            const bool ql = (qll >> j) & 1u, qu = (qul >> j) & 1u;
            if (PFLIB || a.resume) {
                if (qu && x[j] > hi[j]) x[j] = hi[j];
                if (ql && x[j] < lo[j]) x[j] = lo[j];
            }
            bad |= (ql && x[j] < lo[j]) || (qu && x[j] > hi[j]) || (ql && qu && lo[j] >= hi[j]);   // mpfit.py:956-964
        }

        int status = 0, niter = 1, nfev = 1, n_damped = 0;
        double fnorm = -1.0, fnorm1 = -1.0;
        if (!bad) {
            fnorm = sqrt(chi_pass<T, DS>(x, d));                 // mpfit.py:999, :1019
            double diag[FNP], xnew[FNP], p[FNP];
            T A[FNT], g[FNP];
            double iS[FNP];                                 // 1 / column norm (1 for a zero column)
            double delta = 0.0, par = 0.0, xnorm = 0.0, gnorm = 0.0;
            unsigned lpeg = 0, upeg = 0;
            bool need_jac = true;
#pragma unroll
            for (int j = 0; j < FNP; ++j) { diag[j] = 1.0; iS[j] = 1.0; }

#pragma unroll 1
            while (status == 0) {
                if (need_jac) {
                    T ss;
                    normal_pass<T, DS>(x, d, A, g, ss);
                    // pegged parameters: zero the column when the gradient pushes outwards (:1073-1091)
                    lpeg = 0; upeg = 0;
#pragma unroll
                    for (int j = 0; j < FNP; ++j) {
                        const bool lp = ((qll >> j) & 1u) && (x[j] == lo[j]);
                        const bool up = ((qul >> j) & 1u) && (x[j] == hi[j]);
                        lpeg |= (lp ? 1u : 0u) << j; upeg |= (up ? 1u : 0u) << j;
                        const bool zero = (lp && g[j] > (T)0) || (up && g[j] < (T)0);
                        if (zero) {
                            g[j] = 0;
#pragma unroll
                            for (int k = 0; k < FNP; ++k) A[(k >= j) ? tri(k, j) : tri(j, k)] = 0;
                        }
                    }
                    gnorm = 0.0;
                    double acn[FNP];
#pragma unroll
                    for (int j = 0; j < FNP; ++j) {
                        acn[j] = sqrt((double)A[tri(j, j)]);                           // column norms (:1758)
                        iS[j] = (acn[j] > 0.0) ? 1.0 / acn[j] : 1.0;
                        if (acn[j] != 0.0 && fnorm != 0.0)
                            gnorm = fmax(gnorm, fabs((double)g[j] * iS[j] / fnorm));     // :1142-1148
                    }
                    if (niter == 1) {                                                   // :1099-1110
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < FNP; ++j) {
                            diag[j] = (acn[j] == 0.0) ? 1.0 : acn[j];
                            const double t = diag[j] * x[j];
                            s += t * t;
                        }
                        xnorm = sqrt(s);
                        delta = factor * xnorm;
                        if (delta == 0.0) delta = factor;
                    }
                    if (gnorm <= gtol) { status = 4; break; }                           // :1151
                    if (maxiter == 0) { status = 5; break; }
#pragma unroll
                    for (int j = 0; j < FNP; ++j) diag[j] = fmax(diag[j], acn[j]);      // :1160
                    need_jac = false;
                }

                // ---------------------------------------------------------------- lmpar (:2077-2190)
                float L[FNT], Li[FNP], rhs[FNP], z[FNP];
#pragma unroll
                for (int i = 0; i < FNP; ++i) rhs[i] = (float)(-(double)g[i] * iS[i]);
                const unsigned ok = chol7(A, iS, diag, 0.0, L, Li, 16.0f * 1.1920929e-07f);
                fwd7(L, Li, rhs, z);
                bwd7(L, Li, z);
                double dxnorm = 0.0;
#pragma unroll
                for (int j = 0; j < FNP; ++j) { p[j] = (double)z[j] * iS[j]; const double t = diag[j] * p[j]; dxnorm += t * t; }
                dxnorm = sqrt(dxnorm);
                double fp = dxnorm - delta;
                double par_used = 0.0;
                if (fp > 0.1 * delta) {                                   // Gauss-Newton step too long (:2112)
                    double parl = 0.0;
                    if (ok == 0x7fu) {
                        float u[FNP], w[FNP];
#pragma unroll
                        for (int j = 0; j < FNP; ++j) u[j] = (float)(diag[j] * diag[j] * iS[j] * p[j] / dxnorm);
                        fwd7(L, Li, u, w);
                        double t2 = 0.0;
#pragma unroll
                        for (int j = 0; j < FNP; ++j) t2 += (double)w[j] * (double)w[j];
                        if (t2 > 0.0) parl = (fp / delta) / t2;
                    }
                    double gsn = 0.0;
#pragma unroll
                    for (int j = 0; j < FNP; ++j) { const double t = (double)g[j] / diag[j]; gsn += t * t; }
                    gsn = sqrt(gsn);
                    double paru = gsn / delta;
                    if (paru == 0.0) paru = FQ_DWARF / fmin(delta, 0.1);
                    double prr = fmin(fmax(par, parl), paru);
                    if (prr == 0.0) prr = gsn / dxnorm;
#pragma unroll 1
                    for (int it = 0; it < 10; ++it) {
                        if (prr == 0.0) prr = fmax(FQ_DWARF, paru * 0.001);
                        chol7(A, iS, diag, prr, L, Li, 0.0f);
                        fwd7(L, Li, rhs, z);
                        bwd7(L, Li, z);
                        ++n_damped;
                        double dx2 = 0.0;
#pragma unroll
                        for (int j = 0; j < FNP; ++j) { p[j] = (double)z[j] * iS[j]; const double t = diag[j] * p[j]; dx2 += t * t; }
                        dxnorm = sqrt(dx2);
                        const double temp = fp;
                        fp = dxnorm - delta;
                        par_used = prr;
                        if ((fabs(fp) <= 0.1 * delta) || ((parl == 0.0) && (fp <= temp) && (temp < 0.0)) || it == 9) break;
                        float u[FNP], w[FNP];
#pragma unroll
                        for (int j = 0; j < FNP; ++j) u[j] = (float)(diag[j] * diag[j] * iS[j] * p[j] / dxnorm);
                        fwd7(L, Li, u, w);
                        double t2 = 0.0;
#pragma unroll
                        for (int j = 0; j < FNP; ++j) t2 += (double)w[j] * (double)w[j];
                        const double parc = (fp / delta) / t2;
                        if (fp > 0.0) parl = fmax(parl, prr);
                        if (fp < 0.0) paru = fmin(paru, prr);
                        prr = fmax(parl, prr + parc);
                    }
                }
                par = par_used;

                // ---------------------------------------------------------------- bounds (:1184-1231)
                double alpha = 1.0;
                if (qll | qul) {
                    double mx = p[0], mn = p[0];
#pragma unroll
                    for (int j = 1; j < FNP; ++j) { mx = fmax(mx, p[j]); mn = fmin(mn, p[j]); }
#pragma unroll
                    for (int j = 0; j < FNP; ++j) {
                        if ((lpeg >> j) & 1u) p[j] = fmin(fmax(p[j], 0.0), mx);
                        if ((upeg >> j) & 1u) p[j] = fmin(fmax(p[j], mn), 0.0);
                    }
#pragma unroll
                    for (int j = 0; j < FNP; ++j) {
                        if (fabs(p[j]) > FQ_MACHEP) {
                            if (((qll >> j) & 1u) && (x[j] + p[j] < lo[j])) alpha = fmin(alpha, (lo[j] - x[j]) / p[j]);
                            if (((qul >> j) & 1u) && (x[j] + p[j] > hi[j])) alpha = fmin(alpha, (hi[j] - x[j]) / p[j]);
                        }
                    }
                }
                double pn = 0.0;
                bool nonfinite = false;
#pragma unroll
                for (int j = 0; j < FNP; ++j) {
                    p[j] *= alpha;
                    double xn = x[j] + p[j];
                    if (qll | qul) {
                        const double ul = hi[j], ll = lo[j];
                        const double sgnu = (ul >= 0.0) ? 1.0 : -1.0, sgnl = (ll >= 0.0) ? 1.0 : -1.0;
                        const double ulim1 = ul * (1.0 - sgnu * FQ_MACHEP) - ((ul == 0.0) ? FQ_MACHEP : 0.0);
                        const double llim1 = ll * (1.0 + sgnl * FQ_MACHEP) + ((ll == 0.0) ? FQ_MACHEP : 0.0);
                        if (((qul >> j) & 1u) && (xn >= ulim1)) xn = ul;
                        if (((qll >> j) & 1u) && (xn <= llim1)) xn = ll;
                    }
                    xnew[j] = xn;
                    const double t = diag[j] * p[j];
                    pn += t * t;
                    nonfinite |= !(isfinite(p[j]) && isfinite(xn));
                }
                const double pnorm = sqrt(pn);
                if (niter == 1) delta = fmin(delta, pnorm);                              // :1237-1238

                // ---------------------------------------------------------------- trial point (:1245-1273)
                fnorm1 = sqrt(chi_pass<T, DS>(xnew, d));
                ++nfev;
                double actred = -1.0;
                if (0.1 * fnorm1 < fnorm) { const double r = fnorm1 / fnorm; actred = 1.0 - r * r; }
                double pAp = 0.0;                                                        // |J p|^2
#pragma unroll
                for (int i = 0; i < FNP; ++i) {
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < FNP; ++j) s += (double)A[(i >= j) ? tri(i, j) : tri(j, i)] * p[j];
                    pAp += s * p[i];
                }
                // mpfit applies alpha to the (already scaled) step once more here (:1265)
                const double t1sq = alpha * alpha * fmax(pAp, 0.0) / (fnorm * fnorm);
                const double t2sq = alpha * par * pnorm * pnorm / (fnorm * fnorm);
                const double prered = t1sq + t2sq / 0.5;
                const double dirder = -(t1sq + t2sq);
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
                    for (int j = 0; j < FNP; ++j) { x[j] = xnew[j]; const double t = diag[j] * x[j]; s += t * t; }
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
                    if ((fabs(actred) <= FQ_MACHEP) && (prered <= FQ_MACHEP) && (0.5 * ratio <= 1.0)) status = 6;
                    if (delta <= FQ_MACHEP * xnorm) status = 7;
                    if (gnorm <= FQ_MACHEP) status = 8;
                }
                if (status == 0) {
                    if (accepted) need_jac = true;
                    else if (nonfinite || !isfinite(ratio)) status = -16;                 // :1330-1335
                }
            }
            if (status > 0) ++nfev;                                                      // :1351-1355
        } else {
            niter = 0; nfev = 0;
        }

        // ------------------------------------------------------------ results
        const double fn = fmax(fnorm, fnorm1);
        const double chi2 = bad ? -1.0 : fn * fn;                                         // :1357-1359
        niter += base_niter; nfev += base_nfev; n_damped += base_damped;
        double gimg[25];
        const bool want_img = a.write_metrics && (PFLIB || a.fit_img != nullptr);
        if (want_img) {
            Geo<double> q; q.set(x);
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const double dx = q.cx - (double)r;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const double dy = q.cy - (double)c;
                    const double aa = (dx * q.cs - dy * q.sn) * q.iwx;
                    const double bb = (dx * q.sn + dy * q.cs) * q.iwy;
                    gimg[r * 5 + c] = q.H + q.A * exp(-0.5 * (aa * aa + bb * bb));
                }
            }
        }
        if (PFLIB) {
            double* o = a.out_fit + idx * 12;
            int* oi = a.out_int + idx * 4;
            o[0] = (x[2] + (double)cand_h) - 2.5;                                         // pflib.py:461
            o[1] = (x[3] + (double)cand_w) - 2.5;
            o[2] = x[0]; o[3] = x[1]; o[4] = x[4]; o[5] = x[5]; o[6] = x[6];
            if (a.write_metrics) {
                // pflib.py:463-473 + illumina_s_n (:261-281): raster-order sums in float64
                double ssr = 0.0, sst = 0.0, evar = 0.0;
                const double emean = esum / 16.0;
#pragma unroll
                for (int i = 0; i < 25; ++i) {
                    const double di = (double)d[i];
                    const double e1 = di - gimg[i];
                    ssr += e1 * e1;
                    const double e2 = di - dmean;
                    sst += e2 * e2;
                    const int r = i / 5, c = i % 5;
                    if (r == 0 || r == 4 || c == 0 || c == 4) { const double e3 = di - emean; evar += e3 * e3; }
                }
                o[7] = sqrt(ssr / 25.0); o[8] = 1.0 - ssr / sst; o[9] = (dmax - emean) / sqrt(evar / 16.0);
                o[10] = chi2; o[11] = fnorm;
                if (a.fit_img) {
#pragma unroll
                    for (int i = 0; i < 25; ++i) a.fit_img[idx * 25 + i] = gimg[i];
                }
            }
            oi[0] = bad ? 0 : status; oi[1] = niter; oi[2] = nfev; oi[3] = n_damped;
        } else {
#pragma unroll
            for (int j = 0; j < FNP; ++j) a.params[idx * FNP + j] = x[j];
            a.status[idx] = bad ? 0 : status; a.niter[idx] = niter; a.nfev[idx] = nfev; a.chi2[idx] = chi2;
            if (a.n_damped) a.n_damped[idx] = n_damped;
            if (a.fit_img && a.write_metrics) {
#pragma unroll
                for (int i = 0; i < 25; ++i) a.fit_img[idx * 25 + i] = gimg[i];
            }
        }
    }
}

template <typename T, bool PFLIB>
static int launch_one(const FastArgs& a, cudaStream_t st) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lmfast_kernel<T, PFLIB>, FAST_THREADS, 0) != cudaSuccess || per_sm < 1)
        per_sm = 2;
    long long blocks = (long long)sm_count() * per_sm;
    const long long need = (a.n + FAST_THREADS - 1) / FAST_THREADS;
    if (need < blocks) blocks = need < 1 ? 1 : need;
    FSQ_CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st));
    lmfast_kernel<T, PFLIB><<<(unsigned)blocks, FAST_THREADS, 0, st>>>(a);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}

// FSQ_SOLVER_FAST64: one FP64 launch.
template <bool PFLIB>
int launch_fast(FastArgs a, cudaStream_t st) {
    a.resume = 0; a.write_metrics = 1;
    return launch_one<double, PFLIB>(a, st);
}

template int launch_fast<true>(FastArgs, cudaStream_t);
template int launch_fast<false>(FastArgs, cudaStream_t);

// entry used by fsq_lmfit.cu's extern "C" functions
int fast_fit_candidates(const void* frames, int dtype_code, int H, int W, const int32_t* cand_hw,
                        const int32_t* cand_frame, long long n, const long long* n_dev, const fsq_lm_opts* opts,
                        double* out_fit, int32_t* out_int, double* fit_img, unsigned long long* work_counter,
                        cudaStream_t st) {
    FastArgs a;
    memset(&a, 0, sizeof(a));
    a.frames = frames; a.fdtype = dtype_code; a.H = H; a.W = W; a.cand_hw = cand_hw; a.cand_frame = cand_frame;
    a.n = n; a.n_dev = n_dev; a.o = *opts; a.out_fit = out_fit; a.out_int = out_int; a.fit_img = fit_img;
    a.work_counter = work_counter;
    return launch_fast<true>(a, st);
}

int fast_gaussfit_batch(const void* windows, int dtype_code, long long n, const double* p0, const double* lo,
                        const double* hi, const uint8_t* lim_lo, const uint8_t* lim_hi, const fsq_lm_opts* opts,
                        double* params, int32_t* status, int32_t* niter, int32_t* nfev, double* chi2,
                        int32_t* n_damped, double* fit_img, unsigned long long* work_counter, cudaStream_t st) {
    FastArgs a;
    memset(&a, 0, sizeof(a));
    a.windows = windows; a.wdtype = dtype_code; a.p0 = p0; a.lo = lo; a.hi = hi; a.lim_lo = lim_lo; a.lim_hi = lim_hi;
    a.n = n; a.o = *opts; a.params = params; a.status = status; a.niter = niter; a.nfev = nfev; a.chi2 = chi2;
    a.n_damped = n_damped; a.fit_img = fit_img; a.work_counter = work_counter;
    return launch_fast<false>(a, st);
}

}  // namespace fsq
