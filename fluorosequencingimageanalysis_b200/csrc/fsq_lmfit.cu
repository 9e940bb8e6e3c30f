// Batched bounded Levenberg-Marquardt fit of the 7-parameter rotated elliptical Gaussian
// (replaces gaussfitter.gaussfit -> class mpfit; agpy/gaussfitter.py:142-255,
// agpy/mpfit/mpfit.py:600-1388) plus the fit-quality metrics of pflib.py:461-473.
//
// Mapping (B200, FP64 pipe bound -- never tensor cores):
//   * one sub-warp ("group") of G lanes per window: G = 8 for <= 32 pixels (5x5), G = 32 for
//     <= 128 pixels (11x11); lane g owns pixels  s*G + g, s < S = 4  ("slots").
//   * data, residuals, exp() values and the P x 7 finite-difference Jacobian live in
//     registers (J[s][k], static indices; column pivoting swaps registers physically).
//   * Householder QR with column pivoting is done on the register-distributed Jacobian with
//     xor-butterfly shuffle reductions (bitwise identical in every lane of a group, so all
//     control flow stays group-uniform); per-column scalars (rdiag, wa, acnorm) are owned by
//     lane k.  R (7x7), Q^T f and all 7-vectors of the trust-region logic live in shared
//     memory, one FitShared per group, executed redundantly by the lanes of the group.
//   * persistent kernel: every group pulls the next window from an atomic work queue and the
//     LM state machine is ONE loop (jacobian phase / trial phase / finalize phase), so a
//     group whose fit ends early never waits for its warp neighbours' long fits.
//
// Reference behaviour kept literally (App. B of SURVEY.md): forward-difference Jacobian with
// h = sqrt(eps)|x| and the upper-limit sign flip (mpfit.py:1557-1587), pegging by exact
// equality and gradient sign (:1073-1091), lmpar/qrsolv (:2077-2190, :1903-1978) with the
// optional diagonal-view behaviour (`faithful`), step clipping / alpha scaling / snapping
// (:1184-1231), ratio / delta / par updates (:1253-1288), termination tests (:1301-1335),
// .fnorm = max(fnorm, fnorm1)^2 (:1357-1359), covariance (:2274-2336).
#include "fsq_common.cuh"
#include <string.h>

namespace fsq {

constexpr int NP = 7;
constexpr int SLOTS = 4;
constexpr int LM_THREADS = 128;
#define FSQ_MACHEP 2.220446049250313e-16
#define FSQ_DWARF 2.2250738585072014e-308
#define FSQ_SQRT_MACHEP 1.4901161193847656e-08
#define FSQ_DEG2RAD 0.017453292519943295

struct FitShared {
    double r[NP][NP];
    double x[NP], xnew[NP], step[NP], diag[NP], qtf[NP], acnorm[NP];
    double llim[NP], ulim[NP];
    double lw1[NP], lw2[NP], lwa[NP], sdiag[NP], xsave[NP];
    double pad[3];                       // 1040 B -> 260 words: groups of a warp land on distinct banks
};

struct LmArgs {
    // generic window mode
    const void* windows; int wdtype; int win;
    const double* p0; const double* lo; const double* hi;
    const uint8_t* lim_lo; const uint8_t* lim_hi;
    // pflib mode: frames + candidates
    const void* frames; int fdtype; int H; int W;
    const int32_t* cand_hw; const int32_t* cand_frame;
    long long n; const long long* n_dev;
    fsq_lm_opts o;
    double* params; double* perror; int32_t* status; int32_t* niter; int32_t* nfev;
    double* chi2; int32_t* n_qrsolv; double* fit_img;
    double* out_fit; int32_t* out_int;
    unsigned long long* work_counter;
    double* trace; int trace_steps; long long trace_n;   // optional per-trial-step trace (debug/tests)
    // general gaussfit surface (lmfit_kernel<G, false, true>): fixed parameters, residual weights, the circular model,
    // the full covariance matrix -- gaussfitter.py:188-232, mpfit.py:917-948, :1361-1388
    const uint8_t* fixed;      // [n,7] 1 = parinfo 'fixed' (mpfit.py:917-921) or NULL
    const double* err;         // [n,P] residuals are divided by it (gaussfitter.py:218) or NULL
    int circle;                // 1: width_y := width_x, no rotation (gaussfitter.py:104-107); parameters 5, 6 are not fitted
    double* covar;             // [n,7,7] mpfit .covar (zero rows / columns for fixed parameters) or NULL
};

__device__ __forceinline__ int perm_get(unsigned perm, int j) { return (perm >> (4 * j)) & 15; }

__device__ __forceinline__ double enorm7(const double* v, int n = NP) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NP; ++j) if (j < n) s += v[j] * v[j];
    return sqrt(s);
}

// Free-parameter bookkeeping of the general kernel: mpfit works on x = xall[ifree] (mpfit.py:943-948), so every
// vector and matrix of the algorithm is indexed by the POSITION k < n of a free parameter; pmap holds ifree[k] in
// nibble k.  The standard kernels (GEN = false) have n = 7 and the identity map at compile time.
template <bool GEN>
__device__ __forceinline__ int pidx(unsigned pmap, int k) { return GEN ? (int)((pmap >> (4 * k)) & 15u) : k; }

// full parameter vector of the model from the free vector xs[0..n) and the fixed values pbase
template <bool GEN>
__device__ __forceinline__ void expand_params(const double* xs, int n, unsigned pmap, const double (&pbase)[NP], bool circle,
                                              double (&p)[NP]) {
    if (!GEN) {
#pragma unroll
        for (int j = 0; j < NP; ++j) p[j] = xs[j];
        return;
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) p[j] = pbase[j];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        if (k < n) {
            const double v = xs[k];
            const int pj = pidx<true>(pmap, k);
#pragma unroll
            for (int j = 0; j < NP; ++j) if (pj == j) p[j] = v;
        }
    }
    if (circle) p[5] = p[4];                               // gaussfitter.py:105 width_x = width_y
}

template <int S>
struct Rot { double cs, sn; double xp[S], yp[S]; };

// The model is evaluated with the reference's exact operation sequence, every operation
// individually rounded (no FMA contraction, true divisions): the rotation column of the
// finite-difference Jacobian is pure rounding noise at the pflib start point (theta pegged at 0
// with width_x == width_y), so the first LM step -- and with it the end point of most
// non-stable fits -- depends on the last bit of every intermediate (DESIGN.md "Parity").
template <int S>
__device__ __forceinline__ void make_rot(double theta_deg, const double (&px)[S], const double (&py)[S], Rot<S>& r) {
    const double rota = __dmul_rn(FSQ_DEG2RAD, theta_deg);   // gaussfitter.py:115  pi/180. * rota
    sincos(rota, &r.sn, &r.cs);
#pragma unroll
    for (int s = 0; s < S; ++s) {                         // gaussfitter.py:128-129
        r.xp[s] = __dsub_rn(__dmul_rn(px[s], r.cs), __dmul_rn(py[s], r.sn));
        r.yp[s] = __dadd_rn(__dmul_rn(px[s], r.sn), __dmul_rn(py[s], r.cs));
    }
}

// exp(-(((rcen_x-xp)/width_x)^2 + ((rcen_y-yp)/width_y)^2)/2)  -- gaussfitter.py:116-117,133-135
// p[2] is popped as center_y, p[3] as center_x (gaussfitter.py:100)
template <int S>
__device__ __forceinline__ void eval_E(const double (&p)[NP], const Rot<S>& r, double (&E)[S]) {
    const double rcx = __dsub_rn(__dmul_rn(p[3], r.cs), __dmul_rn(p[2], r.sn));
    const double rcy = __dadd_rn(__dmul_rn(p[3], r.sn), __dmul_rn(p[2], r.cs));
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const double a = __ddiv_rn(__dsub_rn(rcx, r.xp[s]), p[4]);
        const double b = __ddiv_rn(__dsub_rn(rcy, r.yp[s]), p[5]);
        E[s] = exp(__dmul_rn(-0.5, __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b))));
    }
}

// height + amplitude * E, then data - model (gaussfitter.py:133, :214), separately rounded
__device__ __forceinline__ double model_of(double height, double amplitude, double e) {
    return __dadd_rn(height, __dmul_rn(amplitude, e));
}

__device__ __forceinline__ double load_as_double(const void* base, int dtype, size_t off) {
    switch (dtype) {
        case FSQ_U8:  return (double)((const uint8_t*)base)[off];
        case FSQ_U16: return (double)((const uint16_t*)base)[off];
        case FSQ_I16: return (double)((const int16_t*)base)[off];
        case FSQ_I32: return (double)((const int32_t*)base)[off];
        case FSQ_I64: return (double)((const long long*)base)[off];
        default:      return ((const double*)base)[off];
    }
}

__device__ __forceinline__ int load_as_int(const void* base, int dtype, size_t off) {
    switch (dtype) {
        case FSQ_U8:  return (int)((const uint8_t*)base)[off];
        case FSQ_U16: return (int)((const uint16_t*)base)[off];
        case FSQ_I16: return (int)((const int16_t*)base)[off];
        default:      return ((const int32_t*)base)[off];
    }
}

// ------------------------------------------------------------------------------------------
// qrsolv -- mpfit.py:1903-1978.  d = sqrt(par)*diag (st.lw1), qtb = st.qtf.  Output: st.step
// (x, original order), st.sdiag; st.r modified in place.  faithful: the diagonal is not
// restored and the solution is scattered into it (numpy.diagonal view, :1915/:1956/:1977).
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void qrsolv(FitShared& st, unsigned perm, bool faithful, const int n) {
    for (int j = 0; j < n; ++j)
        for (int i = j; i < n; ++i) st.r[i][j] = st.r[j][i];
    if (!faithful)
        for (int j = 0; j < n; ++j) st.xsave[j] = st.r[j][j];
    for (int j = 0; j < n; ++j) st.lwa[j] = st.qtf[j];
    for (int j = 0; j < n; ++j) {
        const int l = perm_get(perm, j);
        if (st.lw1[l] == 0.0) break;
        for (int k = j; k < n; ++k) st.sdiag[k] = 0.0;
        st.sdiag[j] = st.lw1[l];
        double qtbpj = 0.0;
        for (int k = j; k < n; ++k) {
            const double sk = st.sdiag[k];
            if (sk == 0.0) break;
            const double rkk = st.r[k][k];
            double sine, cosine;
            if (fabs(rkk) < fabs(sk)) {
                const double cotan = rkk / sk;
                sine = 0.5 / sqrt(.25 + .25 * cotan * cotan);
                cosine = sine * cotan;
            } else {
                const double tang = sk / rkk;
                cosine = 0.5 / sqrt(.25 + .25 * tang * tang);
                sine = cosine * tang;
            }
            st.r[k][k] = cosine * rkk + sine * sk;
            const double wk = st.lwa[k];
            st.lwa[k] = cosine * wk + sine * qtbpj;
            qtbpj = -sine * wk + cosine * qtbpj;
            for (int i = k + 1; i < n; ++i) {
                const double rik = st.r[i][k], si = st.sdiag[i];
                st.r[i][k] = cosine * rik + sine * si;
                st.sdiag[i] = -sine * rik + cosine * si;
            }
        }
        st.sdiag[j] = st.r[j][j];
        if (!faithful) st.r[j][j] = st.xsave[j];
    }
    int nsing = n;
    for (int j = 0; j < n; ++j)
        if (st.sdiag[j] == 0.0) { nsing = j; break; }
    for (int j = nsing; j < n; ++j) st.lwa[j] = 0.0;
    if (nsing >= 1) {
        st.lwa[nsing - 1] = st.lwa[nsing - 1] / st.sdiag[nsing - 1];
        for (int j = nsing - 2; j >= 0; --j) {
            double sum0 = 0.0;
            for (int i = j + 1; i < nsing; ++i) sum0 += st.r[i][j] * st.lwa[i];
            st.lwa[j] = (st.lwa[j] - sum0) / st.sdiag[j];
        }
    }
    for (int k = 0; k < n; ++k) st.step[perm_get(perm, k)] = st.lwa[k];
    if (faithful)
        for (int k = 0; k < n; ++k) { const int l = perm_get(perm, k); st.r[l][l] = st.lwa[k]; }
}

// ------------------------------------------------------------------------------------------
// lmpar -- mpfit.py:2077-2190.  Returns the new par; step (un-negated) in st.step.
// ------------------------------------------------------------------------------------------
__device__ __noinline__ double lmpar(FitShared& st, unsigned perm, double delta, double par,
                                     bool faithful, int& n_qrsolv, const int n) {
    double dmax = 0.0;
    for (int j = 0; j < n; ++j) dmax = fmax(dmax, fabs(st.r[j][j]));
    const double rthresh = dmax * FSQ_MACHEP;
    int nsing = n;
    for (int j = 0; j < n; ++j) st.lw1[j] = st.qtf[j];
    for (int j = 0; j < n; ++j)
        if (fabs(st.r[j][j]) < rthresh) { nsing = j; break; }
    for (int j = nsing; j < n; ++j) st.lw1[j] = 0.0;
    for (int j = nsing - 1; j >= 0; --j) {
        const double t = st.lw1[j] / st.r[j][j];
        st.lw1[j] = t;
        for (int i = 0; i < j; ++i) st.lw1[i] = st.lw1[i] - st.r[i][j] * t;
    }
    for (int j = 0; j < n; ++j) st.step[perm_get(perm, j)] = st.lw1[j];
    for (int j = 0; j < n; ++j) st.lw2[j] = st.diag[j] * st.step[j];
    double dxnorm = enorm7(st.lw2, n);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) return 0.0;                       // Gauss-Newton step accepted (:2112)

    double parl = 0.0;
    if (nsing >= n) {
        for (int j = 0; j < n; ++j) { const int l = perm_get(perm, j); st.lw1[j] = st.diag[l] * st.lw2[l] / dxnorm; }
        st.lw1[0] = st.lw1[0] / st.r[0][0];
        for (int j = 1; j < n; ++j) {
            double sum0 = 0.0;
            for (int i = 0; i < j; ++i) sum0 += st.r[i][j] * st.lw1[i];
            st.lw1[j] = (st.lw1[j] - sum0) / st.r[j][j];
        }
        const double temp = enorm7(st.lw1, n);
        parl = ((fp / delta) / temp) / temp;
    }
    for (int j = 0; j < n; ++j) {
        double sum0 = 0.0;
        for (int i = 0; i <= j; ++i) sum0 += st.r[i][j] * st.qtf[i];
        st.lw1[j] = sum0 / st.diag[perm_get(perm, j)];
    }
    const double gnorm = enorm7(st.lw1, n);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = FSQ_DWARF / fmin(delta, 0.1);
    par = fmax(par, parl);
    par = fmin(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;

    int iter = 0;
    for (;;) {
        ++iter;
        if (par == 0.0) par = fmax(FSQ_DWARF, paru * 0.001);
        double temp = sqrt(par);
        for (int j = 0; j < n; ++j) st.lw1[j] = temp * st.diag[j];
        ++n_qrsolv;
        qrsolv(st, perm, faithful, n);
        for (int j = 0; j < n; ++j) st.lw2[j] = st.diag[j] * st.step[j];
        dxnorm = enorm7(st.lw2, n);
        temp = fp;
        fp = dxnorm - delta;
        if ((fabs(fp) <= 0.1 * delta) || ((parl == 0.0) && (fp <= temp) && (temp < 0.0)) || (iter == 10)) break;
        for (int j = 0; j < n; ++j) { const int l = perm_get(perm, j); st.lw1[j] = st.diag[l] * st.lw2[l] / dxnorm; }
        for (int j = 0; j < n - 1; ++j) {
            const double t = st.lw1[j] / st.sdiag[j];
            st.lw1[j] = t;
            for (int i = j + 1; i < n; ++i) st.lw1[i] = st.lw1[i] - st.r[i][j] * t;
        }
        st.lw1[n - 1] = st.lw1[n - 1] / st.sdiag[n - 1];
        temp = enorm7(st.lw1, n);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = fmax(parl, par);
        if (fp < 0.0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
    return par;
}

// ------------------------------------------------------------------------------------------
// covariance -> perror (mpfit.py:2274-2336, :1361-1388).  Works in st.r (destroyed).
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void covar_perror(FitShared& st, unsigned perm, double* perr /*st.lw1*/, const int n) {
    double (*r)[NP] = st.r;
    int l = -1;
    const double tolr = 1.e-14 * fabs(r[0][0]);
    for (int k = 0; k < n; ++k) {
        if (fabs(r[k][k]) <= tolr) break;
        r[k][k] = 1.0 / r[k][k];
        for (int j = 0; j < k; ++j) {
            const double temp = r[k][k] * r[j][k];
            r[j][k] = 0.0;
            for (int i = 0; i <= j; ++i) r[i][k] = r[i][k] - temp * r[i][j];
        }
        l = k;
    }
    if (l >= 0) {
        for (int k = 0; k <= l; ++k) {
            for (int j = 0; j < k; ++j) {
                const double temp = r[j][k];
                for (int i = 0; i <= j; ++i) r[i][j] = r[i][j] + temp * r[i][k];
            }
            const double temp = r[k][k];
            for (int i = 0; i <= k; ++i) r[i][k] = temp * r[i][k];
        }
    }
    // the full lower triangle of the covariance in the strict lower triangle of r and in wa (mpfit.py:2313-2327);
    // wa = st.lw2.  (Position (ii, jj) with ii > jj is only ever written, never read, by a later (i, j) of this
    // loop: reads are confined to i <= j, the upper triangle.)
    for (int j = 0; j < n; ++j) {
        const int jj = perm_get(perm, j);
        const bool sing = j > l;
        for (int i = 0; i <= j; ++i) {
            if (sing) r[i][j] = 0.0;
            const int ii = perm_get(perm, i);
            if (ii > jj) r[ii][jj] = r[i][j];
            if (ii < jj) r[jj][ii] = r[i][j];
        }
        st.lw2[jj] = r[j][j];
    }
    for (int j = 0; j < n; ++j) {                            // symmetrize (:2330-2333)
        for (int i = 0; i <= j; ++i) r[i][j] = r[j][i];
        r[j][j] = st.lw2[j];
    }
    for (int j = 0; j < n; ++j) {                            // perror = sqrt(diag) where diag >= 0 (:1382-1386)
        const double d = r[j][j];
        perr[j] = (d >= 0.0) ? sqrt(d) : 0.0;
    }
}

// ==========================================================================================
// GEN: the general gaussfit surface (fixed parameters, weights, circular model): n free parameters at run time.
// GEN = false instantiations are the standard 7-parameter kernels, arithmetic unchanged.
template <int G, bool PFLIB, bool GEN = false>
__global__ void __launch_bounds__(LM_THREADS)
lmfit_kernel(const LmArgs a) {
    static_assert(!(PFLIB && GEN), "the frame path is always the 7-parameter model");
    constexpr int S = SLOTS;
    extern __shared__ __align__(16) unsigned char fsq_smem[];
    const int lane32 = threadIdx.x & 31;
    const int g = threadIdx.x % G;
    const int gbase = lane32 - g;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    FitShared& st = reinterpret_cast<FitShared*>(fsq_smem)[threadIdx.x / G];

    const int win = PFLIB ? 5 : a.win;
    const int P = win * win;
    const bool faithful = a.o.faithful != 0;
    const double ftol = a.o.ftol, xtol = a.o.xtol, gtol = a.o.gtol, factor = a.o.factor;
    const int maxiter = a.o.maxiter;
    long long n_total = a.n;
    if (a.n_dev) { const long long nd = *a.n_dev; n_total = nd < a.n ? nd : a.n; }

    double px[S], py[S];
    bool valid[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int pix = s * G + g;
        valid[s] = pix < P;
        const int rr = pix / win;
        px[s] = (double)rr;
        py[s] = (double)(pix - rr * win);
    }

    // ---- per-fit state (registers) ----
    double dat[S], fv[S], E[S];
    double wgt[GEN ? S : 1];                          // residual weights 1 / err (GEN)
    double pbase[NP];                                 // full start vector: the values of the fixed parameters (GEN)
    int n = NP;                                       // free parameters
    unsigned pmap = 0x6543210u;                       // ifree
    const bool circle = GEN && (a.circle != 0);
#pragma unroll
    for (int j = 0; j < NP; ++j) pbase[j] = 0.0;
    int idat[S];
    Rot<S> rot;
    long long idx = 0;
    bool have = false, need_jac = false;
    int status = 0, niter = 0, nfev = 0, n_qrsolv = 0;
    unsigned perm = 0x6543210u, qll = 0, qul = 0, lpeg = 0, upeg = 0;
    double fnorm = 0.0, fnorm1 = -1.0, delta = 0.0, par = 0.0, xnorm = 0.0, gnorm = 0.0;
    int cand_h = 0, cand_w = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) { dat[s] = 0.0; fv[s] = 0.0; E[s] = 0.0; idat[s] = 0; }

    for (;;) {
        // ================================================================ fetch + init
        if (!have) {
            unsigned long long t = 0;
            if (g == 0) t = atomicAdd(a.work_counter, 1ull);
            t = __shfl_sync(gmask, t, 0, G);
            idx = (long long)t;
            if (idx >= n_total) break;
            have = true;
            status = 0; niter = 0; nfev = 0; n_qrsolv = 0; fnorm1 = -1.0; par = 0.0;
            perm = 0x6543210u; lpeg = 0; upeg = 0; gnorm = 0.0;
            double p0v = 0.0, lov = 0.0, hiv = 0.0;       // lane j < 7 owns parameter j
            bool ql = false, qu = false;
            if (PFLIB) {
                cand_h = a.cand_hw[2 * idx]; cand_w = a.cand_hw[2 * idx + 1];
                const int f = a.cand_frame[idx];
                const size_t fbase = (size_t)f * a.H * a.W;
                long long isum = 0; int imax = -2147483647 - 1;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    idat[s] = 0;
                    if (valid[s]) {
                        const int rr = (int)px[s], cc = (int)py[s];
                        idat[s] = load_as_int(a.frames, a.fdtype,
                                              fbase + (size_t)(cand_h - 2 + rr) * a.W + (cand_w - 2 + cc));
                        isum += idat[s];
                        imax = max(imax, idat[s]);
                    }
                    dat[s] = (double)idat[s];
                }
#pragma unroll
                for (int m = G / 2; m >= 1; m >>= 1) {
                    isum += __shfl_xor_sync(gmask, isum, m, G);
                    imax = max(imax, __shfl_xor_sync(gmask, imax, m, G));
                }
                // median of 25 by rank counting (np.median: the 13th smallest)
                int cnt[S];
#pragma unroll
                for (int s = 0; s < S; ++s) cnt[s] = 0;
#pragma unroll
                for (int q = 0; q < 25; ++q) {
                    const int vq = __shfl_sync(gmask, idat[q / G], q % G, G);
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const int pix = s * G + g;
                        cnt[s] += (vq < idat[s]) || (vq == idat[s] && q < pix);
                    }
                }
                int imed = 0;
#pragma unroll
                for (int s = 0; s < S; ++s) if (valid[s] && cnt[s] == 12) imed = idat[s];
#pragma unroll
                for (int m = G / 2; m >= 1; m >>= 1) imed += __shfl_xor_sync(gmask, imed, m, G);
                const double dmax = (double)imax, dmean = (double)isum / 25.0;
                // pflib.py:199-213
                switch (g) {
                    case 0: p0v = (double)imed; lov = 0.0; hiv = 0.0; ql = true; qu = false; break;
                    case 1: p0v = dmax; lov = (dmax - dmean) / 3.0; hiv = 0.0; ql = true; qu = false; break;
                    case 2: case 3: p0v = 2.5; lov = 2.0; hiv = 3.0; ql = true; qu = true; break;
                    case 4: case 5: p0v = 1.0; lov = 0.75; hiv = 2.0; ql = true; qu = true; break;
                    case 6: p0v = 0.0; lov = 0.0; hiv = 360.0; ql = true; qu = true; break;
                    default: break;
                }
                // gaussfitter.py:202-204 start clamp
                if (qu && p0v > hiv) p0v = hiv;
                if (ql && p0v < lov) p0v = lov;
            } else {
                const size_t wbase = (size_t)idx * P;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int pix = s * G + g;
                    dat[s] = valid[s] ? load_as_double(a.windows, a.wdtype, wbase + pix) : 0.0;
                }
                if (g < NP) {
                    p0v = a.p0[idx * NP + g]; lov = a.lo[idx * NP + g]; hiv = a.hi[idx * NP + g];
                    ql = a.lim_lo[idx * NP + g] != 0; qu = a.lim_hi[idx * NP + g] != 0;
                }
                if (GEN) {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const int pix = s * G + g;
                        wgt[GEN ? s : 0] = (a.err && valid[s]) ? a.err[wbase + pix] : 1.0;
                    }
                }
            }
            bool is_fixed = false;
            if (GEN) {
                // free parameters (mpfit.py:917-948): parinfo 'fixed'; the circular model has no parameters 5, 6
                is_fixed = (g < NP) && ((a.fixed && a.fixed[idx * NP + g] != 0) || (circle && g >= 5));
                const unsigned freem = (__ballot_sync(gmask, (g < NP) && !is_fixed) >> gbase) & 0x7fu;
                n = __popc(freem);
                pmap = 0;
                {
                    int k = 0;
#pragma unroll
                    for (int j = 0; j < NP; ++j) if ((freem >> j) & 1u) { pmap |= (unsigned)j << (4 * k); ++k; }
                }
#pragma unroll
                for (int j = 0; j < NP; ++j) pbase[j] = __shfl_sync(gmask, p0v, j, G);
            }
            // mpfit.py:956-964 limit checks -> status 0, niter 0 (the start check covers fixed parameters too,
            // the consistency check only free ones)
            const bool bad1 = (g < NP) && ((ql && p0v < lov) || (qu && p0v > hiv));
            const bool bad2 = (g < NP) && (ql && qu && lov >= hiv) && !is_fixed;
            if (GEN) {
                // compress to the free positions: lane k < n takes the values of parameter ifree[k]
                const int src = (g < n) ? pidx<true>(pmap, g) : 0;
                const double c_p0 = __shfl_sync(gmask, p0v, src, G), c_lo = __shfl_sync(gmask, lov, src, G),
                             c_hi = __shfl_sync(gmask, hiv, src, G);
                const bool c_ql = __shfl_sync(gmask, (int)ql, src, G) != 0, c_qu = __shfl_sync(gmask, (int)qu, src, G) != 0;
                p0v = c_p0; lov = c_lo; hiv = c_hi; ql = c_ql && (g < n); qu = c_qu && (g < n);
            }
            if (g < NP) { st.x[g] = p0v; st.llim[g] = lov; st.ulim[g] = hiv; }
            qll = (__ballot_sync(gmask, ql && g < NP) >> gbase) & 0x7fu;
            qul = (__ballot_sync(gmask, qu && g < NP) >> gbase) & 0x7fu;
            const bool bad = __any_sync(gmask, bad1 || bad2) || (GEN && n == 0);      // 'no free parameters' (:944-946)
            __syncwarp(gmask);
            if (bad) {
                status = 0; niter = 0; nfev = 0; fnorm = -1.0; fnorm1 = -1.0;
                need_jac = false;
                // finalize below with status 0: handled by 'bad_input' flag
                if (PFLIB) {
                    if (g == 0) {
                        for (int q = 0; q < 12; ++q) a.out_fit[idx * 12 + q] = 0.0;
                        a.out_int[idx * 4 + 0] = 0; a.out_int[idx * 4 + 1] = 0;
                        a.out_int[idx * 4 + 2] = 0; a.out_int[idx * 4 + 3] = 0;
                    }
                } else {
                    if (g < NP) {
                        double pv = p0v;
                        if (GEN) {
#pragma unroll
                            for (int j = 0; j < NP; ++j) if (g == j) pv = pbase[j];
                        }
                        a.params[idx * NP + g] = pv;
                        if (a.perror) a.perror[idx * NP + g] = 0.0;
                    }
                    if (GEN && a.covar) for (int q = g; q < NP * NP; q += G) a.covar[idx * NP * NP + q] = 0.0;
                    if (g == 0) {
                        a.status[idx] = 0; a.niter[idx] = 0; a.nfev[idx] = 0; a.chi2[idx] = -1.0;
                        if (a.n_qrsolv) a.n_qrsolv[idx] = 0;
                    }
                }
                have = false;
                continue;
            }
            // first residual (mpfit.py:999), fnorm (:1019)
            {
                double p[NP];
                expand_params<GEN>(st.x, n, pmap, pbase, circle, p);
                make_rot<S>(p[6], px, py, rot);
                eval_E<S>(p, rot, E);
                double ss = 0.0;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    fv[s] = valid[s] ? __dsub_rn(dat[s], model_of(p[0], p[1], E[s])) : 0.0;
                    if (GEN) fv[s] = __ddiv_rn(fv[s], wgt[GEN ? s : 0]);                 // gaussfitter.py:218
                    ss = fma(fv[s], fv[s], ss);
                }
                fnorm = sqrt(group_sum<G>(ss, gmask));
            }
            nfev = 1; niter = 1; need_jac = true;
        }

        // ================================================================ jacobian phase
        if (need_jac) {
            double J[S][NP];
            double p[NP];                                  // full parameter vector of the model
            double xc[NP];                                 // free vector x = xall[ifree] (== p when !GEN)
#pragma unroll
            for (int j = 0; j < NP; ++j) xc[j] = st.x[j];
            expand_params<GEN>(st.x, n, pmap, pbase, circle, p);
            // ---- forward differences, mpfit.py:1512-1612 ----
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                if (GEN && j >= n) {
#pragma unroll
                    for (int s = 0; s < S; ++s) J[s][j] = 0.0;             // no such column: inert in everything below
                    continue;
                }
                const int pj = pidx<GEN>(pmap, j);         // which model parameter position j perturbs
                const double xj = xc[j];
                double h = FSQ_SQRT_MACHEP * fabs(xj);
                if (h == 0.0) h = FSQ_SQRT_MACHEP;
                if (((qul >> j) & 1u) && (xj > st.ulim[j] - h)) h = -h;
                const double xph = xj + h;
                double Ep[S];
                if (pj >= 2) {
                    double q[NP];
#pragma unroll
                    for (int i = 0; i < NP; ++i) q[i] = (GEN ? (pj == i) : (j == i)) ? xph : p[i];
                    if (GEN && circle) q[5] = q[4];
                    if (pj == 6) {
                        Rot<S> r2;
                        make_rot<S>(xph, px, py, r2);
                        eval_E<S>(q, r2, Ep);
                    } else {
                        eval_E<S>(q, rot, Ep);
                    }
                }
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    double gm;
                    if (pj == 0) gm = model_of(xph, p[1], E[s]);
                    else if (pj == 1) gm = model_of(p[0], xph, E[s]);
                    else gm = model_of(p[0], p[1], Ep[s]);
                    double fp_ = valid[s] ? __dsub_rn(dat[s], gm) : 0.0;
                    if (GEN) fp_ = __ddiv_rn(fp_, wgt[GEN ? s : 0]);
                    J[s][j] = __ddiv_rn(__dsub_rn(fp_, fv[s]), h);              // mpfit.py:1599
                }
            }
            nfev += n;
            // ---- pegged parameters, mpfit.py:1073-1091 ----
            lpeg = 0; upeg = 0;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const bool lp = ((qll >> j) & 1u) && (xc[j] == st.llim[j]);
                const bool up = ((qul >> j) & 1u) && (xc[j] == st.ulim[j]);
                if (lp || up) {
                    double d = 0.0;
#pragma unroll
                    for (int s = 0; s < S; ++s) d = fma(fv[s], J[s][j], d);
                    d = group_sum<G>(d, gmask);
                    bool zero = false;
                    if (lp) { lpeg |= 1u << j; if (d > 0.0) zero = true; }
                    if (up) { upeg |= 1u << j; if (d < 0.0) zero = true; }
                    if (zero) {
#pragma unroll
                        for (int s = 0; s < S; ++s) J[s][j] = 0.0;
                    }
                }
            }
            // ---- Householder QR with column pivoting, mpfit.py:1748-1822 ----
            double acn_l, rdiag_l, wa_l;
            {
                double mine = 0.0;
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    double t = 0.0;
#pragma unroll
                    for (int s = 0; s < S; ++s) t = fma(J[s][k], J[s][k], t);
                    t = group_sum<G>(t, gmask);
                    if (g == k) mine = t;
                }
                acn_l = sqrt(mine); rdiag_l = acn_l; wa_l = acn_l;
            }
            perm = 0x6543210u;
            bool qr_active = true;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                if (qr_active && (!GEN || j < n)) {
                    // pivot: first position with the largest remaining norm
                    const double cand = (g >= j && g < n) ? rdiag_l : -1.0;
                    double rmax = cand;
#pragma unroll
                    for (int m = G / 2; m >= 1; m >>= 1) rmax = fmax(rmax, __shfl_xor_sync(gmask, rmax, m, G));
                    const unsigned eq = (__ballot_sync(gmask, cand == rmax) >> gbase) & 0x7fu & ~((1u << j) - 1u);
                    const int kmax = eq ? (__ffs(eq) - 1) : j;
                    if (kmax != j) {
#pragma unroll
                        for (int k = j + 1; k < NP; ++k) {
                            if (kmax == k) {
#pragma unroll
                                for (int s = 0; s < S; ++s) { const double t = J[s][j]; J[s][j] = J[s][k]; J[s][k] = t; }
                            }
                        }
                        const unsigned ej = (perm >> (4 * j)) & 15u, ek = (perm >> (4 * kmax)) & 15u;
                        perm &= ~((15u << (4 * j)) | (15u << (4 * kmax)));
                        perm |= (ek << (4 * j)) | (ej << (4 * kmax));
                        const double rj = group_bcast<G>(rdiag_l, j, gmask);
                        const double wj = group_bcast<G>(wa_l, j, gmask);
                        if (g == kmax) { rdiag_l = rj; wa_l = wj; }
                    }
                    double t = (g >= j) ? J[0][j] * J[0][j] : 0.0;
#pragma unroll
                    for (int s = 1; s < S; ++s) t = fma(J[s][j], J[s][j], t);
                    t = group_sum<G>(t, gmask);
                    double ajnorm = sqrt(t);
                    if (ajnorm == 0.0) {
                        qr_active = false;
                    } else {
                        const double ajj0 = group_bcast<G>(J[0][j], j, gmask);
                        if (ajj0 < 0.0) ajnorm = -ajnorm;
                        const double inv = 1.0 / ajnorm;
                        if (g >= j) J[0][j] *= inv;
#pragma unroll
                        for (int s = 1; s < S; ++s) J[s][j] *= inv;
                        if (g == j) J[0][j] += 1.0;
                        const double ajj = group_bcast<G>(J[0][j], j, gmask);
                        if (ajj != 0.0) {
                            double dot_l = 0.0;
#pragma unroll
                            for (int k = j + 1; k < NP; ++k) {
                                double d = (g >= j) ? J[0][k] * J[0][j] : 0.0;
#pragma unroll
                                for (int s = 1; s < S; ++s) d = fma(J[s][k], J[s][j], d);
                                d = group_sum<G>(d, gmask);
                                if (g == k) dot_l = d;
                            }
                            const double t_l = dot_l / ajj;
                            double ajk_l = 0.0;
#pragma unroll
                            for (int k = j + 1; k < NP; ++k) {
                                const double tk = group_bcast<G>(t_l, k, gmask);
                                if (g >= j) J[0][k] = fma(-J[0][j], tk, J[0][k]);
#pragma unroll
                                for (int s = 1; s < S; ++s) J[s][k] = fma(-J[s][j], tk, J[s][k]);
                                const double v = group_bcast<G>(J[0][k], j, gmask);
                                if (g == k) ajk_l = v;
                            }
                            bool redo = false;
                            if (g > j && g < n && rdiag_l != 0.0) {
                                double temp = ajk_l / rdiag_l;
                                rdiag_l = rdiag_l * sqrt(fmax(1.0 - temp * temp, 0.0));
                                temp = rdiag_l / wa_l;
                                redo = (0.05 * temp * temp) <= FSQ_MACHEP;
                            }
                            const unsigned rb = (__ballot_sync(gmask, redo) >> gbase) & 0x7fu;
                            if (rb) {
#pragma unroll
                                for (int k = j + 1; k < NP; ++k) {
                                    if (rb & (1u << k)) {
                                        double d = (g >= j + 1) ? J[0][k] * J[0][k] : 0.0;
#pragma unroll
                                        for (int s = 1; s < S; ++s) d = fma(J[s][k], J[s][k], d);
                                        d = group_sum<G>(d, gmask);
                                        if (g == k) { rdiag_l = sqrt(d); wa_l = rdiag_l; }
                                    }
                                }
                            }
                        }
                        if (g == j) rdiag_l = -ajnorm;
                    }
                }
            }
            // ---- Q^T f, mpfit.py:1114-1124 ----
            {
                double w4[S];
#pragma unroll
                for (int s = 0; s < S; ++s) w4[s] = fv[s];
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const double t3 = group_bcast<G>(J[0][j], j, gmask);
                    if (t3 != 0.0) {
                        double d = (g >= j) ? J[0][j] * w4[0] : 0.0;
#pragma unroll
                        for (int s = 1; s < S; ++s) d = fma(J[s][j], w4[s], d);
                        d = group_sum<G>(d, gmask);
                        const double tq = d / t3;
                        if (g >= j) w4[0] = fma(-J[0][j], tq, w4[0]);
#pragma unroll
                        for (int s = 1; s < S; ++s) w4[s] = fma(-J[s][j], tq, w4[s]);
                    }
                }
                if (g < NP) st.qtf[g] = w4[0];
            }
            // ---- R (pivot order) to shared memory, mpfit.py:1123, 1127-1132 ----
            if (g < NP) {
#pragma unroll
                for (int k = 0; k < NP; ++k) st.r[g][k] = (g == k) ? rdiag_l : ((g < k) ? J[0][k] : 0.0);
                st.acnorm[g] = acn_l;
            }
            __syncwarp(gmask);
            // ---- first-iteration scaling, mpfit.py:1099-1110 ----
            if (niter == 1) {
                double ss = 0.0;
                for (int j = 0; j < n; ++j) {
                    double d = st.acnorm[j];
                    if (d == 0.0) d = 1.0;
                    st.diag[j] = d;
                    const double t = d * st.x[j];
                    ss += t * t;
                }
                xnorm = sqrt(ss);
                delta = factor * xnorm;
                if (delta == 0.0) delta = factor;
            }
            // ---- scaled gradient norm, mpfit.py:1142-1148 ----
            gnorm = 0.0;
            if (fnorm != 0.0) {
                for (int j = 0; j < n; ++j) {
                    const int l = perm_get(perm, j);
                    const double an = st.acnorm[l];
                    if (an != 0.0) {
                        double sum0 = 0.0;
                        for (int i = 0; i <= j; ++i) sum0 += st.r[i][j] * st.qtf[i];
                        sum0 = sum0 / fnorm;
                        gnorm = fmax(gnorm, fabs(sum0 / an));
                    }
                }
            }
            if (gnorm <= gtol) status = 4;                      // :1151
            else if (maxiter == 0) status = 5;                  // :1154
            else {
                for (int j = 0; j < n; ++j) st.diag[j] = (st.diag[j] > st.acnorm[j]) ? st.diag[j] : st.acnorm[j];  // :1160
            }
            need_jac = false;
            __syncwarp(gmask);
        }

        // ================================================================ trial phase
        if (status == 0) {
            par = lmpar(st, perm, delta, par, faithful, n_qrsolv, n);       // :1167
            __syncwarp(gmask);
            double alpha = 1.0;
            for (int j = 0; j < n; ++j) st.step[j] = -st.step[j];           // :1170
            if (qll | qul) {                                                // :1184-1202
                if (lpeg) {
                    double mx = st.step[0];
                    for (int j = 1; j < n; ++j) mx = fmax(mx, st.step[j]);
                    for (int j = 0; j < n; ++j)
                        if ((lpeg >> j) & 1u) st.step[j] = fmin(fmax(st.step[j], 0.0), mx);
                }
                if (upeg) {
                    double mn = st.step[0];
                    for (int j = 1; j < n; ++j) mn = fmin(mn, st.step[j]);
                    for (int j = 0; j < n; ++j)
                        if ((upeg >> j) & 1u) st.step[j] = fmin(fmax(st.step[j], mn), 0.0);
                }
                for (int j = 0; j < n; ++j) {
                    const double sj = st.step[j], xj = st.x[j];
                    if (fabs(sj) > FSQ_MACHEP) {
                        if (((qll >> j) & 1u) && (xj + sj < st.llim[j])) alpha = fmin(alpha, (st.llim[j] - xj) / sj);
                        if (((qul >> j) & 1u) && (xj + sj > st.ulim[j])) alpha = fmin(alpha, (st.ulim[j] - xj) / sj);
                    }
                }
            }
            double pn = 0.0;
            bool nonfinite = false;
            for (int j = 0; j < n; ++j) {                                   // :1215-1234
                const double sj = st.step[j] * alpha;
                st.step[j] = sj;
                double xn = st.x[j] + sj;
                if (qll | qul) {
                    const double ul = st.ulim[j], ll = st.llim[j];
                    const double sgnu = (ul >= 0.0) ? 1.0 : -1.0, sgnl = (ll >= 0.0) ? 1.0 : -1.0;
                    const double ulim1 = ul * (1.0 - sgnu * FSQ_MACHEP) - ((ul == 0.0) ? FSQ_MACHEP : 0.0);
                    const double llim1 = ll * (1.0 + sgnl * FSQ_MACHEP) + ((ll == 0.0) ? FSQ_MACHEP : 0.0);
                    if (((qul >> j) & 1u) && (xn >= ulim1)) xn = ul;
                    if (((qll >> j) & 1u) && (xn <= llim1)) xn = ll;
                }
                st.xnew[j] = xn;
                const double t = st.diag[j] * sj;
                pn += t * t;
                nonfinite |= !(isfinite(sj) && isfinite(xn) && isfinite(st.x[j]));
            }
            const double pnorm = sqrt(pn);
            if (niter == 1) delta = fmin(delta, pnorm);                     // :1237-1238
            __syncwarp(gmask);
            // ---- evaluate at x + p, mpfit.py:1245-1249 ----
            double f1[S], E1[S];
            Rot<S> rot1;
            double pt[NP];
            expand_params<GEN>(st.xnew, n, pmap, pbase, circle, pt);
            bool same_rot;
            if (GEN) { double pc[NP]; expand_params<GEN>(st.x, n, pmap, pbase, circle, pc); same_rot = pt[6] == pc[6]; }
            else same_rot = pt[6] == st.x[6];
            if (same_rot) rot1 = rot; else make_rot<S>(pt[6], px, py, rot1);
            eval_E<S>(pt, rot1, E1);
            {
                double ss = 0.0;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    f1[s] = valid[s] ? __dsub_rn(dat[s], model_of(pt[0], pt[1], E1[s])) : 0.0;
                    if (GEN) f1[s] = __ddiv_rn(f1[s], wgt[GEN ? s : 0]);
                    ss = fma(f1[s], f1[s], ss);
                }
                fnorm1 = sqrt(group_sum<G>(ss, gmask));
            }
            ++nfev;
            // ---- actual / predicted reduction, mpfit.py:1253-1273 ----
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) { const double q = fnorm1 / fnorm; actred = -(q * q) + 1.0; }
            for (int j = 0; j < n; ++j) st.lw2[j] = 0.0;
            for (int j = 0; j < n; ++j) {
                const double sj = st.step[perm_get(perm, j)];
                for (int i = 0; i <= j; ++i) st.lw2[i] = st.lw2[i] + st.r[i][j] * sj;
            }
            double t1 = 0.0;
            for (int j = 0; j < n; ++j) { const double t = alpha * st.lw2[j]; t1 += t * t; }
            const double temp1 = sqrt(t1) / fnorm;
            const double temp2 = (sqrt(alpha * par) * pnorm) / fnorm;
            const double prered = temp1 * temp1 + (temp2 * temp2) / 0.5;
            const double dirder = -(temp1 * temp1 + temp2 * temp2);
            double ratio = 0.0;
            if (prered != 0.0) ratio = actred / prered;
            // ---- trust region update, mpfit.py:1276-1288 ----
            if (ratio <= 0.25) {
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
            const bool accepted = ratio >= 0.0001;                          // :1291-1298
            if (accepted) {
                double ss = 0.0;
                for (int j = 0; j < n; ++j) {
                    const double xn = st.xnew[j];
                    st.x[j] = xn;
                    const double t = st.diag[j] * xn;
                    ss += t * t;
                }
                xnorm = sqrt(ss);
#pragma unroll
                for (int s = 0; s < S; ++s) { fv[s] = f1[s]; E[s] = E1[s]; }
                rot = rot1;
                fnorm = fnorm1;
                ++niter;
            }
            // ---- convergence tests, mpfit.py:1301-1323 ----
            const bool c1 = (fabs(actred) <= ftol) && (prered <= ftol) && (0.5 * ratio <= 1.0);
            if (c1) status = 1;
            if (delta <= xtol * xnorm) status = 2;
            if (c1 && status == 2) status = 3;
            if (status == 0) {
                if (niter >= maxiter) status = 5;
                if ((fabs(actred) <= FSQ_MACHEP) && (prered <= FSQ_MACHEP) && (0.5 * ratio <= 1.0)) status = 6;
                if (delta <= FSQ_MACHEP * xnorm) status = 7;
                if (gnorm <= FSQ_MACHEP) status = 8;
            }
            if (status == 0) {
                if (accepted) need_jac = true;
                else if (nonfinite || !isfinite(ratio)) status = -16;       // :1330-1335
            }
            if (a.trace && idx < a.trace_n && g == 0) {
                const int stepno = nfev - 2 - n * (niter - (accepted ? 1 : 0));    // trial index (0-based)
                const int slot = nfev;   // unique, increasing
                (void)stepno;
                double* tr = a.trace + ((size_t)idx * a.trace_steps) * 20;
                long long cnt = (long long)tr[0];
                if (cnt + 1 < a.trace_steps) {
                    double* rec = tr + (cnt + 1) * 20;
                    rec[0] = niter; rec[1] = accepted ? 1.0 : 0.0; rec[2] = status; rec[3] = fnorm; rec[4] = fnorm1;
                    rec[5] = delta; rec[6] = par; rec[7] = ratio; rec[8] = alpha; rec[9] = pnorm;
                    for (int j = 0; j < NP; ++j) rec[10 + j] = st.xnew[j];
                    rec[17] = n_qrsolv; rec[18] = actred; rec[19] = prered;
                    tr[0] = (double)(cnt + 1);
                    (void)slot;
                }
            }
            __syncwarp(gmask);
        }

        // ================================================================ finalize
        if (status != 0) {
            if (status > 0) ++nfev;                                         // :1351-1355 final call
            const double fn = fmax(fnorm, fnorm1);                          // :1357-1359
            const double chi2 = fn * fn;
            double pf[NP];
            expand_params<GEN>(st.x, n, pmap, pbase, circle, pf);
            double gimg[S];
#pragma unroll
            for (int s = 0; s < S; ++s) gimg[s] = model_of(pf[0], pf[1], E[s]);   // gaussfitter.py:253
            const bool want_cv = !PFLIB && (a.covar != nullptr);
            const bool want_pe = ((a.o.want_perror != 0) && (PFLIB ? false : (a.perror != nullptr))) || want_cv;
            if (want_pe) {
                if (status > 0) covar_perror(st, perm, st.lw1, n);
                else for (int j = 0; j < NP; ++j) st.lw1[j] = __longlong_as_double(0x7ff8000000000000LL);   // perror is None there
                __syncwarp(gmask);
            }
            if (PFLIB) {
                // ---- pflib.py:461-473 ----
                // in the reference's own order: Python's sum() adds the unfused squares one after another in raster
                // order (:463-465, :470-472), illumina_s_n runs in numpy's arithmetic (fsq_common.cuh) -- pixel q of
                // the window sits in slot q / G of lane q % G, so every lane replays the 25 terms through shuffles
                long long isum = 0;
#pragma unroll
                for (int s = 0; s < S; ++s) if (valid[s]) isum += idat[s];
#pragma unroll
                for (int m = G / 2; m >= 1; m >>= 1) isum += __shfl_xor_sync(gmask, isum, m, G);
                const double mean = __ddiv_rn((double)isum, 25.0);
                double d2[S], m2[S];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const double d = __dsub_rn(dat[s], gimg[s]), m_ = __dsub_rn(dat[s], mean);
                    d2[s] = __dmul_rn(d, d);
                    m2[s] = __dmul_rn(m_, m_);
                }
                double ssr = 0.0, sst = 0.0;
                int win[25];
#pragma unroll
                for (int q = 0; q < 25; ++q) {
                    ssr = __dadd_rn(ssr, __shfl_sync(gmask, d2[q / G], q % G, G));
                    sst = __dadd_rn(sst, __shfl_sync(gmask, m2[q / G], q % G, G));
                    win[q] = __shfl_sync(gmask, idat[q / G], q % G, G);
                }
                const double r_2 = __dsub_rn(1.0, __ddiv_rn(ssr, sst));
                const double rmse = sqrt(__ddiv_rn(ssr, 25.0));
                const double s_n = illumina_sn<5>(5, [&](int r, int c) { return (long long)win[r * 5 + c]; });
                if (g == 0) {
                    double* o = a.out_fit + idx * 12;
                    o[0] = (pf[2] + (double)cand_h) - 2.5;                   // pflib.py:461
                    o[1] = (pf[3] + (double)cand_w) - 2.5;
                    o[2] = pf[0]; o[3] = pf[1]; o[4] = pf[4]; o[5] = pf[5]; o[6] = pf[6];
                    o[7] = rmse; o[8] = r_2; o[9] = s_n; o[10] = chi2; o[11] = fnorm;
                    int* oi = a.out_int + idx * 4;
                    oi[0] = status; oi[1] = niter; oi[2] = nfev; oi[3] = n_qrsolv;
                }
                if (a.fit_img) {
#pragma unroll
                    for (int s = 0; s < S; ++s) if (valid[s]) a.fit_img[idx * 25 + s * G + g] = gimg[s];
                }
            } else {
                if (!GEN) {
                    if (g < NP) {
                        a.params[idx * NP + g] = st.x[g];
                        if (a.perror) a.perror[idx * NP + g] = want_pe ? st.lw1[g] : 0.0;
                    }
                } else if (g == 0) {
                    // mpfit.py:1346-1348, :1375-1386: full-length vectors, zeros where a parameter is fixed
                    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
                    for (int j = 0; j < NP; ++j) a.params[idx * NP + j] = pf[j];
                    if (a.perror) {
                        for (int j = 0; j < NP; ++j) a.perror[idx * NP + j] = (want_pe && status <= 0) ? nanv : 0.0;
                        if (want_pe && status > 0)
                            for (int k = 0; k < n; ++k) a.perror[idx * NP + pidx<true>(pmap, k)] = st.lw1[k];
                    }
                }
                if (want_cv && g == 0) {
                    // mpfit.py:1375-1379: covar[ifree, ifree[i]] = cv[:, i]; None (NaN here) when the fit did not converge
                    double* cv = a.covar + idx * NP * NP;
                    const double fill = (status > 0) ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
                    for (int q = 0; q < NP * NP; ++q) cv[q] = fill;
                    if (status > 0)
                        for (int i = 0; i < n; ++i)
                            for (int k = 0; k < n; ++k) cv[pidx<GEN>(pmap, i) * NP + pidx<GEN>(pmap, k)] = st.r[i][k];
                }
                if (g == 0) {
                    a.status[idx] = status; a.niter[idx] = niter; a.nfev[idx] = nfev; a.chi2[idx] = chi2;
                    if (a.n_qrsolv) a.n_qrsolv[idx] = n_qrsolv;
                }
                if (a.fit_img) {
#pragma unroll
                    for (int s = 0; s < S; ++s) if (valid[s]) a.fit_img[(size_t)idx * P + s * G + g] = gimg[s];
                }
            }
            have = false;
            __syncwarp(gmask);
        }
    }
}

// ---- FMA peak micro-benchmark (roofline denominator) --------------------------------------
static int launch_lm(const LmArgs& a, int G, bool pflib, cudaStream_t st) {
    const int groups_per_block = LM_THREADS / G;
    const size_t smem = sizeof(FitShared) * groups_per_block;
    int blocks_per_sm = 4;
    const int grid = sm_count() * blocks_per_sm;
    FSQ_CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st));
    const bool gen = (a.fixed != nullptr) || (a.err != nullptr) || (a.circle != 0);
    if (pflib) {
        lmfit_kernel<8, true><<<grid, LM_THREADS, smem, st>>>(a);
    } else if (G == 8) {
        if (gen) lmfit_kernel<8, false, true><<<grid, LM_THREADS, smem, st>>>(a);
        else lmfit_kernel<8, false><<<grid, LM_THREADS, smem, st>>>(a);
    } else {
        if (gen) lmfit_kernel<32, false, true><<<grid, LM_THREADS, smem, st>>>(a);
        else lmfit_kernel<32, false><<<grid, LM_THREADS, smem, st>>>(a);
    }
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}

}  // namespace fsq

namespace fsq {
// fsq_lmwarp.cu
int warp_fit_candidates(const void* frames, int dtype_code, int H, int W, const int32_t* cand_hw,
                        const int32_t* cand_frame, long long n, const long long* n_dev, const fsq_lm_opts* opts,
                        double* out_fit, int32_t* out_int, double* fit_img, void* scratch, cudaStream_t st);
long long warp_scratch_bytes(long long n);
int warp_gaussfit_batch(const void* windows, int dtype_code, long long n, int win, const double* p0, const double* lo,
                        const double* hi, const uint8_t* lim_lo, const uint8_t* lim_hi, const fsq_lm_opts* opts,
                        double* params, int32_t* status, int32_t* niter, int32_t* nfev, double* chi2,
                        int32_t* n_damped, double* fit_img, cudaStream_t st);
// fsq_lmfast.cu
int fast_fit_candidates(const void* frames, int dtype_code, int H, int W, const int32_t* cand_hw,
                        const int32_t* cand_frame, long long n, const long long* n_dev, const fsq_lm_opts* opts,
                        double* out_fit, int32_t* out_int, double* fit_img, unsigned long long* work_counter,
                        cudaStream_t st);
int fast_gaussfit_batch(const void* windows, int dtype_code, long long n, const double* p0, const double* lo,
                        const double* hi, const uint8_t* lim_lo, const uint8_t* lim_hi, const fsq_lm_opts* opts,
                        double* params, int32_t* status, int32_t* niter, int32_t* nfev, double* chi2,
                        int32_t* n_damped, double* fit_img, unsigned long long* work_counter, cudaStream_t st);
}  // namespace fsq

using namespace fsq;

extern "C" void fsq_lm_default_opts(fsq_lm_opts* o) {
    if (!o) return;
    o->ftol = 1e-10; o->xtol = 1e-10; o->gtol = 1e-10; o->factor = 100.0; o->maxiter = 200;   // mpfit.py:600-605
    o->faithful = 1; o->want_perror = 0; o->solver = FSQ_SOLVER_MINPACK; o->park_after = 0; o->warps_per_sm = 0;
}

static int check_opts(const fsq_lm_opts* o, const char* who) {
    if (!o) { set_error("%s: opts is NULL", who); return FSQ_E_ARG; }
    // mpfit.py:986-989 "input keywords are inconsistent"
    if (!(o->ftol > 0) || !(o->xtol > 0) || !(o->gtol > 0) || o->maxiter < 0 || !(o->factor > 0)) {
        set_error("%s: input keywords are inconsistent (ftol/xtol/gtol/factor must be > 0, maxiter >= 0)", who);
        return FSQ_E_ARG;
    }
    if (o->solver < FSQ_SOLVER_MINPACK || o->solver > FSQ_SOLVER_FAST) {
        set_error("%s: unknown solver %d", who, o->solver);
        return FSQ_E_ARG;
    }
    return FSQ_OK;
}

extern "C" int fsq_gaussfit_batch(const void* windows, int dtype_code, int64_t n, int win,
                                  const double* p0, const double* lo, const double* hi,
                                  const uint8_t* lim_lo, const uint8_t* lim_hi, const fsq_lm_opts* opts,
                                  double* params, double* perror, int32_t* status, int32_t* niter,
                                  int32_t* nfev, double* chi2, int32_t* n_qrsolv, double* fit_img,
                                  int64_t* work_counter, void* stream) {
    return fsq_gaussfit_batch_ex(windows, dtype_code, n, win, p0, lo, hi, lim_lo, lim_hi, nullptr, nullptr, 0, opts, params,
                                 perror, nullptr, status, niter, nfev, chi2, n_qrsolv, fit_img, work_counter, stream);
}

extern "C" int fsq_gaussfit_batch_ex(const void* windows, int dtype_code, int64_t n, int win,
                                     const double* p0, const double* lo, const double* hi,
                                     const uint8_t* lim_lo, const uint8_t* lim_hi, const uint8_t* fixed,
                                     const double* err, int circle, const fsq_lm_opts* opts,
                                     double* params, double* perror, double* covar, int32_t* status, int32_t* niter,
                                     int32_t* nfev, double* chi2, int32_t* n_qrsolv, double* fit_img,
                                     int64_t* work_counter, void* stream) {
    int rc = check_opts(opts, "fsq_gaussfit_batch");
    if (rc) return rc;
    if (n < 0) { set_error("fsq_gaussfit_batch: n < 0"); return FSQ_E_ARG; }
    if (n == 0) return FSQ_OK;
    if (!windows || !p0 || !lo || !hi || !lim_lo || !lim_hi || !params || !status || !niter || !nfev || !chi2 || !work_counter) {
        set_error("fsq_gaussfit_batch: NULL pointer argument");
        return FSQ_E_ARG;
    }
    if (win < 3 || win > 11) {
        set_error("fsq_gaussfit_batch: window side must be in 3..11 (got %d); mpfit needs m >= n = 7", win);
        return FSQ_E_ARG;
    }
    if (dtype_code != FSQ_F64 && dtype_code != FSQ_I64 && dtype_code != FSQ_U16 && dtype_code != FSQ_I32 &&
        dtype_code != FSQ_U8 && dtype_code != FSQ_I16) {
        set_error("fsq_gaussfit_batch: unsupported dtype code %d", dtype_code);
        return FSQ_E_ARG;
    }
    if (opts->want_perror && !perror) { set_error("fsq_gaussfit_batch: want_perror set but perror is NULL"); return FSQ_E_ARG; }
    if ((fixed || err || circle || covar) && opts->solver != FSQ_SOLVER_MINPACK) {
        set_error("fsq_gaussfit_batch_ex: fixed parameters, err weights, the circular model and covar need FSQ_SOLVER_MINPACK");
        return FSQ_E_ARG;
    }
    if (win * win < 7) { set_error("fsq_gaussfit_batch: number of parameters must not exceed data"); return FSQ_E_ARG; }
    if (opts->solver == FSQ_SOLVER_FAST) {
        if (win != 5 && win != 11) { set_error("fsq_gaussfit_batch: FSQ_SOLVER_FAST takes 5x5 or 11x11 windows (got %d); use FSQ_SOLVER_MINPACK", win); return FSQ_E_ARG; }
        if (opts->want_perror) { set_error("fsq_gaussfit_batch: want_perror needs FSQ_SOLVER_MINPACK"); return FSQ_E_ARG; }
        return warp_gaussfit_batch(windows, dtype_code, n, win, p0, lo, hi, lim_lo, lim_hi, opts, params, status, niter, nfev,
                                   chi2, n_qrsolv, fit_img, (cudaStream_t)stream);
    }
    if (opts->solver != FSQ_SOLVER_MINPACK) {
        if (win != 5) { set_error("fsq_gaussfit_batch: the FAST solvers take 5x5 windows (got %d); use FSQ_SOLVER_MINPACK", win); return FSQ_E_ARG; }
        if (opts->want_perror) { set_error("fsq_gaussfit_batch: want_perror needs FSQ_SOLVER_MINPACK"); return FSQ_E_ARG; }
        return fast_gaussfit_batch(windows, dtype_code, n, p0, lo, hi, lim_lo, lim_hi, opts, params, status, niter, nfev,
                                   chi2, n_qrsolv, fit_img, (unsigned long long*)work_counter, (cudaStream_t)stream);
    }
    LmArgs a;
    memset(&a, 0, sizeof(a));
    a.windows = windows; a.wdtype = dtype_code; a.win = win;
    a.p0 = p0; a.lo = lo; a.hi = hi; a.lim_lo = lim_lo; a.lim_hi = lim_hi;
    a.n = n; a.n_dev = nullptr; a.o = *opts;
    a.params = params; a.perror = perror; a.status = status; a.niter = niter; a.nfev = nfev;
    a.chi2 = chi2; a.n_qrsolv = n_qrsolv; a.fit_img = fit_img;
    a.fixed = fixed; a.err = err; a.circle = circle ? 1 : 0; a.covar = covar;
    a.work_counter = (unsigned long long*)work_counter;
    const int G = (win * win <= 8 * SLOTS) ? 8 : 32;
    return launch_lm(a, G, false, (cudaStream_t)stream);
}

extern "C" int fsq_gaussfit_batch_trace(const void* windows, int dtype_code, int64_t n, int win,
                                        const double* p0, const double* lo, const double* hi,
                                        const uint8_t* lim_lo, const uint8_t* lim_hi, const fsq_lm_opts* opts,
                                        double* params, int32_t* status, int32_t* niter, int32_t* nfev,
                                        double* chi2, int32_t* n_qrsolv, double* trace, int trace_steps,
                                        int64_t trace_n, int64_t* work_counter, void* stream) {
    int rc = check_opts(opts, "fsq_gaussfit_batch_trace");
    if (rc) return rc;
    if (n <= 0 || !windows || !p0 || !lo || !hi || !lim_lo || !lim_hi || !params || !status || !niter || !nfev ||
        !chi2 || !work_counter || !trace || trace_steps < 2 || win < 3 || win > 11) {
        set_error("fsq_gaussfit_batch_trace: bad argument");
        return FSQ_E_ARG;
    }
    LmArgs a;
    memset(&a, 0, sizeof(a));
    a.windows = windows; a.wdtype = dtype_code; a.win = win;
    a.p0 = p0; a.lo = lo; a.hi = hi; a.lim_lo = lim_lo; a.lim_hi = lim_hi;
    a.n = n; a.o = *opts;
    a.params = params; a.status = status; a.niter = niter; a.nfev = nfev; a.chi2 = chi2; a.n_qrsolv = n_qrsolv;
    a.work_counter = (unsigned long long*)work_counter;
    a.trace = trace; a.trace_steps = trace_steps; a.trace_n = trace_n;
    const int G = (win * win <= 8 * SLOTS) ? 8 : 32;
    return launch_lm(a, G, false, (cudaStream_t)stream);
}

extern "C" int64_t fsq_fit_scratch_bytes(int64_t n) { return (int64_t)warp_scratch_bytes(n); }

extern "C" int fsq_fit_candidates(const void* frames, int dtype_code, int n_frames, int H, int W,
                                  const int32_t* cand_hw, const int32_t* cand_frame, int64_t n,
                                  const int64_t* n_dev, const fsq_lm_opts* opts, double* out_fit,
                                  int32_t* out_int, double* fit_img, void* scratch, int64_t scratch_bytes, void* stream) {
    int rc = check_opts(opts, "fsq_fit_candidates");
    if (rc) return rc;
    if (n < 0) { set_error("fsq_fit_candidates: n < 0"); return FSQ_E_ARG; }
    if (n == 0) return FSQ_OK;
    if (!frames || !cand_hw || !cand_frame || !out_fit || !out_int || !scratch) {
        set_error("fsq_fit_candidates: NULL pointer argument");
        return FSQ_E_ARG;
    }
    if (scratch_bytes < fsq_fit_scratch_bytes(n)) {
        set_error("fsq_fit_candidates: scratch of %lld bytes is smaller than fsq_fit_scratch_bytes(%lld) = %lld",
                  (long long)scratch_bytes, (long long)n, (long long)fsq_fit_scratch_bytes(n));
        return FSQ_E_CAPACITY;
    }
    int64_t* work_counter = (int64_t*)scratch;
    if (n_frames <= 0 || H < 5 || W < 5) { set_error("fsq_fit_candidates: bad frame shape"); return FSQ_E_ARG; }
    if (dtype_code != FSQ_U8 && dtype_code != FSQ_U16 && dtype_code != FSQ_I16 && dtype_code != FSQ_I32) {
        set_error("fsq_fit_candidates: unsupported frame dtype code %d", dtype_code);
        return FSQ_E_ARG;
    }
    if (opts->solver == FSQ_SOLVER_FAST && (H > 65535 || W > 65535)) {
        set_error("fsq_fit_candidates: frames larger than 65535 pixels per side are not supported by FSQ_SOLVER_FAST");
        return FSQ_E_ARG;
    }
    if (opts->solver == FSQ_SOLVER_FAST)
        return warp_fit_candidates(frames, dtype_code, H, W, cand_hw, cand_frame, n, (const long long*)n_dev, opts, out_fit,
                                   out_int, fit_img, scratch, (cudaStream_t)stream);
    if (opts->solver != FSQ_SOLVER_MINPACK)
        return fast_fit_candidates(frames, dtype_code, H, W, cand_hw, cand_frame, n, (const long long*)n_dev, opts, out_fit,
                                   out_int, fit_img, (unsigned long long*)work_counter, (cudaStream_t)stream);
    LmArgs a;
    memset(&a, 0, sizeof(a));
    a.frames = frames; a.fdtype = dtype_code; a.H = H; a.W = W;
    a.cand_hw = cand_hw; a.cand_frame = cand_frame; a.n = n; a.n_dev = (const long long*)n_dev; a.o = *opts;
    a.out_fit = out_fit; a.out_int = out_int; a.fit_img = fit_img;
    a.work_counter = (unsigned long long*)work_counter;
    return launch_lm(a, 8, true, (cudaStream_t)stream);
}

