// Frame registration: phase_correlate (phase_correlate.py:11-196, a port of Guizar-Sicairos' efficient sub-pixel
// registration; caller SequenceExperiment.offsets_from_frames, flexlibrary.py:1717-1741) for a batch of image pairs.
//
//   F_ref, F_reg = fft2(ref), fft2(reg)                         cuFFT Z2Z (the two library calls SURVEY.md 8(f) allows)
//   cc = ifft2(F_ref conj(F_reg)); whole-pixel peak = argmax    pc_cross_kernel, cuFFT inverse, pc_argmax_kernel
//   upsample_factor > 1: the upsampled DFT of F_reg conj(F_ref) on a ceil(1.5 usf)^2 grid around the peak by two
//   matrix products with twiddle tables (phase_correlate.py:136-196)   pc_twiddle / pc_dft_rows / pc_dft_cols kernels
//   argmax again, error and phase from the peak value            pc_finish_kernel
//
// Everything is FP64 / complex128 like the numpy reference.  One launch sequence serves all pairs of a stack.
#include "fsq_common.cuh"
#include <cufft.h>
#include <map>
#include <mutex>
#include <tuple>

namespace fsq {

typedef double2 cplx;

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cmul_conj(cplx a, cplx b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a conj(b)

__device__ __forceinline__ double pc_load(const void* base, int dtype, size_t off) {
    switch (dtype) {
        case FSQ_U8:  return (double)((const uint8_t*)base)[off];
        case FSQ_U16: return (double)((const uint16_t*)base)[off];
        case FSQ_I16: return (double)((const int16_t*)base)[off];
        case FSQ_I32: return (double)((const int32_t*)base)[off];
        case FSQ_I64: return (double)((const long long*)base)[off];
        default:      return ((const double*)base)[off];
    }
}

// numpy.fft frequency index of position k: ifftshift(arange(n))[k] - floor(n / 2)   (phase_correlate.py:184-186, 190-192)
__device__ __forceinline__ double pc_freq(int k, int n) { return (double)(((k + n / 2) % n) - n / 2); }

struct PcScratch {
    cplx* fref; cplx* freg; cplx* cc;      // [B, rows, cols] each
    cplx* rowk;                            // [B, U, rows] exp(-i 2 pi / (rows usf) (u - row_off) freq(r))
    cplx* colk;                            // [B, cols, U]
    cplx* t;                               // [B, U, cols]
    cplx* up;                              // [B, U, U]
    double* power;                         // [B, nblk, 2]  per-block partial sums of |F_ref|^2, |F_reg|^2 (added in block order: deterministic)
    double* peak;                          // [B, 4]  whole-pixel peak: flat index, re, im, unused
};

static int64_t pc_carve(PcScratch* s, char* base, long long B, int rows, int cols, int U) {
    int64_t off = 0;
    auto take = [&](int64_t bytes) { char* p = base ? base + off : nullptr; off += (bytes + 255) & ~(int64_t)255; return p; };
    const int64_t plane = (int64_t)rows * cols * sizeof(cplx);
    cplx* fref = (cplx*)take(B * plane);
    cplx* freg = (cplx*)take(B * plane);
    cplx* cc = (cplx*)take(B * plane);
    cplx* rowk = (cplx*)take(B * (int64_t)U * rows * sizeof(cplx));
    cplx* colk = (cplx*)take(B * (int64_t)U * cols * sizeof(cplx));
    cplx* t = (cplx*)take(B * (int64_t)U * cols * sizeof(cplx));
    cplx* up = (cplx*)take(B * (int64_t)U * U * sizeof(cplx));
    const int64_t nblk = ((int64_t)rows * cols + 255) / 256;
    double* power = (double*)take(B * nblk * 2 * sizeof(double));
    double* peak = (double*)take(B * 4 * sizeof(double));
    if (s) { s->fref = fref; s->freg = freg; s->cc = cc; s->rowk = rowk; s->colk = colk; s->t = t; s->up = up; s->power = power; s->peak = peak; }
    return off;
}

__global__ void __launch_bounds__(256)
pc_to_complex_kernel(const void* __restrict__ ref, const void* __restrict__ reg, int dtype, long long n, cplx* __restrict__ fref,
                     cplx* __restrict__ freg) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    fref[i] = make_double2(pc_load(ref, dtype, (size_t)i), 0.0);
    freg[i] = make_double2(pc_load(reg, dtype, (size_t)i), 0.0);
}

// cc = F_ref conj(F_reg) (phase_correlate.py:72) and the two power sums (:87-88, :117-121)
__global__ void __launch_bounds__(256)
pc_cross_kernel(const cplx* __restrict__ fref, const cplx* __restrict__ freg, cplx* __restrict__ cc, long long plane,
                double* __restrict__ power) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    double pr = 0.0, pg = 0.0;
    if (i < plane) {
        const cplx a = fref[b * plane + i], g = freg[b * plane + i];
        cc[b * plane + i] = cmul_conj(a, g);
        pr = a.x * a.x + a.y * a.y;
        pg = g.x * g.x + g.y * g.y;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { pr += __shfl_xor_sync(0xffffffffu, pr, m); pg += __shfl_xor_sync(0xffffffffu, pg, m); }
    __shared__ double sr[8], sg[8];
    if ((threadIdx.x & 31) == 0) { sr[threadIdx.x >> 5] = pr; sg[threadIdx.x >> 5] = pg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, g = 0.0;
        for (int k = 0; k < 8; ++k) { a += sr[k]; g += sg[k]; }
        power[2 * ((size_t)b * gridDim.x + blockIdx.x)] = a;
        power[2 * ((size_t)b * gridDim.x + blockIdx.x) + 1] = g;
    }
}

// numpy.argmax of a complex array: lexicographic (real, imag), first occurrence.  One block per pair.
// `scale` multiplies the values before they are compared / returned (1 / (rows cols) for cuFFT's unnormalised inverse).
__global__ void __launch_bounds__(256)
pc_argmax_kernel(const cplx* __restrict__ v, long long n, double scale, int conj, double* __restrict__ peak) {
    const int b = blockIdx.x;
    const cplx* p = v + (long long)b * n;
    double bre = 0.0, bim = 0.0;
    long long bi = -1;
    for (long long i = threadIdx.x; i < n; i += 256) {
        const double re = p[i].x * scale, im = (conj ? -p[i].y : p[i].y) * scale;
        const bool better = (bi < 0) || (re > bre) || (re == bre && im > bim);       // i increases: ties keep the earlier one
        if (better) { bre = re; bim = im; bi = i; }
    }
    __shared__ double s_re[256], s_im[256];
    __shared__ long long s_i[256];
    s_re[threadIdx.x] = bre; s_im[threadIdx.x] = bim; s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int m = 128; m >= 1; m >>= 1) {
        if (threadIdx.x < m) {
            const int o = threadIdx.x + m;
            const long long io = s_i[o], im_ = s_i[threadIdx.x];
            bool take = false;
            if (io >= 0) {
                if (im_ < 0) take = true;
                else if (s_re[o] > s_re[threadIdx.x]) take = true;
                else if (s_re[o] == s_re[threadIdx.x]) {
                    if (s_im[o] > s_im[threadIdx.x]) take = true;
                    else if (s_im[o] == s_im[threadIdx.x] && io < im_) take = true;
                }
            }
            if (take) { s_re[threadIdx.x] = s_re[o]; s_im[threadIdx.x] = s_im[o]; s_i[threadIdx.x] = io; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { peak[4 * b] = (double)s_i[0]; peak[4 * b + 1] = s_re[0]; peak[4 * b + 2] = s_im[0]; peak[4 * b + 3] = 0.0; }
}

// whole-pixel shifts (phase_correlate.py:73-85) from the peak index
__device__ __forceinline__ void pc_whole_shifts(double flat, int rows, int cols, double* rs, double* cs) {
    const long long f = (long long)flat;
    const int rmax = (int)(f / cols), cmax = (int)(f % cols);
    const int mid_r = rows / 2, mid_c = cols / 2;                       // numpy.fix(n / 2) for n >= 0
    *rs = (double)(rmax > mid_r ? rmax - rows : rmax);
    *cs = (double)(cmax > mid_c ? cmax - cols : cmax);
}

// twiddle tables of the upsampled DFT around the whole-pixel peak (phase_correlate.py:97-107, 180-194)
__global__ void __launch_bounds__(256)
pc_twiddle_kernel(const double* __restrict__ peak, int rows, int cols, int U, double usf, cplx* __restrict__ rowk,
                  cplx* __restrict__ colk) {
    const int b = blockIdx.y;
    double rs, cs;
    pc_whole_shifts(peak[4 * b], rows, cols, &rs, &cs);
    rs = rint(rs * usf) / usf; cs = rint(cs * usf) / usf;               // :97-98
    const double dftshift = (double)(U / 2);                            // fix(U / 2)
    const double row_off = dftshift - rs * usf, col_off = dftshift - cs * usf;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const double two_pi = 6.283185307179586;
    if (i < U * rows) {
        const int u = i / rows, r = i - u * rows;
        const double arg = (-two_pi / ((double)rows * usf)) * (((double)u - row_off) * pc_freq(r, rows));
        double s, c;
        sincos(arg, &s, &c);
        rowk[((size_t)b * U + u) * rows + r] = make_double2(c, s);
    }
    if (i < U * cols) {
        const int c_ = i / U, v = i - c_ * U;
        const double arg = (-two_pi / ((double)cols * usf)) * (pc_freq(c_, cols) * ((double)v - col_off));
        double s, c;
        sincos(arg, &s, &c);
        colk[((size_t)b * cols + c_) * U + v] = make_double2(c, s);
    }
}

// t[b, u, c] = sum_r rowk[b, u, r] * (F_reg conj(F_ref))[b, r, c]; block = 32 columns x 8 u-lanes, rows in tiles of 32
__global__ void __launch_bounds__(256)
pc_dft_rows_kernel(const cplx* __restrict__ fref, const cplx* __restrict__ freg, const cplx* __restrict__ rowk, int rows, int cols,
                   int U, cplx* __restrict__ t) {
    const int b = blockIdx.z, c0 = blockIdx.x * 32, u0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    __shared__ cplx s_data[32][33];
    __shared__ cplx s_k[32][33];
    cplx acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_double2(0.0, 0.0);
    const size_t plane = (size_t)rows * cols;
    for (int r0 = 0; r0 < rows; r0 += 32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rr = ty + 8 * k;
            const int r = r0 + rr, c = c0 + tx;
            cplx d = make_double2(0.0, 0.0);
            if (r < rows && c < cols) d = cmul_conj(freg[b * plane + (size_t)r * cols + c], fref[b * plane + (size_t)r * cols + c]);
            s_data[rr][tx] = d;
            const int u = u0 + rr, rk = r0 + tx;
            s_k[rr][tx] = (u < U && rk < rows) ? rowk[((size_t)b * U + u) * rows + rk] : make_double2(0.0, 0.0);
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
            const cplx d = s_data[rr][tx];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const cplx w = s_k[ty + 8 * k][rr];
                acc[k].x += w.x * d.x - w.y * d.y;
                acc[k].y += w.x * d.y + w.y * d.x;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int u = u0 + ty + 8 * k, c = c0 + tx;
        if (u < U && c < cols) t[((size_t)b * U + u) * cols + c] = acc[k];
    }
}

// up[b, u, v] = sum_c t[b, u, c] * colk[b, c, v]
__global__ void __launch_bounds__(256)
pc_dft_cols_kernel(const cplx* __restrict__ t, const cplx* __restrict__ colk, int cols, int U, cplx* __restrict__ up) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= U * U) return;
    const int u = i / U, v = i - u * U;
    const cplx* tr = t + ((size_t)b * U + u) * cols;
    const cplx* ck = colk + (size_t)b * cols * U + v;
    cplx acc = make_double2(0.0, 0.0);
    for (int c = 0; c < cols; ++c) {
        const cplx a = tr[c], w = ck[(size_t)c * U];
        acc.x += a.x * w.x - a.y * w.y;
        acc.y += a.x * w.y + a.y * w.x;
    }
    up[(size_t)b * U * U + i] = acc;
}

// shifts, error, phase (phase_correlate.py:86-93 for usf == 1, :108-131 otherwise)
__global__ void __launch_bounds__(64)
pc_finish_kernel(const double* __restrict__ peak, const double* __restrict__ peak_up, const double* __restrict__ power_part, int nblk,
                 int B, int rows, int cols, int U, double usf, double* __restrict__ out) {
    const int b = blockIdx.x * 64 + threadIdx.x;
    if (b >= B) return;
    double power[2] = {0.0, 0.0};
    for (int k = 0; k < nblk; ++k) { power[0] += power_part[2 * ((size_t)b * nblk + k)]; power[1] += power_part[2 * ((size_t)b * nblk + k) + 1]; }
    double rs, cs;
    pc_whole_shifts(peak[4 * b], rows, cols, &rs, &cs);
    const int mid_r = rows / 2, mid_c = cols / 2;
    double re, im, rg, rf;
    if (usf == 1.0) {
        re = peak[4 * b + 1]; im = peak[4 * b + 2];
        rf = power[0] / ((double)rows * cols);                           // rfzero (ref), rgzero (reg)
        rg = power[1] / ((double)rows * cols);
    } else {
        rs = rint(rs * usf) / usf; cs = rint(cs * usf) / usf;
        const double dftshift = (double)(U / 2);
        const long long f = (long long)peak_up[4 * b];
        rs += ((double)(f / U) - dftshift) / usf;                        // :113-116
        cs += ((double)(f % U) - dftshift) / usf;
        re = peak_up[4 * b + 1]; im = peak_up[4 * b + 2];
        const double norm = (double)mid_r * (double)mid_c * usf * usf;
        rg = power[0] / norm;                                            // rg00 is built from ref_image_freq (:117-118)
        rf = power[1] / norm;
    }
    const double err = sqrt(fabs(1.0 - (re * re + im * im) / (rg * rf)));
    if (usf != 1.0) {                                                    // :128-131 (only reached on the upsampled path)
        if (mid_r == 1) rs = 0.0;
        if (mid_c == 1) cs = 0.0;
    }
    out[4 * b] = rs; out[4 * b + 1] = cs; out[4 * b + 2] = err; out[4 * b + 3] = atan2(im, re);
}

// cuFFT plans are the one thing this library keeps between calls: one Z2Z plan per (device, rows, cols, batch), created
// on first use, used under a lock (a plan carries its stream)
static std::mutex g_plan_mutex;
static std::map<std::tuple<int, int, int, int>, cufftHandle> g_plans;

static int pc_plan(int rows, int cols, int B, cufftHandle* out) {
    int dev = 0;
    FSQ_CUDA_CHECK(cudaGetDevice(&dev));
    const auto key = std::make_tuple(dev, rows, cols, B);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) { *out = it->second; return FSQ_OK; }
    cufftHandle h;
    int n[2] = {rows, cols};
    const cufftResult r = cufftPlanMany(&h, 2, n, nullptr, 1, rows * cols, nullptr, 1, rows * cols, CUFFT_Z2Z, B);
    if (r != CUFFT_SUCCESS) { set_error("cufftPlanMany(%d x %d, batch %d) failed with code %d", rows, cols, B, (int)r); return FSQ_E_CUDA; }
    g_plans[key] = h;
    *out = h;
    return FSQ_OK;
}

}  // namespace fsq

using namespace fsq;

static int pc_upsampled(int usf) { return (int)ceil((double)usf * 1.5); }

extern "C" int64_t fsq_phase_correlate_scratch_bytes(int n_pairs, int rows, int cols, int upsample_factor) {
    if (n_pairs < 0 || rows < 1 || cols < 1 || upsample_factor < 1) return 0;
    return pc_carve(nullptr, nullptr, n_pairs, rows, cols, upsample_factor > 1 ? pc_upsampled(upsample_factor) : 1);
}

extern "C" int fsq_phase_correlate(const void* ref, const void* reg, int dtype_code, int n_pairs, int rows, int cols,
                                   int upsample_factor, double* out, void* scratch, int64_t scratch_bytes, void* stream) {
    if (n_pairs < 0 || rows < 1 || cols < 1 || upsample_factor < 1) {
        set_error("fsq_phase_correlate: n_pairs >= 0, rows, cols, upsample_factor >= 1 required");
        return FSQ_E_ARG;
    }
    if (n_pairs == 0) return FSQ_OK;
    if (!ref || !reg || !out || !scratch) { set_error("fsq_phase_correlate: NULL pointer argument"); return FSQ_E_ARG; }
    if (dtype_code < FSQ_U8 || dtype_code > FSQ_I64) { set_error("fsq_phase_correlate: unsupported dtype code %d", dtype_code); return FSQ_E_ARG; }
    if (n_pairs > 65535) { set_error("fsq_phase_correlate: at most 65535 pairs per call"); return FSQ_E_ARG; }
    const int U = upsample_factor > 1 ? pc_upsampled(upsample_factor) : 1;
    const int64_t need = pc_carve(nullptr, nullptr, n_pairs, rows, cols, U);
    if (scratch_bytes < need) {
        set_error("fsq_phase_correlate: scratch of %lld bytes is smaller than fsq_phase_correlate_scratch_bytes = %lld",
                  (long long)scratch_bytes, (long long)need);
        return FSQ_E_CAPACITY;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PcScratch s;
    pc_carve(&s, (char*)scratch, n_pairs, rows, cols, U);
    const long long plane = (long long)rows * cols, total = plane * n_pairs;
    const int B = n_pairs;
    pc_to_complex_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ref, reg, dtype_code, total, s.fref, s.freg);
    FSQ_LAUNCH_CHECK();
    {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        cufftHandle plan;
        const int rc = pc_plan(rows, cols, B, &plan);
        if (rc != FSQ_OK) return rc;
        cufftResult r = cufftSetStream(plan, st);
        if (r == CUFFT_SUCCESS) r = cufftExecZ2Z(plan, (cufftDoubleComplex*)s.fref, (cufftDoubleComplex*)s.fref, CUFFT_FORWARD);
        if (r == CUFFT_SUCCESS) r = cufftExecZ2Z(plan, (cufftDoubleComplex*)s.freg, (cufftDoubleComplex*)s.freg, CUFFT_FORWARD);
        if (r != CUFFT_SUCCESS) { set_error("cuFFT forward transform failed with code %d", (int)r); return FSQ_E_CUDA; }
        pc_cross_kernel<<<dim3((unsigned)((plane + 255) / 256), B), 256, 0, st>>>(s.fref, s.freg, s.cc, plane, s.power);
        FSQ_LAUNCH_CHECK();
        r = cufftExecZ2Z(plan, (cufftDoubleComplex*)s.cc, (cufftDoubleComplex*)s.cc, CUFFT_INVERSE);
        if (r != CUFFT_SUCCESS) { set_error("cuFFT inverse transform failed with code %d", (int)r); return FSQ_E_CUDA; }
    }
    pc_argmax_kernel<<<B, 256, 0, st>>>(s.cc, plane, 1.0 / (double)plane, 0, s.peak);
    FSQ_LAUNCH_CHECK();
    double* peak_up = s.peak;      // unused when usf == 1
    if (upsample_factor > 1) {
        const double usf = (double)upsample_factor;
        const int m = U * (rows > cols ? rows : cols);
        pc_twiddle_kernel<<<dim3((unsigned)((m + 255) / 256), B), 256, 0, st>>>(s.peak, rows, cols, U, usf, s.rowk, s.colk);
        FSQ_LAUNCH_CHECK();
        pc_dft_rows_kernel<<<dim3((unsigned)((cols + 31) / 32), (unsigned)((U + 31) / 32), B), 256, 0, st>>>(s.fref, s.freg, s.rowk, rows, cols, U, s.t);
        FSQ_LAUNCH_CHECK();
        pc_dft_cols_kernel<<<dim3((unsigned)((U * U + 255) / 256), B), 256, 0, st>>>(s.t, s.colk, cols, U, s.up);
        FSQ_LAUNCH_CHECK();
        // cross_correlation = up.conj() / (mid_row mid_col usf^2) (:103-108); its argmax; the peak record reuses the tail of `t`
        peak_up = (double*)s.t;
        const double norm = (double)(rows / 2) * (double)(cols / 2) * usf * usf;
        pc_argmax_kernel<<<B, 256, 0, st>>>(s.up, (long long)U * U, 1.0 / norm, 1, peak_up);
        FSQ_LAUNCH_CHECK();
    }
    pc_finish_kernel<<<(unsigned)((B + 63) / 64), 64, 0, st>>>(s.peak, peak_up, s.power, (int)((plane + 255) / 256), B, rows, cols, U,
                                                               (double)upsample_factor, out);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}
