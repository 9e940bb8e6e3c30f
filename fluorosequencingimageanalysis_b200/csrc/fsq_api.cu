// libfsq: error channel, version, and the small per-spot kernels (fit-quality metrics on
// arbitrary (sub_img, fit_img) pairs; photometry on spots).
#include "fsq_common.cuh"
#include <string.h>

namespace fsq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
    }
    return n;
}

// ---- pflib.py:463-473 + illumina_s_n pflib.py:261-281; one thread per pair, sums in the
// reference's (raster / edge-list) order --------------------------------------------------
__global__ void __launch_bounds__(128)
metrics_kernel(const long long* __restrict__ sub, const double* __restrict__ fit, long long n,
               double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long* s = sub + i * 25;
    const double* f = fit + i * 25;
    long long isum = 0;
    for (int q = 0; q < 25; ++q) isum += s[q];
    const double mean = __ddiv_rn((double)isum, 25.0);
    double ssr = 0.0, sst = 0.0;                    // Python sum(): unfused squares added in raster order (pflib.py:463-465, 470-472)
    for (int q = 0; q < 25; ++q) {
        const double d = __dsub_rn((double)s[q], f[q]);
        ssr = __dadd_rn(ssr, __dmul_rn(d, d));
        const double m = __dsub_rn((double)s[q], mean);
        sst = __dadd_rn(sst, __dmul_rn(m, m));
    }
    const double s_n = illumina_sn<5>(5, [&](int r, int c) { return s[r * 5 + c]; });   // numpy's arithmetic, bit for bit
    out[i * 3 + 0] = __dsub_rn(1.0, __ddiv_rn(ssr, sst));
    out[i * 3 + 1] = sqrt(__ddiv_rn(ssr, 25.0));
    out[i * 3 + 2] = s_n;
}

// pflib.illumina_s_n for n square integer windows of any size <= 33 (pflib.py:261-281; Spot sizes are a parameter,
// flexlibrary.py:319-320), one thread per window
__global__ void __launch_bounds__(128)
illumina_sn_kernel(const long long* __restrict__ sub, long long n, int size, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long* s = sub + i * size * size;
    out[i] = illumina_sn<0>(size, [&](int r, int c) { return s[r * size + c]; });
}

__device__ __forceinline__ int load_pix_i(const void* base, int dtype, size_t off) {
    switch (dtype) {
        case FSQ_U8:  return (int)((const uint8_t*)base)[off];
        case FSQ_U16: return (int)((const uint16_t*)base)[off];
        case FSQ_I16: return (int)((const int16_t*)base)[off];
        default:      return ((const int32_t*)base)[off];
    }
}

// ---- flexlibrary.Spot photometry; one warp per spot -------------------------------------------
//  method 0 simple   : sum of the (border-truncated) slice of radius `radius`    flexlibrary.py:160-170
//  method 1 mexican  : sum(crown) - len(crown)*median(brim), slice-LOCAL indices flexlibrary.py:172-210
//  method 2 maximum  : max of the slice                                          flexlibrary.py:264-284
constexpr int PH_MAXR = 12;
constexpr int PH_MAXN = (2 * PH_MAXR + 1) * (2 * PH_MAXR + 1);

__global__ void __launch_bounds__(128)
photometry_kernel(const void* __restrict__ frames, int dtype, int H, int W,
                  const int32_t* __restrict__ hw, const int32_t* __restrict__ fr, long long n,
                  int method, int radius, int brim, double* __restrict__ out) {
    __shared__ int vals[4][PH_MAXN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * 4 + warp;
    if (i >= n) return;
    const int h = hw[2 * i], w = hw[2 * i + 1];
    const size_t fbase = (size_t)fr[i] * H * W;
    const int r0 = max(0, h - radius), r1 = min(H, h + radius + 1);
    const int c0 = max(0, w - radius), c1 = min(W, w + radius + 1);
    const int sh = max(r1 - r0, 0), sw = max(c1 - c0, 0);
    const int ns = sh * sw;
    const int diameter = 2 * radius + 1;
    long long tot = 0, crown_sum = 0;
    int crown_n = 0, mx = -2147483647 - 1, nb = 0;
    // pass 1: sums; brim values compacted into shared memory in slice raster order
    for (int base = 0; base < ns; base += 32) {
        const int q = base + lane;
        bool is_brim = false;
        int v = 0;
        if (q < ns) {
            const int lh = q / sw, lw = q - lh * sw;
            v = load_pix_i(frames, dtype, fbase + (size_t)(r0 + lh) * W + (c0 + lw));
            tot += v;
            mx = max(mx, v);
            const bool crown = (brim <= lh && lh < diameter - brim && brim <= lw && lw < diameter - brim);
            if (crown) { crown_sum += v; ++crown_n; } else is_brim = true;
        }
        const unsigned m = __ballot_sync(0xffffffffu, is_brim);
        if (is_brim) vals[warp][nb + __popc(m & ((1u << lane) - 1u))] = v;
        nb += __popc(m);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, m);
        crown_sum += __shfl_xor_sync(0xffffffffu, crown_sum, m);
        crown_n += __shfl_xor_sync(0xffffffffu, crown_n, m);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    }
    __syncwarp();
    double res;
    if (method == 0) res = (double)tot;
    else if (method == 2) res = (double)mx;
    else {
        // median of the nb brim values by rank counting; even count -> mean of the two middle
        const int k_lo = (nb - 1) / 2, k_hi = nb / 2;
        long long pick = 0;     // sum of the (one or two) middle elements
        for (int a = lane; a < nb; a += 32) {
            const int va = vals[warp][a];
            int c = 0;
            for (int b = 0; b < nb; ++b) { const int vb = vals[warp][b]; c += (vb < va) || (vb == va && b < a); }
            if (c == k_lo) pick += va;
            if (c == k_hi) pick += va;
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) pick += __shfl_xor_sync(0xffffffffu, pick, m);
        const double med = nb > 0 ? (double)pick / 2.0 : __longlong_as_double(0x7ff8000000000000LL);
        res = (double)crown_sum - (double)crown_n * med;
    }
    if (lane == 0) out[i] = res;
}


// ---- start values from image moments (agpy/gaussfitter.py:29-61), one warp per window -------------
//  out[i] = (height, amplitude, x, y, width_x, width_y, 0): height = median of the window, amplitude =
//  max - height, y / x = first argmax of the |data|-weighted marginal first moments (divided by
//  sum|data|, gaussfitter.py:39-40), widths = sqrt(sum|(k - centre) * line| / sum|line|) along the row /
//  column through that argmax ("FIRST moment, not second", :42-46).  Every sum runs in the order numpy's
//  add.reduce uses (8 strided accumulators + sequential remainder for a contiguous 1-D run of >= 8
//  elements, plain sequential otherwise and across rows for axis=0), with explicit _rn intrinsics so
//  that no product is contracted into an FMA: float64 windows give the reference's bits.
constexpr int MO_MAXW = 11;
__global__ void __launch_bounds__(128)
moments_kernel(const void* __restrict__ windows, int dtype, long long n, int win, double* out,
               const double* __restrict__ lo, const double* __restrict__ hi,
               const uint8_t* __restrict__ lim_lo, const uint8_t* __restrict__ lim_hi) {
    __shared__ double px[4][MO_MAXW * MO_MAXW];
    __shared__ double marg[4][2 * MO_MAXW + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * 4 + warp;
    if (i >= n) return;
    const int P = win * win;
    double* d = px[warp];
    double vmax = -1.0 / 0.0;
    bool has_nan = false;
    for (int q = lane; q < P; q += 32) {
        double v;
        const size_t off = (size_t)i * P + q;
        switch (dtype) {
            case FSQ_U8:  v = (double)((const uint8_t*)windows)[off]; break;
            case FSQ_U16: v = (double)((const uint16_t*)windows)[off]; break;
            case FSQ_I16: v = (double)((const int16_t*)windows)[off]; break;
            case FSQ_I32: v = (double)((const int32_t*)windows)[off]; break;
            case FSQ_I64: v = (double)((const long long*)windows)[off]; break;
            default:      v = ((const double*)windows)[off]; break;
        }
        d[q] = v;
        vmax = fmax(vmax, v);                      // (numpy's max propagates NaN; handled through has_nan)
        has_nan |= (v != v);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, m));
    has_nan = __any_sync(0xffffffffu, has_nan);
    __syncwarp();
    // lanes 0..win-1: row marginals (sum over the contiguous axis 1 -> pairwise order); lanes 16..: column
    // marginals (axis 0 -> row after row); lane 15: total = sum|data| over the flat window
    if (lane < win) {
        marg[warp][lane] = np_sum(win, [&](int c) { return __dmul_rn((double)c, fabs(d[lane * win + c])); });
    } else if (lane >= 16 && lane - 16 < win) {
        const int c = lane - 16;
        double s = 0.0;
        for (int r = 0; r < win; ++r) s = __dadd_rn(s, __dmul_rn((double)r, fabs(d[r * win + c])));
        marg[warp][MO_MAXW + c] = s;
    } else if (lane == 15) {
        marg[warp][2 * MO_MAXW] = np_sum(P, [&](int q) { return fabs(d[q]); });
    }
    __syncwarp();
    if (lane == 0) {
        const double total = marg[warp][2 * MO_MAXW];
        int ya = 0, xa = 0;                         // first maximum, like numpy.argmax (NaN wins there too)
        double by = __ddiv_rn(marg[warp][0], total), bx = __ddiv_rn(marg[warp][MO_MAXW], total);
        for (int k = 1; k < win; ++k) {
            const double qy = __ddiv_rn(marg[warp][k], total), qx = __ddiv_rn(marg[warp][MO_MAXW + k], total);
            if (qy > by) { by = qy; ya = k; }
            if (qx > bx) { bx = qx; xa = k; }
        }
        // gaussfitter.py:41-45: the *row* through y is paired with (k - y), the *column* through x with (k - x)
        const double wx = sqrt(__ddiv_rn(
            np_sum(win, [&](int k) { return fabs(__dmul_rn((double)(k - ya), d[ya * win + k])); }),
            np_sum(win, [&](int k) { return fabs(d[ya * win + k]); })));
        const double wy = sqrt(__ddiv_rn(
            np_sum(win, [&](int k) { return fabs(__dmul_rn((double)(k - xa), d[k * win + xa])); }),
            np_sum(win, [&](int k) { return fabs(d[k * win + xa]); })));
        double* o = out + i * 7;
        o[2] = (double)xa; o[3] = (double)ya; o[4] = wx; o[5] = wy; o[6] = 0.0;
    }
    // median by rank counting (numpy.median: mean of the two middle elements for an even count)
    const int k_lo = (P - 1) / 2, k_hi = P / 2;
    double pick = 0.0;
    for (int a = lane; a < P; a += 32) {
        const double va = d[a];
        int c = 0;
        for (int b = 0; b < P; ++b) { const double vb = d[b]; c += (vb < va) || (vb == va && b < a); }
        if (c == k_lo) pick += va;
        if (c == k_hi) pick += va;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) pick += __shfl_xor_sync(0xffffffffu, pick, m);
    if (lane == 0) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        const double med = has_nan ? nan : 0.5 * pick;
        out[i * 7 + 0] = med;
        out[i * 7 + 1] = (has_nan ? nan : vmax) - med;
    }
    if (lo) {                                       // gaussfitter.py:202-204: start values clipped into the limits
        __syncwarp();
        if (lane < 7) {
            double v = out[i * 7 + lane];
            if (lim_hi[lane] && v > hi[lane]) v = hi[lane];
            if (lim_lo[lane] && v < lo[lane]) v = lo[lane];
            out[i * 7 + lane] = v;
        }
    }
}

// ---- luminosity-centroid particle tracking (flexlibrary.py:1173-1317), one warp per spot ----------------------
//  For every spot of frame 0 and every later frame f: take the (2R+1)^2 window of frame f around the spot's last
//  known position (minus the frame's integer offset); if it is cut by the image border the spot is None in this
//  frame; otherwise the new position is the python-2-rounded centre of mass of the window (scipy
//  center_of_mass: integer sums, one float64 division per axis); a position whose size x size square leaves the
//  image is None (Spot.__init__ raises AttributeError, :100-118); if pflib.illumina_s_n of the new square is below
//  the cut-off the spot stays where it was (:1244-1250 -- at the PRIOR coordinates, not offset-adjusted, as the
//  reference does).  A None leaves the last known position in force (:1313-1314).
//  state: 0 None, 1 centroid accepted, 2 fell back to the prior position, 3 the initial spot (frame 0).
__global__ void __launch_bounds__(128)
track_centroid_kernel(const void* __restrict__ frames, int dtype, int n_frames, int H, int W,
                      const int32_t* __restrict__ spots_hw, const int32_t* __restrict__ spot_field,
                      const int32_t* __restrict__ offsets, long long n, int size, int R, double cutoff,
                      int32_t* __restrict__ track_hw, uint8_t* __restrict__ track_state, double* __restrict__ track_sn) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * 4 + warp;
    if (i >= n) return;
    const int half = (size - 1) / 2;
    const int D = 2 * R + 1;
    const size_t field_base = (size_t)(spot_field ? spot_field[i] : 0) * n_frames * H * W;
    int ph = spots_hw[2 * i], pw = spots_hw[2 * i + 1];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int f = 0; f < n_frames; ++f) {
        const size_t fbase = field_base + (size_t)f * H * W;
        int nh = ph, nw = pw, state = 3;
        bool none = false;
        if (f > 0) {
            const int oh = ph - (offsets ? offsets[2 * f] : 0), ow = pw - (offsets ? offsets[2 * f + 1] : 0);
            if (oh - R < 0 || oh + R + 1 > H || ow - R < 0 || ow + R + 1 > W) none = true;          // :1227-1229
            else {
                long long S = 0, Sh = 0, Sw = 0;
                for (int q = lane; q < D * D; q += 32) {
                    const int r = q / D, c = q - r * D;
                    const long long v = load_pix_i(frames, dtype, fbase + (size_t)(oh - R + r) * W + (ow - R + c));
                    S += v; Sh += v * r; Sw += v * c;
                }
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) {
                    S += __shfl_xor_sync(0xffffffffu, S, m);
                    Sh += __shfl_xor_sync(0xffffffffu, Sh, m);
                    Sw += __shfl_xor_sync(0xffffffffu, Sw, m);
                }
                if (S == 0) none = true;            // (the reference divides by zero here and fails in round())
                else {
                    const double ch = __ddiv_rn((double)Sh, (double)S), cw = __ddiv_rn((double)Sw, (double)S);
                    nh = (int)round(__dadd_rn(__dadd_rn(ch, (double)oh), -(double)R));               // :1233-1234
                    nw = (int)round(__dadd_rn(__dadd_rn(cw, (double)ow), -(double)R));
                    state = 1;
                }
            }
        }
        if (!none && !(0 <= nh - half && nh + half < H && 0 <= nw - half && nw + half < W)) none = true;   // Spot.__init__
        double sn = nan;
        if (!none) {
            // pflib.illumina_s_n of the size x size square (pflib.py:261-281): numpy's mean / std arithmetic
            int v = 0;
            bool edge = false;
            if (lane < size * size) {
                const int r = lane / size, c = lane - r * size;
                v = load_pix_i(frames, dtype, fbase + (size_t)(nh - half + r) * W + (nw - half + c));
                edge = (r == 0 || r == size - 1 || c == 0 || c == size - 1);
            }
            int mx = lane < size * size ? v : (-2147483647 - 1);
            long long es = edge ? v : 0;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, m));
                es += __shfl_xor_sync(0xffffffffu, es, m);
            }
            const int ne = 4 * (size - 1);
            const double mean = __ddiv_rn((double)es, (double)ne);
            // squared deviations summed in numpy's order over the reference's element list: top row, bottom row,
            // then the (h, 0), (h, -1) pairs of the middle rows (pflib.py:278-280); ne <= 16
            const double dev = (double)v - mean;
            const double sq = __dmul_rn(dev, dev);
            double ordered[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                int src;
                if (k < size) src = k;
                else if (k < 2 * size) src = (size - 1) * size + (k - size);
                else { const int kk = k - 2 * size; src = (1 + kk / 2) * size + ((kk & 1) ? size - 1 : 0); }
                ordered[k] = __shfl_sync(0xffffffffu, sq, src & 31);
            }
            const double var_sum = np_sum(ne, [&](int k) { return ordered[k]; });
            const double sd = sqrt(__ddiv_rn(var_sum, (double)ne));
            sn = __ddiv_rn((double)mx - mean, sd);
            if (f > 0 && sn < cutoff) { nh = ph; nw = pw; state = 2; }                          // :1244-1250
        }
        if (lane == 0) {
            const long long o = i * n_frames + f;
            track_hw[2 * o] = none ? -1 : nh; track_hw[2 * o + 1] = none ? -1 : nw;
            track_state[o] = none ? 0 : (uint8_t)state;
            track_sn[o] = sn;
        }
        if (!none) { ph = nh; pw = nw; }
    }
}

// ---- FMA-pipe micro-benchmark (roofline denominator of bench.py) -------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters) {
    // 8 independent chains per thread, multiplier / addend read from memory so nothing folds
    T a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = out[threadIdx.x] + (T)k;
    const T b = out[1] + (T)1.0000001, c = out[2] + (T)1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
        }
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace fsq

using namespace fsq;

extern "C" int fsq_version(void) { return FSQ_VERSION; }
extern "C" const char* fsq_last_error(void) { return g_err; }

extern "C" int fsq_illumina_s_n(const int64_t* sub, int64_t n, int size, double* out, void* stream) {
    if (n < 0 || size < 1 || size > 33) { set_error("fsq_illumina_s_n: n >= 0 and 1 <= size <= 33 required (got %lld, %d)", (long long)n, size); return FSQ_E_ARG; }
    if (n == 0) return FSQ_OK;
    illumina_sn_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const long long*)sub, n, size, out);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}

extern "C" int fsq_metrics(const int64_t* sub, const double* fit, int64_t n, double* out, void* stream) {
    if (n < 0) { set_error("fsq_metrics: n < 0"); return FSQ_E_ARG; }
    if (n == 0) return FSQ_OK;
    if (!sub || !fit || !out) { set_error("fsq_metrics: NULL pointer argument"); return FSQ_E_ARG; }
    metrics_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const long long*)sub, fit, n, out);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}

extern "C" int fsq_photometry(const void* frames, int dtype_code, int n_frames, int H, int W,
                              const int32_t* spots_hw, const int32_t* spot_frame, int64_t n,
                              int method, int radius, int brim, double* out, void* stream) {
    if (n < 0) { set_error("fsq_photometry: n < 0"); return FSQ_E_ARG; }
    if (n == 0) return FSQ_OK;
    if (!frames || !spots_hw || !spot_frame || !out) { set_error("fsq_photometry: NULL pointer argument"); return FSQ_E_ARG; }
    if (method < 0 || method > 2) { set_error("fsq_photometry: Uknown method specified."); return FSQ_E_ARG; }
    if (radius < 0 || radius > PH_MAXR) { set_error("fsq_photometry: radius must be in 0..%d", PH_MAXR); return FSQ_E_ARG; }
    if (dtype_code != FSQ_U8 && dtype_code != FSQ_U16 && dtype_code != FSQ_I16 && dtype_code != FSQ_I32) {
        set_error("fsq_photometry: unsupported frame dtype code %d", dtype_code);
        return FSQ_E_ARG;
    }
    (void)n_frames;
    photometry_kernel<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(frames, dtype_code, H, W, spots_hw,
                                                                                 spot_frame, n, method, radius, brim, out);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}

extern "C" int fsq_fma_peak(int fp64, double* flops_out_host, void* stream) {
    if (!flops_out_host) { set_error("fsq_fma_peak: NULL"); return FSQ_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = sm_count() * 8, threads = 256, iters = 4096;
    void* buf = nullptr;
    FSQ_CUDA_CHECK(cudaMalloc(&buf, size_t(blocks) * threads * 8));
    FSQ_CUDA_CHECK(cudaMemsetAsync(buf, 0, size_t(blocks) * threads * 8, st));
    cudaEvent_t e0, e1;
    FSQ_CUDA_CHECK(cudaEventCreate(&e0));
    FSQ_CUDA_CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        FSQ_CUDA_CHECK(cudaEventRecord(e0, st));
        if (fp64) fma_peak_kernel<double><<<blocks, threads, 0, st>>>((double*)buf, iters);
        else fma_peak_kernel<float><<<blocks, threads, 0, st>>>((float*)buf, iters);
        FSQ_CUDA_CHECK(cudaEventRecord(e1, st));
        FSQ_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        FSQ_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    const double flop = 2.0 * 64.0 * double(iters) * double(blocks) * double(threads);
    *flops_out_host = flop / (double(best) * 1e-3);
    return FSQ_OK;
}

extern "C" int fsq_moments(const void* windows, int dtype_code, int64_t n, int win, const double* lo, const double* hi,
                           const uint8_t* lim_lo, const uint8_t* lim_hi, double* p0_out, void* stream) {
    if (n < 0) { set_error("fsq_moments: n < 0"); return FSQ_E_ARG; }
    if (n == 0) return FSQ_OK;
    if (!windows || !p0_out) { set_error("fsq_moments: NULL pointer argument"); return FSQ_E_ARG; }
    if (win < 1 || win > MO_MAXW) { set_error("fsq_moments: window side must be in 1..%d (got %d)", MO_MAXW, win); return FSQ_E_ARG; }
    if (dtype_code < FSQ_U8 || dtype_code > FSQ_I64) { set_error("fsq_moments: unsupported dtype code %d", dtype_code); return FSQ_E_ARG; }
    if (lo && !(hi && lim_lo && lim_hi)) { set_error("fsq_moments: lo, hi, lim_lo, lim_hi go together"); return FSQ_E_ARG; }
    moments_kernel<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(windows, dtype_code, n, win, p0_out,
                                                                             lo, hi, lim_lo, lim_hi);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}

extern "C" int fsq_track_centroid(const void* frames, int dtype_code, int n_fields, int n_frames, int H, int W,
                                  const int32_t* spots_hw, const int32_t* spot_field, const int32_t* offsets, int64_t n,
                                  int spot_size, int search_radius, double s_n_cutoff,
                                  int32_t* track_hw, uint8_t* track_state, double* track_sn, void* stream) {
    if (n < 0 || n_fields <= 0 || n_frames <= 0 || H <= 0 || W <= 0) { set_error("fsq_track_centroid: bad sizes"); return FSQ_E_ARG; }
    if (n == 0) return FSQ_OK;
    if (!frames || !spots_hw || !track_hw || !track_state || !track_sn) { set_error("fsq_track_centroid: NULL pointer argument"); return FSQ_E_ARG; }
    if (spot_size % 2 == 0) { set_error("Spot.size must be odd."); return FSQ_E_ARG; }                      // flexlibrary.py:98-99
    if (spot_size < 3 || spot_size > 5) { set_error("fsq_track_centroid: spot_size must be 3 or 5 (got %d)", spot_size); return FSQ_E_ARG; }
    if (search_radius < 1 || search_radius > 15) { set_error("fsq_track_centroid: search_radius must be in 1..15 (got %d)", search_radius); return FSQ_E_ARG; }
    if (dtype_code < FSQ_U8 || dtype_code > FSQ_I32) { set_error("fsq_track_centroid: unsupported dtype code %d", dtype_code); return FSQ_E_ARG; }
    track_centroid_kernel<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        frames, dtype_code, n_frames, H, W, spots_hw, spot_field, offsets, (long long)n, spot_size, search_radius, s_n_cutoff,
        track_hw, track_state, track_sn);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}
