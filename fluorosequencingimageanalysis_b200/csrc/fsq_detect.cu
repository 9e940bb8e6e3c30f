// Candidate detection for a batch of frames (replaces pflib._psf_candidates, pflib.py:217-258).
//
//   pass A  detect_cm_kernel    tile-staged: raw tile (+halo, scipy 'reflect' borders) -> shared
//                               memory -> s x s median (99-comparator selection network for
//                               s = 5) -> mf = v - min(med, v) -> k x k integer correlation in
//                               int64 -> clamp -> cm (saturated u32 scratch) + exact per-frame
//                               integer moments  sum(cm), sum(cm^2) (three 64-bit limbs)
//   pass T  detect_thr_kernel   thr = mean + c_std * std from the exact moments (128-bit integer
//                               variance numerator, one sqrt)                      pflib.py:250
//   pass B  detect_rowmask / scans / detect_emit   per-row ballot masks -> exclusive scans ->
//                               raster-ordered (h, w) list, 2-px border excluded   pflib.py:252-257
//
// Algorithmic bytes per frame: H*W*sizeof(pixel) read + 8*N_cand written.
#include "fsq_common.cuh"
#include "fsq_median.cuh"
#include <stdlib.h>

namespace fsq {

constexpr int TW = 64;     // output tile width
constexpr int TH = 32;     // output tile height
constexpr int NT = 256;    // threads per block (pass A)
constexpr int MAXS = 9;    // largest median window side
constexpr int MAXK = 9;    // largest correlation template side
constexpr int RAW_MAX = (TH + (MAXK - 1) + (MAXS - 1)) * (TW + (MAXK - 1) + (MAXS - 1));
constexpr int MF_MAX = (TH + (MAXK - 1)) * (TW + (MAXK - 1));
constexpr int CM_SPLIT = 20;               // cm = a * 2^20 + b
constexpr int NSUM = 8;                    // u64 slots per frame in the moment scratch

struct KParam { int k[MAXK * MAXK]; };

struct DetectScratch {
    uint32_t* cm32;                 // [F*H*W]
    unsigned long long* sums;       // [F*NSUM]  s1, saa, sab, sbb, range-flag
    int32_t* rowcount;              // [F*H]
    int32_t* rowoff;                // [F*H]
    int64_t* framebase;             // [F+1]
    uint32_t* masks;                // [F*H*ceil(W/32)]
    int32_t* flags;                 // [4]
};

static inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

static int64_t carve(DetectScratch* s, char* base, int F, int H, int W) {
    int64_t off = 0;
    const int64_t npx = int64_t(F) * H * W;
    const int64_t nrow = int64_t(F) * H;
    const int64_t nw = (W + 31) / 32;
    auto take = [&](int64_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
    char* p;
    p = take(npx * 4);               if (s) s->cm32 = (uint32_t*)p;
    p = take(int64_t(F) * NSUM * 8); if (s) s->sums = (unsigned long long*)p;
    p = take(nrow * 4);              if (s) s->rowcount = (int32_t*)p;
    p = take(nrow * 4);              if (s) s->rowoff = (int32_t*)p;
    p = take((int64_t(F) + 1) * 8);  if (s) s->framebase = (int64_t*)p;
    p = take(nrow * nw * 4);         if (s) s->masks = (uint32_t*)p;
    p = take(64);                    if (s) s->flags = (int32_t*)p;
    return off;
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // scipy.ndimage mode='reflect' == numpy 'symmetric':  d c b a | a b c d | d c b a
    if ((unsigned)i < (unsigned)n) return i;           // interior: no integer modulo
    if (n == 1) return 0;
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}


// Generic rank-(n/2) selection by counting (used only for non-default median sizes).
__device__ int median_generic(const int* raw, int RW, int y0, int x0, int s) {
    const int n = s * s;
    const int want = n / 2;
    for (int a = 0; a < n; ++a) {
        const int va = raw[(y0 + a / s) * RW + x0 + a % s];
        int c = 0;
        for (int b = 0; b < n; ++b) {
            const int vb = raw[(y0 + b / s) * RW + x0 + b % s];
            c += (vb < va) || (vb == va && b < a);
        }
        if (c == want) return va;
    }
    return 0;
}

template <typename PixT, int S, int K>
__global__ void __launch_bounds__(NT)
detect_cm_kernel(const PixT* __restrict__ frames, int H, int W, KParam kp, int s_rt, int k_rt,
                 uint32_t* __restrict__ cm32, unsigned long long* __restrict__ sums) {
    const int s = S ? S : s_rt;
    const int k = K ? K : k_rt;
    const int mlo = s / 2, mhi = s - 1 - mlo, kh = k / 2;
    const int RW = TW + 2 * kh + mlo + mhi, RH = TH + 2 * kh + mlo + mhi;
    const int MW = TW + 2 * kh, MH = TH + 2 * kh;
    __shared__ int raw[RAW_MAX];
    __shared__ int mf[MF_MAX];
    __shared__ unsigned long long red[4][NT / 32];

    const int tid = threadIdx.x;
    const int frame = blockIdx.z;
    const int ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
    const PixT* img = frames + size_t(frame) * H * W;

    // stage the raw tile (reflected at the image border)
    for (int idx = tid; idx < RH * RW; idx += NT) {
        const int ry = idx / RW, rx = idx - ry * RW;
        const int gy = reflect_idx(ty0 - kh - mlo + ry, H);
        const int gx = reflect_idx(tx0 - kh - mlo + rx, W);
        raw[idx] = (int)img[size_t(gy) * W + gx];
    }
    __syncthreads();

    // background removal: mf = v - min(median, v); zero outside the image (correlate pads with 0)
    for (int idx = tid; idx < MH * MW; idx += NT) {
        const int my = idx / MW, mx = idx - my * MW;
        const int gy = ty0 - kh + my, gx = tx0 - kh + mx;
        int out = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const int v = raw[(my + mlo) * RW + mx + mlo];
            int med;
            if (S == 5) {
                int p[25];
#pragma unroll
                for (int i = 0; i < 5; ++i)
#pragma unroll
                    for (int j = 0; j < 5; ++j) p[i * 5 + j] = raw[(my + i) * RW + mx + j];
                med = median25(p);
            } else {
                med = median_generic(raw, RW, my, mx, s);
            }
            out = v - min(med, v);
        }
        mf[idx] = out;
    }
    __syncthreads();

    // correlation, clamp, store, exact moments
    unsigned long long s1 = 0, saa = 0, sab = 0, sbb = 0, bad = 0;
    for (int idx = tid; idx < TH * TW; idx += NT) {
        const int oy = idx / TW, ox = idx - oy * TW;
        const int gy = ty0 + oy, gx = tx0 + ox;
        if (gy < H && gx < W) {
            long long acc = 0;
            if (K == 5) {
#pragma unroll
                for (int i = 0; i < 5; ++i)
#pragma unroll
                    for (int j = 0; j < 5; ++j)
                        acc += (long long)kp.k[i * 5 + j] * (long long)mf[(oy + i) * MW + ox + j];
            } else {
                for (int i = 0; i < k; ++i)
                    for (int j = 0; j < k; ++j)
                        acc += (long long)kp.k[i * k + j] * (long long)mf[(oy + i) * MW + ox + j];
            }
            const unsigned long long cm = acc > 0 ? (unsigned long long)acc : 0ull;
            cm32[(size_t(frame) * H + gy) * W + gx] = cm > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cm;
            const unsigned long long a = cm >> CM_SPLIT, b = cm & ((1ull << CM_SPLIT) - 1);
            bad |= (a >> 21);                  // cm >= 2^41 would overflow the limb sums
            s1 += cm; saa += a * a; sab += a * b; sbb += b * b;
        }
    }
    // block reduction -> one atomic per block and limb
    unsigned long long v[4] = {s1, saa, sab, sbb};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], m);
        if ((tid & 31) == 0) red[q][tid >> 5] = v[q];
    }
    bad = __any_sync(0xffffffffu, bad != 0);
    if (bad && (tid & 31) == 0) atomicOr(&sums[size_t(frame) * NSUM + 4], 1ull);
    __syncthreads();
    if (tid < 4) {
        unsigned long long t = 0;
        for (int w = 0; w < NT / 32; ++w) t += red[tid][w];
        atomicAdd(&sums[size_t(frame) * NSUM + tid], t);
    }
}

// -------------------------------------------------------------------------------------------
// Fast path of pass A for unsigned 8/16-bit pixels, 5x5 median, 5x5 template (the defaults).
//   * the raw tile is staged as packed u16 pairs;
//   * every thread computes the medians of TWO horizontally adjacent pixels at once: element
//     (i, j) of both windows sits in one 32-bit register (byte-permuted from aligned words) and
//     the 99-comparator network runs on __vminu2 / __vmaxu2, one instruction per pair;
//   * mf = v - min(med, v) is one saturating packed subtract;
//   * RING templates (pflib.default_correlation_matrix: border ring kr, inner ring ki, centre
//     kc) use  cm = kr*S25 + (ki-kr)*S9 + (kc-ki)*c  with the 5x5 / 3x3 box sums slid down a
//     column strip in registers; any other 5x5 template takes the 25-tap int64 form.
// Same integers as the generic kernel (tests/test_gpu_detect.py compares both with the oracle).
// -------------------------------------------------------------------------------------------
struct U16x2 { unsigned v; };
__device__ __forceinline__ U16x2 min(U16x2 a, U16x2 b) { U16x2 r; r.v = __vminu2(a.v, b.v); return r; }
__device__ __forceinline__ U16x2 max(U16x2 a, U16x2 b) { U16x2 r; r.v = __vmaxu2(a.v, b.v); return r; }

}  // namespace fsq
#include "fsq_median_pair.cuh"
namespace fsq {

constexpr int PTW = 64, PTH = 32;                 // output tile
constexpr int PRW = PTW + 8, PRH = PTH + 8;       // raw tile (halo 4), PRW even
constexpr int PMW = PTW + 4, PMH = PTH + 4;       // mf tile (halo 2)
constexpr int PSTRIP = 8;                         // output rows per thread in the correlation stage
#ifndef DETECT_MEDIAN_DEFAULT
#define DETECT_MEDIAN_DEFAULT 2                   // see detect_cm_packed_kernel
#endif
// Threads per block of the packed kernel.  The median stage has (PMH / 2) (PMW / 2) = 612 work items: 256 threads need
// three sweeps with 20 % of the last one idle, 320 threads two sweeps with 4 % idle (the correlation stage then uses
// the first 256 threads).  Measured on B200 (profiles/r02_detect_kernel_320threads.txt): the same 105 us per 42-frame
// chunk either way -- with 256 threads the ALU pipe is 79 % busy and is the limit, with 320 threads it is 61 % busy and
// the block-wide barriers between the stages are (10 warps per barrier) -- so the idle lanes of the last sweep cost
// nothing: other warps use the pipe.  256 it stays.
#ifndef DETECT_NTP
#define DETECT_NTP 256
#endif
constexpr int NTP = DETECT_NTP;
static_assert(NTP >= 256 && NTP % 32 == 0, "the correlation stage maps 256 threads onto the tile");

// Shared-memory rows of the staged raw tile: 80 pixels = 40 words.  The tile needs columns tx0 - 4 .. tx0 + 67 (PRW = 72),
// but tx0 - 4 is only 8-byte aligned; a row that starts at tx0 - 8 is 16-byte aligned (tx0 is a multiple of 64, W of 8),
// which is what the bulk-async copy engine asks for.  RO = word offset of raw column 0 inside a staged row.
constexpr int RP = (PRW + 8) / 2;
constexpr int RO = 2;

// ---- bulk-async (TMA unit, cp.async.bulk -> UBLKCP) staging of interior raw tiles: one thread arms an mbarrier with the
// tile's byte count, PRH threads issue one 160-byte row copy each straight into shared memory, the block waits on the
// barrier's phase -- no register round trip, no per-thread address arithmetic.  Border tiles need scipy's reflection and
// are staged index by index.  (The tensor-map form, cp.async.bulk.tensor / UTMALDG, traps with "illegal instruction" on
// this pool's boxes for every descriptor tried -- tools/micro/tma_probe.cu, profiles/r02_tma_probe.txt -- the row form
// of the same engine works.)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");              // the async proxy sees the initialised barrier
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

// MEDIAN: 0 = one window pair per step (99-comparator network); 1 = two vertically adjacent window pairs per step
// (fsq_median_pair.cuh); 2 = the same on row windows sorted once by a pre-pass into shared memory (each sorted row
// window serves the five output rows that contain it)
template <typename PixT, bool RING, int MEDIAN>
__global__ void __launch_bounds__(NTP)
detect_cm_packed_kernel(const PixT* __restrict__ frames, int H, int W, KParam kp,
                        uint32_t* __restrict__ cm32, unsigned long long* __restrict__ sums, int use_bulk) {
    constexpr bool PAIR = (MEDIAN == 1);
    __shared__ __align__(128) unsigned raw2[PRH * RP];            // u16 pairs, rows of 80 pixels (see RP / RO)
    __shared__ __align__(8) unsigned long long stage_bar;
    __shared__ unsigned srow[MEDIAN == 2 ? PRH * (PMW / 2) * 5 : 1];      // sorted 5-pixel row windows of pixel pairs
    __shared__ __align__(16) unsigned short mf[PMH * PMW];
    __shared__ unsigned long long red[4][NTP / 32];

    const int tid = threadIdx.x;
    const int frame = blockIdx.z;
    const int ty0 = blockIdx.y * PTH, tx0 = blockIdx.x * PTW;
    const PixT* img = frames + size_t(frame) * H * W;

    // stage the raw tile, two pixels per word.  Interior tiles of 16-bit frames: one bulk-async row copy of 80 pixels per
    // tile row (rows start at tx0 - 8, 16-byte aligned when W % 8 == 0 and the frame base is), or -- when that alignment
    // is not given -- four pixels per 64-bit load; border tiles reflect index by index
    const bool inside = (sizeof(PixT) == 2) && (ty0 >= 4) && (ty0 + PTH + 4 <= H);
    const bool bulk = use_bulk && inside && (W % 8 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15u) == 0) && (tx0 >= 8) && (tx0 + PTW + 8 <= W);
    const bool interior = inside && (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 7u) == 0) && (tx0 >= 4) && (tx0 + PTW + 4 <= W);
    if (bulk) {
        if (tid == 0) { mbar_init(&stage_bar, 1); mbar_expect_tx(&stage_bar, (unsigned)(PRH * RP * 4)); }
        __syncthreads();                                                      // barrier armed before any copy completes on it
        if (tid < PRH) bulk_copy_g2s(raw2 + tid * RP, img + size_t(ty0 - 4 + tid) * W + (tx0 - 8), (unsigned)(RP * 4), &stage_bar);
        mbar_wait(&stage_bar, 0);
    } else if (interior) {
        for (int idx = tid; idx < PRH * (PRW / 4); idx += NTP) {
            const int ry = idx / (PRW / 4), rq = idx - ry * (PRW / 4);
            const uint2 v = *reinterpret_cast<const uint2*>(img + size_t(ty0 - 4 + ry) * W + (tx0 - 4) + 4 * rq);
            *reinterpret_cast<uint2*>(raw2 + ry * RP + RO + 2 * rq) = v;
        }
    } else
    for (int idx = tid; idx < PRH * (PRW / 2); idx += NTP) {
        const int ry = idx / (PRW / 2), rw = idx - ry * (PRW / 2), rx = rw * 2;
        const int gy = reflect_idx(ty0 - 4 + ry, H);
        const int gx0 = reflect_idx(tx0 - 4 + rx, W), gx1 = reflect_idx(tx0 - 4 + rx + 1, W);
        const unsigned lo = (unsigned)img[size_t(gy) * W + gx0], hi = (unsigned)img[size_t(gy) * W + gx1];
        raw2[ry * RP + RO + rw] = lo | (hi << 16);
    }
    __syncthreads();

    // background removal for pixel pairs: mf = v - min(median, v); zero outside the image
    if (MEDIAN == 2) {
        for (int idx = tid; idx < PRH * (PMW / 2); idx += NTP) {               // pre-pass: sort every row window once
            const int ry = idx / (PMW / 2), mxp = idx - ry * (PMW / 2);
            const unsigned* row = raw2 + ry * RP + RO + mxp;
            const unsigned w0 = row[0], w1 = row[1], w2 = row[2];
            U16x2 r[5];
            r[0].v = w0; r[1].v = __byte_perm(w0, w1, 0x5432); r[2].v = w1; r[3].v = __byte_perm(w1, w2, 0x5432); r[4].v = w2;
            sort5<U16x2>(r);
#pragma unroll
            for (int k = 0; k < 5; ++k) srow[idx * 5 + k] = r[k].v;
        }
        __syncthreads();
        for (int idx = tid; idx < (PMH / 2) * (PMW / 2); idx += NTP) {
            const int mp = idx / (PMW / 2), mxp = idx - mp * (PMW / 2);
            const int my = 2 * mp, mx = 2 * mxp;
            U16x2 p[30];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int k = 0; k < 5; ++k) p[i * 5 + k].v = srow[((my + i) * (PMW / 2) + mxp) * 5 + k];
            const unsigned v_top = raw2[(my + 2) * RP + RO + mxp + 1], v_bot = raw2[(my + 3) * RP + RO + mxp + 1];
            U16x2 m_top, m_bot;
            median25_pair_sorted<U16x2>(p, m_top, m_bot);
            unsigned out[2] = {__vsubus2(v_top, m_top.v), __vsubus2(v_bot, m_bot.v)};      // max(v - med, 0) per half
            const int gx = tx0 - 2 + mx;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int gy = ty0 - 2 + my + q;
                const bool yok = (gy >= 0) && (gy < H);
                if (!(yok && gx >= 0 && gx < W)) out[q] &= 0xffff0000u;
                if (!(yok && gx + 1 >= 0 && gx + 1 < W)) out[q] &= 0x0000ffffu;
                *reinterpret_cast<unsigned*>(mf + (my + q) * PMW + mx) = out[q];
            }
        }
    } else if (PAIR) {
        // two vertically adjacent pixel pairs per step: their windows share four of five rows, and of those 20 values
        // only the middle six can be either median -- fsq_median_pair.cuh (generated, verified on all 0/1 inputs):
        // 108 instead of 198 min/max per window
        static_assert(PMH % 2 == 0 && PRH >= PMH + 4, "row pairs need an even mf tile");
        for (int idx = tid; idx < (PMH / 2) * (PMW / 2); idx += NTP) {
            const int mp = idx / (PMW / 2), mx = (idx - mp * (PMW / 2)) * 2;
            const int my = 2 * mp;
            U16x2 p[30];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const unsigned* row = raw2 + (my + i) * RP + RO + (mx >> 1);
                const unsigned w0 = row[0], w1 = row[1], w2 = row[2];
                p[i * 5 + 0].v = w0;
                p[i * 5 + 1].v = __byte_perm(w0, w1, 0x5432);
                p[i * 5 + 2].v = w1;
                p[i * 5 + 3].v = __byte_perm(w1, w2, 0x5432);
                p[i * 5 + 4].v = w2;
            }
            const unsigned v_top = p[12].v, v_bot = p[17].v;
            U16x2 m_top, m_bot;
            median25_pair<U16x2>(p, m_top, m_bot);
            unsigned out[2] = {__vsubus2(v_top, m_top.v), __vsubus2(v_bot, m_bot.v)};      // max(v - med, 0) per half
            const int gx = tx0 - 2 + mx;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int gy = ty0 - 2 + my + q;
                const bool yok = (gy >= 0) && (gy < H);
                if (!(yok && gx >= 0 && gx < W)) out[q] &= 0xffff0000u;
                if (!(yok && gx + 1 >= 0 && gx + 1 < W)) out[q] &= 0x0000ffffu;
                *reinterpret_cast<unsigned*>(mf + (my + q) * PMW + mx) = out[q];
            }
        }
    } else
    for (int idx = tid; idx < PMH * (PMW / 2); idx += NTP) {
        const int my = idx / (PMW / 2), mx = (idx - my * (PMW / 2)) * 2;      // mf coords of the pair's first pixel
        U16x2 p[25];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const unsigned* row = raw2 + (my + i) * RP + RO + (mx >> 1);     // raw col = mx .. mx+5
            const unsigned w0 = row[0], w1 = row[1], w2 = row[2];
            p[i * 5 + 0].v = w0;
            p[i * 5 + 1].v = __byte_perm(w0, w1, 0x5432);
            p[i * 5 + 2].v = w1;
            p[i * 5 + 3].v = __byte_perm(w1, w2, 0x5432);
            p[i * 5 + 4].v = w2;
        }
        const unsigned v = p[12].v;
        const U16x2 med = median25<U16x2>(p);
        unsigned out = __vsubus2(v, med.v);                                    // max(v - med, 0) per half
        const int gy = ty0 - 2 + my, gx = tx0 - 2 + mx;
        const bool yok = (gy >= 0) && (gy < H);
        if (!(yok && gx >= 0 && gx < W)) out &= 0xffff0000u;
        if (!(yok && gx + 1 >= 0 && gx + 1 < W)) out &= 0x0000ffffu;
        *reinterpret_cast<unsigned*>(mf + my * PMW + mx) = out;
    }
    __syncthreads();

    // correlation down column strips, clamp, store, exact moments
    unsigned long long s1 = 0, saa = 0, sab = 0, sbb = 0, bad = 0;
    if (tid < PTW * (PTH / PSTRIP)) {
        const int ox = tid & (PTW - 1);                // 64 columns
        const int oy0 = (tid / PTW) * PSTRIP;          // 4 strips of 8 rows
        const int gx = tx0 + ox;
        if (RING) {
            const long long kr = kp.k[0], kd1 = (long long)kp.k[6] - kp.k[0], kd2 = (long long)kp.k[12] - kp.k[6];
            int h5[5], h3[5];                          // row sums of the 5 rows under the window
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned short* r = mf + (oy0 + i) * PMW + ox;
                const int m1 = r[1] + r[2] + r[3];
                h3[i] = m1; h5[i] = m1 + r[0] + r[4];
            }
#pragma unroll
            for (int q = 0; q < PSTRIP; ++q) {
                const unsigned short* r = mf + (oy0 + q + 4) * PMW + ox;
                const int m1 = r[1] + r[2] + r[3];
                h3[(q + 4) % 5] = m1; h5[(q + 4) % 5] = m1 + r[0] + r[4];
                const int S25 = h5[0] + h5[1] + h5[2] + h5[3] + h5[4];
                const int S9 = h3[(q + 1) % 5] + h3[(q + 2) % 5] + h3[(q + 3) % 5];
                const int c = mf[(oy0 + q + 2) * PMW + ox + 2];
                const long long acc = kr * S25 + kd1 * S9 + kd2 * c;
                const int gy = ty0 + oy0 + q;
                if (gy < H && gx < W) {
                    const unsigned long long cm = acc > 0 ? (unsigned long long)acc : 0ull;
                    cm32[(size_t(frame) * H + gy) * W + gx] = cm > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cm;
                    const unsigned long long a = cm >> CM_SPLIT, b = cm & ((1ull << CM_SPLIT) - 1);
                    bad |= (a >> 21);
                    s1 += cm; saa += a * a; sab += a * b; sbb += b * b;
                }
            }
        } else {
#pragma unroll 1
            for (int q = 0; q < PSTRIP; ++q) {
                const int oy = oy0 + q, gy = ty0 + oy;
                if (gy < H && gx < W) {
                    long long acc = 0;
#pragma unroll
                    for (int i = 0; i < 5; ++i)
#pragma unroll
                        for (int j = 0; j < 5; ++j)
                            acc += (long long)kp.k[i * 5 + j] * (long long)mf[(oy + i) * PMW + ox + j];
                    const unsigned long long cm = acc > 0 ? (unsigned long long)acc : 0ull;
                    cm32[(size_t(frame) * H + gy) * W + gx] = cm > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cm;
                    const unsigned long long a = cm >> CM_SPLIT, b = cm & ((1ull << CM_SPLIT) - 1);
                    bad |= (a >> 21);
                    s1 += cm; saa += a * a; sab += a * b; sbb += b * b;
                }
            }
        }
    }
    unsigned long long v[4] = {s1, saa, sab, sbb};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], m);
        if ((tid & 31) == 0) red[q][tid >> 5] = v[q];
    }
    bad = __any_sync(0xffffffffu, bad != 0);
    if (bad && (tid & 31) == 0) atomicOr(&sums[size_t(frame) * NSUM + 4], 1ull);
    __syncthreads();
    if (tid < 4) {
        unsigned long long t = 0;
        for (int w = 0; w < NTP / 32; ++w) t += red[tid][w];
        atomicAdd(&sums[size_t(frame) * NSUM + tid], t);
    }
}

__global__ void detect_thr_kernel(const unsigned long long* __restrict__ sums, int F, long long N,
                                  double c_std, double* __restrict__ thr, int32_t* flags) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const unsigned long long s1 = sums[size_t(f) * NSUM + 0];
    const unsigned long long saa = sums[size_t(f) * NSUM + 1];
    const unsigned long long sab = sums[size_t(f) * NSUM + 2];
    const unsigned long long sbb = sums[size_t(f) * NSUM + 3];
    if (sums[size_t(f) * NSUM + 4]) atomicOr(&flags[0], 1);
    typedef unsigned __int128 u128;
    const u128 s2 = ((u128)saa << (2 * CM_SPLIT)) + ((u128)sab << (CM_SPLIT + 1)) + (u128)sbb;
    const u128 d = (u128)(unsigned long long)N * s2 - (u128)s1 * (u128)s1;   // N*sum(x^2) - sum(x)^2 >= 0
    const double dd = (double)(unsigned long long)(d >> 64) * 18446744073709551616.0 +
                      (double)(unsigned long long)d;
    const double mean = (double)s1 / (double)N;
    const double sd = sqrt(dd) / (double)N;
    const double t = mean + c_std * sd;
    thr[f] = t;
    if (t > 4294967294.0) atomicOr(&flags[0], 2);   // beyond the saturated u32 scratch
}

// one warp per image row: threshold test -> ballot masks + row count
__global__ void __launch_bounds__(256)
detect_rowmask_kernel(const uint32_t* __restrict__ cm32, const double* __restrict__ thr,
                      long long nrows, int H, int W, uint32_t* __restrict__ masks,
                      int32_t* __restrict__ rowcount) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    const int f = int(row / H), y = int(row - (long long)f * H);
    const double t = thr[f];
    const int nw = (W + 31) / 32;
    const bool yok = (y >= 2) && (y < H - 2);
    const uint32_t* src = cm32 + size_t(row) * W;
    int count = 0;
    for (int wi = 0; wi < nw; ++wi) {
        const int x = wi * 32 + lane;
        bool pred = false;
        if (yok && x >= 2 && x < W - 2) pred = !((double)src[x] < t);     // kept when NOT '<'
        const unsigned m = __ballot_sync(0xffffffffu, pred);
        if (lane == 0) masks[size_t(row) * nw + wi] = m;
        count += __popc(m);
    }
    if (lane == 0) rowcount[row] = count;
}

// exclusive scan of one frame's row counts (block per frame)
__global__ void __launch_bounds__(256)
detect_rowscan_kernel(const int32_t* __restrict__ rowcount, int H, int32_t* __restrict__ rowoff,
                      int64_t* __restrict__ n_cand) {
    __shared__ int buf[256];
    __shared__ int carry;
    const int f = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < H; base += 256) {
        const int y = base + tid;
        const int v = y < H ? rowcount[size_t(f) * H + y] : 0;
        buf[tid] = v;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {
            const int t = tid >= d ? buf[tid - d] : 0;
            __syncthreads();
            buf[tid] += t;
            __syncthreads();
        }
        if (y < H) rowoff[size_t(f) * H + y] = carry + buf[tid] - v;
        __syncthreads();
        if (tid == 255) carry += buf[255];
        __syncthreads();
    }
    if (tid == 0) n_cand[f] = carry;
}

// exclusive scan of the per-frame totals (single block)
__global__ void __launch_bounds__(1024)
detect_framescan_kernel(int64_t* __restrict__ n_cand, int F, int64_t* __restrict__ framebase) {
    __shared__ long long buf[1024];
    __shared__ long long carry;
    const int tid = threadIdx.x;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < F; base += 1024) {
        const int f = base + tid;
        const long long v = f < F ? n_cand[f] : 0;
        buf[tid] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const long long t = tid >= d ? buf[tid - d] : 0;
            __syncthreads();
            buf[tid] += t;
            __syncthreads();
        }
        if (f < F) framebase[f] = carry + buf[tid] - v;
        __syncthreads();
        if (tid == 1023) carry += buf[1023];
        __syncthreads();
    }
    if (tid == 0) { framebase[F] = carry; n_cand[F] = carry; }
}

__global__ void __launch_bounds__(256)
detect_emit_kernel(const uint32_t* __restrict__ masks, const int32_t* __restrict__ rowoff,
                   const int64_t* __restrict__ framebase, long long nrows, int H, int W,
                   int32_t* __restrict__ cand_hw, int32_t* __restrict__ cand_frame, long long cap) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    const int f = int(row / H), y = int(row - (long long)f * H);
    const int nw = (W + 31) / 32;
    long long pos = framebase[f] + rowoff[row];
    for (int wi = 0; wi < nw; ++wi) {
        const unsigned m = masks[size_t(row) * nw + wi];
        if (m == 0) continue;
        if ((m >> lane) & 1u) {
            const long long idx = pos + __popc(m & ((1u << lane) - 1u));
            if (idx < cap) {
                cand_hw[2 * idx] = y;
                cand_hw[2 * idx + 1] = wi * 32 + lane;
                cand_frame[idx] = f;
            }
        }
        pos += __popc(m);
    }
}

static bool is_ring_template(const KParam& kp) {
    // border ring = k[0], inner ring = k[6], centre = k[12] (pflib.default_correlation_matrix, pflib.py:48-52)
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            const bool border = (i == 0 || i == 4 || j == 0 || j == 4);
            const bool centre = (i == 2 && j == 2);
            const int want = border ? kp.k[0] : centre ? kp.k[12] : kp.k[6];
            if (kp.k[i * 5 + j] != want) return false;
        }
    return true;
}

template <typename PixT>
static int launch_cm(const void* frames, int F, int H, int W, const KParam& kp, int s, int k,
                     const DetectScratch& sc, cudaStream_t st, bool allow_packed) {
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, F);
    constexpr bool kPackable = (sizeof(PixT) <= 2) && (PixT(-1) > PixT(0));      // u8, u16
    if (kPackable && allow_packed && s == 5 && k == 5) {
        static_assert(PTW == TW && PTH == TH, "tile shapes of the two pass-A kernels must agree");
        // developer switch (tests): FSQ_DETECT_PAIR = 0 / 1 / 2 selects the median formulation (see the kernel)
        static const int median = getenv("FSQ_DETECT_PAIR") ? atoi(getenv("FSQ_DETECT_PAIR")) : DETECT_MEDIAN_DEFAULT;
        const bool ring = is_ring_template(kp);
        // developer switch (A/B): FSQ_DETECT_BULK = 0 stages interior tiles with 64-bit loads instead of bulk-async row copies
        static const int use_bulk = getenv("FSQ_DETECT_BULK") ? atoi(getenv("FSQ_DETECT_BULK")) : 1;
#define FSQ_LAUNCH_PACKED(R, M) detect_cm_packed_kernel<PixT, R, M><<<grid, NTP, 0, st>>>((const PixT*)frames, H, W, kp, sc.cm32, sc.sums, use_bulk)
        if (ring) { if (median == 2) FSQ_LAUNCH_PACKED(true, 2); else if (median == 1) FSQ_LAUNCH_PACKED(true, 1); else FSQ_LAUNCH_PACKED(true, 0); }
        else      { if (median == 2) FSQ_LAUNCH_PACKED(false, 2); else if (median == 1) FSQ_LAUNCH_PACKED(false, 1); else FSQ_LAUNCH_PACKED(false, 0); }
#undef FSQ_LAUNCH_PACKED
    } else if (s == 5 && k == 5)
        detect_cm_kernel<PixT, 5, 5><<<grid, NT, 0, st>>>((const PixT*)frames, H, W, kp, s, k, sc.cm32, sc.sums);
    else
        detect_cm_kernel<PixT, 0, 0><<<grid, NT, 0, st>>>((const PixT*)frames, H, W, kp, s, k, sc.cm32, sc.sums);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}

}  // namespace fsq

using namespace fsq;

extern "C" int64_t fsq_detect_scratch_bytes(int n_frames, int H, int W) {
    if (n_frames <= 0 || H <= 0 || W <= 0) return 0;
    return carve(nullptr, nullptr, n_frames, H, W);
}

extern "C" int fsq_detect(const void* frames, int dtype_code, int n_frames, int H, int W,
                          const int64_t* K_host, int ksize, int mf_size, double c_std,
                          int32_t* cand_hw, int32_t* cand_frame, int64_t* n_cand, double* thr,
                          int64_t cap, void* scratch, int64_t scratch_bytes, void* stream) {
    if (!frames || !K_host || !cand_hw || !cand_frame || !n_cand || !thr || !scratch) {
        set_error("fsq_detect: NULL pointer argument");
        return FSQ_E_ARG;
    }
    if (n_frames <= 0 || H <= 0 || W <= 0 || cap < 0) {
        set_error("fsq_detect: bad shape n_frames=%d H=%d W=%d cap=%lld", n_frames, H, W, (long long)cap);
        return FSQ_E_ARG;
    }
    if (ksize < 1 || ksize > MAXK || (ksize % 2) == 0) {
        set_error("fsq_detect: correlation_matrix must be square with an odd side <= %d (got %d)", MAXK, ksize);
        return FSQ_E_ARG;
    }
    if (mf_size < 1 || mf_size > MAXS) {
        set_error("fsq_detect: median_filter_size must be in 1..%d (got %d)", MAXS, mf_size);
        return FSQ_E_ARG;
    }
    if (n_frames > 65535) {
        set_error("fsq_detect: at most 65535 frames per call (got %d); split the batch", n_frames);
        return FSQ_E_ARG;
    }
    KParam kp;
    for (int i = 0; i < MAXK * MAXK; ++i) kp.k[i] = 0;
    for (int i = 0; i < ksize * ksize; ++i) {
        if (K_host[i] > 2147483647LL || K_host[i] < -2147483647LL) {
            set_error("fsq_detect: correlation_matrix entries must fit in int32");
            return FSQ_E_RANGE;
        }
        kp.k[i] = (int)K_host[i];
    }
    const int64_t need = carve(nullptr, nullptr, n_frames, H, W);
    if (scratch_bytes < need) {
        set_error("fsq_detect: scratch too small (%lld < %lld)", (long long)scratch_bytes, (long long)need);
        return FSQ_E_CAPACITY;
    }
    DetectScratch sc;
    carve(&sc, (char*)scratch, n_frames, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    // developer switch (tests): FSQ_DETECT_GENERIC=1 forces the scalar pass-A kernel
    static const bool allow_packed = !(getenv("FSQ_DETECT_GENERIC") && getenv("FSQ_DETECT_GENERIC")[0] == '1');
    FSQ_CUDA_CHECK(cudaMemsetAsync(sc.sums, 0, size_t(n_frames) * NSUM * 8, st));
    FSQ_CUDA_CHECK(cudaMemsetAsync(sc.flags, 0, 64, st));
    // Pass A (correlation map + moments), the thresholds and the row masks run chunk by chunk, so that the u32 map of a
    // chunk (4 B per pixel) is still in the 126 MB L2 when the row-mask pass reads it back: with all frames of a large
    // batch in one go the map was written to and re-read from DRAM (37 MB instead of 22 MB of traffic per 40-frame
    // stack, ncu).  Thresholds are per frame, so the chunking changes no result.
    const long long px = (long long)H * W;
    long long chunk = (64LL << 20) / (px * 6);
    if (chunk < 1) chunk = 1;
    const int nw = (W + 31) / 32;
    const size_t esize = dtype_code == FSQ_U8 ? 1 : (dtype_code == FSQ_I32 ? 4 : 2);
    for (long long f0 = 0; f0 < n_frames; f0 += chunk) {
        const int fc = (int)(n_frames - f0 < chunk ? n_frames - f0 : chunk);
        DetectScratch c = sc;
        c.cm32 = sc.cm32 + f0 * px;
        c.sums = sc.sums + f0 * NSUM;
        c.masks = sc.masks + f0 * H * nw;
        c.rowcount = sc.rowcount + f0 * H;
        const void* fr = (const char*)frames + (size_t)f0 * px * esize;
        int rc;
        switch (dtype_code) {
            case FSQ_U8:  rc = launch_cm<uint8_t>(fr, fc, H, W, kp, mf_size, ksize, c, st, allow_packed); break;
            case FSQ_U16: rc = launch_cm<uint16_t>(fr, fc, H, W, kp, mf_size, ksize, c, st, allow_packed); break;
            case FSQ_I16: rc = launch_cm<int16_t>(fr, fc, H, W, kp, mf_size, ksize, c, st, allow_packed); break;
            case FSQ_I32: rc = launch_cm<int32_t>(fr, fc, H, W, kp, mf_size, ksize, c, st, allow_packed); break;
            default:
                set_error("fsq_detect: unsupported dtype code %d", dtype_code);
                return FSQ_E_ARG;
        }
        if (rc != FSQ_OK) return rc;
        detect_thr_kernel<<<(fc + 127) / 128, 128, 0, st>>>(c.sums, fc, (long long)H * W, c_std, thr + f0, sc.flags);
        FSQ_LAUNCH_CHECK();
        const long long crow = (long long)fc * H;
        detect_rowmask_kernel<<<(unsigned)((crow + 7) / 8), 256, 0, st>>>(c.cm32, thr + f0, crow, H, W, c.masks, c.rowcount);
        FSQ_LAUNCH_CHECK();
    }
    const long long nrows = (long long)n_frames * H;
    const unsigned rb = (unsigned)((nrows + 7) / 8);
    detect_rowscan_kernel<<<n_frames, 256, 0, st>>>(sc.rowcount, H, sc.rowoff, n_cand);
    FSQ_LAUNCH_CHECK();
    detect_framescan_kernel<<<1, 1024, 0, st>>>(n_cand, n_frames, sc.framebase);
    FSQ_LAUNCH_CHECK();
    detect_emit_kernel<<<rb, 256, 0, st>>>(sc.masks, sc.rowoff, sc.framebase, nrows, H, W, cand_hw, cand_frame, (long long)cap);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}

// flags[0] bit0: cm >= 2^41 somewhere; bit1: threshold beyond the u32 scratch.  Host-visible
// through fsq_detect_flags (synchronises the stream).
extern "C" int fsq_detect_flags(const void* scratch, int n_frames, int H, int W, void* stream) {
    DetectScratch sc;
    carve(&sc, (char*)scratch, n_frames, H, W);
    int32_t h = 0;
    FSQ_CUDA_CHECK(cudaMemcpyAsync(&h, sc.flags, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    FSQ_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    if (h) {
        set_error("fsq_detect: data outside the exact range (flags=%d: 1=correlation >= 2^41, 2=threshold > 2^32)", h);
        return FSQ_E_RANGE;
    }
    return FSQ_OK;
}

extern "C" int fsq_detect_copy_cm32(const void* scratch, int n_frames, int H, int W, uint32_t* cm32_out,
                                    void* stream) {
    if (!scratch || !cm32_out) { set_error("fsq_detect_copy_cm32: NULL"); return FSQ_E_ARG; }
    DetectScratch sc;
    carve(&sc, (char*)scratch, n_frames, H, W);
    FSQ_CUDA_CHECK(cudaMemcpyAsync(cm32_out, sc.cm32, size_t(n_frames) * H * W * 4,
                                   cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return FSQ_OK;
}
