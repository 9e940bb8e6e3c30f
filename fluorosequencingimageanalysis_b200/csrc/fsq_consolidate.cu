// R^2 gate, rival consolidation and re-key of pflib.find_peptides (pflib.py:466-468, 479-519) on the packed
// per-candidate records of a batch of frames -- the last per-candidate Python loop of the reference.
//
// The reference walks its dictionary of accepted fits in insertion (= raster) order; for the PSF at pixel
// (h, w) it visits every other accepted PSF whose candidate pixel lies within +-(radius + 2) of (h, w), in
// raster order, and whenever the two FITTED centres are within `radius` of each other it deletes the one with
// the lower-or-equal R^2 (`>` at :508, so a tie deletes the visiting PSF and ends its scan).  The outcome
// depends on the visiting order, so it cannot be replaced by "keep the local R^2 maximum".
//
// What makes it parallel: two PSFs interact only if they are rivals (pixels within radius + 2 on both axes AND
// centres within radius).  The sequential process therefore factorises exactly over the connected components
// of the rival graph, which are tiny (the ~3-6 accepted candidates around one spot).  So:
//   cons_count / cons_scan / cons_scatter   order-preserving compaction of the accepted candidates (R^2 gate)
//   cons_union     lock-free union-find over rival pairs (larger root hooked under the smaller one, so a
//                  component's root is its FIRST member in raster order)
//   cons_flatten   label = root
//   cons_process   one thread per component replays the reference's loop over the component's members in
//                  raster order (members are found by scanning forward while the row gap stays <= radius + 2)
//   cons_finish    states, python-2 rounding of the fitted centre -> new key, the reference's collision assert
//                  (:518) as a flag, survivors per frame
#include "fsq_common.cuh"

namespace fsq {

struct __align__(8) ConsRec {
    int f, h, w, idx;          // frame, candidate pixel, index into the candidate arrays
    double h0, w0, r2;         // fitted centre (image coordinates), R^2
};
static_assert(sizeof(ConsRec) == 40, "ConsRec must be 40 bytes");

struct ConsScratch {
    long long* m_total;        // [1] number of accepted candidates (header)
    int* blockcount;           // [nb + 1] accepted per block of 256 candidates -> exclusive offsets
    int* parent;               // [n] union-find forest over accepted positions, then the flattened labels
    unsigned char* alive;      // [n]
    ConsRec* rec;              // [n]
};

static inline long long cons_align(long long x) { return (x + 255) & ~255LL; }

static long long cons_carve(ConsScratch* s, char* base, long long n) {
    long long off = 0;
    auto take = [&](long long bytes) { char* p = base ? base + off : nullptr; off += cons_align(bytes); return p; };
    const long long nb = (n + 255) / 256;
    char* p;
    p = take(64);                          if (s) s->m_total = (long long*)p;
    p = take((nb + 1) * 4);                if (s) s->blockcount = (int*)p;
    p = take(n * 4);                       if (s) s->parent = (int*)p;
    p = take(n);                           if (s) s->alive = (unsigned char*)p;
    p = take(n * (long long)sizeof(ConsRec)); if (s) s->rec = (ConsRec*)p;
    return off;
}

__device__ __forceinline__ long long cons_n(long long n, const long long* n_dev) {
    if (!n_dev) return n;
    const long long nd = *n_dev;
    return nd < n ? nd : n;
}

// accepted = NOT (r_2 < threshold): the reference `continue`s only on '<' (pflib.py:466), so a NaN R^2 stays in
__global__ void __launch_bounds__(256)
cons_count_kernel(const double* __restrict__ fit, long long n, const long long* __restrict__ n_dev, double thr,
                  const int32_t* __restrict__ cand_hw, unsigned char* __restrict__ psf_state,
                  int32_t* __restrict__ psf_key, int* __restrict__ blockcount) {
    const long long nt = cons_n(n, n_dev);
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    bool keep = false;
    if (i < nt) {
        keep = !(fit[i * 12 + 8] < thr);
        psf_state[i] = 0;
        psf_key[2 * i] = cand_hw[2 * i]; psf_key[2 * i + 1] = cand_hw[2 * i + 1];
    }
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) blockcount[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024)
cons_scan_kernel(int* __restrict__ blockcount, long long nb, long long* __restrict__ m_total) {
    __shared__ long long buf[1024];
    __shared__ long long carry;
    const int tid = threadIdx.x;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nb; base += 1024) {
        const long long b = base + tid;
        const long long v = b < nb ? blockcount[b] : 0;
        buf[tid] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const long long t = tid >= d ? buf[tid - d] : 0;
            __syncthreads();
            buf[tid] += t;
            __syncthreads();
        }
        if (b < nb) blockcount[b] = (int)(carry + buf[tid] - v);
        __syncthreads();
        if (tid == 1023) carry += buf[1023];
        __syncthreads();
    }
    if (tid == 0) *m_total = carry;
}

__global__ void __launch_bounds__(256)
cons_scatter_kernel(const double* __restrict__ fit, long long n, const long long* __restrict__ n_dev, double thr,
                    const int32_t* __restrict__ cand_hw, const int32_t* __restrict__ cand_frame,
                    const int* __restrict__ blockoff, ConsScratch s) {
    __shared__ int wsum[8];
    const long long nt = cons_n(n, n_dev);
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool keep = (i < nt) && !(fit[i * 12 + 8] < thr);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int k = 0; k < warp; ++k) before += wsum[k];
    if (keep) {
        const int pos = blockoff[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
        ConsRec r;
        r.f = cand_frame[i]; r.h = cand_hw[2 * i]; r.w = cand_hw[2 * i + 1]; r.idx = (int)i;
        r.h0 = fit[i * 12 + 0]; r.w0 = fit[i * 12 + 1]; r.r2 = fit[i * 12 + 8];
        s.rec[pos] = r;
        s.parent[pos] = pos;
        s.alive[pos] = 1;
    }
}

// (h_0 - h_0')^2 + (w_0 - w_0')^2 > radius^2 in the reference's arithmetic (no contraction)
__device__ __forceinline__ bool cons_far(const ConsRec& a, const ConsRec& b, double rr) {
    const double dh = a.h0 - b.h0, dw = a.w0 - b.w0;
    return __dadd_rn(__dmul_rn(dh, dh), __dmul_rn(dw, dw)) > rr;
}

__device__ __forceinline__ int cons_find(const volatile int* parent, int x) {
    int p;
    while ((p = parent[x]) != x) x = p;
    return x;
}

__global__ void __launch_bounds__(128)
cons_union_kernel(ConsScratch s, int reach, double rr) {
    const long long m = *s.m_total;
    const long long a = (long long)blockIdx.x * 128 + threadIdx.x;
    if (a >= m) return;
    const ConsRec ra = s.rec[a];
    for (long long b = a - 1; b >= 0; --b) {                   // earlier rivals; later ones find this PSF themselves
        const ConsRec rb = s.rec[b];
        if (rb.f != ra.f || rb.h < ra.h - reach) break;
        if (abs(rb.w - ra.w) > reach || cons_far(ra, rb, rr)) continue;
        int x = (int)a, y = (int)b;
        for (;;) {
            x = cons_find(s.parent, x); y = cons_find(s.parent, y);
            if (x == y) break;
            if (x < y) { const int t = x; x = y; y = t; }      // hook the larger root under the smaller
            if (atomicCAS(&s.parent[x], x, y) == x) break;
        }
    }
}

__global__ void __launch_bounds__(128)
cons_flatten_kernel(ConsScratch s, int* __restrict__ label) {
    const long long m = *s.m_total;
    const long long a = (long long)blockIdx.x * 128 + threadIdx.x;
    if (a >= m) return;
    label[a] = cons_find(s.parent, (int)a);
}

// one thread per component: the reference's loop (pflib.py:479-512) over the members in raster order
__global__ void __launch_bounds__(128)
cons_process_kernel(ConsScratch s, const int* __restrict__ label, int reach, double rr) {
    const long long m = *s.m_total;
    const long long a = (long long)blockIdx.x * 128 + threadIdx.x;
    if (a >= m || label[a] != (int)a) return;
    const int f = s.rec[a].f;
    int maxrow = s.rec[a].h;
    for (long long k = a; k < m; ++k) {
        const ConsRec rk = s.rec[k];
        if (rk.f != f || rk.h > maxrow + reach) break;         // no member can lie further down (rows of rivals differ by <= reach)
        if (label[k] != (int)a) continue;
        maxrow = rk.h;
        if (!s.alive[k]) continue;                             // "skip pixels that have had their psfs deleted" (:481)
        long long j = k;
        while (j > 0 && s.rec[j - 1].f == f && s.rec[j - 1].h >= rk.h - reach) --j;
        for (; j < m; ++j) {                                   // itertools.product(h_range, w_range): raster order
            const ConsRec rj = s.rec[j];
            if (rj.f != f || rj.h > rk.h + reach) break;
            if (j == k || abs(rj.w - rk.w) > reach) continue;
            if (cons_far(rk, rj, rr)) continue;                // (other components' alive flags are never read)
            if (!s.alive[j]) continue;
            if (rk.r2 > rj.r2) s.alive[j] = 0;                 // :508-509
            else { s.alive[k] = 0; break; }                    // :510-512
        }
    }
}

__global__ void __launch_bounds__(128)
cons_finish_kernel(ConsScratch s, int reach, unsigned char* __restrict__ psf_state, int32_t* __restrict__ psf_key,
                   unsigned long long* __restrict__ n_psf, int32_t* __restrict__ flags) {
    const long long m = *s.m_total;
    const long long a = (long long)blockIdx.x * 128 + threadIdx.x;
    if (a >= m) return;
    const ConsRec ra = s.rec[a];
    if (!s.alive[a]) { psf_state[ra.idx] = 1; return; }
    const int kh = (int)round(ra.h0), kw = (int)round(ra.w0);  // python-2 round(): half away from zero (:515)
    const bool moved = (kh != ra.h) || (kw != ra.w);
    psf_state[ra.idx] = moved ? 3 : 2;
    psf_key[2 * ra.idx] = kh; psf_key[2 * ra.idx + 1] = kw;
    if (n_psf) atomicAdd(&n_psf[ra.f], 1ull);
    if (moved) {
        // :518 asserts that the new key is free at the moment of the move: earlier survivors sit at their FINAL
        // keys by then, later ones still at their candidate pixels.  (Cannot happen when the fitted centre lies
        // within 0.5 px of the candidate pixel, as it does for pflib fits; checked within the rival reach.)
        long long b = a;
        while (b > 0 && s.rec[b - 1].f == ra.f && s.rec[b - 1].h >= kh - reach - 1) --b;
        for (; b < m; ++b) {
            const ConsRec rb = s.rec[b];
            if (rb.f != ra.f || rb.h > kh + reach + 1) break;
            if (b == a || !s.alive[b]) continue;
            const int bh = b < a ? (int)round(rb.h0) : rb.h, bw = b < a ? (int)round(rb.w0) : rb.w;
            if (bh == kh && bw == kw) atomicOr(flags, 1);
        }
    }
}

// ---- final PSFs packed in the reference's dictionary order ------------------------------------------------
//  position of survivor i of frame f:  kept its key (state 2):  Mbase[f]     + U(i)
//                                      re-keyed   (state 3):  Ubase[f + 1] + M(i)
//  with U(i) / M(i) = number of state-2 / state-3 candidates before i in the whole batch and Ubase / Mbase their
//  exclusive per-frame offsets: frames in sequence, inside a frame the PSFs that kept their key in raster order,
//  then the re-keyed ones (`del` + `setdefault` appends, pflib.py:516-519).
struct PackScratch {
    int* blockU; int* blockM;            // [nb]   per block of 256 candidates -> exclusive offsets
    long long* frameU; long long* frameM;  // [F + 1] per frame               -> exclusive offsets
};

static long long pack_carve(PackScratch* s, char* base, long long n, int F) {
    long long off = 0;
    auto take = [&](long long bytes) { char* p = base ? base + off : nullptr; off += cons_align(bytes); return p; };
    const long long nb = (n + 255) / 256;
    char* p;
    p = take(nb * 4);                 if (s) s->blockU = (int*)p;
    p = take(nb * 4);                 if (s) s->blockM = (int*)p;
    p = take(((long long)F + 1) * 8); if (s) s->frameU = (long long*)p;
    p = take(((long long)F + 1) * 8); if (s) s->frameM = (long long*)p;
    return off;
}

__global__ void __launch_bounds__(256)
pack_count_kernel(const unsigned char* __restrict__ state, const int32_t* __restrict__ cand_frame, long long n,
                  const long long* __restrict__ n_dev, PackScratch s) {
    const long long nt = cons_n(n, n_dev);
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const int st = i < nt ? state[i] : 0;
    if (st == 2) atomicAdd((unsigned long long*)&s.frameU[cand_frame[i]], 1ull);
    if (st == 3) atomicAdd((unsigned long long*)&s.frameM[cand_frame[i]], 1ull);
    const int cu = __syncthreads_count(st == 2);
    const int cm = __syncthreads_count(st == 3);
    if (threadIdx.x == 0) { s.blockU[blockIdx.x] = cu; s.blockM[blockIdx.x] = cm; }
}

template <typename T>
__device__ void block_exclusive_scan(T* v, long long count, long long* total_out) {     // one block of 1024 threads
    __shared__ long long buf[1024];
    __shared__ long long carry;
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < count; base += 1024) {
        const long long b = base + tid;
        const long long x = b < count ? (long long)v[b] : 0;
        buf[tid] = x;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const long long t = tid >= d ? buf[tid - d] : 0;
            __syncthreads();
            buf[tid] += t;
            __syncthreads();
        }
        if (b < count) v[b] = (T)(carry + buf[tid] - x);
        __syncthreads();
        if (tid == 1023) carry += buf[1023];
        __syncthreads();
    }
    if (tid == 0 && total_out) *total_out = carry;
    __syncthreads();
}

__global__ void __launch_bounds__(1024)
pack_scan_kernel(PackScratch s, long long nb, int F, long long* __restrict__ psf_base) {
    block_exclusive_scan(s.blockU, nb, nullptr);
    block_exclusive_scan(s.blockM, nb, nullptr);
    block_exclusive_scan(s.frameU, F, &s.frameU[F]);
    block_exclusive_scan(s.frameM, F, &s.frameM[F]);
    for (int f = threadIdx.x; f <= F; f += 1024) psf_base[f] = s.frameU[f] + s.frameM[f];
}

__global__ void __launch_bounds__(256)
pack_scatter_kernel(const unsigned char* __restrict__ state, const int32_t* __restrict__ key,
                    const int32_t* __restrict__ cand_frame, const double* __restrict__ fit, long long n,
                    const long long* __restrict__ n_dev, PackScratch s, double* __restrict__ psf_fit,
                    int32_t* __restrict__ psf_int, long long cap_psf) {
    __shared__ int wu[8], wm[8];
    const long long nt = cons_n(n, n_dev);
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int st = i < nt ? state[i] : 0;
    const unsigned mu = __ballot_sync(0xffffffffu, st == 2), mm = __ballot_sync(0xffffffffu, st == 3);
    if (lane == 0) { wu[warp] = __popc(mu); wm[warp] = __popc(mm); }
    __syncthreads();
    if (st < 2) return;
    int bu = 0, bm = 0;
    for (int k = 0; k < warp; ++k) { bu += wu[k]; bm += wm[k]; }
    const unsigned below = (1u << lane) - 1u;
    const int f = cand_frame[i];
    long long pos;
    if (st == 2) pos = s.frameM[f] + s.blockU[blockIdx.x] + bu + __popc(mu & below);
    else         pos = s.frameU[f + 1] + s.blockM[blockIdx.x] + bm + __popc(mm & below);
    if (pos >= cap_psf) return;
    const double2* src = reinterpret_cast<const double2*>(fit + i * 12);
    double2* dst = reinterpret_cast<double2*>(psf_fit + pos * 12);
#pragma unroll
    for (int q = 0; q < 6; ++q) dst[q] = src[q];
    *reinterpret_cast<int4*>(psf_int + pos * 4) = make_int4(f, key[2 * i], key[2 * i + 1], (int)i);
}

}  // namespace fsq

using namespace fsq;

extern "C" int64_t fsq_consolidate_scratch_bytes(int64_t n) {
    return cons_carve(nullptr, nullptr, n > 0 ? n : 0) + cons_align((n > 0 ? n : 0) * 4);
}

extern "C" int fsq_consolidate(const int32_t* cand_hw, const int32_t* cand_frame, const double* out_fit, int64_t n,
                               const int64_t* n_dev, int n_frames, double r_2_threshold, int consolidation_radius,
                               uint8_t* psf_state, int32_t* psf_key, int64_t* n_psf, int32_t* flags,
                               void* scratch, int64_t scratch_bytes, void* stream) {
    if (consolidation_radius < 2) { set_error("consolidation_radius must be at least 2"); return FSQ_E_ARG; }   // pflib.py:431-432
    if (n < 0 || n > 2147483647LL) { set_error("fsq_consolidate: n out of range"); return FSQ_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (flags) FSQ_CUDA_CHECK(cudaMemsetAsync(flags, 0, 4, st));
    if (n_psf && n_frames > 0) FSQ_CUDA_CHECK(cudaMemsetAsync(n_psf, 0, sizeof(int64_t) * (size_t)n_frames, st));
    if (n == 0) return FSQ_OK;
    if (!cand_hw || !cand_frame || !out_fit || !psf_state || !psf_key || !flags || !scratch) {
        set_error("fsq_consolidate: NULL pointer argument");
        return FSQ_E_ARG;
    }
    if (scratch_bytes < fsq_consolidate_scratch_bytes(n)) {
        set_error("fsq_consolidate: scratch too small (%lld < %lld)", (long long)scratch_bytes, (long long)fsq_consolidate_scratch_bytes(n));
        return FSQ_E_CAPACITY;
    }
    ConsScratch s;
    const long long used = cons_carve(&s, (char*)scratch, n);
    int* label = (int*)((char*)scratch + used);
    const long long nb = (n + 255) / 256;
    const unsigned g128 = (unsigned)((n + 127) / 128);
    const int reach = consolidation_radius + 2;                                          // pflib.py:492-495
    const double rr = (double)consolidation_radius * (double)consolidation_radius;
    const long long* nd = (const long long*)n_dev;
    cons_count_kernel<<<(unsigned)nb, 256, 0, st>>>(out_fit, n, nd, r_2_threshold, cand_hw, psf_state, psf_key, s.blockcount);
    FSQ_LAUNCH_CHECK();
    cons_scan_kernel<<<1, 1024, 0, st>>>(s.blockcount, nb, s.m_total);
    FSQ_LAUNCH_CHECK();
    cons_scatter_kernel<<<(unsigned)nb, 256, 0, st>>>(out_fit, n, nd, r_2_threshold, cand_hw, cand_frame, s.blockcount, s);
    FSQ_LAUNCH_CHECK();
    cons_union_kernel<<<g128, 128, 0, st>>>(s, reach, rr);
    FSQ_LAUNCH_CHECK();
    cons_flatten_kernel<<<g128, 128, 0, st>>>(s, label);
    FSQ_LAUNCH_CHECK();
    cons_process_kernel<<<g128, 128, 0, st>>>(s, label, reach, rr);
    FSQ_LAUNCH_CHECK();
    cons_finish_kernel<<<g128, 128, 0, st>>>(s, reach, psf_state, psf_key, (unsigned long long*)n_psf, flags);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}

extern "C" int64_t fsq_pack_psfs_scratch_bytes(int64_t n, int n_frames) {
    return pack_carve(nullptr, nullptr, n > 0 ? n : 0, n_frames > 0 ? n_frames : 0);
}

extern "C" int fsq_pack_psfs(const uint8_t* psf_state, const int32_t* psf_key, const int32_t* cand_frame,
                             const double* out_fit, int64_t n, const int64_t* n_dev, int n_frames,
                             double* psf_fit, int32_t* psf_int, int64_t* psf_base, int64_t cap_psf,
                             void* scratch, int64_t scratch_bytes, void* stream) {
    if (n < 0 || n > 2147483647LL || n_frames <= 0 || cap_psf < 0) { set_error("fsq_pack_psfs: bad sizes"); return FSQ_E_ARG; }
    if (!psf_base || !scratch) { set_error("fsq_pack_psfs: NULL pointer argument"); return FSQ_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { FSQ_CUDA_CHECK(cudaMemsetAsync(psf_base, 0, sizeof(int64_t) * ((size_t)n_frames + 1), st)); return FSQ_OK; }
    if (!psf_state || !psf_key || !cand_frame || !out_fit || !psf_fit || !psf_int) {
        set_error("fsq_pack_psfs: NULL pointer argument");
        return FSQ_E_ARG;
    }
    if (scratch_bytes < pack_carve(nullptr, nullptr, n, n_frames)) {
        set_error("fsq_pack_psfs: scratch too small");
        return FSQ_E_CAPACITY;
    }
    PackScratch s;
    pack_carve(&s, (char*)scratch, n, n_frames);
    const long long nb = (n + 255) / 256;
    FSQ_CUDA_CHECK(cudaMemsetAsync(s.frameU, 0, sizeof(long long) * ((size_t)n_frames + 1), st));
    FSQ_CUDA_CHECK(cudaMemsetAsync(s.frameM, 0, sizeof(long long) * ((size_t)n_frames + 1), st));
    const long long* nd = (const long long*)n_dev;
    pack_count_kernel<<<(unsigned)nb, 256, 0, st>>>(psf_state, cand_frame, n, nd, s);
    FSQ_LAUNCH_CHECK();
    pack_scan_kernel<<<1, 1024, 0, st>>>(s, nb, n_frames, (long long*)psf_base);
    FSQ_LAUNCH_CHECK();
    pack_scatter_kernel<<<(unsigned)nb, 256, 0, st>>>(psf_state, psf_key, cand_frame, out_fit, n, nd, s, psf_fit, psf_int, (long long)cap_psf);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}
