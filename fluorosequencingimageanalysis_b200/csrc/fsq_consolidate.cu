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
// of the rival graph, which are tiny (the ~5 accepted candidates around one spot).  So:
//   cons_count / cons_scan / cons_scatter   order-preserving compaction of the accepted candidates (R^2 gate)
//   cons_adj        every accepted candidate lists its rivals, in raster order (one scan of its window rows;
//                   headers are fetched eight at a time, predicates are branch-free)
//   cons_component  breadth-first walk over the rival lists; a candidate that meets an EARLIER member leaves at
//                   once, so only the first member of each component finishes the walk -- and then replays the
//                   reference's loop over the members in raster order (no atomics, no iteration to convergence)
//   cons_fallback   frames whose rival graph exceeds the fixed limits of those two kernels (> 16 rivals, > 64 members:
//                   synthetic / extremely dense inputs) are redone by one thread each, scanning window rows directly
//   cons_finish     states, python-2 rounding of the fitted centre -> new key, the reference's collision assert
//                   (:518) as a flag, survivors per frame
// (A first version -- lock-free union-find + one thread per component scanning window rows -- ran with 2.3 of 32
//  lanes active and took 0.36 ms per 40-frame batch; ncu: profiles/r01j_cons_unionfind_kernels.txt.)
#include "fsq_common.cuh"

namespace fsq {

// One accepted candidate = a 16-byte header (what the window scans read: eight of them are fetched at a time, the
// scans are latency chains otherwise) + the three doubles the rival test needs (read only for pixels in the window).
struct __align__(16) ConsHdr { int f, h, w, idx; };     // frame, candidate pixel, index into the candidate arrays
struct __align__(8) ConsVal { double h0, w0, r2; };      // fitted centre (image coordinates), R^2
constexpr int CONS_BATCH = 8;
constexpr int CONS_MAXDEG = 16;     // rivals per accepted candidate (config-1 data: mean 3.6, max 7)
constexpr int CONS_MAXCOMP = 64;    // members per component (config-1 data: mean 4.2, max 12)
                                    // a frame beyond either limit is redone by cons_fallback_kernel (one thread, exact)

struct ConsScratch {
    long long* m_total;        // [1] number of accepted candidates (header)
    int* blockcount;           // [nb + 1] accepted per block of 256 candidates -> exclusive offsets
    unsigned char* alive;      // [n]
    unsigned char* deg;        // [n] number of rivals
    short* adj;                // [n][CONS_MAXDEG] rival positions relative to the own position, raster order
    int* frame_over;           // [F] frame holds a candidate / component beyond CONS_MAXDEG / CONS_MAXCOMP
    ConsHdr* hdr;              // [n]
    ConsVal* val;              // [n]
    int* heads;                // [n] accepted candidates without an earlier rival: the only ones that can be the first
                               //     member of a component (cons_adj appends them in any order; m_total[1] = their number)
};

static inline long long cons_align(long long x) { return (x + 255) & ~255LL; }

static long long cons_carve(ConsScratch* s, char* base, long long n, int F) {
    long long off = 0;
    auto take = [&](long long bytes) { char* p = base ? base + off : nullptr; off += cons_align(bytes); return p; };
    const long long nb = (n + 255) / 256;
    char* p;
    p = take(64);                          if (s) s->m_total = (long long*)p;
    p = take((nb + 1) * 4);                if (s) s->blockcount = (int*)p;
    p = take(n);                           if (s) s->alive = (unsigned char*)p;
    p = take(n);                           if (s) s->deg = (unsigned char*)p;
    p = take(n * 2 * CONS_MAXDEG);         if (s) s->adj = (short*)p;
    p = take(((long long)F + 1) * 4);      if (s) s->frame_over = (int*)p;
    p = take(n * (long long)sizeof(ConsHdr)); if (s) s->hdr = (ConsHdr*)p;
    p = take(n * (long long)sizeof(ConsVal)); if (s) s->val = (ConsVal*)p;
    p = take(n * 4);                       if (s) s->heads = (int*)p;
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
    if (tid == 0) { m_total[0] = carry; m_total[1] = 0; m_total[2] = 0; }   // [1]: length of the component-head list (cons_adj_kernel
                                                                           // appends); [2]: "some accepted fit has its centre more than
                                                                           // half a pixel from its candidate pixel" (cons_scatter_kernel)
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
        ConsHdr hd;
        hd.f = cand_frame[i]; hd.h = cand_hw[2 * i]; hd.w = cand_hw[2 * i + 1]; hd.idx = (int)i;
        ConsVal v;
        v.h0 = fit[i * 12 + 0]; v.w0 = fit[i * 12 + 1]; v.r2 = fit[i * 12 + 8];
        // pflib fits keep their centre within half a pixel of the candidate pixel (centre limits [2, 3] in the 5x5 window);
        // anything else (NaN included) makes cons_finish_kernel run the collision check of pflib.py:518 -- see there
        if (!(fabs(v.h0 - (double)hd.h) <= 0.5 && fabs(v.w0 - (double)hd.w) <= 0.5)) s.m_total[2] = 1;
        *reinterpret_cast<int4*>(&s.hdr[pos]) = *reinterpret_cast<const int4*>(&hd);
        s.val[pos] = v;
        s.alive[pos] = 1;
    }
}

// (h_0 - h_0')^2 + (w_0 - w_0')^2 > radius^2 in the reference's arithmetic (no contraction)
__device__ __forceinline__ bool cons_far(const ConsVal& a, const ConsVal& b, double rr) {
    const double dh = a.h0 - b.h0, dw = a.w0 - b.w0;
    return __dadd_rn(__dmul_rn(dh, dh), __dmul_rn(dw, dw)) > rr;
}

__device__ __forceinline__ ConsHdr cons_ld_hdr(const ConsHdr* hdr, long long i) {
    const int4 q = __ldg(reinterpret_cast<const int4*>(hdr + i));
    ConsHdr r; r.f = q.x; r.h = q.y; r.w = q.z; r.idx = q.w;
    return r;
}

// rival lists: accepted candidates of the same frame whose pixel lies within +-reach on both axes (the window of
// pflib.py:492-495) and whose fitted centre lies within the radius (:505), in raster order
__global__ void __launch_bounds__(128)
cons_adj_kernel(ConsScratch s, int reach, double rr) {
    const long long m = *s.m_total;
    const long long a = (long long)blockIdx.x * 128 + threadIdx.x;
    if (a >= m) return;
    const ConsHdr ha = cons_ld_hdr(s.hdr, a);
    const ConsVal va = s.val[a];
    // When every accepted centre lies within half a pixel of its candidate pixel (m_total[2] == 0, pflib fits), two
    // candidates whose pixels are reach = radius + 2 rows apart have centres more than the radius apart: the outermost
    // window rows cannot hold a rival and are not scanned (13 -> 11 rows at the default radius).
    const int rreach = (s.m_total[2] == 0 && reach > 1) ? reach - 1 : reach;
    // first accepted position inside the window rows (the list is sorted by frame, row, column)
    long long lo = a;
    for (long long base = a - 1; base >= 0; base -= CONS_BATCH) {
        ConsHdr hb[CONS_BATCH];
#pragma unroll
        for (int q = 0; q < CONS_BATCH; ++q) hb[q] = cons_ld_hdr(s.hdr, base - q >= 0 ? base - q : 0);
        bool last_in = false;
#pragma unroll
        for (int q = 0; q < CONS_BATCH; ++q) {
            const long long b = base - q;
            const bool in = (b >= 0) && (hb[q].f == ha.f) && (hb[q].h >= ha.h - rreach);
            lo = in ? b : lo;                                  // `in` is monotone along the scan
            last_in = in;
        }
        if (!last_in) break;
    }
    int deg = 0;
    bool over = false, earlier = false;
    short* adj = s.adj + a * CONS_MAXDEG;
    for (long long base = lo; base < m; base += CONS_BATCH) {
        ConsHdr hj[CONS_BATCH];
#pragma unroll
        for (int q = 0; q < CONS_BATCH; ++q) hj[q] = cons_ld_hdr(s.hdr, base + q < m ? base + q : m - 1);
        bool last_in = false;
#pragma unroll
        for (int q = 0; q < CONS_BATCH; ++q) {
            const long long j = base + q;
            const bool in = (j < m) && (hj[q].f == ha.f) && (hj[q].h <= ha.h + rreach);
            last_in = in;
            if (in && j != a && abs(hj[q].w - ha.w) <= reach) {
                const ConsVal vj = s.val[j];
                if (!cons_far(va, vj, rr)) {
                    const long long off = j - a;
                    earlier |= (off < 0);
                    if (deg < CONS_MAXDEG && off >= -32768 && off <= 32767) adj[deg] = (short)off;
                    else over = true;
                    ++deg;
                }
            }
        }
        if (!last_in) break;
    }
    // an overflowing candidate (too many rivals, or one out of int16 reach) flags its frame for cons_fallback and
    // publishes NO rival list: a partly written list must never be walked (the scratch is uninitialised memory)
    if (over) { atomicOr(&s.frame_over[ha.f], 1); deg = 0; }
    s.deg[a] = (unsigned char)deg;
    // a candidate with an earlier rival cannot be the first member of its component: only the others start a walk in
    // cons_component_kernel, packed (one walk per lane instead of one per ~4 lanes).  The order of the list does not
    // matter -- components are independent.  (The compiler aggregates the atomic over the warp.)
    if (!earlier && !over) s.heads[atomicAdd(reinterpret_cast<unsigned long long*>(s.m_total + 1), 1ull)] = (int)a;
}

// one walk per accepted candidate; only the first member of a component completes it and replays pflib.py:479-512
// (Loading a rival list as two 16-byte vectors and replaying on a private copy of the members' R^2 / alive bits was
//  measured: 70 us instead of 58 us per 40-frame batch -- the local arrays cost more than the loads they save.)
__global__ void __launch_bounds__(128)
cons_component_kernel(ConsScratch s) {
    const long long nh = s.m_total[1];
    const long long hpos = (long long)blockIdx.x * 128 + threadIdx.x;
    if (hpos >= nh) return;
    const long long a = s.heads[hpos];
    if (s.frame_over[s.hdr[a].f]) return;                      // the whole frame is redone by cons_fallback_kernel
    int member[CONS_MAXCOMP];
    int cnt = 1;
    member[0] = (int)a;
    for (int i = 0; i < cnt; ++i) {
        const int k = member[i];
        const int dk = s.deg[k];
        const short* adj = s.adj + (long long)k * CONS_MAXDEG;
        for (int r = 0; r < dk; ++r) {
            const int j = k + adj[r];
            if (j < (int)a) return;                            // an earlier member exists: it does the work
            bool seen = false;
            for (int t = 0; t < cnt; ++t) seen |= (member[t] == j);
            if (!seen) {
                if (cnt < CONS_MAXCOMP) member[cnt++] = j;
                else atomicOr(&s.frame_over[s.hdr[a].f], 1);
            }
        }
    }
    // raster order = ascending position (insertion sort; the walk leaves the list nearly sorted)
    for (int i = 1; i < cnt; ++i) {
        const int v = member[i];
        int t = i - 1;
        while (t >= 0 && member[t] > v) { member[t + 1] = member[t]; --t; }
        member[t + 1] = v;
    }
    for (int i = 0; i < cnt; ++i) {
        const int k = member[i];
        if (!s.alive[k]) continue;                             // "skip pixels that have had their psfs deleted" (:481)
        const double r2k = s.val[k].r2;
        const int dk = s.deg[k];
        const short* adj = s.adj + (long long)k * CONS_MAXDEG;
        for (int r = 0; r < dk; ++r) {                         // itertools.product(h_range, w_range): raster order
            const int j = k + adj[r];
            if (!s.alive[j]) continue;
            if (r2k > s.val[j].r2) s.alive[j] = 0;             // :508-509
            else { s.alive[k] = 0; break; }                    // :510-512
        }
    }
}

// Exact fallback for frames whose rival graph exceeds the limits of the two kernels above (extremely dense or
// synthetic inputs): one thread replays pflib.py:479-512 over the whole frame, scanning window rows directly.
__global__ void __launch_bounds__(32)
cons_fallback_kernel(ConsScratch s, int F, int reach, double rr) {
    const int f = blockIdx.x * 32 + threadIdx.x;
    if (f >= F || !s.frame_over[f]) return;
    const long long m = *s.m_total;
    long long beg, end;                                        // [beg, end) = accepted candidates of frame f (sorted by frame)
    {
        long long a = 0, b = m;
        while (a < b) { const long long c = (a + b) / 2; if (s.hdr[c].f < f) a = c + 1; else b = c; }
        beg = a;
        b = m;
        while (a < b) { const long long c = (a + b) / 2; if (s.hdr[c].f <= f) a = c + 1; else b = c; }
        end = a;
    }
    for (long long k = beg; k < end; ++k) s.alive[k] = 1;
    long long lo = beg;
    for (long long k = beg; k < end; ++k) {
        if (!s.alive[k]) continue;
        const ConsHdr hk = s.hdr[k];
        const ConsVal vk = s.val[k];
        while (s.hdr[lo].h < hk.h - reach) ++lo;
        for (long long j = lo; j < end; ++j) {
            const ConsHdr hj = s.hdr[j];
            if (hj.h > hk.h + reach) break;
            if (j == k || abs(hj.w - hk.w) > reach) continue;
            const ConsVal vj = s.val[j];
            if (cons_far(vk, vj, rr) || !s.alive[j]) continue;
            if (vk.r2 > vj.r2) s.alive[j] = 0;
            else { s.alive[k] = 0; break; }
        }
    }
}

__global__ void __launch_bounds__(128)
cons_finish_kernel(ConsScratch s, int reach, unsigned char* __restrict__ psf_state, int32_t* __restrict__ psf_key,
                   unsigned long long* __restrict__ n_psf, int32_t* __restrict__ flags) {
    const long long m = *s.m_total;
    const long long a = (long long)blockIdx.x * 128 + threadIdx.x;
    if (a >= m) return;
    const ConsHdr ha = cons_ld_hdr(s.hdr, a);
    if (!s.alive[a]) { psf_state[ha.idx] = 1; return; }
    const ConsVal va = s.val[a];
    const int kh = (int)round(va.h0), kw = (int)round(va.w0);  // python-2 round(): half away from zero (:515)
    const bool moved = (kh != ha.h) || (kw != ha.w);
    psf_state[ha.idx] = moved ? 3 : 2;
    psf_key[2 * ha.idx] = kh; psf_key[2 * ha.idx + 1] = kw;
    if (n_psf) {                                               // one atomic per (warp, frame): candidates are sorted by frame
        const unsigned peers = __match_any_sync(__activemask(), ha.f);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&n_psf[ha.f], (unsigned long long)__popc(peers));
    }
    if (moved && s.m_total[2] != 0) {
        // :518 asserts that the new key is free at the moment of the move: earlier survivors sit at their FINAL
        // keys by then, later ones still at their candidate pixels.  The scan is skipped when EVERY accepted fit of the
        // batch has its centre within half a pixel of its candidate pixel on both axes (m_total[2] == 0, the case for pflib
        // fits): two survivors that end on one key then have centres within one pixel per axis of each other
        // (<= 1.42 px <= the radius, which is at least 2) and candidate pixels within two (<= radius + 2, the window), i.e.
        // they are rivals -- and of two rivals the consolidation never leaves both alive.  Otherwise: checked within the
        // rival reach.
        long long b = a;
        while (b > 0 && s.hdr[b - 1].f == ha.f && s.hdr[b - 1].h >= kh - reach - 1) --b;
        bool stop = false;
        for (; b < m && !stop; b += CONS_BATCH) {
            ConsHdr hb[CONS_BATCH];
#pragma unroll
            for (int q = 0; q < CONS_BATCH; ++q) hb[q] = cons_ld_hdr(s.hdr, b + q < m ? b + q : m - 1);
#pragma unroll
            for (int q = 0; q < CONS_BATCH; ++q) {
                const long long c = b + q;
                if (stop || c >= m || hb[q].f != ha.f || hb[q].h > kh + reach + 1) { stop = true; continue; }
                if (c == a || abs(hb[q].w - kw) > reach + 1 || !s.alive[c]) continue;
                const int bh = c < a ? (int)round(s.val[c].h0) : hb[q].h, bw = c < a ? (int)round(s.val[c].w0) : hb[q].w;
                if (bh == kh && bw == kw) atomicOr(flags, 1);
            }
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
    if (st >= 2) {                                             // one atomic per (warp, frame, class)
        const int key = cand_frame[i] * 2 + (st - 2);
        const unsigned peers = __match_any_sync(__activemask(), key);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1))
            atomicAdd((unsigned long long*)(st == 2 ? &s.frameU[cand_frame[i]] : &s.frameM[cand_frame[i]]), (unsigned long long)__popc(peers));
    }
    const int cu = __syncthreads_count(st == 2);
    const int cm = __syncthreads_count(st == 3);
    if (threadIdx.x == 0) { s.blockU[blockIdx.x] = cu; s.blockM[blockIdx.x] = cm; }
}

// exclusive scans of two arrays at once (one block of 1024 threads; the two share every barrier)
template <typename T>
__device__ void block_exclusive_scan2(T* u, T* v, long long count, long long* u_total, long long* v_total) {
    __shared__ long long bu[1024], bv[1024];
    __shared__ long long cu, cv;
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) { cu = 0; cv = 0; }
    __syncthreads();
    for (long long base = 0; base < count; base += 1024) {
        const long long b = base + tid;
        const long long xu = b < count ? (long long)u[b] : 0, xv = b < count ? (long long)v[b] : 0;
        bu[tid] = xu; bv[tid] = xv;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const long long tu = tid >= d ? bu[tid - d] : 0, tv = tid >= d ? bv[tid - d] : 0;
            __syncthreads();
            bu[tid] += tu; bv[tid] += tv;
            __syncthreads();
        }
        if (b < count) { u[b] = (T)(cu + bu[tid] - xu); v[b] = (T)(cv + bv[tid] - xv); }
        __syncthreads();
        if (tid == 1023) { cu += bu[1023]; cv += bv[1023]; }
        __syncthreads();
    }
    if (tid == 0) { if (u_total) *u_total = cu; if (v_total) *v_total = cv; }
    __syncthreads();
}

__global__ void __launch_bounds__(1024)
pack_scan_kernel(PackScratch s, long long nb, int F, long long* __restrict__ psf_base) {
    block_exclusive_scan2(s.blockU, s.blockM, nb, nullptr, nullptr);
    block_exclusive_scan2(s.frameU, s.frameM, F, &s.frameU[F], &s.frameM[F]);
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

extern "C" int64_t fsq_consolidate_scratch_bytes(int64_t n, int n_frames) {
    return cons_carve(nullptr, nullptr, n > 0 ? n : 0, n_frames > 0 ? n_frames : 0);
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
    if (n_frames <= 0) { set_error("fsq_consolidate: n_frames must be positive"); return FSQ_E_ARG; }
    if (scratch_bytes < fsq_consolidate_scratch_bytes(n, n_frames)) {
        set_error("fsq_consolidate: scratch too small (%lld < %lld)", (long long)scratch_bytes, (long long)fsq_consolidate_scratch_bytes(n, n_frames));
        return FSQ_E_CAPACITY;
    }
    ConsScratch s;
    cons_carve(&s, (char*)scratch, n, n_frames);
    FSQ_CUDA_CHECK(cudaMemsetAsync(s.frame_over, 0, sizeof(int) * (size_t)n_frames, st));
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
    cons_adj_kernel<<<g128, 128, 0, st>>>(s, reach, rr);
    FSQ_LAUNCH_CHECK();
    cons_component_kernel<<<g128, 128, 0, st>>>(s);
    FSQ_LAUNCH_CHECK();
    cons_fallback_kernel<<<(unsigned)((n_frames + 31) / 32), 32, 0, st>>>(s, n_frames, reach, rr);
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
