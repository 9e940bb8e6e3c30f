// Greedy cross-frame particle tracking -- replaces Experiment.greedy_particle_tracking (flexlibrary.py:680-1027),
// the tracker of the experiment path (basic_experiment_script.py: one field, one frame per Edman cycle).
//
// The reference bins every spot by its rounded, drift-corrected position, then walks the frames: the spots of
// frame f-1 join an "ancestor cache" (unpaired spots of ANY earlier frame stay in it, so a spot may skip frames; a
// newer spot in the same pixel replaces an older one), every (ancestor, spot of frame f) pair closer than
// candidate_radius is collected -- ancestors in raster order, candidates in raster order inside the ancestor's
// window -- the pairs are sorted by distance with a STABLE sort, and the sorted list is walked greedily: a pair
// links iff neither end is linked yet.
//
// A greedy walk over a strictly ordered edge list selects exactly the edges that are, at some point, the first
// remaining edge of BOTH their ends ("locally dominant"), so the walk is replayed in parallel rounds: every free
// ancestor picks its best free candidate by the key (distance, candidate's raster position) -- the order of the
// stable sort restricted to that ancestor --, the candidate checks that this ancestor is its own best by
// (distance, ancestor's raster position), and mutual picks link (compare-and-swap on the candidate's ancestor slot,
// so that a candidate whose better suitor was linked in the same round cannot be taken twice).  One thread block
// per field, frames in sequence inside the kernel, two H x W index grids per field (ancestor cache, current frame).
#include "fsq_common.cuh"

namespace fsq {

struct TrackArgs {
    const double* spot_hw;       // [n,2] spot positions (Spot.h, Spot.w)
    const int32_t* seg_start;    // [n_fields * n_frames + 1] spots are sorted by (field, frame)
    const double* cum_off;       // [n_fields, n_frames, 2] cumulative offsets (accumulate_offsets), or NULL
    int n_fields, n_frames, H, W, radius;
    double spot_radius;
    int32_t* anc; int32_t* desc; // [n] links (global spot indices, -1 = none)
    int32_t* bin_hw;             // [n,2] rounded drift-corrected position
    uint8_t* discarded;          // [n]
    int32_t* flags;              // [n_fields] bit 0: two spots of one frame in one pixel (the reference asserts)
    int32_t* grid_cache;         // [n_fields, H, W] scratch, -1 filled
    int32_t* grid_frame;         // [n_fields, H, W] scratch, -1 filled
    double* pos;                 // [n,2] scratch: drift-corrected positions
};

__device__ __forceinline__ double track_dist(const double* pos, int a, int d) {
    const double dh = pos[2 * a] - pos[2 * d], dw = pos[2 * a + 1] - pos[2 * d + 1];
    return sqrt(__dadd_rn(__dmul_rn(dh, dh), __dmul_rn(dw, dw)));          // scipy.spatial.distance.euclidean
}

__global__ void __launch_bounds__(256)
track_greedy_kernel(const TrackArgs a) {
    __shared__ int s_changed;
    const int field = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const int F = a.n_frames, H = a.H, W = a.W, R = a.radius + 2;
    const int32_t* seg = a.seg_start + (size_t)field * F;
    const int first = seg[0], last = seg[F];
    int32_t* cache = a.grid_cache + (size_t)field * H * W;
    int32_t* grid = a.grid_frame + (size_t)field * H * W;
    const double radius = (double)a.radius;

    // ---- drift-corrected positions, drop-outs (flexlibrary.py:626-677), bins (:850-864)
    for (int f = 0; f < F; ++f) {
        const double ch = a.cum_off ? a.cum_off[((size_t)field * F + f) * 2] : 0.0;
        const double cw = a.cum_off ? a.cum_off[((size_t)field * F + f) * 2 + 1] : 0.0;
        for (int i = seg[f] + tid; i < seg[f + 1]; i += NT) {
            const double oh = a.spot_hw[2 * i] + ch, ow = a.spot_hw[2 * i + 1] + cw;
            bool drop = false;
            for (int g = 0; g < F; ++g) {
                const double gh = oh - (a.cum_off ? a.cum_off[((size_t)field * F + g) * 2] : 0.0);
                const double gw = ow - (a.cum_off ? a.cum_off[((size_t)field * F + g) * 2 + 1] : 0.0);
                if (!(a.spot_radius <= gh && gh < (double)H - 0.5 - a.spot_radius &&
                      a.spot_radius <= gw && gw < (double)W - 0.5 - a.spot_radius)) { drop = true; break; }
            }
            const int bh = (int)round(oh), bw = (int)round(ow);             // python-2 round(): half away from zero
            // (with cumulative offsets that start at (0, 0) a kept spot always rounds into the frame; the guard only
            //  protects the index grids against offsets that do not)
            if (!drop && (bh < 0 || bh >= H || bw < 0 || bw >= W)) drop = true;
            a.pos[2 * i] = oh; a.pos[2 * i + 1] = ow;
            a.discarded[i] = drop ? 1 : 0;
            a.bin_hw[2 * i] = drop ? -1 : bh;
            a.bin_hw[2 * i + 1] = drop ? -1 : bw;
            a.anc[i] = -1; a.desc[i] = -1;
        }
    }
    __syncthreads();
    // ---- frames in sequence
    for (int f = 0; f < F; ++f) {
        // the spots of frame f-1 leave the frame grid and join the ancestor cache (a newer spot replaces an older one)
        if (f > 0) {
            for (int i = seg[f - 1] + tid; i < seg[f] ; i += NT) {
                if (a.discarded[i]) continue;
                const int cell = a.bin_hw[2 * i] * W + a.bin_hw[2 * i + 1];
                grid[cell] = -1;
                cache[cell] = i;                                           // (an older, still unpaired spot here is dropped)
            }
        }
        __syncthreads();
        // the spots of frame f fill the frame grid; two in one pixel is the reference's assertion (:853-858)
        for (int i = seg[f] + tid; i < seg[f + 1]; i += NT) {
            if (a.discarded[i]) continue;
            const int cell = a.bin_hw[2 * i] * W + a.bin_hw[2 * i + 1];
            if (atomicCAS(&grid[cell], -1, i) != -1) atomicOr(&a.flags[field], 1);
        }
        __syncthreads();
        if (f == 0) continue;
        // ---- parallel replay of the greedy walk over the distance-sorted pairs (:941-984)
        for (;;) {
            if (tid == 0) s_changed = 0;
            __syncthreads();
            for (int i = first + tid; i < seg[f]; i += NT) {               // every spot of an earlier frame ...
                if (a.discarded[i] || a.desc[i] >= 0) continue;
                const int ah = a.bin_hw[2 * i], aw = a.bin_hw[2 * i + 1];
                if (cache[ah * W + aw] != i) continue;                     // ... that is still in the ancestor cache
                // its best free candidate: first by distance, then by the candidate's raster position
                int best = -1; double bd = 0.0;
                for (int dh = max(ah - R, 0); dh <= min(ah + R, H - 1); ++dh)
                    for (int dw = max(aw - R, 0); dw <= min(aw + R, W - 1); ++dw) {
                        const int d = grid[dh * W + dw];
                        if (d < 0 || a.anc[d] >= 0) continue;
                        const double dist = track_dist(a.pos, i, d);
                        if (!(dist < radius)) continue;
                        if (best < 0 || dist < bd) { best = d; bd = dist; }   // raster scan: ties keep the first
                    }
                if (best < 0) continue;
                // is this ancestor the candidate's own best: first by distance, then by the ancestor's raster position
                const int eh = a.bin_hw[2 * best], ew = a.bin_hw[2 * best + 1];
                int mine = -1; double md = 0.0;
                for (int qh = max(eh - R, 0); qh <= min(eh + R, H - 1); ++qh)
                    for (int qw = max(ew - R, 0); qw <= min(ew + R, W - 1); ++qw) {
                        const int q = cache[qh * W + qw];
                        if (q < 0 || a.desc[q] >= 0) continue;
                        const double dist = track_dist(a.pos, q, best);
                        if (!(dist < radius)) continue;
                        if (mine < 0 || dist < md) { mine = q; md = dist; }
                    }
                if (mine != i) continue;
                if (atomicCAS(&a.anc[best], -1, i) == -1) {                // link (:968-980)
                    a.desc[i] = best;
                    cache[ah * W + aw] = -1;
                    s_changed = 1;
                }
            }
            __syncthreads();
            const int again = s_changed;
            __syncthreads();
            if (!again) break;
        }
    }
    __syncthreads();
    // leave the scratch grids as they were found (-1 everywhere)
    for (int i = first + tid; i < last; i += NT) {
        if (a.discarded[i]) continue;
        const int cell = a.bin_hw[2 * i] * W + a.bin_hw[2 * i + 1];
        cache[cell] = -1; grid[cell] = -1;
    }
}

}  // namespace fsq

using namespace fsq;

extern "C" int64_t fsq_track_greedy_scratch_bytes(int n_fields, int H, int W, int64_t n) {
    if (n_fields <= 0 || H <= 0 || W <= 0) return 0;
    return (int64_t)2 * n_fields * H * W * 4 + (n > 0 ? n : 0) * 16 + 256;
}

extern "C" int fsq_track_greedy(const double* spot_hw, const int32_t* seg_start, const double* cum_offsets,
                                int n_fields, int n_frames, int H, int W, int64_t n, int candidate_radius,
                                double spot_radius, int32_t* anc, int32_t* desc, int32_t* bin_hw, uint8_t* discarded,
                                int32_t* flags, void* scratch, int64_t scratch_bytes, void* stream) {
    if (n_fields <= 0 || n_frames <= 0 || H <= 0 || W <= 0 || n < 0 || candidate_radius < 0 || n > 2147483647LL) {
        set_error("fsq_track_greedy: bad sizes");
        return FSQ_E_ARG;
    }
    if (!seg_start || !flags || !scratch) { set_error("fsq_track_greedy: NULL pointer argument"); return FSQ_E_ARG; }
    if (n > 0 && (!spot_hw || !anc || !desc || !bin_hw || !discarded)) { set_error("fsq_track_greedy: NULL pointer argument"); return FSQ_E_ARG; }
    if (scratch_bytes < fsq_track_greedy_scratch_bytes(n_fields, H, W, n)) { set_error("fsq_track_greedy: scratch too small"); return FSQ_E_CAPACITY; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t gbytes = (size_t)n_fields * H * W * 4;
    TrackArgs a;
    a.spot_hw = spot_hw; a.seg_start = seg_start; a.cum_off = cum_offsets;
    a.n_fields = n_fields; a.n_frames = n_frames; a.H = H; a.W = W; a.radius = candidate_radius; a.spot_radius = spot_radius;
    a.anc = anc; a.desc = desc; a.bin_hw = bin_hw; a.discarded = discarded; a.flags = flags;
    a.grid_cache = (int32_t*)scratch;
    a.grid_frame = (int32_t*)((char*)scratch + gbytes);
    a.pos = (double*)((char*)scratch + 2 * gbytes);
    FSQ_CUDA_CHECK(cudaMemsetAsync(scratch, 0xff, 2 * gbytes, st));
    FSQ_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(int32_t) * (size_t)n_fields, st));
    track_greedy_kernel<<<n_fields, 256, 0, st>>>(a);
    FSQ_LAUNCH_CHECK();
    return FSQ_OK;
}
