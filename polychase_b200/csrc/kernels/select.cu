// K6 + K7: deterministic ordering and greedy min-distance suppression of corner
// candidates.
//
// Replaces the std::sort + sequential grid-bucket loop of the reference detector
// (/root/reference/cpp/feature_detection/gftt.cc:7-12,98-164).  The reference walks the
// candidates strongest-first (value desc, address desc) and keeps one iff no already-kept
// candidate lies within min_distance (dx*dx + dy*dy < min_distance^2 on integer pixel
// coordinates; its 3x3 bucket search with cell = cvRound(min_distance) covers exactly that
// disc for integer coordinates).  Whether a candidate is kept therefore depends only on
// the decisions of *stronger* candidates inside the disc, which gives an exact parallel
// formulation: a candidate becomes REJECTED as soon as a stronger in-disc candidate is
// KEPT, and KEPT once every stronger in-disc candidate is decided and none is kept.
// A persistent cooperative kernel iterates that to its fixed point (grid-wide barrier per
// round, no host round trips); blocked candidates poll their few blockers inside a round, so on 4K
// frames one round decides everything.
//
// With max_corners > 0 the reference stops after max_corners kept corners (gftt.cc:160-162),
// i.e. it returns a prefix of the unlimited result.  Because decisions only depend on stronger
// candidates, the fixed point is run on value-closed prefixes of the candidate set (value
// thresholds from the 12-bit histogram the NMS kernel fills): the strongest ~1.5*max_corners, then
// down to ~4*max_corners, then everything, stopping once max_corners corners are kept.  Kept keys
// are counted into a 12-bit-bin histogram of their value as they are accepted; compact_top_kernel
// finds the bin that holds the max_corners-th strongest key and scatters the keys at or above it
// (max_corners plus part of one bin) grouped by bin; select_rank_emit_kernel ranks every key inside
// its bin group ((value, address) descending): rank r < max_corners is keypoint r.
// With max_corners == 0 every kept key is sorted (CUB radix sort).
#include <algorithm>
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace pc {

enum : uint8_t { ST_NONE = 0, ST_UNDECIDED = 1, ST_KEPT = 2, ST_REJECTED = 3 };
constexpr int kGreedySpins = 256;     // polls of a blocked candidate's blockers (~0.5 us each) before it waits for the next round

__device__ __forceinline__ unsigned long long key_at(const float* __restrict__ eig, int eig_pitch, int w, int x,
                                                     int y) {
    return ((unsigned long long)float_to_ordered_uint(eig[(size_t)y * eig_pitch + x]) << 32) | (unsigned)(y * w + x);
}

// Decision for one undecided candidate from the current state map: 0 = still blocked by an
// undecided stronger neighbour, ST_KEPT or ST_REJECTED.  R <= 4 takes the word-wide path: the
// 9 rows x 3 aligned words that cover the disc's bounding box are loaded up front
// (independent L2 loads), then only the non-zero state bytes are looked at.
constexpr int kMaxBlockers = 4;
__device__ __noinline__ int decide_scan(unsigned long long key, int x, int y, const float* __restrict__ eig,
                                      int eig_pitch, const uint8_t* state, int state_pitch, int w, int h, int R,
                                      double md2, int (&blockers)[kMaxBlockers], int& num_blockers) {
    bool blocked = false;
    num_blockers = 0;
    if (R <= 4) {
        const int xl = x - 4, wx0 = xl & ~3;                 // may be negative: masked below
        uint32_t wd[9][3];
#pragma unroll
        for (int r = 0; r < 9; r++) {
            const int ny = y - 4 + r;
            const bool row_ok = (unsigned)ny < (unsigned)h;
            const uint8_t* row = state + (size_t)(row_ok ? ny : y) * state_pitch;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int wx = wx0 + 4 * k;
                // state_pitch is a multiple of 128 >= w: an aligned word that starts inside the
                // pitch is readable; bytes at x >= w are skipped below
                wd[r][k] = (row_ok && wx >= 0 && wx < state_pitch) ? __ldcg(reinterpret_cast<const uint32_t*>(row + wx)) : 0u;
            }
        }
#pragma unroll
        for (int r = 0; r < 9; r++) {
            const int dy = r - 4;
            if (dy < -R || dy > R) continue;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                uint32_t v = wd[r][k];
                while (v) {
                    const int b = (__ffs(v) - 1) >> 3;
                    const uint32_t ns = (v >> (8 * b)) & 0xffu;
                    v &= ~(0xffu << (8 * b));
                    const int nx = wx0 + 4 * k + b, dx = nx - x;
                    if (dx < -R || dx > R || (dx == 0 && dy == 0) || nx < 0 || nx >= w) continue;
                    if ((double)(dx * dx + dy * dy) >= md2) continue;
                    if (ns != ST_UNDECIDED && ns != ST_KEPT) continue;
                    if (key_at(eig, eig_pitch, w, nx, y + dy) > key) {
                        if (ns == ST_KEPT) return ST_REJECTED;
                        blocked = true;
                        if (num_blockers < kMaxBlockers) blockers[num_blockers] = (y + dy) * state_pitch + nx;
                        num_blockers++;
                    }
                }
            }
        }
    } else {
        for (int dy = -R; dy <= R; dy++) {
            const int ny = y + dy;
            if (ny < 0 || ny >= h) continue;
            for (int dx = -R; dx <= R; dx++) {
                const int nx = x + dx;
                if (nx < 0 || nx >= w || (dx == 0 && dy == 0)) continue;
                if ((double)(dx * dx + dy * dy) >= md2) continue;
                const uint8_t ns = __ldcg(state + (size_t)ny * state_pitch + nx);
                if (ns != ST_UNDECIDED && ns != ST_KEPT) continue;
                if (key_at(eig, eig_pitch, w, nx, ny) > key) {
                    if (ns == ST_KEPT) return ST_REJECTED;
                    blocked = true;
                    num_blockers = kMaxBlockers + 1;         // generic path: no blocker list
                }
            }
        }
    }
    return blocked ? 0 : ST_KEPT;
}

// Front end of decide_scan for R <= 4.  A warp executes the union of its lanes' neighbour loops, so a
// load of the neighbour's eigenvalue inside those loops (one L2 round trip each, one after the other)
// made a warp's 32 candidates cost ~40 us.  Here the loops only collect the (few) decided-or-pending
// neighbours inside the disc; their eigenvalues are then fetched with independent loads and compared.
constexpr int kMaxNbr = 12;
__device__ __forceinline__ int decide(unsigned long long key, int x, int y, const float* __restrict__ eig,
                                      int eig_pitch, const uint8_t* state, int state_pitch, int w, int h, int R,
                                      double md2, int (&blockers)[kMaxBlockers], int& num_blockers) {
    if (R > 4) return decide_scan(key, x, y, eig, eig_pitch, state, state_pitch, w, h, R, md2, blockers, num_blockers);
    num_blockers = 0;
    const int xl = x - 4, wx0 = xl & ~3;                     // may be negative: masked below
    uint32_t wd[9][3];
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int ny = y - 4 + r;
        const bool row_ok = (unsigned)ny < (unsigned)h;
        const uint8_t* row = state + (size_t)(row_ok ? ny : y) * state_pitch;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int wx = wx0 + 4 * k;
            wd[r][k] = (row_ok && wx >= 0 && wx < state_pitch) ? __ldcg(reinterpret_cast<const uint32_t*>(row + wx)) : 0u;
        }
    }
    int nbr[kMaxNbr];                                        // (dy + 4) << 20 | nx << 4 | state
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int dy = r - 4;
        if (dy < -R || dy > R) continue;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint32_t v = wd[r][k];
            while (v) {
                const int b = (__ffs(v) - 1) >> 3;
                const uint32_t ns = (v >> (8 * b)) & 0xffu;
                v &= ~(0xffu << (8 * b));
                const int nx = wx0 + 4 * k + b, dx = nx - x;
                if (dx < -R || dx > R || (dx == 0 && dy == 0) || nx < 0 || nx >= w) continue;
                if ((double)(dx * dx + dy * dy) >= md2) continue;
                if (ns != ST_UNDECIDED && ns != ST_KEPT) continue;
                if (cnt < kMaxNbr) nbr[cnt] = (r << 20) | (nx << 4) | (int)ns;
                cnt++;
            }
        }
    }
    if (cnt > kMaxNbr)                                       // a crowd: the sequential scan handles any count
        return decide_scan(key, x, y, eig, eig_pitch, state, state_pitch, w, h, R, md2, blockers, num_blockers);
    float ev[kMaxNbr];
#pragma unroll
    for (int i = 0; i < kMaxNbr; i++) {
        ev[i] = 0.f;
        if (i < cnt) {
            const int ny = y - 4 + (nbr[i] >> 20), nx = (nbr[i] >> 4) & 0xffff;
            ev[i] = eig[(size_t)ny * eig_pitch + nx];
        }
    }
    bool blocked = false, rejected = false;
#pragma unroll
    for (int i = 0; i < kMaxNbr; i++) {
        if (i < cnt) {
            const int ny = y - 4 + (nbr[i] >> 20), nx = (nbr[i] >> 4) & 0xffff;
            const unsigned long long nkey = ((unsigned long long)float_to_ordered_uint(ev[i]) << 32) | (unsigned)(ny * w + nx);
            if (nkey > key) {
                if ((nbr[i] & 0xf) == ST_KEPT) rejected = true;
                blocked = true;
                if (num_blockers < kMaxBlockers) blockers[num_blockers] = ny * state_pitch + nx;
                num_blockers++;
            }
        }
    }
    if (rejected) return ST_REJECTED;
    return blocked ? 0 : ST_KEPT;
}

// Block-wide "largest bin b whose suffix count reaches `want`" over a histogram in global memory,
// PER_THREAD consecutive bins per thread (descending search; PER_THREAD <= 64).  Returns the bin
// (0 if the whole histogram holds fewer than `want`) and, through *count_out, the number of
// entries in bins >= that bin.
template <int THREADS, int PER_THREAD>
__device__ __forceinline__ int suffix_threshold_bin(const int* hist, int want, int* s_warp, int* s_res, int* count_out) {
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int4* hp = reinterpret_cast<const int4*>(hist) + t * (PER_THREAD / 4);
    int s = 0;
#pragma unroll
    for (int q = 0; q < PER_THREAD / 4; q++) {
        const int4 v = __ldcg(hp + q);
        s += v.x + v.y + v.z + v.w;
    }
    int v = s;                                              // suffix sum inside the warp (lanes >= lane)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += nb;
    }
    if (lane == 0) s_warp[wid] = v;
    if (t == 0) { s_res[0] = -1; s_res[1] = 0; }
    __syncthreads();
    int above = 0;
    for (int q = wid + 1; q < THREADS / 32; q++) above += s_warp[q];
    const int suf_incl = v + above, suf_excl = suf_incl - s;
    if (suf_excl < want && want <= suf_incl) { s_res[0] = t; s_res[1] = suf_excl; }   // boundary thread
    if (t == 0) s_res[2] = suf_incl;                        // whole histogram
    __syncthreads();
    const int tb = s_res[0];
    if (tb < 0) {                                           // fewer than `want` entries: take everything
        *count_out = s_res[2];
        return 0;
    }
    // the first warp resolves the boundary thread's bins: 2 per lane, suffix scan
    if (wid == 0) {
        const int excl = s_res[1];
        const int b0 = 2 * lane < PER_THREAD ? __ldcg(hist + tb * PER_THREAD + 2 * lane) : 0;
        const int b1 = 2 * lane + 1 < PER_THREAD ? __ldcg(hist + tb * PER_THREAD + 2 * lane + 1) : 0;
        int sv = b0 + b1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int nb = __shfl_down_sync(0xffffffffu, sv, o);
            if (lane + o < 32) sv += nb;
        }
        const int incl1 = excl + sv - b0;                   // entries in bins >= 2*lane+1 of thread tb
        const int incl0 = excl + sv;                        // entries in bins >= 2*lane
        const int excl1 = incl0 - b0 - b1;                  // entries in bins > 2*lane+1
        if (excl1 < want && want <= incl1) { s_res[3] = tb * PER_THREAD + 2 * lane + 1; s_res[4] = incl1; }
        else if (incl1 < want && want <= incl0) { s_res[3] = tb * PER_THREAD + 2 * lane; s_res[4] = incl0; }
    }
    __syncthreads();
    *count_out = s_res[4];
    return s_res[3];
}

// The barrier between the rounds of the fixed point: the whole (cooperative) grid, or -- with max_corners > 0, where a
// few thousand strong candidates decide everything -- one thread-block cluster.  A cooperative launch needs every one
// of its 2 x 148 blocks resident at the same time, so in the pipeline it waits for blocks of the concurrent LK batch
// to retire on ALL SMs while its early blocks spin at the barrier (measured: 0.21 ms per frame in the pipeline against
// 0.03 ms alone); a 16-CTA cluster needs 16 free slots in one GPC and leaves the rest of the chip to LK.
template <bool CLUSTER>
struct RoundBarrier {
    __device__ __forceinline__ void sync() {
        if (CLUSTER) cg::this_cluster().sync();
        else cg::this_grid().sync();
    }
};

// One fixed-point run over list[0..n): returns the number of candidates still undecided (0 =
// converged).  Every block of the grid must call it (barrier per round).
template <bool CLUSTER>
__device__ __forceinline__ int greedy_rounds(RoundBarrier<CLUSTER>& grid, int* block_undecided,
                                             const unsigned long long* __restrict__ list, int n,
                                             const float* __restrict__ eig, int eig_pitch, uint8_t* state,
                                             int state_pitch, int w, int h, int R, double md2,
                                             unsigned long long* __restrict__ accepted, int* accepted_count,
                                             int* kept_hist, int* round_counters) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const int gstride = gridDim.x * blockDim.x;
    int last = 0;
    for (int round = 0; round < kMaxGreedyRounds; round++) {
        if (threadIdx.x == 0) *block_undecided = 0;
        __syncthreads();
        int undecided = 0;
        // whole warps iterate together (the append below is warp-aggregated)
        for (int ib = gtid - lane; ib < n; ib += gstride) {
            const int i = ib + lane;
            unsigned long long key = 0ull;
            int d = -1;                                      // -1: nothing to decide for this lane
            if (i < n) {
                key = list[i];
                const int addr = (int)(key & 0xffffffffu);
                const int y = addr / w, x = addr - y * w;
                uint8_t* sp = state + (size_t)y * state_pitch + x;
                if (__ldcg(sp) == ST_UNDECIDED) {
                    int blockers[kMaxBlockers], nb;
                    d = decide(key, x, y, eig, eig_pitch, state, state_pitch, w, h, R, md2, blockers, nb);
                    // Publish a decision the moment it exists, BEFORE any code that waits: the lanes of a
                    // warp reconverge after the polling loop below, so a store placed behind it would be
                    // held back until the warp's slowest lane stops polling -- and that lane may be
                    // polling for exactly this store.
                    if (d == ST_REJECTED) *(volatile uint8_t*)sp = ST_REJECTED;
                    else if (d == ST_KEPT) *(volatile uint8_t*)sp = ST_KEPT;
                    if (d == 0 && nb <= kMaxBlockers) {
                        // The blockers are stronger candidates that other (co-resident) threads are deciding
                        // right now: watch just those few state bytes for a bounded time instead of paying
                        // a grid barrier + rescan per dependency level.  States only move UNDECIDED -> final,
                        // so "a blocker got KEPT" / "all blockers got REJECTED" are final answers too.
                        for (int spin = 0; spin < kGreedySpins && d == 0; spin++) {
                            __nanosleep(40);
                            bool pending = false;
                            for (int k = 0; k < nb; k++) {
                                const uint8_t ns = __ldcg(state + blockers[k]);
                                if (ns == ST_KEPT) d = ST_REJECTED;
                                else if (ns == ST_UNDECIDED) pending = true;
                            }
                            if (d == 0 && !pending) d = ST_KEPT;
                            if (d != 0) *(volatile uint8_t*)sp = (uint8_t)d;     // published inside the loop
                        }
                    }
                    if (d == 0) undecided++;
                }
            }
            // the list append can wait for the warp to reconverge: one atomic per warp, not per corner
            const unsigned kept = __ballot_sync(0xffffffffu, d == ST_KEPT);
            if (kept) {
                const int leader = __ffs(kept) - 1;
                int b = 0;
                if (lane == leader) b = atomicAdd(accepted_count, __popc(kept));
                b = __shfl_sync(0xffffffffu, b, leader);
                if (d == ST_KEPT) {
                    accepted[b + __popc(kept & ((1u << lane) - 1))] = key;
                    if (kept_hist) atomicAdd(&kept_hist[(unsigned)(key >> 52)], 1);
                }
            }
        }
        if (undecided) atomicAdd(block_undecided, undecided);
        __syncthreads();
        if (threadIdx.x == 0 && *block_undecided) atomicAdd(&round_counters[round], *block_undecided);
        grid.sync();
        last = *((volatile int*)&round_counters[round]);
        if (last == 0) break;
    }
    return last;
}

// The whole suppression stage in one cooperative launch.
//   max_corners == 0: one fixed-point run over every candidate.
//   max_corners  > 0: the reference returns the max_corners strongest kept corners, and a decision only
//   depends on stronger candidates, so the fixed point is run on value-closed prefixes of the candidate
//   set: first the candidates at or above the value bin at which the suffix count of the NMS histogram
//   reaches 1.5 x max_corners (on textured frames suppression removes few of the strongest corners:
//   8 466 candidates give 8 000 kept corners at 4K), then the band down to 4 x max_corners, then
//   everything, stopping as soon as max_corners corners are kept.  Every block recomputes the bins
//   (4096 ints); each stage appends its band of the candidate list to `strong`.
template <bool CLUSTER>
__global__ void __launch_bounds__(256) greedy_suppress_kernel(
    const unsigned long long* __restrict__ cand, const int* __restrict__ cand_count, int cand_cap,
    const float* __restrict__ eig, int eig_pitch, uint8_t* state, int state_pitch, int w, int h, int R,
    double md2, unsigned long long* __restrict__ accepted, int* accepted_count, int* kept_hist, int* round_counters,
    int* remaining, const int* value_hist, unsigned long long* __restrict__ strong, int* strong_count,
    int max_corners) {
    RoundBarrier<CLUSTER> grid;
    __shared__ int block_undecided;
    __shared__ int s_warp[8], s_res[8];
    const int n = min(*cand_count, cand_cap);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    int last = 0;
    if (max_corners > 0) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const long long wants[2] = {(long long)max_corners + max_corners / 2, 4ll * max_corners};
        unsigned above = 0xffffffffu;                        // bins >= `above` were handled by an earlier stage
        int list_begin = 0;
        bool enough = false;
        for (int stage = 0; stage < 2 && !enough && above > 0; stage++) {
            int total;
            const int want = (int)min(wants[stage], (long long)n);
            const unsigned thr = (unsigned)suffix_threshold_bin<256, 16>(value_hist, want, s_warp, s_res, &total);
            __syncthreads();                                 // s_warp / s_res are reused below
            // append this stage's band with one global atomic per block and chunk
            for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
                const int i = base + threadIdx.x;
                unsigned long long key = 0;
                bool keep = false;
                if (i < n) {
                    key = cand[i];
                    const unsigned bin = (unsigned)(key >> 52);
                    keep = bin >= thr && bin < above;
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) s_warp[wid] = __popc(m);
                __syncthreads();
                if (threadIdx.x == 0) {
                    int tot = 0;
#pragma unroll
                    for (int q = 0; q < 8; q++) { const int c = s_warp[q]; s_warp[q] = tot; tot += c; }
                    s_res[0] = tot ? atomicAdd(strong_count, tot) : 0;
                }
                __syncthreads();
                if (keep) strong[s_res[0] + s_warp[wid] + __popc(m & ((1u << lane) - 1))] = key;
                __syncthreads();
            }
            grid.sync();
            const int list_end = min(*((volatile int*)strong_count), cand_cap);
            last = greedy_rounds(grid, &block_undecided, strong + list_begin, list_end - list_begin, eig, eig_pitch, state,
                                 state_pitch, w, h, R, md2, accepted, accepted_count, kept_hist,
                                 round_counters + stage * kMaxGreedyRounds);
            // every block is past the barrier of the last round: the kept count is final and uniform
            enough = *((volatile int*)accepted_count) >= max_corners;
            list_begin = list_end;
            above = thr;
        }
        if (!enough && above > 0)                            // candidates below the last band remain
            last = greedy_rounds(grid, &block_undecided, cand, n, eig, eig_pitch, state, state_pitch, w, h, R, md2,
                                 accepted, accepted_count, kept_hist, round_counters + 2 * kMaxGreedyRounds);
    } else {
        last = greedy_rounds(grid, &block_undecided, cand, n, eig, eig_pitch, state, state_pitch, w, h, R, md2,
                             accepted, accepted_count, kept_hist, round_counters);
    }
    if (gtid == 0) *remaining = last;
}

__global__ void accept_all_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ cand_count,
                                  int cand_cap, unsigned long long* __restrict__ accepted, int* accepted_count,
                                  int* kept_hist, int* remaining) {
    const int n = min(*cand_count, cand_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long key = cand[i];
        accepted[i] = key;
        if (kept_hist) atomicAdd(&kept_hist[(unsigned)(key >> 52)], 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { *accepted_count = n; *remaining = 0; }
}

// ---- final selection: the max_corners strongest kept keys, in order, as keypoints -------------
// The 12-bit value histogram of the kept keys is a counting sort's first half: the rank of a key is
// (number of keys in higher bins) + (its rank inside its own bin).
// compact_top_kernel: every CTA builds the exclusive suffix sums of the histogram in shared memory,
// finds the bin `thr` that holds the k-th strongest key (k = min(max_corners, kept)) and scatters its
// share of the keys at or above `thr` into `top`, grouped by bin (group start = suffix sum, slot inside
// the group from a per-bin cursor).  m = keys at or above `thr` = k plus part of one bin.
// select_rank_emit_kernel: one warp per key of top[0..m) counts the larger keys of its own group
// (a few hundred entries): rank < k -> keypoint slot `rank`.
constexpr int TOP_BINS = 4096;

// sel[2] = threshold bin, sel[3] = m.  bin_cursor[TOP_BINS] is zero on entry; bin_start is written by
// CTA 0 (suffix sums of the bins >= thr).
__global__ void __launch_bounds__(256) compact_top_kernel(const unsigned long long* __restrict__ accepted,
                                                          const int* __restrict__ accepted_count,
                                                          const int* __restrict__ kept_hist, int max_corners,
                                                          int kps_cap, unsigned long long* __restrict__ top,
                                                          int* __restrict__ sel, int* __restrict__ kps_count,
                                                          int* __restrict__ bin_cursor, int* __restrict__ bin_start) {
    __shared__ int s_suf[TOP_BINS];                          // keys in bins > b
    __shared__ int s_warp[8];
    __shared__ int s_thr, s_m;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int n = *accepted_count;
    const int k = min(min(max_corners, n), kps_cap);
    if (blockIdx.x == 0 && t == 0) *kps_count = k;
    if (k == 0) {
        if (blockIdx.x == 0 && t == 0) { sel[2] = 0; sel[3] = 0; }
        return;
    }
    // thread t owns bins [16 t, 16 t + 16)
    int hv[16];
    {
        const int4* hp = reinterpret_cast<const int4*>(kept_hist) + t * 4;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int4 v = __ldcg(hp + q);
            hv[4 * q] = v.x; hv[4 * q + 1] = v.y; hv[4 * q + 2] = v.z; hv[4 * q + 3] = v.w;
        }
    }
    int own = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) own += hv[q];
    int v = own;                                             // suffix sum inside the warp (lanes >= lane)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += nb;
    }
    if (lane == 0) s_warp[wid] = v;
    if (t == 0) { s_thr = 0; s_m = 0; }
    __syncthreads();
    int above = 0;
    for (int q = wid + 1; q < 8; q++) above += s_warp[q];
    int run = v + above - own;                               // keys in bins above this thread's bins
    if (t == 0) s_m = v + above;                             // whole histogram (taken when it holds < k: cannot happen, k <= n)
    __syncthreads();
#pragma unroll
    for (int q = 15; q >= 0; q--) {
        s_suf[16 * t + q] = run;
        if (run < k && k <= run + hv[q]) { s_thr = 16 * t + q; s_m = run + hv[q]; }
        run += hv[q];
    }
    __syncthreads();
    const unsigned thr = (unsigned)s_thr;
    if (blockIdx.x == 0) {
        if (t == 0) { sel[2] = s_thr; sel[3] = s_m; }
#pragma unroll
        for (int q = 0; q < 16; q++) bin_start[16 * t + q] = s_suf[16 * t + q];
    }
    for (int i = blockIdx.x * blockDim.x + t; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long key = accepted[i];
        const unsigned bin = (unsigned)(key >> 52);
        if (bin >= thr) top[s_suf[bin] + atomicAdd(&bin_cursor[bin], 1)] = key;
    }
}

// Ranks the keys of the short list inside their value bin and emits the first k as keypoints.  top[] holds the bins'
// groups in descending bin order, unordered inside a group; a benchmark frame's 8 000 strongest corners fall into a
// handful of 12-bit bins of ~2 000 keys each, so counting "greater keys of my bin" key by key (one warp per key: the
// first version, 6.8 M warp instructions and the whole chip for 12 us) is quadratic where it hurts.  Here a block
// takes a bin, splits it into 1024 sub-bins by the next ten bits of the value (shared-memory histogram, suffix sums,
// one scatter into `scratch` -- the accepted[] buffer, dead by now), and a key then only compares itself with the
// couple of keys of its own sub-bin: linear in the bin size.
constexpr int RANK_SUB = 1024;
__global__ void __launch_bounds__(256) select_rank_emit_kernel(
    const unsigned long long* __restrict__ top, unsigned long long* __restrict__ scratch, const int* __restrict__ sel,
    const int* __restrict__ kps_count, const int* __restrict__ kept_hist, const int* __restrict__ bin_start, int w,
    float* __restrict__ kps) {
    __shared__ int s_cnt[RANK_SUB];                          // keys per sub-bin
    __shared__ int s_cur[RANK_SUB];                          // fill cursors of the scatter
    __shared__ int s_off[RANK_SUB];                          // keys of the bin in greater sub-bins
    __shared__ int s_warp[8];
    __shared__ int s_g0[RANK_SUB / 16], s_n[RANK_SUB / 16];   // this block's bins: start and size (one round trip, not one per bin)
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int k = *kps_count;
    const int thr = sel[2];
    if (k == 0) return;
    const int my_bins = (TOP_BINS - 1 - (int)blockIdx.x - thr) / (int)gridDim.x + 1;     // bins of this block at or above thr
    for (int i = t; i < RANK_SUB / 16; i += 256) {
        const int bin = TOP_BINS - 1 - (int)blockIdx.x - i * (int)gridDim.x;
        const bool in = i < my_bins && bin >= thr && bin >= 0;
        s_g0[i] = in ? __ldg(bin_start + bin) : 0;
        s_n[i] = in ? __ldg(kept_hist + bin) : 0;
    }
    __syncthreads();
    for (int bi = 0; bi < min(my_bins, RANK_SUB / 16); bi++) {
        const int g0 = s_g0[bi], n = s_n[bi];
        if (n == 0 || g0 >= k) continue;                     // empty, or every key of it ranks behind the k-th (block-uniform)
        for (int i = t; i < RANK_SUB; i += 256) { s_cnt[i] = 0; s_cur[i] = 0; }
        __syncthreads();
        for (int i = t; i < n; i += 256) atomicAdd(&s_cnt[(int)(top[g0 + i] >> 42) & (RANK_SUB - 1)], 1);
        __syncthreads();
        // suffix sums: thread t owns sub-bins [4 t, 4 t + 4)
        int c4[4], own = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) { c4[q] = s_cnt[4 * t + q]; own += c4[q]; }
        int v = own;                                         // suffix sum inside the warp (lanes >= lane)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int nb = __shfl_down_sync(0xffffffffu, v, o);
            if (lane + o < 32) v += nb;
        }
        if (lane == 0) s_warp[wid] = v;
        __syncthreads();
        int above = 0;
        for (int q = wid + 1; q < 8; q++) above += s_warp[q];
        int run = v + above - own;                           // keys in sub-bins above this thread's
#pragma unroll
        for (int q = 3; q >= 0; q--) { s_off[4 * t + q] = run; run += c4[q]; }
        __syncthreads();
        for (int i = t; i < n; i += 256) {
            const unsigned long long key = top[g0 + i];
            const int sb = (int)(key >> 42) & (RANK_SUB - 1);
            scratch[g0 + s_off[sb] + atomicAdd(&s_cur[sb], 1)] = key;
        }
        __syncthreads();                                     // (block-scope visibility of the scatter)
        for (int i = t; i < n; i += 256) {
            const unsigned long long key = scratch[g0 + i];
            const int sb = (int)(key >> 42) & (RANK_SUB - 1);
            const int r0 = s_off[sb], c = s_cnt[sb];
            int cnt = 0;
            for (int j = 0; j < c; j++) cnt += scratch[g0 + r0 + j] > key ? 1 : 0;
            const int rank = g0 + r0 + cnt;
            if (rank < k) {
                const int addr = (int)(key & 0xffffffffu);
                const int y = addr / w;
                kps[2 * rank] = (float)(addr - y * w);
                kps[2 * rank + 1] = (float)y;
            }
        }
        __syncthreads();                                     // s_* are reused by the next bin
    }
}

__global__ void keys_to_keypoints_kernel(const unsigned long long* __restrict__ sorted,
                                         const int* __restrict__ accepted_count, int w, int max_corners,
                                         float* __restrict__ kps, int kps_cap, int* __restrict__ kps_count) {
    int n = *accepted_count;
    if (max_corners > 0) n = min(n, max_corners);
    n = min(n, kps_cap);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *kps_count = n;
    if (i < n) {
        const int addr = (int)(sorted[i] & 0xffffffffu);
        const int y = addr / w;
        kps[2 * i] = (float)(addr - y * w);
        kps[2 * i + 1] = (float)y;
    }
}

size_t select_cub_temp_bytes(int cap) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeysDescending(nullptr, bytes, (const unsigned long long*)nullptr,
                                             (unsigned long long*)nullptr, cap);
    return bytes;
}

static void launch_greedy(const unsigned long long* cand, const int* cand_count, int cand_cap, const float* eig,
                          int eig_pitch, uint8_t* state, int state_pitch, int w, int h, double min_distance,
                          const SelectWorkspace& ws, int* kept_hist, int max_corners, int sm_count, cudaStream_t s) {
    int R = (int)ceil(min_distance) - 1;
    double md2 = min_distance * min_distance;
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, greedy_suppress_kernel<false>, 256, 0);
        blocks_per_sm = std::min(std::max(blocks_per_sm, 1), 2);
    }
    int nblocks = sm_count * blocks_per_sm;
    const int* value_hist = ws.hist;
    unsigned long long* strong = ws.strong;
    int* strong_count = ws.sel + 1;
    int* round_counters = ws.round_counters;
    // max_corners > 0: one 16-CTA cluster (8 where 16 cannot be placed); PC_GREEDY_GRID=1 forces the cooperative grid
    static int cluster = -1;
    if (cluster < 0) {
        cluster = 0;
        if (!getenv("PC_GREEDY_GRID")) {
            int want = 16;
            if (cudaFuncSetAttribute(greedy_suppress_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
                cudaGetLastError();
                want = 8;
            }
            for (; want >= 8 && !cluster; want -= 8) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(want);
                cfg.blockDim = dim3(256);
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = want;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                int n = 0;
                if (cudaOccupancyMaxActiveClusters(&n, greedy_suppress_kernel<true>, &cfg) == cudaSuccess && n >= 1) cluster = want;
                else cudaGetLastError();
            }
        }
    }
    if (max_corners > 0 && cluster > 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cluster);
        cfg.blockDim = dim3(256);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, greedy_suppress_kernel<true>, cand, cand_count, cand_cap, eig, eig_pitch, state, state_pitch, w,
                           h, R, md2, ws.accepted, ws.accepted_count, kept_hist, round_counters, ws.remaining, value_hist,
                           strong, strong_count, max_corners);
        return;
    }
    void* args[] = {(void*)&cand, (void*)&cand_count, (void*)&cand_cap, (void*)&eig, (void*)&eig_pitch, (void*)&state,
                    (void*)&state_pitch, (void*)&w, (void*)&h, (void*)&R, (void*)&md2, (void*)&ws.accepted,
                    (void*)&ws.accepted_count, (void*)&kept_hist, (void*)&round_counters, (void*)&ws.remaining,
                    (void*)&value_hist, (void*)&strong, (void*)&strong_count, (void*)&max_corners};
    cudaLaunchCooperativeKernel((void*)greedy_suppress_kernel<false>, dim3(nblocks), dim3(256), args, 0, s);
}

void launch_select(const unsigned long long* cand, const int* cand_count, int cand_cap, const float* eig,
                   int eig_pitch, uint8_t* state, int state_pitch, int w, int h, double min_distance,
                   int max_corners, SelectWorkspace ws, float* kps_out, int kps_cap, int* kps_count, int sm_count,
                   cudaStream_t s) {
    // accepted_count, the histograms, round counters, cursors and sel are zero on entry (launch_min_eig
    // clears the detector's counter block and the frame's counters)
    const bool limited = max_corners > 0;
    int* kept_hist = limited ? ws.kept_hist : nullptr;
    // the unlimited path sorts the whole accepted[] buffer: unused slots must be zero (they sort last)
    if (!limited) cudaMemsetAsync(ws.accepted, 0, sizeof(unsigned long long) * (size_t)ws.cap, s);
    if (min_distance >= 1.0) {
        launch_greedy(cand, cand_count, cand_cap, eig, eig_pitch, state, state_pitch, w, h, min_distance, ws, kept_hist,
                      limited ? max_corners : 0, sm_count, s);
    } else {
        accept_all_kernel<<<sm_count * 2, 256, 0, s>>>(cand, cand_count, cand_cap, ws.accepted, ws.accepted_count,
                                                       kept_hist, ws.remaining);
    }
    // keys: [63:32] ordered value, [31:0] address (< w*h); zero keys sort last
    if (limited) {
        // the short list lives in the (otherwise unused on this path) sort output buffer
        compact_top_kernel<<<sm_count, 256, 0, s>>>(ws.accepted, ws.accepted_count, ws.kept_hist, max_corners, kps_cap,
                                                    ws.sorted, ws.sel, kps_count, ws.bin_cursor, ws.bin_start);
        select_rank_emit_kernel<<<64, 256, 0, s>>>(ws.sorted, ws.accepted, ws.sel, kps_count, ws.kept_hist, ws.bin_start, w,
                                                   kps_out);
    } else {
        size_t temp = ws.cub_temp_bytes;
        cub::DeviceRadixSort::SortKeysDescending(ws.cub_temp, temp, ws.accepted, ws.sorted, ws.cap, 0, 64, s);
        keys_to_keypoints_kernel<<<(kps_cap + 255) / 256, 256, 0, s>>>(ws.sorted, ws.accepted_count, w, max_corners,
                                                                       kps_out, kps_cap, kps_count);
    }
}

}  // namespace pc
