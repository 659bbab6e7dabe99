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
// round, no host round trips).
//
// With max_corners > 0 the reference stops after max_corners kept corners (gftt.cc:160-162),
// i.e. it returns a prefix of the unlimited result.  Because decisions only depend on stronger
// candidates, the fixed point is first run on the strongest ~4*max_corners candidates (a value
// threshold from a 12-bit histogram); only if that yields fewer than max_corners kept corners
// does a second pass process everything.  The max_corners strongest kept keys are then
// extracted exactly with a 4-pass radix select and sorted (64-bit radix sort on
// (value, address) descending).  With max_corners == 0 every kept key is sorted.
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace pc {

enum : uint8_t { ST_NONE = 0, ST_UNDECIDED = 1, ST_KEPT = 2, ST_REJECTED = 3 };

// list/count: candidates to decide in this launch.  If enough_at > 0 and that many corners are
// already kept, the launch is a no-op (second, full pass of the max_corners path).
__global__ void __launch_bounds__(256) greedy_suppress_kernel(
    const unsigned long long* __restrict__ cand, const int* __restrict__ cand_count, int cand_cap,
    const float* __restrict__ eig, int eig_pitch, uint8_t* state, int state_pitch, int w, int h, int R,
    double md2, unsigned long long* __restrict__ accepted, int* accepted_count, int* round_counters,
    int* remaining, int enough_at) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int block_undecided;
    if (enough_at > 0) {        // every block must take the same decision: read, barrier, then decide
        const int kept = *((volatile int*)accepted_count);
        grid.sync();
        if (kept >= enough_at) return;
    }
    const int n = min(*cand_count, cand_cap);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gstride = gridDim.x * blockDim.x;
    int last = 0;
    for (int round = 0; round < kMaxGreedyRounds; round++) {
        if (threadIdx.x == 0) block_undecided = 0;
        __syncthreads();
        int undecided = 0;
        for (int i = gtid; i < n; i += gstride) {
            const unsigned long long key = cand[i];
            const int addr = (int)(key & 0xffffffffu);
            const int y = addr / w, x = addr - y * w;
            uint8_t* sp = state + (size_t)y * state_pitch + x;
            if (__ldcg(sp) != ST_UNDECIDED) continue;
            bool blocked = false, rejected = false;
            for (int dy = -R; dy <= R && !rejected; dy++) {
                const int ny = y + dy;
                if (ny < 0 || ny >= h) continue;
                for (int dx = -R; dx <= R; dx++) {
                    const int nx = x + dx;
                    if (nx < 0 || nx >= w || (dx == 0 && dy == 0)) continue;
                    if ((double)(dx * dx + dy * dy) >= md2) continue;
                    const uint8_t ns = __ldcg(state + (size_t)ny * state_pitch + nx);
                    if (ns != ST_UNDECIDED && ns != ST_KEPT) continue;
                    const unsigned long long nkey =
                        ((unsigned long long)float_to_ordered_uint(eig[(size_t)ny * eig_pitch + nx]) << 32) |
                        (unsigned)(ny * w + nx);
                    if (nkey > key) {
                        if (ns == ST_KEPT) { rejected = true; break; }
                        blocked = true;
                    }
                }
            }
            if (rejected) {
                *sp = ST_REJECTED;
            } else if (!blocked) {
                *sp = ST_KEPT;
                accepted[atomicAdd(accepted_count, 1)] = key;
            } else {
                undecided++;
            }
        }
        if (undecided) atomicAdd(&block_undecided, undecided);
        __syncthreads();
        if (threadIdx.x == 0 && block_undecided) atomicAdd(&round_counters[round], block_undecided);
        grid.sync();
        last = *((volatile int*)&round_counters[round]);
        if (last == 0) break;
    }
    if (gtid == 0) *remaining = last;
}

__global__ void accept_all_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ cand_count,
                                  int cand_cap, unsigned long long* __restrict__ accepted, int* accepted_count,
                                  int* remaining) {
    const int n = min(*cand_count, cand_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) accepted[i] = cand[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) { *accepted_count = n; *remaining = 0; }
}

// sel[0] <- the largest 12-bit value bin b (ordered value >> 20) such that at least `want`
// candidates have bin >= b (0 if there are fewer candidates than that): the strong threshold.
__global__ void __launch_bounds__(1024) strong_threshold_kernel(const int* __restrict__ hist, int want, int* sel) {
    __shared__ int wsum[32];
    const int t = threadIdx.x;
    const int4 hv = __ldcg(reinterpret_cast<const int4*>(hist) + t);      // bins 4t .. 4t+3
    const int s = hv.x + hv.y + hv.z + hv.w;
    int v = s;
    const int lane = t & 31, wid = t >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += n;
    }
    if (lane == 0) wsum[wid] = v;
    if (t == 0) { sel[0] = 0; sel[1] = 0; }
    __syncthreads();
    int above = 0;                                  // candidates in warps above this one
    for (int k = wid + 1; k < 32; k++) above += wsum[k];
    const int suf_incl = v + above, suf_excl = suf_incl - s;
    if (suf_excl < want && want <= suf_incl) {      // the boundary bin is in this thread's range
        const int h4[4] = {hv.x, hv.y, hv.z, hv.w};
        int acc = suf_excl;
        for (int k = 3; k >= 0; k--) {
            acc += h4[k];
            if (acc >= want) { sel[0] = t * 4 + k; break; }
        }
    }
}

__global__ void __launch_bounds__(256) compact_strong_kernel(const unsigned long long* __restrict__ cand,
                                                             const int* __restrict__ cand_count, int cand_cap,
                                                             const int* __restrict__ sel,
                                                             unsigned long long* __restrict__ strong,
                                                             int* __restrict__ strong_count) {
    const int n = min(*cand_count, cand_cap);
    const unsigned thr = (unsigned)sel[0];
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        unsigned long long key = 0;
        bool keep = false;
        if (i < n) {
            key = cand[i];
            keep = (unsigned)(key >> 52) >= thr;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
            int b = 0;
            if (lane == leader) b = atomicAdd(strong_count, __popc(m));
            b = __shfl_sync(0xffffffffu, b, leader);
            if (keep) strong[b + __popc(m & ((1u << lane) - 1))] = key;
        }
    }
}

// Exact selection of the k = min(max_corners, n) largest keys (one CTA): 8 radix passes of 8
// bits with shared-memory histograms find the k-th largest key, then every key >= it is copied
// to `out` (zero padded to out_cap so that a fixed-size sort can follow).
__global__ void __launch_bounds__(1024) topk_select_kernel(const unsigned long long* __restrict__ keys,
                                                           const int* __restrict__ n_ptr, int max_corners,
                                                           unsigned long long* __restrict__ out, int out_cap) {
    __shared__ int h[256];
    __shared__ int wsum[8];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_krem, s_out;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int n = *n_ptr;
    const int k = min(max_corners, n);
    for (int i = t; i < out_cap; i += 1024) out[i] = 0ull;
    if (t == 0) { s_prefix = 0ull; s_krem = k; s_out = 0; }
    __syncthreads();
    unsigned long long kth = 0ull;
    if (n > k) {
        for (int d = 7; d >= 0; d--) {
            if (t < 256) h[t] = 0;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            const int shift = 8 * d;
            for (int i = t; i < n; i += 1024) {
                const unsigned long long key = keys[i];
                const bool match = d == 7 ? true : ((key >> (shift + 8)) == (prefix >> (shift + 8)));
                if (match) atomicAdd(&h[(unsigned)(key >> shift) & 0xffu], 1);
            }
            __syncthreads();
            const int krem = s_krem;
            int s = 0, v = 0;
            if (t < 256) {
                s = h[t];
                v = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int nb = __shfl_down_sync(0xffffffffu, v, o);
                    if (lane + o < 32) v += nb;
                }
                if (lane == 0) wsum[wid] = v;
            }
            __syncthreads();
            if (t < 256) {
                int above = 0;
                for (int q = wid + 1; q < 8; q++) above += wsum[q];
                const int suf_incl = v + above, suf_excl = suf_incl - s;    // bins >= t / bins > t
                if (suf_excl < krem && krem <= suf_incl) {
                    s_prefix = prefix | ((unsigned long long)t << shift);
                    s_krem = krem - suf_excl;                              // rank inside the chosen bin
                }
            }
            __syncthreads();
        }
        kth = s_prefix;
    }
    for (int base = 0; base < n; base += 1024) {
        const int i = base + t;
        unsigned long long key = 0ull;
        bool keep = false;
        if (i < n) {
            key = keys[i];
            keep = key >= kth;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        int b = 0;
        if (m) {
            const int leader = __ffs(m) - 1;
            if (lane == leader) b = atomicAdd(&s_out, __popc(m));
            b = __shfl_sync(0xffffffffu, b, leader);
            if (keep) {
                const int slot = b + __popc(m & ((1u << lane) - 1));
                if (slot < out_cap) out[slot] = key;
            }
        }
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

static void launch_greedy(const unsigned long long* list, const int* count, int cap, const float* eig, int eig_pitch,
                          uint8_t* state, int state_pitch, int w, int h, double min_distance,
                          const SelectWorkspace& ws, int* round_counters, int enough_at, int sm_count,
                          cudaStream_t s) {
    int R = (int)ceil(min_distance) - 1;
    double md2 = min_distance * min_distance;
    int blocks_per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, greedy_suppress_kernel, 256, 0);
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (blocks_per_sm > 2) blocks_per_sm = 2;
    int nblocks = sm_count * blocks_per_sm;
    void* args[] = {(void*)&list, (void*)&count, (void*)&cap, (void*)&eig, (void*)&eig_pitch, (void*)&state,
                    (void*)&state_pitch, (void*)&w, (void*)&h, (void*)&R, (void*)&md2, (void*)&ws.accepted,
                    (void*)&ws.accepted_count, (void*)&round_counters, (void*)&ws.remaining, (void*)&enough_at};
    cudaLaunchCooperativeKernel((void*)greedy_suppress_kernel, dim3(nblocks), dim3(256), args, 0, s);
}

void launch_select(const unsigned long long* cand, const int* cand_count, int cand_cap, const float* eig,
                   int eig_pitch, uint8_t* state, int state_pitch, int w, int h, double min_distance,
                   int max_corners, SelectWorkspace ws, float* kps_out, int kps_cap, int* kps_count, int sm_count,
                   cudaStream_t s) {
    cudaMemsetAsync(ws.accepted_count, 0, sizeof(int), s);
    const bool limited = max_corners > 0 && max_corners <= ws.topk_cap;
    // the unlimited path sorts the whole accepted[] buffer: unused slots must be zero (they sort last)
    if (!limited) cudaMemsetAsync(ws.accepted, 0, sizeof(unsigned long long) * (size_t)ws.cap, s);
    if (min_distance >= 1.0) {
        cudaMemsetAsync(ws.round_counters, 0, sizeof(int) * 2 * kMaxGreedyRounds, s);
        if (limited) {
            // pass 1: the strongest ~4*max_corners candidates
            strong_threshold_kernel<<<1, 1024, 0, s>>>(ws.hist, 4 * max_corners, ws.sel);
            compact_strong_kernel<<<sm_count, 256, 0, s>>>(cand, cand_count, cand_cap, ws.sel, ws.strong, ws.sel + 1);
            launch_greedy(ws.strong, ws.sel + 1, cand_cap, eig, eig_pitch, state, state_pitch, w, h, min_distance, ws,
                          ws.round_counters, 0, sm_count, s);
            // pass 2 (no-op when pass 1 already kept max_corners corners): everything else
            launch_greedy(cand, cand_count, cand_cap, eig, eig_pitch, state, state_pitch, w, h, min_distance, ws,
                          ws.round_counters + kMaxGreedyRounds, max_corners, sm_count, s);
        } else {
            launch_greedy(cand, cand_count, cand_cap, eig, eig_pitch, state, state_pitch, w, h, min_distance, ws,
                          ws.round_counters, 0, sm_count, s);
        }
    } else {
        accept_all_kernel<<<sm_count * 2, 256, 0, s>>>(cand, cand_count, cand_cap, ws.accepted, ws.accepted_count,
                                                       ws.remaining);
    }
    // keys: [63:32] ordered value, [31:0] address (< w*h); zero keys sort last
    size_t temp = ws.cub_temp_bytes;
    const unsigned long long* sorted = ws.sorted;
    if (limited) {
        topk_select_kernel<<<1, 1024, 0, s>>>(ws.accepted, ws.accepted_count, max_corners, ws.topk, max_corners);
        cub::DeviceRadixSort::SortKeysDescending(ws.cub_temp, temp, ws.topk, ws.sorted, max_corners, 0, 64, s);
    } else {
        cub::DeviceRadixSort::SortKeysDescending(ws.cub_temp, temp, ws.accepted, ws.sorted, ws.cap, 0, 64, s);
    }
    const int nthreads = max_corners > 0 ? (max_corners < kps_cap ? max_corners : kps_cap) : kps_cap;
    keys_to_keypoints_kernel<<<(nthreads + 255) / 256, 256, 0, s>>>(sorted, ws.accepted_count, w, max_corners, kps_out,
                                                                    kps_cap, kps_count);
}

}  // namespace pc
