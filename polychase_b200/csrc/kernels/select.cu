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
// round, no host round trips); the kept set is then ordered by a 64-bit radix sort on
// (value, address) descending and cut at max_corners -- the reference's early exit
// (gftt.cc:160-162) is a prefix of the unlimited result in that order.
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace pc {

enum : uint8_t { ST_NONE = 0, ST_UNDECIDED = 1, ST_KEPT = 2, ST_REJECTED = 3 };

__global__ void __launch_bounds__(256) greedy_suppress_kernel(
    const unsigned long long* __restrict__ cand, const int* __restrict__ cand_count, int cand_cap,
    const float* __restrict__ eig, int eig_pitch, uint8_t* state, int state_pitch, int w, int h, int R,
    double md2, unsigned long long* __restrict__ accepted, int* accepted_count, int* round_counters,
    int* remaining) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int block_undecided;
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

void launch_select(const unsigned long long* cand, const int* cand_count, int cand_cap, const float* eig,
                   int eig_pitch, uint8_t* state, int state_pitch, int w, int h, double min_distance,
                   int max_corners, SelectWorkspace ws, float* kps_out, int kps_cap, int* kps_count, int sm_count,
                   cudaStream_t s) {
    cudaMemsetAsync(ws.accepted, 0, sizeof(unsigned long long) * (size_t)ws.cap, s);
    cudaMemsetAsync(ws.accepted_count, 0, sizeof(int), s);
    if (min_distance >= 1.0) {
        cudaMemsetAsync(ws.round_counters, 0, sizeof(int) * kMaxGreedyRounds, s);
        int R = (int)ceil(min_distance) - 1;
        double md2 = min_distance * min_distance;
        int blocks_per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, greedy_suppress_kernel, 256, 0);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        if (blocks_per_sm > 4) blocks_per_sm = 4;
        int nblocks = sm_count * blocks_per_sm;
        void* args[] = {(void*)&cand, (void*)&cand_count, (void*)&cand_cap, (void*)&eig, (void*)&eig_pitch,
                        (void*)&state, (void*)&state_pitch, (void*)&w, (void*)&h, (void*)&R, (void*)&md2,
                        (void*)&ws.accepted, (void*)&ws.accepted_count, (void*)&ws.round_counters,
                        (void*)&ws.remaining};
        cudaLaunchCooperativeKernel((void*)greedy_suppress_kernel, dim3(nblocks), dim3(256), args, 0, s);
    } else {
        accept_all_kernel<<<sm_count * 2, 256, 0, s>>>(cand, cand_count, cand_cap, ws.accepted, ws.accepted_count,
                                                       ws.remaining);
    }
    // keys: [63:32] ordered value, [31:0] address (< w*h).  Unused slots are zero and sort last.
    int addr_bits = 1;
    while ((1ll << addr_bits) < (long long)w * h) addr_bits++;
    size_t temp = ws.cub_temp_bytes;
    cub::DeviceRadixSort::SortKeysDescending(ws.cub_temp, temp, ws.accepted, ws.sorted, ws.cap, 0, 64, s);
    (void)addr_bits;
    const int nthreads = max_corners > 0 ? (max_corners < kps_cap ? max_corners : kps_cap) : kps_cap;
    keys_to_keypoints_kernel<<<(nthreads + 255) / 256, 256, 0, s>>>(ws.sorted, ws.accepted_count, w, max_corners,
                                                                    kps_out, kps_cap, kps_count);
}

}  // namespace pc
