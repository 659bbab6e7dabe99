// K4 + K5: Shi-Tomasi min-eigenvalue map, per-grid-cell maximum, per-cell relative
// threshold, 3x3 non-maximum suppression and candidate compaction.
//
// Replaces cv::cornerMinEigenVal(block 3, ksize 3) + the per-cell minMaxLoc / threshold /
// dilate / equality scan of the reference detector
// (/root/reference/cpp/feature_detection/gftt.cc:31-86).  The float32 operation order is
// that of OpenCV's AVX2-dispatched Sobel (what an x86-64 build of the reference runs;
// SURVEY.md Appendix A.4, pinned against cv2 4.13.0 by oracle/restate.c):
//   rx = p[x+1] - p[x-1]                      dx = fma(rx[y-1] + rx[y+1], s, rx[y] * 2s)
//   rs = fma(p[x+1], s, fma(p[x], 2s, p[x-1]*s))   (plain mul/add in the w%32 tail columns)
//   dy = rs[y+1] - rs[y-1]
//   cov = (dx*dx, dx*dy, dy*dy), 3x3 box sums formed in double and rounded once,
//   eig = (a + c) - sqrt((a - c)*(a - c) + b*b),  a = Sxx/2, b = Sxy, c = Syy/2, no FMA.
// All borders are BORDER_REFLECT_101 (of the gray image for Sobel, of cov for the box).
//
// Known deviation (documented in DESIGN.md): OpenCV forms the vertical box sum with a
// running double accumulator carried down each column from row 0; its rounding history
// changes ~2 pixels per million by a few ulp.  This kernel forms each 3x3 sum
// independently (the correctly rounded value).
#include "common.cuh"
#include "kernels.h"

namespace pc {

constexpr int ME_TW = 64, ME_TH = 16;
constexpr int ME_GW = ME_TW + 4, ME_GH = ME_TH + 4;      // gray region (halo 2)
constexpr int ME_DW = ME_TW + 2, ME_DH = ME_TH + 2;      // cov region (halo 1)

__device__ __forceinline__ int cell_of(int x, int y, const DetectGrid& g) {
    return (y / g.block_h) * g.grid_cols + (x / g.block_w);
}

__global__ void __launch_bounds__(256) min_eig_kernel(const uint8_t* __restrict__ gray, int w, int h, int pitch,
                                                      float* __restrict__ eig, int eig_pitch, DetectGrid grid,
                                                      int* __restrict__ cell_max) {
    __shared__ uint8_t g[ME_GH][ME_GW + 4];
    __shared__ float rx[ME_GH][ME_DW], rs[ME_GH][ME_DW];
    __shared__ float cxx[ME_DH][ME_DW], cxy[ME_DH][ME_DW], cyy[ME_DH][ME_DW];
    __shared__ int red[8];

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * ME_TW, y0 = blockIdx.y * ME_TH;   // tile origin (output coords)
    const float s = (float)(1.0 / (4.0 * 3.0 * 255.0));
    const float s2 = 2.0f * s;
    const int wvec = (w / 32) * 32;

    // gray region [x0-2, x0+TW+2) x [y0-2, y0+TH+2), reflected at the image borders
    for (int idx = tid; idx < ME_GH * ME_GW; idx += 256) {
        const int r = idx / ME_GW, c = idx - r * ME_GW;
        const int gy = reflect101(y0 - 2 + r, h), gx = reflect101(x0 - 2 + c, w);
        g[r][c] = gray[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    // horizontal pass on every region row, for columns x0-1 .. x0+TW
    for (int idx = tid; idx < ME_GH * ME_DW; idx += 256) {
        const int r = idx / ME_DW, c = idx - r * ME_DW;             // c <-> global x = x0-1+c
        const float pm = (float)g[r][c], pc_ = (float)g[r][c + 1], pp = (float)g[r][c + 2];
        rx[r][c] = __fsub_rn(pp, pm);
        const int gx = x0 - 1 + c;
        rs[r][c] = (gx < wvec) ? __fmaf_rn(pp, s, __fmaf_rn(pc_, s2, __fmul_rn(pm, s)))
                               : __fadd_rn(__fadd_rn(__fmul_rn(pm, s), __fmul_rn(pc_, s2)), __fmul_rn(pp, s));
    }
    __syncthreads();
    // vertical pass + covariance products on [x0-1, x0+TW] x [y0-1, y0+TH]
    for (int idx = tid; idx < ME_DH * ME_DW; idx += 256) {
        const int r = idx / ME_DW, c = idx - r * ME_DW;             // r <-> global y = y0-1+r ; region row r+1
        const float dx = __fmaf_rn(__fadd_rn(rx[r][c], rx[r + 2][c]), s, __fmul_rn(rx[r + 1][c], s2));
        const float dy = __fsub_rn(rs[r + 2][c], rs[r][c]);
        cxx[r][c] = __fmul_rn(dx, dx);
        cxy[r][c] = __fmul_rn(dx, dy);
        cyy[r][c] = __fmul_rn(dy, dy);
    }
    __syncthreads();
    // cov outside the image is the reflection of cov inside (not cov of reflected gray)
    auto lc = [&](int gx) { return reflect101(gx, w) - (x0 - 1); };  // local column of global x
    auto lr = [&](int gy) { return reflect101(gy, h) - (y0 - 1); };

    const int col = tid & (ME_TW - 1), rg = tid / ME_TW;            // 4 row groups of 4 rows
    const int gx = x0 + col;
    int local_max = 0x80000000;
    if (gx < w) {
        const int cm = lc(gx - 1), cc = col + 1, cp = lc(gx + 1);
        double Rxx[3], Rxy[3], Ryy[3];
        auto rowsum = [&](int gy, double& oxx, double& oxy, double& oyy) {
            const int r = lr(gy);
            oxx = __dadd_rn(__dadd_rn((double)cxx[r][cm], (double)cxx[r][cc]), (double)cxx[r][cp]);
            oxy = __dadd_rn(__dadd_rn((double)cxy[r][cm], (double)cxy[r][cc]), (double)cxy[r][cp]);
            oyy = __dadd_rn(__dadd_rn((double)cyy[r][cm], (double)cyy[r][cc]), (double)cyy[r][cp]);
        };
        const int gy0 = y0 + rg * 4;
        if (gy0 < h) {
            rowsum(gy0 - 1, Rxx[0], Rxy[0], Ryy[0]);
            rowsum(gy0, Rxx[1], Rxy[1], Ryy[1]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int gy = gy0 + k;
            if (gy >= h) break;
            rowsum(gy + 1, Rxx[2], Rxy[2], Ryy[2]);
            const float sxx = (float)__dadd_rn(__dadd_rn(Rxx[0], Rxx[1]), Rxx[2]);
            const float sxy = (float)__dadd_rn(__dadd_rn(Rxy[0], Rxy[1]), Rxy[2]);
            const float syy = (float)__dadd_rn(__dadd_rn(Ryy[0], Ryy[1]), Ryy[2]);
            const float a = __fmul_rn(sxx, 0.5f), b = sxy, c = __fmul_rn(syy, 0.5f);
            const float t = __fsub_rn(a, c);
            const float v = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b))));
            eig[(size_t)gy * eig_pitch + gx] = v;
            const int ov = float_to_ordered_int(v);
            const int cell = cell_of(gx, gy, grid);
            // tiles that straddle a cell boundary fall back to per-pixel atomics below
            if (cell == cell_of(x0, y0, grid)) local_max = max(local_max, ov);
            else atomicMax(&cell_max[cell], ov);
            Rxx[0] = Rxx[1]; Rxx[1] = Rxx[2];
            Rxy[0] = Rxy[1]; Rxy[1] = Rxy[2];
            Ryy[0] = Ryy[1]; Ryy[1] = Ryy[2];
        }
    }
    // block max of the pixels that share the tile origin's cell -> one atomic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((tid & 31) == 0) red[tid >> 5] = local_max;
    __syncthreads();
    if (tid == 0) {
        int m = red[0];
#pragma unroll
        for (int k = 1; k < 8; k++) m = max(m, red[k]);
        if (m != (int)0x80000000) atomicMax(&cell_max[cell_of(x0, y0, grid)], m);
    }
}

__global__ void init_cell_max_kernel(int* cell_max, int n, int* counters, int n_counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cell_max[i] = 0x80000000;
    if (i < n_counters) counters[i] = 0;
}

void launch_min_eig(Image8 gray, float* eig, int eig_pitch, DetectGrid g, int* cell_max, cudaStream_t s) {
    const int ncell = g.grid_rows * g.grid_cols;
    init_cell_max_kernel<<<(ncell + 255) / 256, 256, 0, s>>>(cell_max, ncell, nullptr, 0);
    dim3 grid((gray.w + ME_TW - 1) / ME_TW, (gray.h + ME_TH - 1) / ME_TH);
    min_eig_kernel<<<grid, 256, 0, s>>>(gray.data, gray.w, gray.h, gray.pitch, eig, eig_pitch, g, cell_max);
}

// ---- K5: threshold (per cell), 3x3 NMS, candidate list + state map ---------------------
// cv::threshold(THRESH_TOZERO, maxVal*quality): thresh is the double product rounded to
// float; value kept iff value > thresh (gftt.cc:61-65).  A pixel is a candidate iff it is
// interior, its thresholded value is non-zero and equals the 3x3 max (gftt.cc:70-86).
__global__ void __launch_bounds__(256) nms_candidates_kernel(const float* __restrict__ eig, int eig_pitch, int w,
                                                             int h, DetectGrid grid,
                                                             const int* __restrict__ cell_max, double quality,
                                                             uint8_t* __restrict__ state, int state_pitch,
                                                             unsigned long long* __restrict__ cand, int cand_cap,
                                                             int* __restrict__ cand_count,
                                                             int* __restrict__ value_hist) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    bool is_cand = false;
    float v = 0.f;
    if (x < w && y < h) {
        auto thr_at = [&](int cx, int cy) {
            const float m = ordered_int_to_float(__ldg(&cell_max[cell_of(cx, cy, grid)]));
            return (float)((double)m * quality);
        };
        auto tz = [&](float val, float thr) { return val > thr ? val : 0.f; };
        if (x >= 1 && y >= 1 && x < w - 1 && y < h - 1) {
            const float thr_c = thr_at(x, y);
            v = tz(eig[(size_t)y * eig_pitch + x], thr_c);
            if (v != 0.f) {
                const int bx = x % grid.block_w, by = y % grid.block_h;
                const bool inner_cell = bx > 0 && bx < grid.block_w - 1 && by > 0 && by < grid.block_h - 1;
                is_cand = true;
#pragma unroll
                for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                    for (int dx = -1; dx <= 1; dx++) {
                        if (dx == 0 && dy == 0) continue;
                        const float thr = inner_cell ? thr_c : thr_at(x + dx, y + dy);
                        const float nv = tz(eig[(size_t)(y + dy) * eig_pitch + x + dx], thr);
                        if (nv > v) is_cand = false;
                    }
            }
        }
        state[(size_t)y * state_pitch + x] = is_cand ? 1 : 0;
    }
    // warp-aggregated append
    const unsigned mask = __ballot_sync(0xffffffffu, is_cand);
    if (mask) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(mask) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(cand_count, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (is_cand) {
            const int slot = base + __popc(mask & ((1u << lane) - 1));
            const uint32_t ov = float_to_ordered_uint(v);
            if (slot < cand_cap) cand[slot] = ((unsigned long long)ov << 32) | (unsigned)(y * w + x);
            atomicAdd(&value_hist[ov >> 20], 1);
        }
    }
}

void launch_nms_candidates(const float* eig, int eig_pitch, int w, int h, DetectGrid g, const int* cell_max,
                           double quality_level, uint8_t* state, int state_pitch, unsigned long long* cand,
                           int cand_cap, int* cand_count, int* value_hist, cudaStream_t s) {
    cudaMemsetAsync(cand_count, 0, sizeof(int), s);
    cudaMemsetAsync(value_hist, 0, sizeof(int) * 4096, s);
    dim3 grid((w + 63) / 64, (h + 3) / 4);
    nms_candidates_kernel<<<grid, 256, 0, s>>>(eig, eig_pitch, w, h, g, cell_max, quality_level, state, state_pitch,
                                               cand, cand_cap, cand_count, value_hist);
}

}  // namespace pc
