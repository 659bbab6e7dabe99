// K1 + K2: RGB8 -> gray8 and the 5x5 binomial pyrDown, bit-exact integer arithmetic.
//
// Replaces cv::cvtColor(RGB2GRAY) (/root/reference/cpp/opticalflow.cc:259,298) and the
// image half of cv::buildOpticalFlowPyramid (/root/reference/cpp/opticalflow.cc:180-187).
//   gray = (R*9798 + G*19235 + B*3735 + 2^14) >> 15
//   down = ([1 4 6 4 1] x [1 4 6 4 1] + 128) >> 8 at even coordinates, BORDER_REFLECT_101
// Both are HBM-bound streaming kernels: 16-byte vector loads/stores on the gray conversion,
// shared-memory tiles with word-wide fills on the pyramid levels.  The Scharr derivative
// images OpenCV materialises per level (4 B/px) are never written: the LK kernel derives
// them from the level inside its window (lk.cu).
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace pc {

__device__ __forceinline__ uint32_t gray_of(uint32_t r, uint32_t g, uint32_t b) {
    return (r * 9798u + g * 19235u + b * 3735u + (1u << 14)) >> 15;
}

// 16 pixels per thread: three 16-byte loads -> one 16-byte store.
__global__ void __launch_bounds__(256) rgb_to_gray_vec16(const uint8_t* __restrict__ rgb, size_t stride,
                                                         uint8_t* __restrict__ gray, int w, int h, int pitch) {
    const int groups = w >> 4;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (gx >= groups) return;
    const uint4* src = reinterpret_cast<const uint4*>(rgb + (size_t)y * stride) + (size_t)gx * 3;
    uint4 v0 = ld_stream(src), v1 = ld_stream(src + 1), v2 = ld_stream(src + 2);
    uint32_t wds[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
    uint32_t out[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {       // 4 pixels = 12 bytes = 3 words
        const uint32_t a = wds[3 * q], b = wds[3 * q + 1], c = wds[3 * q + 2];
        const uint32_t p0 = gray_of(a & 255u, (a >> 8) & 255u, (a >> 16) & 255u);
        const uint32_t p1 = gray_of(a >> 24, b & 255u, (b >> 8) & 255u);
        const uint32_t p2 = gray_of((b >> 16) & 255u, b >> 24, c & 255u);
        const uint32_t p3 = gray_of((c >> 8) & 255u, (c >> 16) & 255u, c >> 24);
        out[q] = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
    }
    *reinterpret_cast<uint4*>(gray + (size_t)y * pitch + ((size_t)gx << 4)) = make_uint4(out[0], out[1], out[2], out[3]);
}

// Generic path (any width / alignment): one pixel per thread.
__global__ void __launch_bounds__(256) rgb_to_gray_scalar(const uint8_t* __restrict__ rgb, size_t stride,
                                                          uint8_t* __restrict__ gray, int x0, int w, int h,
                                                          int pitch) {
    const int x = x0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* p = rgb + (size_t)y * stride + (size_t)x * 3;
    gray[(size_t)y * pitch + x] = (uint8_t)gray_of(p[0], p[1], p[2]);
}

__global__ void __launch_bounds__(256) copy_gray_kernel(const uint8_t* __restrict__ src, size_t stride,
                                                        uint8_t* __restrict__ dst, int w, int h, int pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x < w) dst[(size_t)y * pitch + x] = src[(size_t)y * stride + x];
}

void launch_rgb_to_gray(const uint8_t* rgb, size_t stride, Image8 gray, cudaStream_t s) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(rgb) & 15) == 0) && (stride % 16 == 0);
    int done = 0;
    if (aligned && gray.w >= 16) {
        const int groups = gray.w >> 4;
        dim3 grid((groups + 255) / 256, gray.h);
        rgb_to_gray_vec16<<<grid, 256, 0, s>>>(rgb, stride, gray.data, gray.w, gray.h, gray.pitch);
        done = groups << 4;
    }
    if (done < gray.w) {
        dim3 grid((gray.w - done + 255) / 256, gray.h);
        rgb_to_gray_scalar<<<grid, 256, 0, s>>>(rgb, stride, gray.data, done, gray.w, gray.h, gray.pitch);
    }
}

void launch_copy_gray(const uint8_t* src, size_t stride, Image8 gray, cudaStream_t s) {
    dim3 grid((gray.w + 255) / 256, gray.h);
    copy_gray_kernel<<<grid, 256, 0, s>>>(src, stride, gray.data, gray.w, gray.h, gray.pitch);
}

// ---- pyrDown -------------------------------------------------------------------------
constexpr int PD_TW = 64, PD_TH = 16;                // output tile
constexpr int PD_IW = 2 * PD_TW + 8;                 // 136 input bytes per row (word aligned origin)
constexpr int PD_IH = 2 * PD_TH + 3;                 // 35 input rows

__global__ void __launch_bounds__(256) pyr_down_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch,
                                                       uint8_t* __restrict__ dst, int dw, int dh, int dpitch) {
    __shared__ __align__(16) uint8_t tile[PD_IH][PD_IW];
    __shared__ uint16_t hsum[PD_IH][PD_TW];
    const int tid = threadIdx.x;
    const int ax0 = 2 * PD_TW * blockIdx.x - 4;      // global x of tile column 0 (multiple of 4)
    const int ay0 = 2 * PD_TH * blockIdx.y - 2;      // global y of tile row 0
    constexpr int WPR = PD_IW / 4;
    for (int idx = tid; idx < PD_IH * WPR; idx += 256) {
        const int r = idx / WPR, wi = idx - r * WPR;
        const int gy = reflect101(ay0 + r, sh);
        const int gx = ax0 + 4 * wi;
        const uint8_t* row = src + (size_t)gy * spitch;
        uint32_t v;
        if (gx >= 0 && gx + 3 < sw) {
            v = __ldg(reinterpret_cast<const uint32_t*>(row + gx));
        } else {
            v = (uint32_t)row[reflect101(gx, sw)] | ((uint32_t)row[reflect101(gx + 1, sw)] << 8) |
                ((uint32_t)row[reflect101(gx + 2, sw)] << 16) | ((uint32_t)row[reflect101(gx + 3, sw)] << 24);
        }
        *reinterpret_cast<uint32_t*>(&tile[r][4 * wi]) = v;
    }
    __syncthreads();
    for (int idx = tid; idx < PD_IH * PD_TW; idx += 256) {
        const int r = idx / PD_TW, j = idx - r * PD_TW;
        const uint8_t* t = &tile[r][2 * j + 2];
        hsum[r][j] = (uint16_t)(t[0] + 4 * t[1] + 6 * t[2] + 4 * t[3] + t[4]);
    }
    __syncthreads();
    const int i = tid / 16, j4 = (tid % 16) * 4;
    const int oy = PD_TH * blockIdx.y + i, ox = PD_TW * blockIdx.x + j4;
    if (oy >= dh || ox >= dw) return;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int j = j4 + k, r = 2 * i;
        o[k] = (hsum[r][j] + 4 * hsum[r + 1][j] + 6 * hsum[r + 2][j] + 4 * hsum[r + 3][j] + hsum[r + 4][j] + 128) >> 8;
    }
    uint8_t* out = dst + (size_t)oy * dpitch + ox;
    if (ox + 3 < dw) {
        *reinterpret_cast<uint32_t*>(out) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
    } else {
        for (int k = 0; k < 4 && ox + k < dw; k++) out[k] = (uint8_t)o[k];
    }
}

void launch_pyr_down(Image8 src, Image8 dst, cudaStream_t s) {
    dim3 grid((dst.w + PD_TW - 1) / PD_TW, (dst.h + PD_TH - 1) / PD_TH);
    pyr_down_kernel<<<grid, 256, 0, s>>>(src.data, src.w, src.h, src.pitch, dst.data, dst.w, dst.h, dst.pitch);
}

// ---- apron ---------------------------------------------------------------------------
struct PadPlanes {
    uint8_t* data[kMaxLevels];
    int w[kMaxLevels], h[kMaxLevels], pitch[kMaxLevels];
};

// Work items of a plane: the top and bottom bands as aligned 4-byte words (corners included), then
// the left and right bands one byte per thread (16 consecutive lanes cover one row's run, so a warp
// touches two rows instead of 32).  Both parts read only interior pixels, so they do not depend on
// each other's writes.
__global__ void __launch_bounds__(256) pad_border_kernel(PadPlanes P) {
    const int L = blockIdx.y;
    const int w = P.w[L], h = P.h[L], pitch = P.pitch[L];
    uint8_t* img = P.data[L];
    const int words = (w + 2 * kPadX + 3) / 4;             // per band row, from x = -kPadX
    const int n_rows = 2 * kPadY * words;
    const int n_cols = 2 * kPadX * h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows + n_cols; i += gridDim.x * blockDim.x) {
        if (i < n_cols) {
            const int y = i / (2 * kPadX);
            const int c = i - y * (2 * kPadX);
            const int x = c < kPadX ? c - kPadX : w + (c - kPadX);
            uint8_t* row = img + (ptrdiff_t)y * pitch;
            row[x] = row[reflect101(x, w)];
        } else {
            const int j = i - n_cols;
            const int r = j / words;
            const int x0 = (j - r * words) * 4 - kPadX;
            const int y = r < kPadY ? r - kPadY : h + (r - kPadY);
            const uint8_t* src = img + (ptrdiff_t)reflect101(y, h) * pitch;
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) v |= (uint32_t)src[reflect101(x0 + k, w)] << (8 * k);
            *reinterpret_cast<uint32_t*>(img + (ptrdiff_t)y * pitch + x0) = v;
        }
    }
}

void launch_pad_border(const Image8* planes, int levels, cudaStream_t s) {
    PadPlanes P{};
    int most = 0;
    for (int L = 0; L < levels; L++) {
        P.data[L] = planes[L].data; P.w[L] = planes[L].w; P.h[L] = planes[L].h; P.pitch[L] = planes[L].pitch;
        most = std::max(most, 2 * kPadY * ((planes[L].w + 2 * kPadX + 3) / 4) + 2 * kPadX * planes[L].h);
    }
    dim3 grid(std::min((most + 255) / 256, 592), levels);
    pad_border_kernel<<<grid, 256, 0, s>>>(P);
}

}  // namespace pc
