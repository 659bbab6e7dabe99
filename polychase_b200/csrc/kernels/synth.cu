// Synthetic frame generator for benchmarks/tests: homography warp of a u8 texture into an
// RGB8 frame (R=G=B), bilinear sampling with REFLECT_101 borders.  Not on the reference
// path -- it only manufactures inputs of the shape BASELINE.json's configs name
// (SURVEY.md section 8d), device-side so a 4K clip can be staged in HBM in seconds.
#include "common.cuh"
#include "kernels.h"

namespace pc {

struct H9 { double m[9]; };

__global__ void __launch_bounds__(256) synth_warp_kernel(const uint8_t* __restrict__ tex, int w, int h, int tpitch,
                                                         H9 Hi, uint8_t* __restrict__ rgb, size_t stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w) return;
    const double den = Hi.m[6] * x + Hi.m[7] * y + Hi.m[8];
    const double u = (Hi.m[0] * x + Hi.m[1] * y + Hi.m[2]) / den;
    const double v = (Hi.m[3] * x + Hi.m[4] * y + Hi.m[5]) / den;
    const double fu0 = floor(u), fv0 = floor(v);
    const int u0 = (int)fu0, v0 = (int)fv0;
    const float fu = (float)(u - fu0), fv = (float)(v - fv0);
    const int x0 = reflect101(u0, w), x1 = reflect101(u0 + 1, w);
    const int y0 = reflect101(v0, h), y1 = reflect101(v0 + 1, h);
    const float t00 = tex[(size_t)y0 * tpitch + x0], t01 = tex[(size_t)y0 * tpitch + x1];
    const float t10 = tex[(size_t)y1 * tpitch + x0], t11 = tex[(size_t)y1 * tpitch + x1];
    const float top = t00 * (1.f - fu) + t01 * fu, bot = t10 * (1.f - fu) + t11 * fu;
    const float o = rintf(top * (1.f - fv) + bot * fv);
    const uint8_t g = (uint8_t)fminf(fmaxf(o, 0.f), 255.f);
    uint8_t* p = rgb + (size_t)y * stride + (size_t)x * 3;
    p[0] = g; p[1] = g; p[2] = g;
}

void launch_synth_warp(const uint8_t* tex, int w, int h, int tex_pitch, const double Hinv[9], uint8_t* rgb,
                       size_t stride, cudaStream_t s) {
    H9 Hi;
    for (int i = 0; i < 9; i++) Hi.m[i] = Hinv[i];
    dim3 grid((w + 255) / 256, h);
    synth_warp_kernel<<<grid, 256, 0, s>>>(tex, w, h, tex_pitch, Hi, rgb, stride);
}

}  // namespace pc
