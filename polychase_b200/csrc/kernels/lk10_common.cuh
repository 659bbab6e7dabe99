// Device helpers shared by the 10x10 LK kernels (lk10.cu, lk10q.cu): the pentad layout, OpenCV's
// 128-bit SIMD accumulation order, DP2A bilinear taps.  See the header of lk10.cu.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace pc {
namespace lk10 {

constexpr int WIN = 10;
constexpr int W_BITS = 14;
constexpr int LK_WARPS = 4;
constexpr int LK_CACHED_BLOCKS = 4;   // 5 (96 registers, 48 B of spills) measures 481 us against 426 us
constexpr int PENTAD = 5;
constexpr int PTS_PER_WARP = 6;
constexpr unsigned FULL = 0xffffffffu;

struct LevelRef {
    const uint8_t* img;
    int w, h, pitch;
};

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
    const float sc = (float)(1 << W_BITS);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(oma, omb), sc));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, omb), sc));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(oma, b), sc));
    w11 = (1 << W_BITS) - w00 - w01 - w10;
}

// c + a.lo16 * b.byte0 + a.hi16 * b.byte1 with SIGNED 16-bit weights (w11 can be -1 after
// rounding) and UNSIGNED pixel bytes: the mixed-sign form only exists in PTX.
__device__ __forceinline__ int dp2a_su(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint32_t pack_weights(int lo, int hi) {
    return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16);
}

// tail + ((q0 + q2) + (q1 + q3)) over the five chain totals of a pentad (every lane of the
// pentad gets the same bits: float addition is commutative, so the pairwise exchange gives each
// SIMD lane the same two partial sums)
__device__ __forceinline__ float pentad_total(float v, int base, int role) {
    const float a = __fadd_rn(v, __shfl_sync(FULL, v, base + ((role ^ 2) & 3)));       // q0+q2 | q1+q3 (roles 0..3)
    const float b = __fadd_rn(a, __shfl_sync(FULL, a, base + ((role ^ 1) & 3)));       // (q0+q2)+(q1+q3)
    const float quad = __shfl_sync(FULL, b, base);
    const float tail = __shfl_sync(FULL, v, base + 4);
    return __fadd_rn(tail, quad);
}

// ---- row access: aligned 32-bit loads + funnel shifts -------------------------------------------
// Pyramid levels carry a REFLECT_101 apron (kPadX / kPadY, kernels.h), so a window that hangs over
// the image edge is read like any other.  A lane's two column sets are 4 apart (SIMD chains:
// columns c and c+4) or adjacent (tail chain: columns 8 and 9), so everything it needs from one
// source row lies in the 12 bytes that start at the 4-byte boundary below its first byte: three
// aligned word loads off ONE row pointer (immediate offsets 0/4/8) and funnel shifts replace
// per-byte gathers.  `step` is 8 * (first byte & 3); `simd` selects the column-set distance.
__device__ __forceinline__ void load_row_words(const uint8_t* rowp, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(rowp);
    w0 = __ldg(q); w1 = __ldg(q + 1); w2 = __ldg(q + 2);
}

// ---- template: Ival / Ix / Iy of the lane's 20 pixels + its chain of A11, A12, A22 terms ------
// Pixel arrays are indexed [2 * row + s], s = 0 for the lane's first column (xa), 1 for the second.
// MASK: derivative taps outside the image are zero (cv::buildOpticalFlowPyramid pads the derivative
// image with BORDER_CONSTANT); a warp takes this variant only when one of its keypoints' 13x13
// source patches leaves the image.
template <bool MASK>
__device__ __forceinline__ void template_pass(const LevelRef& A, int ipx, int ipy, int xa, int simd, int w00,
                                                   int w01, int w10, int w11, int (&Ival)[20], int (&Ix)[20],
                                                   int (&Iy)[20], float& a11, float& a12, float& a22) {
    const int X0 = ipx + xa - 1;                      // first byte of column set 0 (taps X-1 .. X+2)
    int off = (ipy - 1) * A.pitch + (X0 & ~3);      // 32-bit row offsets: one wide add per row address
    const int step = (X0 & 3) * 8;
    const int step2 = simd ? 0 : 8;                   // tail: set 1 starts one byte after set 0
    const uint32_t wa = pack_weights(w00, w01), wb = pack_weights(w10, w11);
    bool tap_in[2][2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int X = ipx + xa + (s ? (simd ? 4 : 1) : 0);
        tap_in[s][0] = (unsigned)X < (unsigned)A.w;
        tap_in[s][1] = (unsigned)(X + 1) < (unsigned)A.w;
    }
    uint32_t E[2][3] = {}, O[2][3] = {};
    int ival_top[2] = {0, 0};
    int tx[2] = {0, 0}, ty[2] = {0, 0};
    a11 = 0.f; a12 = 0.f; a22 = 0.f;
#pragma unroll
    for (int rr = 0; rr < WIN + 3; rr++) {            // source rows ipy-1 .. ipy+11
        uint32_t w0, w1, w2;
        load_row_words(A.img + (ptrdiff_t)off, w0, w1, w2);
        off += A.pitch;
        uint32_t Q[2];
        Q[0] = __funnelshift_r(w0, w1, step);         // bytes g0 g1 g2 g3 of set 0
        const uint32_t N = __funnelshift_r(w1, w2, step);
        Q[1] = __funnelshift_r(simd ? N : Q[0], N, step2);
        uint32_t P[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            E[s][0] = E[s][1]; E[s][1] = E[s][2]; E[s][2] = Q[s] & 0x00ff00ffu;          // g0 | g2 << 16
            O[s][0] = O[s][1]; O[s][1] = O[s][2]; O[s][2] = (Q[s] >> 8) & 0x00ff00ffu;   // g1 | g3 << 16
            P[s] = Q[s] >> 8;                                                             // g1 | g2 << 8 (dp2a.lo)
        }
        if (rr >= 2) {
            const int y = rr - 2;
            if (y < WIN) {
#pragma unroll
                for (int s = 0; s < 2; s++) Ival[2 * y + s] = dp2a_su(wb, P[s], ival_top[s]) >> (W_BITS - 5);
            }
        }
        if (rr >= 1 && rr <= WIN) {
#pragma unroll
            for (int s = 0; s < 2; s++) ival_top[s] = dp2a_su(wa, P[s], 1 << (W_BITS - 5 - 1));
        }
        if (rr >= 2) {
            const int t = rr - 2;                     // derivative tap row: image row ipy + t
            const bool row_in = !MASK || (unsigned)(ipy + t) < (unsigned)A.h;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const uint32_t VE = 3u * (E[s][0] + E[s][2]) + 10u * E[s][1];
                const uint32_t VO = 3u * (O[s][0] + O[s][2]) + 10u * O[s][1];
                const uint32_t UE = E[s][2] + 0x01000100u - E[s][0];
                const uint32_t UO = O[s][2] + 0x01000100u - O[s][0];
                int dx0 = (int)(VE >> 16) - (int)(VE & 0xffffu);
                int dx1 = (int)(VO >> 16) - (int)(VO & 0xffffu);
                const int u0 = UE & 0xffffu, u2 = UE >> 16, u1 = UO & 0xffffu, u3 = UO >> 16;
                int dy0 = 3 * (u0 + u2) + 10 * u1 - 4096;
                int dy1 = 3 * (u1 + u3) + 10 * u2 - 4096;
                if (MASK) {
                    if (!(row_in && tap_in[s][0])) { dx0 = 0; dy0 = 0; }
                    if (!(row_in && tap_in[s][1])) { dx1 = 0; dy1 = 0; }
                }
                if (t >= 1) {
                    const int y = t - 1;
                    Ix[2 * y + s] = (tx[s] + dx0 * w10 + dx1 * w11) >> W_BITS;
                    Iy[2 * y + s] = (ty[s] + dy0 * w10 + dy1 * w11) >> W_BITS;
                }
                if (t < WIN) {
                    tx[s] = dx0 * w00 + dx1 * w01 + (1 << (W_BITS - 1));
                    ty[s] = dy0 * w00 + dy1 * w01 + (1 << (W_BITS - 1));
                }
            }
            if (t >= 1) {
                const int y = t - 1;
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const int ix = Ix[2 * y + s], iy = Iy[2 * y + s];
                    a11 = __fadd_rn(a11, (float)(ix * ix));
                    a12 = __fadd_rn(a12, (float)(ix * iy));
                    a22 = __fadd_rn(a22, (float)(iy * iy));
                }
            }
        }
    }
}

// ---- one pass over the target window: b1/b2 chain terms (ERR = false) or the L1 error -------
template <bool ERR>
__device__ __forceinline__ void window_pass(const LevelRef& B, int inx, int iny, int xa, int simd, int w00,
                                                 int w01, int w10, int w11, const int (&Ival)[20],
                                                 const int (&Ix)[20], const int (&Iy)[20], float& bx, float& by,
                                                 int& esum) {
    const int X0 = inx + xa;
    int off = iny * B.pitch + (X0 & ~3);           // 32-bit row offsets: one IMAD.WIDE per row address
    const int step = (X0 & 3) * 8;
    const int step2 = simd ? step : 8;
    const uint32_t wa = pack_weights(w00, w01), wb = pack_weights(w10, w11);
    int top[2] = {0, 0};
    bx = 0.f; by = 0.f; esum = 0;
#pragma unroll
    for (int rr = 0; rr <= WIN; rr++) {               // target rows iny .. iny+10
        uint32_t w0, w1, w2;
        load_row_words(B.img + (ptrdiff_t)off, w0, w1, w2);
        off += B.pitch;
        uint32_t P[2];
        P[0] = __funnelshift_r(w0, w1, step);         // bytes X0, X0+1 in the low half (dp2a.lo)
        P[1] = __funnelshift_r(simd ? w1 : P[0], w2, step2);
        if (rr >= 1) {
            const int y = rr - 1;
            const int d0 = (dp2a_su(wb, P[0], top[0]) >> (W_BITS - 5)) - Ival[2 * y];
            const int d1 = (dp2a_su(wb, P[1], top[1]) >> (W_BITS - 5)) - Ival[2 * y + 1];
            if (ERR) {
                esum += abs(d0) + abs(d1);
            } else {
                // SIMD chains add the two columns' integer products before converting (pmaddwd);
                // the tail chain converts and adds them one by one (x + (+0) == x for the others)
                const int d1s = simd ? d1 : 0, d1t = simd ? 0 : d1;
                bx = __fadd_rn(bx, (float)(d0 * Ix[2 * y] + d1s * Ix[2 * y + 1]));
                by = __fadd_rn(by, (float)(d0 * Iy[2 * y] + d1s * Iy[2 * y + 1]));
                bx = __fadd_rn(bx, (float)(d1t * Ix[2 * y + 1]));
                by = __fadd_rn(by, (float)(d1t * Iy[2 * y + 1]));
            }
        }
        if (rr < WIN) {
            top[0] = dp2a_su(wa, P[0], 1 << (W_BITS - 5 - 1));
            top[1] = dp2a_su(wa, P[1], 1 << (W_BITS - 5 - 1));
        }
    }
}

}  // namespace lk10
}  // namespace pc
