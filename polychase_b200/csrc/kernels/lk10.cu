// K8 specialised for the reference's window (10x10, cpp/opticalflow.h:28): same arithmetic and
// accumulation order as lk.cu (which stays the generic path for other windows), restructured to
// cut the instruction count the ncu profile showed the generic kernel is bound by.
//
//   * Lane -> pixel mapping follows OpenCV's accumulation structure: lanes 0..19 own two
//     "pair items" (row y, column c) + (row y, column c+4) -- the two pixels whose integer
//     products OpenCV adds before converting to float -- lanes 20..29 own one row's tail
//     columns 8,9.  Producers therefore emit ready float terms; fifteen (template) / ten
//     (iteration) chain lanes only replay the ordered float adds.
//   * Interior windows (the common case) stage the 11x11 target patch / 13x13 source patch in
//     shared memory with 4-byte loads; bilinear taps are formed with funnel shifts + DP2A
//     (2x(16-bit weight x 8-bit pixel) per instruction).  Scharr derivatives are computed once
//     per patch pixel (121) instead of once per tap (400).
//   * Windows touching the border take the reflect-101 / zero-derivative path of lk.cu.
// Integer stages are exact, float stages use the same explicit-rounding intrinsics, so the two
// kernels produce identical bits (tests/test_gpu_analyze.py compares both against the oracle).
#include "common.cuh"
#include "kernels.h"

namespace pc {

namespace {

constexpr int WIN = 10;
constexpr int W_BITS = 14;
constexpr int LK_WARPS = 8;
constexpr int ROWW = 5;                 // staged patch row pitch in 32-bit words (20 bytes)

struct LevelRef {
    const uint8_t* img;
    int w, h, pitch;
};

struct __align__(16) WarpSmem {
    float terms[3 * 5 * 20];            // [q][chain][20] ordered float terms (16-byte aligned rows)
    float chain[16];
    uint32_t patch[13 * ROWW + 3];      // staged image patch (13 rows template / 11 rows iteration)
    uint32_t deriv[11 * 11 + 3];        // packed Scharr (dx | dy << 16) on the 11x11 tap grid
};

__device__ __forceinline__ int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }
__device__ __forceinline__ int pix_reflect(const LevelRef& L, int x, int y) {
    return L.img[(size_t)reflect101(y, L.h) * L.pitch + reflect101(x, L.w)];
}

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
    const float sc = (float)(1 << W_BITS);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(oma, omb), sc));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, omb), sc));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(oma, b), sc));
    w11 = (1 << W_BITS) - w00 - w01 - w10;
}

// Stage `rows` rows of 20 bytes starting at the 4-byte aligned address at or below (x0, y0).
// Each lane moves words `lane`, `lane + 32` (and `lane + 64` for the 13-row template patch);
// (row, word) of those indices are lane constants computed once (StageIdx).
struct StageIdx { int r0, w0, r1, w1; };
__device__ __forceinline__ void stage_patch(uint32_t* dst, const LevelRef& L, int x0, int y0, int rows, int lane,
                                            const StageIdx& si) {
    const uint32_t* base = reinterpret_cast<const uint32_t*>(L.img + (size_t)y0 * L.pitch + (x0 & ~3));
    const int pw = L.pitch >> 2;
    const uint32_t a = __ldg(base + si.r0 * pw + si.w0);
    if (lane + 32 < rows * ROWW) dst[lane + 32] = __ldg(base + si.r1 * pw + si.w1);
    dst[lane] = a;
    if (rows * ROWW > 64 && lane == 0) dst[64] = __ldg(base + 12 * pw + 4);
}

// c + a.lo16 * b.byte0 + a.hi16 * b.byte1 with SIGNED 16-bit weights (w11 can be -1 after
// rounding) and UNSIGNED pixel bytes: the mixed-sign form only exists in PTX.
__device__ __forceinline__ int dp2a_su(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// two adjacent bytes (x, x+1) of staged row `row`, as the low 16 bits
__device__ __forceinline__ uint32_t two_bytes(const uint32_t* patch, int row, int s) {
    const uint32_t* p = patch + row * ROWW + (s >> 2);
    return __funnelshift_r(p[0], p[1], (s & 3) * 8);
}

// Generic (border-safe) bilinear sample x32, as in lk.cu
__device__ __forceinline__ int sample_reflect(const LevelRef& L, int X, int Y, int w00, int w01, int w10, int w11) {
    return descale(pix_reflect(L, X, Y) * w00 + pix_reflect(L, X + 1, Y) * w01 + pix_reflect(L, X, Y + 1) * w10 +
                       pix_reflect(L, X + 1, Y + 1) * w11,
                   W_BITS - 5);
}

// ordered accumulation of 20 terms by the chain lanes (t is 16-byte aligned: 5 x LDS.128)
__device__ __forceinline__ float chain20(const float* t) {
    const float4* v = reinterpret_cast<const float4*>(t);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const float4 f = v[i];
        acc = __fadd_rn(acc, f.x); acc = __fadd_rn(acc, f.y); acc = __fadd_rn(acc, f.z); acc = __fadd_rn(acc, f.w);
    }
    return acc;
}

__global__ void __launch_bounds__(LK_WARPS * 32, 3) lk10_kernel(LKBatch batch, LKParams prm) {
    __shared__ WarpSmem s_all[LK_WARPS];
    const LKPair& pr = batch.pair[blockIdx.y];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pi = blockIdx.x * LK_WARPS + wib;
    const int npts = min(*pr.n_pts, batch.cap);
    if (pi >= npts) return;
    WarpSmem& S = s_all[wib];

    // lane -> owned pixels.  role 0: pair lane (4 px), 1: tail lane (2 px), 2: idle
    const int role = lane < 20 ? 0 : (lane < 30 ? 1 : 2);
    const int npx = role == 0 ? 4 : (role == 1 ? 2 : 0);
    int pxx[4], pxy[4];
    {
        const int c = lane & 3, ya = lane >> 2;
        pxx[0] = c; pxy[0] = ya; pxx[1] = c + 4; pxy[1] = ya;
        pxx[2] = c; pxy[2] = ya + 5; pxx[3] = c + 4; pxy[3] = ya + 5;
        if (role == 1) { pxx[0] = 8; pxy[0] = lane - 20; pxx[1] = 9; pxy[1] = lane - 20; }
    }
    // where this lane's float terms go: template layout [q][chain][20], iteration [q][chain][10|20]
    // chain 0..3 = SIMD lanes c, chain 4 = scalar tail
    const int chain_id = role == 0 ? (lane & 3) : 4;
    const StageIdx si = {lane / ROWW, lane % ROWW, (lane + 32) / ROWW, (lane + 32) % ROWW};

    const float ptx = pr.pts[2 * pi], pty = pr.pts[2 * pi + 1];
    const float halfw = (WIN - 1) * 0.5f;
    const int nlevels = min(min(pr.a.levels, pr.b.levels), prm.max_level + 1);
    const double eps2 = prm.eps * prm.eps;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    float nextx = 0.f, nexty = 0.f;
    int status = 1;
    float err = 0.f;

    for (int level = nlevels - 1; level >= 0; level--) {
        const LevelRef A = {pr.a.data[level], pr.a.w[level], pr.a.h[level], pr.a.pitch[level]};
        const LevelRef B = {pr.b.data[level], pr.b.w[level], pr.b.h[level], pr.b.pitch[level]};
        const float scale = 1.f / (float)(1 << level);
        float prevx = __fmul_rn(ptx, scale), prevy = __fmul_rn(pty, scale);
        if (level == nlevels - 1) { nextx = prevx; nexty = prevy; }
        else { nextx = __fmul_rn(nextx, 2.f); nexty = __fmul_rn(nexty, 2.f); }
        prevx = __fsub_rn(prevx, halfw); prevy = __fsub_rn(prevy, halfw);
        const int ipx = __float2int_rd(prevx), ipy = __float2int_rd(prevy);
        if (ipx < -WIN || ipx >= A.w || ipy < -WIN || ipy >= A.h) {
            if (level == 0) { status = 0; err = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        bilinear_weights(__fsub_rn(prevx, (float)ipx), __fsub_rn(prevy, (float)ipy), w00, w01, w10, w11);

        // ---- template ------------------------------------------------------------------------
        int Ival[4], Ixv[4], Iyv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { Ival[k] = 0; Ixv[k] = 0; Iyv[k] = 0; }
        // the 13x13 source patch [ipx-1, ipx+11] x [ipy-1, ipy+11] lies inside the level
        const bool t_inside = ipx >= 1 && ipy >= 1 && ipx + WIN + 1 < A.w && ipy + WIN + 1 < A.h;
        if (t_inside) {
            stage_patch(S.patch, A, ipx - 1, ipy - 1, 13, lane, si);
            __syncwarp();
            const int o = (ipx - 1) & 3;
            // Scharr on the 11x11 tap grid (grid (i,j) <-> patch pixel (i+1, j+1))
            for (int g = lane; g < 121; g += 32) {
                const int j = g / 11, i = g - j * 11;
                const int s = o + i;                       // byte offset of patch column i in a row
                const uint32_t* r0 = S.patch + j * ROWW + (s >> 2);
                const int sh = (s & 3) * 8;
                // 3 bytes (i, i+1, i+2) of rows j, j+1, j+2
                const uint32_t a = __funnelshift_r(r0[0], r0[1], sh);
                const uint32_t b = __funnelshift_r(r0[ROWW], r0[ROWW + 1], sh);
                const uint32_t c = __funnelshift_r(r0[2 * ROWW], r0[2 * ROWW + 1], sh);
                const int a0 = a & 255, a1 = (a >> 8) & 255, a2 = (a >> 16) & 255;
                const int b0 = b & 255, b2 = (b >> 16) & 255;
                const int c0 = c & 255, c1 = (c >> 8) & 255, c2 = (c >> 16) & 255;
                const int t0m = 3 * (a0 + c0) + 10 * b0, t0p = 3 * (a2 + c2) + 10 * b2;
                const int t1m = c0 - a0, t1c = c1 - a1, t1p = c2 - a2;
                const int dx = t0p - t0m, dy = 3 * (t1m + t1p) + 10 * t1c;
                S.deriv[g] = ((uint32_t)dx & 0xffffu) | ((uint32_t)dy << 16);
            }
            __syncwarp();
            const uint32_t wa = ((uint32_t)w00 & 0xffffu) | ((uint32_t)w01 << 16), wb = ((uint32_t)w10 & 0xffffu) | ((uint32_t)w11 << 16);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k < npx) {
                    const int x = pxx[k], y = pxy[k];
                    const int s = o + x + 1;               // window pixel (x,y) = patch pixel (x+1, y+1)
                    const uint32_t p0 = two_bytes(S.patch, y + 1, s), p1 = two_bytes(S.patch, y + 2, s);
                    Ival[k] = descale(dp2a_su(wa, p0, dp2a_su(wb, p1, 0)), W_BITS - 5);
                    const uint32_t d00 = S.deriv[y * 11 + x], d01 = S.deriv[y * 11 + x + 1];
                    const uint32_t d10 = S.deriv[(y + 1) * 11 + x], d11 = S.deriv[(y + 1) * 11 + x + 1];
                    const int sx = (int)(short)(d00 & 0xffff) * w00 + (int)(short)(d01 & 0xffff) * w01 +
                                   (int)(short)(d10 & 0xffff) * w10 + (int)(short)(d11 & 0xffff) * w11;
                    const int sy = ((int)d00 >> 16) * w00 + ((int)d01 >> 16) * w01 + ((int)d10 >> 16) * w10 +
                                   ((int)d11 >> 16) * w11;
                    Ixv[k] = descale(sx, W_BITS);
                    Iyv[k] = descale(sy, W_BITS);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k < npx) {
                    const int X = ipx + pxx[k], Y = ipy + pxy[k];
                    int blk[4][4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint8_t* row = A.img + (size_t)reflect101(Y - 1 + j, A.h) * A.pitch;
#pragma unroll
                        for (int i = 0; i < 4; i++) blk[j][i] = row[reflect101(X - 1 + i, A.w)];
                    }
                    Ival[k] = descale(blk[1][1] * w00 + blk[1][2] * w01 + blk[2][1] * w10 + blk[2][2] * w11, W_BITS - 5);
                    int dxs[2][2], dys[2][2];
#pragma unroll
                    for (int b = 0; b < 2; b++)
#pragma unroll
                        for (int a = 0; a < 2; a++) {
                            const bool in = (unsigned)(X + a) < (unsigned)A.w && (unsigned)(Y + b) < (unsigned)A.h;
                            const int t0m = 3 * (blk[b][a] + blk[b + 2][a]) + 10 * blk[b + 1][a];
                            const int t0p = 3 * (blk[b][a + 2] + blk[b + 2][a + 2]) + 10 * blk[b + 1][a + 2];
                            const int t1m = blk[b + 2][a] - blk[b][a];
                            const int t1c = blk[b + 2][a + 1] - blk[b][a + 1];
                            const int t1p = blk[b + 2][a + 2] - blk[b][a + 2];
                            dxs[b][a] = in ? (t0p - t0m) : 0;
                            dys[b][a] = in ? (3 * (t1m + t1p) + 10 * t1c) : 0;
                        }
                    Ixv[k] = descale(dxs[0][0] * w00 + dxs[0][1] * w01 + dxs[1][0] * w10 + dxs[1][1] * w11, W_BITS);
                    Iyv[k] = descale(dys[0][0] * w00 + dys[0][1] * w01 + dys[1][0] * w10 + dys[1][1] * w11, W_BITS);
                }
            }
        }
        // ordered terms of A11, A12, A22: each pixel is its own term (OpenCV adds the SIMD half
        // for column c, then the one for column c+4, row by row; the tail adds column 8 then 9)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k < npx) {
                const int slot = role == 0 ? (2 * pxy[k] + (k & 1)) : (2 * pxy[k] + k);
                float* t = S.terms + chain_id * 20 + slot;
                t[0 * 100] = (float)(Ixv[k] * Ixv[k]);
                t[1 * 100] = (float)(Ixv[k] * Iyv[k]);
                t[2 * 100] = (float)(Iyv[k] * Iyv[k]);
            }
        }
        __syncwarp();
        if (lane < 15) S.chain[lane] = chain20(S.terms + (lane / 5) * 100 + (lane % 5) * 20);
        __syncwarp();
        float Asum[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const float* c = S.chain + 5 * q;
            Asum[q] = __fadd_rn(c[4], __fadd_rn(__fadd_rn(c[0], c[2]), __fadd_rn(c[1], c[3])));
        }
        __syncwarp();
        // iteration layout reuses [q][chain][20]: SIMD chains hold 10 row terms + 10 zeros (x + 0 = x)
        for (int i = lane; i < 80; i += 32) {
            const int q = i / 40, r = i - q * 40;
            S.terms[q * 100 + (r / 10) * 20 + 10 + (r % 10)] = 0.f;
        }
        const float A11 = __fmul_rn(Asum[0], FLT_SCALE), A12 = __fmul_rn(Asum[1], FLT_SCALE),
                    A22 = __fmul_rn(Asum[2], FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float rad = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * WIN * WIN));
        if ((double)minEig < prm.min_eig || D < 1.1920928955078125e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        float nx = __fsub_rn(nextx, halfw), ny = __fsub_rn(nexty, halfw);
        float pdx = 0.f, pdy = 0.f;

        // ---- iterations ----------------------------------------------------------------------
        for (int j = 0; j < prm.iters; j++) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -WIN || inx >= B.w || iny < -WIN || iny >= B.h) {
                if (level == 0) status = 0;
                break;
            }
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            const bool inside = inx >= 0 && iny >= 0 && inx + WIN < B.w && iny + WIN < B.h;
            int diff[4];
            if (inside) {
                stage_patch(S.patch, B, inx, iny, 11, lane, si);
                __syncwarp();
                const int o = inx & 3;
                const uint32_t wa = ((uint32_t)w00 & 0xffffu) | ((uint32_t)w01 << 16), wb = ((uint32_t)w10 & 0xffffu) | ((uint32_t)w11 << 16);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    diff[k] = 0;
                    if (k < npx) {
                        const int s = o + pxx[k];
                        const uint32_t p0 = two_bytes(S.patch, pxy[k], s), p1 = two_bytes(S.patch, pxy[k] + 1, s);
                        diff[k] = descale(dp2a_su(wa, p0, dp2a_su(wb, p1, 0)), W_BITS - 5) - Ival[k];
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    diff[k] = 0;
                    if (k < npx) diff[k] = sample_reflect(B, inx + pxx[k], iny + pxy[k], w00, w01, w10, w11) - Ival[k];
                }
            }
            // terms: pair lanes add the two pixels' integer products first (pmaddwd), then convert
            if (role == 0) {
                float* t = S.terms + chain_id * 20;
                t[pxy[0]] = (float)(diff[0] * Ixv[0] + diff[1] * Ixv[1]);
                t[pxy[2]] = (float)(diff[2] * Ixv[2] + diff[3] * Ixv[3]);
                t[100 + pxy[0]] = (float)(diff[0] * Iyv[0] + diff[1] * Iyv[1]);
                t[100 + pxy[2]] = (float)(diff[2] * Iyv[2] + diff[3] * Iyv[3]);
            } else if (role == 1) {
                float* t = S.terms + 80 + 2 * pxy[0];
                t[0] = (float)(diff[0] * Ixv[0]);
                t[1] = (float)(diff[1] * Ixv[1]);
                t[100] = (float)(diff[0] * Iyv[0]);
                t[101] = (float)(diff[1] * Iyv[1]);
            }
            __syncwarp();
            if (lane < 10) S.chain[lane] = chain20(S.terms + (lane / 5) * 100 + (lane % 5) * 20);
            __syncwarp();
            float bsum[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const float* c = S.chain + 5 * q;
                bsum[q] = __fadd_rn(c[4], __fadd_rn(__fadd_rn(c[0], c[2]), __fadd_rn(c[1], c[3])));
            }
            __syncwarp();
            const float b1 = __fmul_rn(bsum[0], FLT_SCALE), b2 = __fmul_rn(bsum[1], FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            nextx = __fadd_rn(nx, halfw); nexty = __fadd_rn(ny, halfw);
            if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps2) break;
            if (j > 0 && fabs((double)__fadd_rn(dx, pdx)) < 0.01 && fabs((double)__fadd_rn(dy, pdy)) < 0.01) {
                nextx = __fsub_rn(nextx, __fmul_rn(dx, 0.5f));
                nexty = __fsub_rn(nexty, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {
            const float fx = __fsub_rn(nextx, halfw), fy = __fsub_rn(nexty, halfw);
            const int inx = __float2int_rd(fx), iny = __float2int_rd(fy);
            if (inx < -WIN || inx >= B.w || iny < -WIN || iny >= B.h) {
                status = 0;
                continue;
            }
            bilinear_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), w00, w01, w10, w11);
            int esum = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < npx) esum += abs(sample_reflect(B, inx + pxx[k], iny + pxy[k], w00, w01, w10, w11) - Ival[k]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
            err = __fdiv_rn(__fmul_rn((float)esum, 1.f), (float)(32 * WIN * WIN));
        }
    }
    if (lane == 0) {
        pr.next[2 * pi] = nextx;
        pr.next[2 * pi + 1] = nexty;
        pr.status[pi] = (uint8_t)status;
        pr.err[pi] = err;
    }
}

}  // namespace

void launch_lk10(const LKBatch& batch, const LKParams& p, cudaStream_t s) {
    dim3 grid((batch.cap + LK_WARPS - 1) / LK_WARPS, batch.num_pairs);
    lk10_kernel<<<grid, LK_WARPS * 32, 0, s>>>(batch, p);
}

}  // namespace pc
