// K8 + K9: pyramidal Lucas-Kanade tracking of one frame's keypoints into another frame,
// and the status==1 filter.
//
// Replaces cv::calcOpticalFlowPyrLK as the reference calls it
// (/root/reference/cpp/opticalflow.cc:119-125: win 10x10, maxLevel 3, COUNT+EPS(30, 0.01),
// flags 0, minEigThreshold 1e-4) and the filter loop of GenerateOpticalFlowForAPair
// (/root/reference/cpp/opticalflow.cc:130-147).  Arithmetic per SURVEY.md Appendix A.3,
// pinned bit-for-bit (status, positions, err) against cv2 4.13.0 by oracle/restate.c:
//   * bilinear weights are 14-bit integers (cvRound), patches are integers (x32),
//   * Scharr derivatives are formed from the level image inside the window (OpenCV's
//     per-level derivative images, 4 B/px, are never materialised),
//   * the float sums A11/A12/A22 and b1/b2 are accumulated in the order of OpenCV's 128-bit
//     SIMD loop: four lane accumulators over columns {k, k+4} of each row's first 8 columns,
//     a scalar accumulator over the remaining columns, total = scalar + ((q0+q2)+(q1+q3)).
//
// One warp per (pair, keypoint).  Lanes own window pixels p = lane + 32k; per-pixel integer
// products go through shared memory so that ten "chain" lanes can replay OpenCV's
// accumulation order exactly.  All eight pairs of a frame are one launch (blockIdx.y).
#include "common.cuh"
#include "kernels.h"

namespace pc {

namespace {

constexpr int W_BITS = 14;
constexpr int LK_WARPS = 8;

struct LevelRef {
    const uint8_t* img;
    int w, h, pitch;
};

__device__ __forceinline__ int pix_reflect(const LevelRef& L, int x, int y) {
    return L.img[(size_t)reflect101(y, L.h) * L.pitch + reflect101(x, L.w)];
}
__device__ __forceinline__ int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
    const float sc = (float)(1 << W_BITS);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(oma, omb), sc));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, omb), sc));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(oma, b), sc));
    w11 = (1 << W_BITS) - w00 - w01 - w10;
}

// Bilinear sample (x32) of level L at integer corner (X,Y) with the given weights.
__device__ __forceinline__ int sample_patch(const LevelRef& L, bool inside, int X, int Y, int w00, int w01, int w10,
                                            int w11) {
    int p00, p01, p10, p11;
    if (inside) {
        const uint8_t* r0 = L.img + (size_t)Y * L.pitch + X;
        const uint8_t* r1 = r0 + L.pitch;
        p00 = r0[0]; p01 = r0[1]; p10 = r1[0]; p11 = r1[1];
    } else {
        p00 = pix_reflect(L, X, Y); p01 = pix_reflect(L, X + 1, Y);
        p10 = pix_reflect(L, X, Y + 1); p11 = pix_reflect(L, X + 1, Y + 1);
    }
    return descale(p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11, W_BITS - 5);
}

template <int WIN>
__global__ void __launch_bounds__(LK_WARPS * 32) lk_kernel(LKBatch batch, LKParams prm) {
    constexpr int NPX = WIN * WIN;
    constexpr int PX = (NPX + 31) / 32;
    constexpr int NVEC = (WIN / 8) * 8;
    constexpr int NGRP = WIN / 8;
    __shared__ int s_prod[LK_WARPS][3][NPX];
    __shared__ float s_chain[LK_WARPS][16];

    const LKPair& pr = batch.pair[blockIdx.y];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pi = blockIdx.x * LK_WARPS + wib;
    const int npts = min(*pr.n_pts, batch.cap);
    if (pi >= npts) return;
    int* prodx = s_prod[wib][0];
    int* prody = s_prod[wib][1];
    int* prodz = s_prod[wib][2];
    float* chain = s_chain[wib];

    const float ptx = pr.pts[2 * pi], pty = pr.pts[2 * pi + 1];
    const float halfw = (WIN - 1) * 0.5f;
    const int nlevels = min(min(pr.a.levels, pr.b.levels), prm.max_level + 1);
    const double eps2 = prm.eps * prm.eps;
    float nextx = 0.f, nexty = 0.f;
    int status = 1;
    float err = 0.f;

    int pxx[PX], pxy[PX];          // window coordinates of the pixels this lane owns
#pragma unroll
    for (int k = 0; k < PX; k++) {
        const int p = lane + 32 * k;
        pxy[k] = p / WIN;
        pxx[k] = p - pxy[k] * WIN;
    }

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

        // ---- template: I (x32), Ix, Iy over the window; products for the A sums --------
        int Ival[PX], Ixv[PX], Iyv[PX];
#pragma unroll
        for (int k = 0; k < PX; k++) {
            const int p = lane + 32 * k;
            Ival[k] = 0; Ixv[k] = 0; Iyv[k] = 0;
            if (p < NPX) {
                const int X = ipx + pxx[k], Y = ipy + pxy[k];
                // 4x4 block around the 2x2 taps, reflected like the padded level image
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
                        // derivative images are zero outside the (unpadded) level
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
                prodx[p] = Ixv[k] * Ixv[k];
                prody[p] = Ixv[k] * Iyv[k];
                prodz[p] = Iyv[k] * Iyv[k];
            }
        }
        __syncwarp();
        // chain lanes: quantity q = lane / 5 (0: A11, 1: A12, 2: A22), chain c = lane % 5
        if (lane < 15) {
            const int q = lane / 5, c = lane - q * 5;
            const int* src = s_prod[wib][q];
            float acc = 0.f;
            if (c < 4) {
                for (int y = 0; y < WIN; y++)
#pragma unroll
                    for (int g = 0; g < NGRP; g++) {
                        acc = __fadd_rn((float)src[y * WIN + 8 * g + c], acc);
                        acc = __fadd_rn((float)src[y * WIN + 8 * g + c + 4], acc);
                    }
            } else {
                for (int y = 0; y < WIN; y++)
#pragma unroll
                    for (int x = NVEC; x < WIN; x++) acc = __fadd_rn(acc, (float)src[y * WIN + x]);
            }
            chain[lane] = acc;
        }
        __syncwarp();
        float Asum[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const float* c = chain + 5 * q;
            float tot = c[4];
            if (NGRP > 0) tot = __fadd_rn(tot, __fadd_rn(__fadd_rn(c[0], c[2]), __fadd_rn(c[1], c[3])));
            Asum[q] = tot;
        }
        __syncwarp();
        const float FLT_SCALE = 1.f / (float)(1 << 20);
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
        for (int j = 0; j < prm.iters; j++) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -WIN || inx >= B.w || iny < -WIN || iny >= B.h) {
                if (level == 0) status = 0;
                break;
            }
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            const bool inside = inx >= 0 && iny >= 0 && inx + WIN < B.w && iny + WIN < B.h;
#pragma unroll
            for (int k = 0; k < PX; k++) {
                const int p = lane + 32 * k;
                if (p < NPX) {
                    const int diff = sample_patch(B, inside, inx + pxx[k], iny + pxy[k], w00, w01, w10, w11) - Ival[k];
                    prodx[p] = diff * Ixv[k];
                    prody[p] = diff * Iyv[k];
                }
            }
            __syncwarp();
            if (lane < 10) {
                const int q = lane / 5, c = lane - q * 5;
                const int* src = s_prod[wib][q];
                float acc = 0.f;
                if (c < 4) {
                    for (int y = 0; y < WIN; y++)
#pragma unroll
                        for (int g = 0; g < NGRP; g++)
                            acc = __fadd_rn(acc, (float)(src[y * WIN + 8 * g + c] + src[y * WIN + 8 * g + c + 4]));
                } else {
                    for (int y = 0; y < WIN; y++)
#pragma unroll
                        for (int x = NVEC; x < WIN; x++) acc = __fadd_rn(acc, (float)src[y * WIN + x]);
                }
                chain[lane] = acc;
            }
            __syncwarp();
            float bsum[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const float* c = chain + 5 * q;
                float tot = c[4];
                if (NGRP > 0) tot = __fadd_rn(tot, __fadd_rn(__fadd_rn(c[0], c[2]), __fadd_rn(c[1], c[3])));
                bsum[q] = tot;
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
            const bool inside = inx >= 0 && iny >= 0 && inx + WIN < B.w && iny + WIN < B.h;
            int esum = 0;
#pragma unroll
            for (int k = 0; k < PX; k++) {
                const int p = lane + 32 * k;
                if (p < NPX)
                    esum += abs(sample_patch(B, inside, inx + pxx[k], iny + pxy[k], w00, w01, w10, w11) - Ival[k]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
            // all partial sums are integers < 2^24: any float summation order is exact
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

// K9: order-preserving compaction of status==1 rows (one block per pair).  A warp owns CP_PER * 32
// consecutive rows of each 1024 * CP_PER chunk and reads them as CP_PER coalesced groups of 32 (a
// thread owning CP_PER consecutive rows would make every warp load touch 32 different lines: one SM's
// L1 then bounds the kernel); slots come from ballots inside the warp and one shared-memory pass over
// the 32 warp totals per chunk.
constexpr int CP_PER = 8;
__global__ void __launch_bounds__(1024) lk_compact_kernel(LKBatch batch) {
    const LKPair& pr = batch.pair[blockIdx.x];
    __shared__ int warp_tot[2][32];
    const int n = min(*pr.n_pts, batch.cap);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1;
    int base = 0, buf = 0;
    for (int start = 0; start < n; start += 1024 * CP_PER, buf ^= 1) {
        const int w0 = start + wid * (32 * CP_PER);          // first row of this warp
        unsigned bal[CP_PER];
        float2 t[CP_PER];
        float e[CP_PER];
        int mine = 0;                                        // rows kept by the warp (uniform)
#pragma unroll
        for (int k = 0; k < CP_PER; k++) {
            const int i = w0 + 32 * k + lane;
            const bool keep = i < n && pr.status[i] == 1;
            bal[k] = __ballot_sync(0xffffffffu, keep);
            mine += __popc(bal[k]);
            t[k] = make_float2(0.f, 0.f);
            e[k] = 0.f;
            if (keep) {                                      // every load before the first store
                t[k] = __ldg(reinterpret_cast<const float2*>(pr.next + 2 * i));
                e[k] = __ldg(pr.err + i);
            }
        }
        if (lane == 0) warp_tot[buf][wid] = mine;
        __syncthreads();
        const int wt = warp_tot[buf][lane];                  // 32 warps: one total per lane
        int wincl = wt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wincl, o);
            if (lane >= o) wincl += v;
        }
        int slot = base + __shfl_sync(0xffffffffu, wincl - wt, wid);
        const int total = __shfl_sync(0xffffffffu, wincl, 31);
#pragma unroll
        for (int k = 0; k < CP_PER; k++) {
            if (bal[k] & (1u << lane)) {
                const int s = slot + __popc(bal[k] & lt);
                pr.out_idx[s] = (uint32_t)(w0 + 32 * k + lane);
                *reinterpret_cast<float2*>(pr.out_tgt + 2 * s) = t[k];
                pr.out_err[s] = e[k];
            }
            slot += __popc(bal[k]);
        }
        base += total;
    }
    if (threadIdx.x == 0) *pr.out_count = base;
}

template <int WIN>
void launch_lk_t(const LKBatch& batch, const LKParams& p, cudaStream_t s) {
    dim3 grid((batch.cap + LK_WARPS - 1) / LK_WARPS, batch.num_pairs);
    lk_kernel<WIN><<<grid, LK_WARPS * 32, 0, s>>>(batch, p);
}

}  // namespace

bool lk_window_supported(int win) { return win >= 3 && win <= 16; }

void launch_lk10(const LKBatch& batch, const LKParams& p, cudaStream_t s);   // lk10.cu

void launch_lk(const LKBatch& batch, const LKParams& p, cudaStream_t s) {
    switch (p.win) {
#define PC_LK_CASE(W) case W: launch_lk_t<W>(batch, p, s); break;
        PC_LK_CASE(3) PC_LK_CASE(4) PC_LK_CASE(5) PC_LK_CASE(6) PC_LK_CASE(7) PC_LK_CASE(8) PC_LK_CASE(9)
        case 10: launch_lk10(batch, p, s); break;
        PC_LK_CASE(11) PC_LK_CASE(12) PC_LK_CASE(13) PC_LK_CASE(14) PC_LK_CASE(15) PC_LK_CASE(16)
#undef PC_LK_CASE
        default: break;
    }
}

void launch_lk_compact(const LKBatch& batch, cudaStream_t s) {
    lk_compact_kernel<<<batch.num_pairs, 1024, 0, s>>>(batch);
}

}  // namespace pc
