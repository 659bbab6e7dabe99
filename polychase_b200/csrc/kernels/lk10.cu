// K8 specialised for the reference's window (10x10, cpp/opticalflow.h:28): the arithmetic and the
// float accumulation order of cv::calcOpticalFlowPyrLK's 128-bit SIMD path (restated in
// oracle/restate.c::orc_lk, call site /root/reference/cpp/opticalflow.cc:119-125), laid out so
// that the ordered float sums never leave registers.
//
// OpenCV accumulates every 2x2-system sum (A11, A12, A22 of the template, b1, b2 of each
// iteration) in FIVE independent float chains: SIMD lane c (c = 0..3) takes columns c and c+4 of
// every window row, top to bottom, and a scalar tail takes columns 8 and 9; the total is
// tail + ((q0 + q2) + (q1 + q3)).  Here five CUDA lanes (a "pentad") own one keypoint and each
// lane IS one of those chains: it owns the two columns of its chain (20 pixels: Ival, Ix, Iy in
// registers), loads exactly the bytes those pixels need, and adds its terms in OpenCV's order.
// Only the five chain totals cross lanes (shuffles).  Six keypoints share a warp (lanes 30, 31
// idle).  Compared with the previous warp-per-keypoint layout (ordered terms exchanged through
// shared memory, ten chain lanes replaying the adds) this needs ~4x fewer warp instructions per
// keypoint; see profiles/.
//
//   * bilinear taps: the two horizontally adjacent bytes of a tap are packed into 16 bits and
//     combined with DP2A (2 x (signed 16-bit weight x unsigned byte) per instruction), top row
//     then bottom row, rounding constant folded into the first accumulator.
//   * template: Scharr derivatives are formed separably on bytes packed two per register
//     (16-bit fields: |3a+10b+3c| <= 4080, differences biased by +256), once per tap row.
//   * borders: the pyramid levels carry a REFLECT_101 apron, so windows that hang over the image
//     edge need no index arithmetic; derivative taps outside the image are zeroed (the padding
//     OpenCV's buildOpticalFlowPyramid applies) in a masked variant of the template pass that a
//     warp takes only when one of its keypoints' patches leaves the image at that level.
// Integer stages are exact and float stages use explicit-rounding intrinsics (file is compiled
// with -fmad=false), so results are bit-identical to the oracle (tests/test_gpu_analyze.py).
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "lk10_common.cuh"

namespace pc {

namespace {

using namespace lk10;


// ---- template cache ---------------------------------------------------------------------------
// A frame's keypoints are the source of up to eight pairs (four in the batch of its own frame, one in
// each of the batches of frames +1, +2, +4, +8), and the template of a (keypoint, level) -- the
// lane's 20 pixels of Ival / Ix / Iy and the three structure-tensor sums -- depends on the source
// frame only.  lk10_template_kernel computes them once per frame; the LK kernel then loads a lane's
// 32 packed words (8 x 16 bytes, the pentad's five lanes read five consecutive 128-byte lines)
// instead of re-running the 13-row template pass (~1200 instructions) for every pair.
//   words 0..9 : Ival[2k] | Ival[2k+1] << 16   (0 <= Ival <= 8160)
//   words 10..29: (Ix[i] & 0xffff) | Iy[i] << 16 (|Ix|, |Iy| <= 4080)
__global__ void __launch_bounds__(LK_WARPS * 32, 4) lk10_template_kernel(PyramidView a, const float* __restrict__ pts,
                                                                         const int* __restrict__ n_pts, int cap, int nlev,
                                                                         uint4* __restrict__ words,
                                                                         float* __restrict__ sums,
                                                                         const int* __restrict__ order) {
    const int level = blockIdx.y;
    if (level >= min(a.levels, nlev)) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / PENTAD, role = lane - grp * PENTAD;
    const int base = grp < PTS_PER_WARP ? grp * PENTAD : 0;
    const int npts = min(*n_pts, cap);
    const int slot = (blockIdx.x * LK_WARPS + wib) * PTS_PER_WARP + grp;
    const bool valid = grp < PTS_PER_WARP && slot < npts;
    if (!__any_sync(FULL, valid)) return;
    const int pi = (valid && order) ? __ldg(order + slot) : slot;      // which keypoint this pentad works on
    const int simd_flag = role < 4 ? 1 : 0;
    const int xa = simd_flag ? role : 8;
    float ptx = 0.f, pty = 0.f;
    if (valid) { ptx = pts[2 * pi]; pty = pts[2 * pi + 1]; }
    const float halfw = (WIN - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    const LevelRef A = {a.data[level], a.w[level], a.h[level], a.pitch[level]};
    const float scale = 1.f / (float)(1 << level);
    const float prevx = __fsub_rn(__fmul_rn(ptx, scale), halfw), prevy = __fsub_rn(__fmul_rn(pty, scale), halfw);
    const int ipx = __float2int_rd(prevx), ipy = __float2int_rd(prevy);
    const bool act = valid && !(ipx < -WIN || ipx >= A.w || ipy < -WIN || ipy >= A.h);
    int w00, w01, w10, w11;
    bilinear_weights(__fsub_rn(prevx, (float)ipx), __fsub_rn(prevy, (float)ipy), w00, w01, w10, w11);
    int Ival[20], Ix[20], Iy[20];
    float a11 = 0.f, a12 = 0.f, a22 = 0.f;
    const bool t_inside = ipx >= 1 && ipy >= 1 && ipx + WIN + 1 < A.w && ipy + WIN + 1 < A.h;
    const bool any_border = __any_sync(FULL, act && !t_inside);
    if (act) {
        if (any_border) template_pass<true>(A, ipx, ipy, xa, simd_flag, w00, w01, w10, w11, Ival, Ix, Iy, a11, a12, a22);
        else template_pass<false>(A, ipx, ipy, xa, simd_flag, w00, w01, w10, w11, Ival, Ix, Iy, a11, a12, a22);
    }
    __syncwarp();
    const float A11 = __fmul_rn(pentad_total(a11, base, role), FLT_SCALE);
    const float A12 = __fmul_rn(pentad_total(a12, base, role), FLT_SCALE);
    const float A22 = __fmul_rn(pentad_total(a22, base, role), FLT_SCALE);
    if (!act) return;
    uint32_t wv[32];
#pragma unroll
    for (int k = 0; k < 10; k++) wv[k] = (uint32_t)Ival[2 * k] | ((uint32_t)Ival[2 * k + 1] << 16);
#pragma unroll
    for (int i = 0; i < 20; i++) wv[10 + i] = ((uint32_t)Ix[i] & 0xffffu) | ((uint32_t)Iy[i] << 16);
    wv[30] = 0u; wv[31] = 0u;
    uint4* dst = words + (((size_t)level * cap + pi) * PENTAD + role) * 8;
#pragma unroll
    for (int q = 0; q < 8; q++) dst[q] = make_uint4(wv[4 * q], wv[4 * q + 1], wv[4 * q + 2], wv[4 * q + 3]);
    if (role == 0) {
        float* sp = sums + ((size_t)level * cap + pi) * 4;
        *reinterpret_cast<float4*>(sp) = make_float4(A11, A12, A22, 0.f);
    }
}

// CACHED: every pair of the batch brings its source templates (LKPair::tmpl); the template pass is then
// not even compiled in, which leaves the iteration loop the whole register budget.
template <bool CACHED>
__global__ void __launch_bounds__(LK_WARPS * 32, CACHED ? LK_CACHED_BLOCKS : 4) lk10_kernel(LKBatch batch, LKParams prm) {
    // large skips take several times more iterations (slow pairs are appended last): schedule
    // them first so the launch does not end on a tail of long blocks
    const LKPair& pr = batch.pair[gridDim.y - 1 - blockIdx.y];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / PENTAD, role = lane - grp * PENTAD;
    const int base = grp < PTS_PER_WARP ? grp * PENTAD : 0;       // first lane of this pentad (shuffle source)
    const int npts = min(*pr.n_pts, batch.cap);
    const int slot = (blockIdx.x * LK_WARPS + wib) * PTS_PER_WARP + grp;
    const bool valid = grp < PTS_PER_WARP && slot < npts;
    if (!__any_sync(FULL, valid)) return;
    const int pi = (valid && pr.order) ? __ldg(pr.order + slot) : slot;  // which keypoint this pentad works on

    const int simd_flag = role < 4 ? 1 : 0;
    const int xa = simd_flag ? role : 8;              // first column of the lane's chain (second: +4, tail +1)

    float ptx = 0.f, pty = 0.f;
    if (valid) { ptx = pr.pts[2 * pi]; pty = pr.pts[2 * pi + 1]; }
    const float halfw = (WIN - 1) * 0.5f;
    const int nlevels = min(min(pr.a.levels, pr.b.levels), prm.max_level + 1);
    const double eps2 = prm.eps * prm.eps;
    const float eps2_lo = (float)(eps2 * (1.0 - 1e-6)), eps2_hi = (float)(eps2 * (1.0 + 1e-6));
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    float nextx = 0.f, nexty = 0.f;
    int status = 1;
    float err = 0.f;
    int Ival[20], Ix[20], Iy[20];

    for (int level = nlevels - 1; level >= 0; level--) {
        const LevelRef A = {pr.a.data[level], pr.a.w[level], pr.a.h[level], pr.a.pitch[level]};
        const LevelRef B = {pr.b.data[level], pr.b.w[level], pr.b.h[level], pr.b.pitch[level]};
        const float scale = 1.f / (float)(1 << level);
        float prevx = __fmul_rn(ptx, scale), prevy = __fmul_rn(pty, scale);
        if (level == nlevels - 1) { nextx = prevx; nexty = prevy; }
        else { nextx = __fmul_rn(nextx, 2.f); nexty = __fmul_rn(nexty, 2.f); }
        prevx = __fsub_rn(prevx, halfw); prevy = __fsub_rn(prevy, halfw);
        const int ipx = __float2int_rd(prevx), ipy = __float2int_rd(prevy);
        bool act = valid;                            // this pentad works on this level
        if (act && (ipx < -WIN || ipx >= A.w || ipy < -WIN || ipy >= A.h)) {
            if (level == 0) { status = 0; err = 0.f; }
            act = false;
        }
        int w00, w01, w10, w11;
        bilinear_weights(__fsub_rn(prevx, (float)ipx), __fsub_rn(prevy, (float)ipy), w00, w01, w10, w11);

        // ---- template: loaded from the frame's cache, or computed here ---------------------------
        float A11, A12, A22;
        if (CACHED) {
            A11 = 0.f; A12 = 0.f; A22 = 0.f;
            if (act) {
                const size_t tslot = (size_t)level * pr.tmpl.cap + pi;
                const uint4* q = pr.tmpl.words + (tslot * PENTAD + role) * 8;
                uint32_t wv[32];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint4 v = __ldg(q + k);
                    wv[4 * k] = v.x; wv[4 * k + 1] = v.y; wv[4 * k + 2] = v.z; wv[4 * k + 3] = v.w;
                }
                const float4 sv = __ldg(reinterpret_cast<const float4*>(pr.tmpl.sums + tslot * 4));
                A11 = sv.x; A12 = sv.y; A22 = sv.z;
#pragma unroll
                for (int k = 0; k < 10; k++) { Ival[2 * k] = (int)(wv[k] & 0xffffu); Ival[2 * k + 1] = (int)(wv[k] >> 16); }
#pragma unroll
                for (int i = 0; i < 20; i++) { Ix[i] = (int)(short)(wv[10 + i] & 0xffffu); Iy[i] = (int)wv[10 + i] >> 16; }
            }
        } else {
            float a11 = 0.f, a12 = 0.f, a22 = 0.f;
            // the 13x13 source patch [ipx-1, ipx+11] x [ipy-1, ipy+11] lies inside the level
            const bool t_inside = ipx >= 1 && ipy >= 1 && ipx + WIN + 1 < A.w && ipy + WIN + 1 < A.h;
            const bool any_border = __any_sync(FULL, act && !t_inside);
            if (act) {
                if (any_border) template_pass<true>(A, ipx, ipy, xa, simd_flag, w00, w01, w10, w11, Ival, Ix, Iy, a11, a12, a22);
                else template_pass<false>(A, ipx, ipy, xa, simd_flag, w00, w01, w10, w11, Ival, Ix, Iy, a11, a12, a22);
            }
            __syncwarp();
            A11 = __fmul_rn(pentad_total(a11, base, role), FLT_SCALE);
            A12 = __fmul_rn(pentad_total(a12, base, role), FLT_SCALE);
            A22 = __fmul_rn(pentad_total(a22, base, role), FLT_SCALE);
        }
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float rad = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * WIN * WIN));
        if (act && ((double)minEig < prm.min_eig || D < 1.1920928955078125e-07f)) {
            if (level == 0) status = 0;
            act = false;
        }
        D = __fdiv_rn(1.f, D);
        float nx = __fsub_rn(nextx, halfw), ny = __fsub_rn(nexty, halfw);
        float pdx = 0.f, pdy = 0.f;

        // ---- iterations (the warp loops until its slowest pentad has converged) ----------------
        bool iterating = act && prm.iters > 0;
        for (int j = 0; j < prm.iters; j++) {
            if (!__any_sync(FULL, iterating)) break;
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (iterating && ((unsigned)(inx + WIN) >= (unsigned)(B.w + WIN) || (unsigned)(iny + WIN) >= (unsigned)(B.h + WIN))) {
                if (level == 0) status = 0;                  // inx < -WIN || inx >= B.w || iny < -WIN || iny >= B.h
                iterating = false;
            }
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            float bx = 0.f, by = 0.f;
            int unused = 0;
            if (iterating) window_pass<false>(B, inx, iny, xa, simd_flag, w00, w01, w10, w11, Ival, Ix, Iy, bx, by, unused);
            __syncwarp();
            const float b1 = __fmul_rn(pentad_total(bx, base, role), FLT_SCALE);
            const float b2 = __fmul_rn(pentad_total(by, base, role), FLT_SCALE);
            if (iterating) {
                const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
                const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
                nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
                nextx = __fadd_rn(nx, halfw); nexty = __fadd_rn(ny, halfw);
                // OpenCV tests dx*dx + dy*dy <= eps^2 in double; the float sum decides it except within
                // 1e-6 (relative) of the threshold, where the double expression is evaluated
                const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                bool small = d2 <= eps2_lo;
                if (!small && d2 < eps2_hi)
                    small = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps2;
                // fabs((double)v) < 0.01  <=>  fabsf(v) <= 0.01f: 0.01f is the largest float below 0.01
                if (small) {
                    iterating = false;
                } else if (j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
                    nextx = __fsub_rn(nextx, __fmul_rn(dx, 0.5f));
                    nexty = __fsub_rn(nexty, __fmul_rn(dy, 0.5f));
                    iterating = false;
                }
                pdx = dx; pdy = dy;
            }
        }
        // ---- error of the final position (level 0 only) -----------------------------------------
        if (level == 0) {
            bool want = act && status;
            const float fx = __fsub_rn(nextx, halfw), fy = __fsub_rn(nexty, halfw);
            const int inx = __float2int_rd(fx), iny = __float2int_rd(fy);
            if (want && (inx < -WIN || inx >= B.w || iny < -WIN || iny >= B.h)) {
                status = 0;
                want = false;
            }
            bilinear_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), w00, w01, w10, w11);
            int esum = 0;
            float f0, f1;
            if (want) window_pass<true>(B, inx, iny, xa, simd_flag, w00, w01, w10, w11, Ival, Ix, Iy, f0, f1, esum);
            __syncwarp();
            int tot = 0;
#pragma unroll
            for (int k = 0; k < PENTAD; k++) tot += __shfl_sync(FULL, esum, base + k);
            // every partial sum is an integer < 2^24: the float summation order of the CPU is exact
            if (want) err = __fdiv_rn(__fmul_rn((float)tot, 1.f), (float)(32 * WIN * WIN));
        }
    }
    if (valid && role == 0) {
        pr.next[2 * pi] = nextx;
        pr.next[2 * pi + 1] = nexty;
        pr.status[pi] = (uint8_t)status;
        pr.err[pi] = err;
    }
}

}  // namespace

void launch_lk10_templates(const PyramidView& a, const float* pts, const int* n_pts, int cap, const LKParams& p,
                           uint4* words, float* sums, cudaStream_t s, const int* order) {
    const int per_block = LK_WARPS * PTS_PER_WARP;
    const int nlev = std::min(a.levels, p.max_level + 1);
    dim3 grid((cap + per_block - 1) / per_block, nlev);
    lk10_template_kernel<<<grid, LK_WARPS * 32, 0, s>>>(a, pts, n_pts, cap, p.max_level + 1, words, sums, order);
}


// ---- spatial order of a frame's keypoints (counting sort by 64 x 64-pixel cell) ---------------------------------
namespace {
constexpr int kSpatialCellsMax = 8192;
__global__ void __launch_bounds__(1024) spatial_order_kernel(const float* __restrict__ pts, const int* __restrict__ n_pts, int cap,
                                                             int w, int h, int shift, int* __restrict__ order) {
    __shared__ int s_count[kSpatialCellsMax];
    __shared__ int s_warp[32];
    const int n = min(*n_pts, cap);
    const int ncx = ((w - 1) >> shift) + 1, ncy = ((h - 1) >> shift) + 1, ncell = ncx * ncy;
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) s_count[c] = 0;
    __syncthreads();
    auto cell_of_pt = [&](int i) {
        const int x = min(max(__float2int_rd(pts[2 * i]), 0), w - 1), y = min(max(__float2int_rd(pts[2 * i + 1]), 0), h - 1);
        return (y >> shift) * ncx + (x >> shift);
    };
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&s_count[cell_of_pt(i)], 1);
    __syncthreads();
    // exclusive scan of the cell counts, in place: every thread owns a contiguous run of cells
    const int per = (ncell + blockDim.x - 1) / blockDim.x;
    const int c0 = threadIdx.x * per, c1 = min(c0 + per, ncell);
    int mine = 0;
    for (int c = c0; c < c1; c++) mine += s_count[c];
    int incl = mine;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = s_warp[lane], t = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= d) t += u;
        }
        s_warp[lane] = t - v;                       // exclusive prefix of the warps
    }
    __syncthreads();
    int run = s_warp[wid] + incl - mine;
    for (int c = c0; c < c1; c++) {
        const int k = s_count[c];
        s_count[c] = run;                            // becomes the cell's fill cursor
        run += k;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) order[atomicAdd(&s_count[cell_of_pt(i)], 1)] = i;
}
}  // namespace

void launch_spatial_order(const float* pts, const int* n_pts, int cap, int w, int h, int* order, cudaStream_t s) {
    int shift = 6;                                    // 64-pixel cells, coarser while there would be too many of them
    while ((((w - 1) >> shift) + 1) * (((h - 1) >> shift) + 1) > kSpatialCellsMax) shift++;
    spatial_order_kernel<<<1, 1024, 0, s>>>(pts, n_pts, cap, w, h, shift, order);
}

void launch_lk10q(const LKBatch& batch, const LKParams& p, cudaStream_t s);   // lk10q.cu

void launch_lk10(const LKBatch& batch, const LKParams& p, cudaStream_t s) {
    if (batch.queue != nullptr && batch.num_pairs > 0) {
        launch_lk10q(batch, p, s);
        return;
    }
    const int per_block = LK_WARPS * PTS_PER_WARP;
    dim3 grid((batch.cap + per_block - 1) / per_block, batch.num_pairs);
    bool cached = batch.num_pairs > 0;
    for (int k = 0; k < batch.num_pairs; k++) cached = cached && batch.pair[k].tmpl.words != nullptr;
    if (cached) lk10_kernel<true><<<grid, LK_WARPS * 32, 0, s>>>(batch, p);
    else lk10_kernel<false><<<grid, LK_WARPS * 32, 0, s>>>(batch, p);
}

}  // namespace pc
