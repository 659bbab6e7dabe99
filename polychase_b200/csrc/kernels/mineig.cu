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
//
// Layout (HBM-streaming, no shared memory): one warp owns a (30*NC)-column x 32-row output tile
// and marches down it.  Each lane holds NC adjacent columns (one aligned gray load per row, one
// aligned eig vector store per row); the +-1 column neighbours come from the adjacent lanes by
// shuffle, the +-1 row neighbours are rolling registers.  Lanes 0 and 31 only feed their
// neighbours (the halo columns of the tile).  NC = 4 with the march loop unrolled by 3 (the rolling
// windows rotate through registers instead of being moved) measures 63 us at 4K against 76 us for
// NC = 2 rolled: the halo shuffles and conversions are shared by twice the columns, which outweighs
// the drop to 16 resident warps per SM.
//
// Rows outside the image: the march simply continues over the REFLECT_101 row indices.  For
// the one row each side that the box filter needs (cov row -1 := cov row 1, cov row h := cov
// row h-2) the mirrored rolling window yields the same dx and the negated dy bit for bit, so
// cov_xx / cov_yy are already right and cov_xy only needs its sign flipped back.  Columns
// outside the image take the reflected cov value of the neighbouring lane/column explicitly.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace pc {

constexpr int ME_NC = 4;                       // columns per lane
constexpr int ME_COLS = 30 * ME_NC;            // output columns per warp tile (lanes 1..30)
constexpr int ME_ROWS = 32;                    // output rows per warp tile
constexpr int ME_WARPS = 4;

__device__ __forceinline__ int cell_of(int x, int y, const DetectGrid& g) {
    return (y / g.block_h) * g.grid_cols + (x / g.block_w);
}

// ME_NC gray bytes of this lane's columns (packed little-endian), REFLECT_101 in both directions
__device__ __forceinline__ uint32_t load_gray(const uint8_t* __restrict__ gray, int w, int h, int pitch, int x,
                                              int gy_logical, bool fast) {
    const uint8_t* row = gray + (size_t)reflect101(gy_logical, h) * pitch;
    if (fast) {
        if (ME_NC == 4) return __ldg(reinterpret_cast<const uint32_t*>(row + x));
        return __ldg(reinterpret_cast<const uint16_t*>(row + x));
    }
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < ME_NC; j++) v |= (uint32_t)row[reflect101(x + j, w)] << (8 * j);
    return v;
}

__global__ void __launch_bounds__(ME_WARPS * 32, ME_NC == 2 ? 6 : 4)
min_eig_kernel(const uint8_t* __restrict__ gray, int w, int h, int pitch, float* __restrict__ eig, int eig_pitch,
               DetectGrid grid, int* __restrict__ cell_max, int tiles_x, int tiles_y) {
    constexpr int NC = ME_NC;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * ME_WARPS + (threadIdx.x >> 5);
    if (tile >= tiles_x * tiles_y) return;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int xb = tx * ME_COLS - NC + NC * lane;          // this lane's columns xb .. xb+NC-1 (xb % NC == 0)
    const int y0 = ty * ME_ROWS;
    const int y_end = min(y0 + ME_ROWS, h);
    const float s = (float)(1.0 / (4.0 * 3.0 * 255.0));
    const float s2 = 2.0f * s;
    const int wvec = (w / 32) * 32;
    const bool fast = xb >= 0 && xb + NC - 1 < w;
    const bool live = xb + NC >= 0 && xb <= w;             // lanes whose columns can matter at all
    const bool out_lane = lane >= 1 && lane <= 30 && xb < w;
    const int jr = (w - 1) - xb;                            // position of the last image column in this lane
    // warp-uniform special cases (tile_x0 .. tile_x0 + 32*NC - 1 are the columns the warp touches)
    const int tile_x0 = tx * ME_COLS - NC;
    const bool tail_tile = tile_x0 + 32 * NC > wvec;        // some column uses the non-FMA row filter
    const bool left_tile = tile_x0 < 0, right_tile = tile_x0 + 32 * NC >= w;

    // grid cell bookkeeping: the common case is one cell column per warp tile
    const int cx_first = min(max(xb, 0), w - 1) / grid.block_w, cx_last = min(max(xb + NC - 1, 0), w - 1) / grid.block_w;
    const int cx_ref = __shfl_sync(FULL, cx_first, 1);
    const bool my_uniform = !out_lane || (cx_first == cx_ref && cx_last == cx_ref);
    const bool uniform_x = __all_sync(FULL, my_uniform);
    // eig >= -tiny and never NaN, so a float max is enough for the running cell maximum
    float run_max = -INFINITY;
    int run_cy = y0 / grid.block_h;
    int next_cy_y = (run_cy + 1) * grid.block_h;            // first row of the next cell row

    float rx[3][NC], rsm[3][NC];               // rolling horizontal passes (rows k-2, k-1, k)
    double Rp[3][NC], T[3][NC];                // rowsum of the previous cov row, and prev-prev + prev
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int j = 0; j < NC; j++) { Rp[c][j] = 0.0; T[c][j] = 0.0; }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int j = 0; j < NC; j++) { rx[r][j] = 0.f; rsm[r][j] = 0.f; }

    uint32_t word = live ? load_gray(gray, w, h, pitch, xb, y0 - 2, fast) : 0u;
    const int steps = (y_end - y0) + 4;
#pragma unroll 3
    for (int k = 0; k < steps; k++) {
        // prefetch the next gray row while this one is processed
        uint32_t next_word = 0u;
        if (live && k + 1 < steps) next_word = load_gray(gray, w, h, pitch, xb, y0 - 1 + k, fast);

        // ---- horizontal pass on gray row (logical) y0 - 2 + k ---------------------------------
        const uint32_t wl = __shfl_up_sync(FULL, word, 1), wr = __shfl_down_sync(FULL, word, 1);
        float p[NC + 2];
        p[0] = (float)((wl >> (8 * (NC - 1))) & 255u);
#pragma unroll
        for (int j = 0; j < NC; j++) p[j + 1] = (float)((word >> (8 * j)) & 255u);
        p[NC + 1] = (float)(wr & 255u);
#pragma unroll
        for (int j = 0; j < NC; j++) {
            rx[0][j] = rx[1][j]; rx[1][j] = rx[2][j];
            rsm[0][j] = rsm[1][j]; rsm[1][j] = rsm[2][j];
            const float pm = p[j], pc_ = p[j + 1], pp = p[j + 2];
            rx[2][j] = __fsub_rn(pp, pm);
            rsm[2][j] = __fmaf_rn(pp, s, __fmaf_rn(pc_, s2, __fmul_rn(pm, s)));
        }
        if (tail_tile) {                                    // warp-uniform and rare: one branch, not 4 x 5 predicated slots
#pragma unroll
            for (int j = 0; j < NC; j++)
                if (xb + j >= wvec)
                    rsm[2][j] = __fadd_rn(__fadd_rn(__fmul_rn(p[j], s), __fmul_rn(p[j + 1], s2)), __fmul_rn(p[j + 2], s));
        }
        word = next_word;
        if (k < 2) continue;

        // ---- cov row (logical) cy = y0 - 3 + k ------------------------------------------------
        const int cy = y0 - 3 + k;
        const bool flip = cy < 0 || cy >= h;                // mirrored window: dy came out negated
        float c[3][NC];
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const float dx = __fmaf_rn(__fadd_rn(rx[0][j], rx[2][j]), s, __fmul_rn(rx[1][j], s2));
            const float dy = __fsub_rn(rsm[2][j], rsm[0][j]);
            c[0][j] = __fmul_rn(dx, dx);
            const float xy = __fmul_rn(dx, dy);
            c[1][j] = flip ? -xy : xy;
            c[2][j] = __fmul_rn(dy, dy);
        }
        // ---- 3-tap horizontal sums in double (REFLECT_101 of cov at the image's side borders), consumed at once by
        // the 3-tap vertical sums: a row sum R only lives while its column is handled (fewer live doubles, fewer spills)
        float sum3[3][NC];                                  // Sxx, Sxy, Syy of output row y = cy - 1, rounded once
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            double d[NC + 2];
            d[0] = (double)__shfl_up_sync(FULL, c[ch][NC - 1], 1);
            d[NC + 1] = (double)__shfl_down_sync(FULL, c[ch][0], 1);
#pragma unroll
            for (int j = 0; j < NC; j++) d[j + 1] = (double)c[ch][j];
            if (left_tile && xb == 0) d[0] = d[2];          // cov[-1] := cov[1]
            if (right_tile && (unsigned)jr < (unsigned)NC) {   // cov[w] := cov[w-2]
#pragma unroll
                for (int j = 0; j < NC; j++)
                    if (j == jr) d[j + 2] = d[j];
            }
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const double R = __dadd_rn(__dadd_rn(d[j], d[j + 1]), d[j + 2]);
                sum3[ch][j] = (float)__dadd_rn(T[ch][j], R);
                T[ch][j] = __dadd_rn(Rp[ch][j], R);
                Rp[ch][j] = R;
            }
        }
        // ---- eigenvalue, store (output row y = cy - 1) -----------------------------------------------
        const int y = cy - 1;
        if (k >= 4) {
            float v[NC];
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const float a = __fmul_rn(sum3[0][j], 0.5f), b = sum3[1][j], cc = __fmul_rn(sum3[2][j], 0.5f);
                const float t = __fsub_rn(a, cc);
                v[j] = __fsub_rn(__fadd_rn(a, cc), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b))));
            }
            if (y >= next_cy_y) {                           // warp-uniform: the tile crossed a cell row
                if (uniform_x) {
                    const int m = __reduce_max_sync(FULL, float_to_ordered_int(run_max));
                    if (lane == 0 && m != float_to_ordered_int(-INFINITY))
                        atomicMax(&cell_max[run_cy * grid.grid_cols + cx_ref], m);
                }
                run_max = -INFINITY;
                run_cy = y / grid.block_h;
                next_cy_y = (run_cy + 1) * grid.block_h;
            }
            if (out_lane) {
                float* dst = eig + (size_t)y * eig_pitch + xb;
                if (xb + NC - 1 < w) {
                    if (NC == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[NC - 2], v[NC - 1]);
                    else *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
                } else {
#pragma unroll
                    for (int j = 0; j < NC; j++)
                        if (xb + j < w) dst[j] = v[j];
                }
                if (uniform_x) {
#pragma unroll
                    for (int j = 0; j < NC; j++)
                        if (!right_tile || xb + j < w) run_max = fmaxf(run_max, v[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < NC; j++)
                        if (xb + j < w)                    // (the tile straddles a cell border: rare, the cell is found per pixel)
                            atomicMax(&cell_max[run_cy * grid.grid_cols + (xb + j) / grid.block_w], float_to_ordered_int(v[j]));
                }
            }
        }
    }
    if (uniform_x) {
        const int m = __reduce_max_sync(FULL, float_to_ordered_int(run_max));
        if (lane == 0 && m != float_to_ordered_int(-INFINITY)) atomicMax(&cell_max[run_cy * grid.grid_cols + cx_ref], m);
    }
}

__global__ void init_cell_max_kernel(int* cell_max, int n, int* counters, int n_counters, int* frame_counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cell_max[i] = 0x80000000;
    if (i < n_counters) counters[i] = 0;
    if (frame_counters && i < 3) frame_counters[i] = 0;
}

void launch_min_eig(Image8 gray, float* eig, int eig_pitch, DetectGrid g, int* cell_max, int* zero_block,
                    int zero_ints, int* frame_counters, cudaStream_t s) {
    const int ncell = g.grid_rows * g.grid_cols;
    const int n_init = std::max(std::max(ncell, zero_block ? zero_ints : 0), 3);
    init_cell_max_kernel<<<(n_init + 255) / 256, 256, 0, s>>>(cell_max, ncell, zero_block, zero_block ? zero_ints : 0,
                                                              frame_counters);
    const int tiles_x = (gray.w + ME_COLS - 1) / ME_COLS, tiles_y = (gray.h + ME_ROWS - 1) / ME_ROWS;
    const int blocks = (tiles_x * tiles_y + ME_WARPS - 1) / ME_WARPS;
    min_eig_kernel<<<blocks, ME_WARPS * 32, 0, s>>>(gray.data, gray.w, gray.h, gray.pitch, eig, eig_pitch, g, cell_max,
                                                    tiles_x, tiles_y);
}

// ---- K5: threshold (per cell), 3x3 NMS, candidate list + state map ---------------------
// cv::threshold(THRESH_TOZERO, maxVal*quality): thresh is the double product rounded to
// float; value kept iff value > thresh (gftt.cc:61-65).  A pixel is a candidate iff it is
// interior, its thresholded value is non-zero and equals the 3x3 max of the thresholded map
// (gftt.cc:70-86).
//
// One warp owns a 128-column strip of NMS_ROWS rows and marches down it: per row one 16-byte
// eig load per lane (4 columns), thresholded on load with the threshold of the pixel's own
// cell (table in shared memory), +-1 column from the neighbour lanes (strip-edge lanes load
// the one extra column), rows in rolling registers.  Writes the u8 state map (4 bytes per lane
// per row) and appends candidates (warp-aggregated).
constexpr int NMS_ROWS = 16;
constexpr int NMS_WARPS = 4;
constexpr int NMS_MAX_CELLS = 1024;    // grid_rows * grid_cols limit (validated in csrc/abi/capi.cu)

__global__ void __launch_bounds__(NMS_WARPS * 32) nms_candidates_kernel(
    const float* __restrict__ eig, int eig_pitch, int w, int h, DetectGrid grid, const int* __restrict__ cell_max,
    double quality, uint8_t* __restrict__ state, int state_pitch, unsigned long long* __restrict__ cand, int cand_cap,
    int* __restrict__ cand_count, int* __restrict__ value_hist, int tiles_x, int tiles_y) {
    const unsigned FULL = 0xffffffffu;
    __shared__ float thr_tab[NMS_MAX_CELLS];
    __shared__ int s_hist[4096];                              // block-private copy of value_hist
    const int ncell = grid.grid_rows * grid.grid_cols;
    for (int i = threadIdx.x; i < ncell; i += blockDim.x)
        thr_tab[i] = (float)((double)ordered_int_to_float(__ldg(&cell_max[i])) * quality);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // persistent blocks: the 16 KB histogram is zeroed and flushed once per block, not once per 4 tiles
    for (int tile = blockIdx.x * NMS_WARPS + wib; tile < tiles_x * tiles_y; tile += gridDim.x * NMS_WARPS) {
    const bool tile_ok = true;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int xb = tx * 128 + 4 * lane;
    const int y0 = ty * NMS_ROWS, y_end = tile_ok ? min(y0 + NMS_ROWS, h) : y0;
    const bool live = tile_ok && xb < w;
    unsigned long long cbits = 0;                             // this lane's candidates of the tile: bit 4 * (y - y0) + j
    int cxj[6];                                   // cell column of x = xb-1 .. xb+4
#pragma unroll
    for (int j = 0; j < 6; j++) cxj[j] = min(max(xb - 1 + j, 0), w - 1) / grid.block_w;

    // raw values of one row: this lane's 4 columns + (strip-edge lanes only) the column beside them
    struct RawRow { float4 v; float edge; };
    auto fetch = [&](int y) {
        RawRow r;
        r.v = make_float4(0.f, 0.f, 0.f, 0.f);
        r.edge = 0.f;
        if ((unsigned)y >= (unsigned)h || !live) return r;
        const float* row = eig + (size_t)y * eig_pitch;
        if (xb + 3 < w) r.v = *reinterpret_cast<const float4*>(row + xb);
        else {
            r.v.x = row[xb];
            if (xb + 1 < w) r.v.y = row[xb + 1];
            if (xb + 2 < w) r.v.z = row[xb + 2];
        }
        if (lane == 0 && xb >= 1) r.edge = row[xb - 1];
        if (lane == 31 && xb + 4 < w) r.edge = row[xb + 4];
        return r;
    };
    // thresholded values for columns xb-1 .. xb+4 (0 outside the image: never a max, and border
    // pixels are never candidates)
    // the common case: the whole tile (with its one-pixel ring) lies in one grid cell, so one
    // threshold serves every pixel and the per-row cell arithmetic disappears
    const int tx_lo = max(tx * 128 - 1, 0), tx_hi = min(tx * 128 + 128, w - 1);
    const int ty_lo = max(y0 - 1, 0), ty_hi = min(y_end, h - 1);
    const bool one_cell = tx_lo / grid.block_w == tx_hi / grid.block_w && ty_lo / grid.block_h == ty_hi / grid.block_h;
    const float thr_one = thr_tab[(ty_lo / grid.block_h) * grid.grid_cols + tx_lo / grid.block_w];
    auto finish = [&](int y, const RawRow& r, float (&t)[6]) {
#pragma unroll
        for (int j = 0; j < 6; j++) t[j] = 0.f;
        if ((unsigned)y >= (unsigned)h) return;   // warp-uniform
        const float raw[4] = {r.v.x, r.v.y, r.v.z, r.v.w};
        float th[6];
        if (one_cell) {
#pragma unroll
            for (int j = 0; j < 6; j++) th[j] = thr_one;
        } else {
            const int crow = (y / grid.block_h) * grid.grid_cols;
#pragma unroll
            for (int j = 0; j < 6; j++) th[j] = thr_tab[crow + cxj[j]];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) t[j + 1] = (live && xb + j < w && raw[j] > th[j + 1]) ? raw[j] : 0.f;
        float left = __shfl_up_sync(FULL, t[4], 1), right = __shfl_down_sync(FULL, t[1], 1);
        if (lane == 0) left = (live && xb >= 1 && r.edge > th[0]) ? r.edge : 0.f;
        if (lane == 31) right = (xb + 4 < w && r.edge > th[5]) ? r.edge : 0.f;
        t[0] = left;
        t[5] = right;
    };

    float a[6], b[6], c[6];                       // rows y-1, y, y+1
    finish(y0 - 1, fetch(y0 - 1), a);
    finish(y0, fetch(y0), b);
    RawRow ahead = fetch(y0 + 1);
    for (int y = y0; y < y_end; y++) {
        const RawRow cur = ahead;
        ahead = fetch(y + 2);                     // in flight while this row is processed
        finish(y + 1, cur, c);
        uint32_t bits = 0;
        bool is_c[4];
        float colmax[6];                           // max over the three rows, per column (centre included:
#pragma unroll                                     //  v >= max(all nine) <=> v >= max(the other eight))
        for (int j = 0; j < 6; j++) colmax[j] = fmaxf(fmaxf(a[j], b[j]), c[j]);
        const bool row_in = y >= 1 && y < h - 1;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float vmid = b[j + 1];
            const float m = fmaxf(fmaxf(colmax[j], colmax[j + 1]), colmax[j + 2]);
            const int x = xb + j;
            is_c[j] = vmid != 0.f && vmid >= m && x >= 1 && x < w - 1 && row_in;
            bits |= is_c[j] ? (1u << (8 * j)) : 0u;
        }
        if (live) {
            uint8_t* sp = state + (size_t)y * state_pitch + xb;
            if (xb + 3 < w) *reinterpret_cast<uint32_t*>(sp) = bits;
            else
                for (int j = 0; j < 4 && xb + j < w; j++) sp[j] = (bits >> (8 * j)) & 1u;
        }
        // the row's candidates are only remembered here (4 bits per lane and row); they are appended once per tile
        cbits |= (unsigned long long)((bits & 1u) | ((bits >> 7) & 2u) | ((bits >> 14) & 4u) | ((bits >> 21) & 8u)) << (4 * (y - y0));
#pragma unroll
        for (int j = 0; j < 6; j++) { a[j] = b[j]; b[j] = c[j]; }
    }
    // one append per warp tile (order inside the list is irrelevant: keys are sorted later): a warp scan of the
    // lanes' counts, one atomic for the warp, then every lane writes its own keys -- the value of a candidate is its
    // raw eigenvalue (it passed its threshold), re-read from the map (an L1 / L2 hit).  Appending row by row through a
    // shared-memory stage cost a quarter of the kernel's instructions (profiles/r2_w_ncu_all_kernels_cold.txt).
    static_assert(NMS_ROWS * 4 <= 64, "candidate bits of a tile fit one 64-bit word per lane");
    const int mine = __popcll(cbits);
    if (__any_sync(FULL, mine > 0)) {
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += n;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        int base = 0;
        if (lane == 0) base = atomicAdd(cand_count, total);
        base = __shfl_sync(FULL, base, 0);
        int slot = base + incl - mine;
        while (cbits) {
            const int bit = __ffsll((long long)cbits) - 1;
            cbits &= cbits - 1;
            const int y = y0 + (bit >> 2), x = xb + (bit & 3);
            const uint32_t ov = float_to_ordered_uint(__ldg(eig + (size_t)y * eig_pitch + x));
            if (slot < cand_cap) cand[slot] = ((unsigned long long)ov << 32) | (unsigned)(y * w + x);
            slot++;
            atomicAdd(&s_hist[ov >> 20], 1);
        }
        __syncwarp();
    }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
        const int v = s_hist[i];
        if (v) atomicAdd(&value_hist[i], v);
    }
}

void launch_nms_candidates(const float* eig, int eig_pitch, int w, int h, DetectGrid g, const int* cell_max,
                           double quality_level, uint8_t* state, int state_pitch, unsigned long long* cand,
                           int cand_cap, int* cand_count, int* value_hist, cudaStream_t s) {
    const int tiles_x = (w + 127) / 128, tiles_y = (h + NMS_ROWS - 1) / NMS_ROWS;
    int blocks = (tiles_x * tiles_y + NMS_WARPS - 1) / NMS_WARPS;
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    blocks = std::min(blocks, sm_count * 8);                  // persistent blocks (64 registers, 20 KB of shared memory each)
    nms_candidates_kernel<<<blocks, NMS_WARPS * 32, 0, s>>>(eig, eig_pitch, w, h, g, cell_max, quality_level, state,
                                                            state_pitch, cand, cand_cap, cand_count, value_hist,
                                                            tiles_x, tiles_y);
}

}  // namespace pc
