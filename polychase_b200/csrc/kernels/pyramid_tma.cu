// K1 + K2 fused, with TMA tile movement (sm_100a): RGB8 -> gray8 + pyramid levels 1..3 in two launches.
//
// Replaces cv::cvtColor(RGB2GRAY) (/root/reference/cpp/opticalflow.cc:259,298) and the image half of
// cv::buildOpticalFlowPyramid (/root/reference/cpp/opticalflow.cc:180-187); integer arithmetic, bit-exact:
//   gray = (R*9798 + G*19235 + B*3735 + 2^14) >> 15
//   down = ([1 4 6 4 1] x [1 4 6 4 1] + 128) >> 8 at even coordinates, BORDER_REFLECT_101
//
//   gray_l1_tma_kernel   one CTA per 64 x 16 tile of level 1: the 136 x 35 pixel RGB patch under it (14 KB)
//                        arrives by one 2-D TMA load (cp.async.bulk.tensor + mbarrier; out-of-image elements
//                        are zero filled), is converted to gray in shared memory (the level-0 pixels are never
//                        re-read from HBM), reduced 5x5, and the 128 x 32 level-0 block and the level-1 tile
//                        leave by TMA stores (clipped at the image edge by the tensor map).
//   l2_l3_tma_kernel     one CTA per 32 x 8 tile of level 3: a 144 x 43 patch of level 1 by TMA, the 68 x 20
//                        region of level 2 over it (its 64 x 16 centre is stored), then the level-3 tile.
// REFLECT_101 is applied when the patch is read in shared memory, by edge CTAs only.
// TMA boxes must start on a 16-byte boundary of the innermost dimension (an unaligned inner coordinate raises an
// illegal-instruction fault on sm_100: scripts/probes/tma_probe.cu), so both load boxes start a few bytes left of
// the patch (A_RGB_SKIP / B_P1_SKIP); rows are unconstrained.
// Algorithmic bytes per 4K frame: 24.9 MB RGB read + 11.0 MB pyramid written; the RGB halo re-reads (16 %) and
// the level-1 patch re-reads hit L2.
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace pc {

namespace {

// ---- TMA / mbarrier primitives (PTX ISA 8.x, sm_90+) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int x, int y, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(smem_u32(smem_src)), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// generic-proxy writes to shared memory must be made visible to the async proxy before a TMA store reads them
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ uint32_t gray_of(uint32_t r, uint32_t g, uint32_t b) {
    return (r * 9798u + g * 19235u + b * 3735u + (1u << 14)) >> 15;
}

// ---- kernel A: RGB -> level 0 + level 1 ----------------------------------------------------------------
constexpr int A_TW = 64, A_TH = 16;              // level-1 tile
constexpr int A_IW = 2 * A_TW + 8;               // 136 level-0 columns under it (origin 4 left of the block: word aligned)
constexpr int A_IH = 2 * A_TH + 3;               // 35 rows
constexpr int A_RGB_ROW = 416;                   // box row: 4 skipped bytes + 3 * 136 = 408, rounded to the 16-byte box granularity
constexpr int A_RGB_SKIP = 4;                    // the box starts at byte 384 bx - 16 (16-byte aligned), the patch at 384 bx - 12
constexpr int A_THREADS = 256;

struct SmemA {
    alignas(128) uint8_t rgb[A_IH][A_RGB_ROW];
    alignas(128) uint8_t out0[2 * A_TH][2 * A_TW];   // the level-0 block this CTA owns
    alignas(128) uint8_t out1[A_TH][A_TW];
    alignas(16) uint8_t gray[A_IH][A_IW];
    uint16_t hsum[A_IH][A_TW];
    alignas(8) uint64_t bar;
};

__global__ void __launch_bounds__(A_THREADS) gray_l1_tma_kernel(const __grid_constant__ CUtensorMap tm_rgb,
                                                                const __grid_constant__ CUtensorMap tm_l0,
                                                                const __grid_constant__ CUtensorMap tm_l1, int w, int h) {
    __shared__ SmemA S;
    const int tid = threadIdx.x;
    const int ax0 = 2 * A_TW * blockIdx.x - 4;   // level-0 x of patch column 0
    const int ay0 = 2 * A_TH * blockIdx.y - 2;
    if (tid == 0) mbar_init(&S.bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&S.bar, A_IH * A_RGB_ROW);
        tma_load_2d(&S.rgb[0][0], &tm_rgb, (3 * ax0 - A_RGB_SKIP) / 4, ay0, &S.bar);   // x in 4-byte elements: 96 bx - 4
    }
    mbar_wait(&S.bar, 0);
    // gray of the whole patch, four pixels (12 bytes = 3 words) per step.  The 15-bit coefficients are split into
    // bytes (9798 = 38 * 256 + 70, 19235 = 75 * 256 + 35, 3735 = 14 * 256 + 151) so that a pixel costs two DP4A on
    // its three bytes as they lie in memory (the fourth byte meets a zero coefficient: no masking, no byte
    // extraction) instead of three extractions and three multiplies.  Edge CTAs read the pixels that lie outside
    // the image through REFLECT_101; everything inside takes the same fast path.
    const bool edge = ax0 < 0 || ay0 < 0 || ax0 + A_IW > w || ay0 + A_IH > h;
    constexpr int GPR = A_IW / 4;                // 34 groups per row
    constexpr uint32_t CH = 38u | (75u << 8) | (14u << 16), CL = 70u | (35u << 8) | (151u << 16);
    // fixed mapping (no per-step division): thread -> (row of a 7-row pass, group); 238 of the 256 threads work
    constexpr int RPP = A_THREADS / GPR;         // 7 rows per pass, 5 passes
    const int g = tid % GPR, r0 = tid / GPR;
    const bool cols_in = ax0 + 4 * g >= 0 && ax0 + 4 * g + 3 < w;
    const bool own_col = g >= 1 && g <= 2 * A_TW / 4;
    for (int r = r0; r < A_IH && r0 < RPP; r += RPP) {
        uint32_t out;
        const bool inside = !edge || (cols_in && (unsigned)(ay0 + r) < (unsigned)h);
        if (inside) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(&S.rgb[r][A_RGB_SKIP + 12 * g]);
            const uint32_t a = p[0], b = p[1], c = p[2];
            const uint32_t x0 = a, x1 = __funnelshift_r(a, b, 24), x2 = __funnelshift_r(b, c, 16), x3 = c >> 8;
            const uint32_t p0 = (__dp4a(x0, CH, 0u) * 256u + __dp4a(x0, CL, 1u << 14)) >> 15;
            const uint32_t p1 = (__dp4a(x1, CH, 0u) * 256u + __dp4a(x1, CL, 1u << 14)) >> 15;
            const uint32_t p2 = (__dp4a(x2, CH, 0u) * 256u + __dp4a(x2, CL, 1u << 14)) >> 15;
            const uint32_t p3 = (__dp4a(x3, CH, 0u) * 256u + __dp4a(x3, CL, 1u << 14)) >> 15;
            out = __byte_perm(__byte_perm(p0, p1, 0x0040), __byte_perm(p2, p3, 0x0040), 0x5410);
        } else {
            // (rows / columns further out than any output needs may map outside the patch: clamped, never used)
            const int sr = clampi(reflect101(ay0 + r, h) - ay0, 0, A_IH - 1);
            out = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int sc = clampi(reflect101(ax0 + 4 * g + k, w) - ax0, 0, A_IW - 1);
                const uint8_t* q = &S.rgb[sr][A_RGB_SKIP + 3 * sc];
                out |= gray_of(q[0], q[1], q[2]) << (8 * k);
            }
        }
        *reinterpret_cast<uint32_t*>(&S.gray[r][4 * g]) = out;
        if (own_col && r >= 2 && r < 2 + 2 * A_TH) *reinterpret_cast<uint32_t*>(&S.out0[r - 2][4 * (g - 1)]) = out;
    }
    __syncthreads();
    // horizontal 1-4-6-4-1 pass, two outputs per step in 16-bit fields: outputs 2q' and 2q'+1 are centred on bytes
    // 0 and 2 of gray word q = q' + 1; E = (b0, b2), O = (b1, b3) of a word as packed 16-bit pairs
    for (int idx = tid; idx < A_IH * (A_TW / 2); idx += A_THREADS) {
        const int r = idx / (A_TW / 2), qp = idx - r * (A_TW / 2);
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(&S.gray[r][0]) + qp;
        const uint32_t wm = gw[0], w0 = gw[1], wp = gw[2];
        const uint32_t Em = wm & 0x00ff00ffu, Om = (wm >> 8) & 0x00ff00ffu;
        const uint32_t E0 = w0 & 0x00ff00ffu, O0 = (w0 >> 8) & 0x00ff00ffu;
        const uint32_t Ep = wp & 0x00ff00ffu;
        const uint32_t tm2 = __funnelshift_r(Em, E0, 16);        // (b2 of q-1, b0 of q)
        const uint32_t tm1 = __funnelshift_r(Om, O0, 16);        // (b3 of q-1, b1 of q)
        const uint32_t tp2 = __funnelshift_r(E0, Ep, 16);        // (b2 of q, b0 of q+1)
        const uint32_t sum = tm2 + tp2 + 4u * (tm1 + O0) + 6u * E0;   // <= 16 * 255 per field
        *reinterpret_cast<uint32_t*>(&S.hsum[r][2 * qp]) = sum;
    }
    __syncthreads();
    // vertical pass, the same packing (16 * 4080 + 128 < 65536): four outputs = two packed columns pairs per thread
    {
        const int i = tid / 16, j4 = (tid % 16) * 4, r = 2 * i;
        uint32_t o = 0;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int j = j4 + 2 * k;
            auto hs = [&](int rr) { return *reinterpret_cast<const uint32_t*>(&S.hsum[rr][j]); };
            const uint32_t v = hs(r) + hs(r + 4) + 4u * (hs(r + 1) + hs(r + 3)) + 6u * hs(r + 2) + 0x00800080u;
            o |= (((v >> 8) & 0xffu) | ((v >> 16) & 0xff00u)) << (16 * k);
        }
        *reinterpret_cast<uint32_t*>(&S.out1[i][j4]) = o;
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        tma_store_2d(&tm_l0, 2 * A_TW * blockIdx.x, 2 * A_TH * blockIdx.y, &S.out0[0][0]);
        tma_store_2d(&tm_l1, A_TW * blockIdx.x, A_TH * blockIdx.y, &S.out1[0][0]);
        tma_store_commit_and_wait();
    }
}

// ---- kernel B: level 1 -> level 2 (+ level 3) ------------------------------------------------------------
constexpr int B_T3W = 32, B_T3H = 8;             // level-3 tile
constexpr int B_R2W = 2 * B_T3W + 4;             // 68 x 20 region of level 2 (its 64 x 16 centre is owned)
constexpr int B_R2H = 2 * B_T3H + 4;
constexpr int B_P1W = 160;                       // box row: 10 skipped bytes + the 140-column level-1 patch (2 * 68 + 4), rounded to 16
constexpr int B_P1_SKIP = 10;                    // the box starts at x = 128 bx - 16, the patch at 128 bx - 6
constexpr int B_P1H = 2 * B_R2H + 3;             // 43 rows
constexpr int B_THREADS = 256;

struct SmemB {
    alignas(128) uint8_t l1[B_P1H][B_P1W];
    alignas(128) uint8_t out2[2 * B_T3H][2 * B_T3W];
    alignas(128) uint8_t out3[B_T3H][B_T3W];
    alignas(16) uint8_t l2[B_R2H][B_R2W];
    uint16_t h1[B_P1H][B_R2W];
    uint16_t h2[B_R2H][B_T3W];
    alignas(8) uint64_t bar;
};

__global__ void __launch_bounds__(B_THREADS) l2_l3_tma_kernel(const __grid_constant__ CUtensorMap tm_l1,
                                                              const __grid_constant__ CUtensorMap tm_l2,
                                                              const __grid_constant__ CUtensorMap tm_l3, int w1, int h1,
                                                              int w2, int h2, int with_l3) {
    __shared__ SmemB S;
    const int tid = threadIdx.x;
    const int x2_0 = 2 * B_T3W * blockIdx.x - 2, y2_0 = 2 * B_T3H * blockIdx.y - 2;   // level-2 origin of the region
    const int x1_0 = 2 * x2_0 - 2, y1_0 = 2 * y2_0 - 2;                               // level-1 origin of the patch
    if (tid == 0) mbar_init(&S.bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&S.bar, B_P1H * B_P1W);
        tma_load_2d(&S.l1[0][0], &tm_l1, x1_0 - B_P1_SKIP, y1_0, &S.bar);
    }
    mbar_wait(&S.bar, 0);
    const bool edge = x1_0 < 0 || y1_0 < 0 || x1_0 + 2 * B_R2W + 4 > w1 || y1_0 + B_P1H > h1;
    // horizontal pass of level 1 -> level 2: region column c is level-2 x = x2_0 + c (REFLECT_101 in level 2)
    for (int idx = tid; idx < B_P1H * B_R2W; idx += B_THREADS) {
        const int r = idx / B_R2W, c = idx - r * B_R2W;
        uint32_t s;
        if (!edge) {
            const uint8_t* t = &S.l1[r][B_P1_SKIP + 2 * c];
            s = t[0] + 4 * t[1] + 6 * t[2] + 4 * t[3] + t[4];
        } else {
            // patch row r stands for level-1 row y1_0 + r of the *mapped* level-2 row; rows are mapped in the vertical pass
            const int x2 = reflect101(x2_0 + c, w2);
            const uint8_t* row = S.l1[r] + B_P1_SKIP;
            auto col = [&](int x1) { return clampi(reflect101(x1, w1) - x1_0, 0, B_P1W - B_P1_SKIP - 1); };
            s = row[col(2 * x2 - 2)] + 4 * row[col(2 * x2 - 1)] + 6 * row[col(2 * x2)] + 4 * row[col(2 * x2 + 1)] +
                row[col(2 * x2 + 2)];
        }
        S.h1[r][c] = (uint16_t)s;
    }
    __syncthreads();
    for (int idx = tid; idx < B_R2H * B_R2W; idx += B_THREADS) {
        const int r = idx / B_R2W, c = idx - r * B_R2W;
        uint32_t v;
        if (!edge) {
            v = S.h1[2 * r][c] + 4 * S.h1[2 * r + 1][c] + 6 * S.h1[2 * r + 2][c] + 4 * S.h1[2 * r + 3][c] + S.h1[2 * r + 4][c];
        } else {
            const int y2 = reflect101(y2_0 + r, h2);
            auto rw = [&](int y1) { return clampi(reflect101(y1, h1) - y1_0, 0, B_P1H - 1); };
            v = S.h1[rw(2 * y2 - 2)][c] + 4 * S.h1[rw(2 * y2 - 1)][c] + 6 * S.h1[rw(2 * y2)][c] + 4 * S.h1[rw(2 * y2 + 1)][c] +
                S.h1[rw(2 * y2 + 2)][c];
        }
        const uint8_t px = (uint8_t)((v + 128) >> 8);
        S.l2[r][c] = px;
        if (r >= 2 && r < 2 + 2 * B_T3H && c >= 2 && c < 2 + 2 * B_T3W) S.out2[r - 2][c - 2] = px;
    }
    __syncthreads();
    if (with_l3) {
        for (int idx = tid; idx < B_R2H * B_T3W; idx += B_THREADS) {
            const int r = idx / B_T3W, j = idx - r * B_T3W;
            const uint8_t* t = &S.l2[r][2 * j];
            S.h2[r][j] = (uint16_t)(t[0] + 4 * t[1] + 6 * t[2] + 4 * t[3] + t[4]);
        }
        __syncthreads();
        if (tid < B_T3H * B_T3W) {
            const int i = tid / B_T3W, j = tid - i * B_T3W, r = 2 * i;
            S.out3[i][j] = (uint8_t)((S.h2[r][j] + 4 * S.h2[r + 1][j] + 6 * S.h2[r + 2][j] + 4 * S.h2[r + 3][j] +
                                      S.h2[r + 4][j] + 128) >> 8);
        }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        tma_store_2d(&tm_l2, 2 * B_T3W * blockIdx.x, 2 * B_T3H * blockIdx.y, &S.out2[0][0]);
        if (with_l3) tma_store_2d(&tm_l3, B_T3W * blockIdx.x, B_T3H * blockIdx.y, &S.out3[0][0]);
        tma_store_commit_and_wait();
    }
}

// ---- tensor maps -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 2-D row-major tensor of `elem_bytes`-byte elements: dims (w_elems, h), row pitch in bytes, box (bw, bh)
bool make_tmap(CUtensorMap* tm, const void* base, int elem_bytes, uint64_t w_elems, uint64_t h, uint64_t pitch_bytes,
               uint32_t bw, uint32_t bh) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {w_elems, h};
    const cuuint64_t strides[1] = {pitch_bytes};
    const cuuint32_t box[2] = {bw, bh};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    return fn(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Builds gray + levels 1..min(levels-1, 3) of the pyramid from an RGB8 device image.  Returns the number of levels
// written (0: the fast path does not apply -- alignment, width, or no driver entry point -- and nothing was launched).
int launch_pyramid_tma(const uint8_t* rgb, size_t stride, const Image8* lv, int levels, cudaStream_t s, int* launches) {
    *launches = 0;
    const int w = lv[0].w, h = lv[0].h;
    if (levels < 2 || (reinterpret_cast<uintptr_t>(rgb) & 15) != 0 || stride % 16 != 0 || w % 4 != 0 || w < A_IW || h < A_IH)
        return 0;
    for (int L = 0; L < levels && L < 4; L++)
        if ((reinterpret_cast<uintptr_t>(lv[L].data) & 15) != 0 || lv[L].pitch % 16 != 0) return 0;
    CUtensorMap tm_rgb, tm0, tm1, tm2, tm3;
    if (!make_tmap(&tm_rgb, rgb, 4, (uint64_t)w * 3 / 4, (uint64_t)h, stride, A_RGB_ROW / 4, A_IH)) return 0;
    if (!make_tmap(&tm0, lv[0].data, 1, (uint64_t)w, (uint64_t)h, (uint64_t)lv[0].pitch, 2 * A_TW, 2 * A_TH)) return 0;
    // level 1 is stored by kernel A (box 64 x 16) and loaded by kernel B (box 144 x 43): two maps over the same plane
    CUtensorMap tm1_load;
    if (!make_tmap(&tm1, lv[1].data, 1, (uint64_t)lv[1].w, (uint64_t)lv[1].h, (uint64_t)lv[1].pitch, A_TW, A_TH)) return 0;
    const bool two_more = levels >= 3 && lv[1].w >= B_P1W && lv[1].h >= B_P1H;
    if (two_more) {
        if (!make_tmap(&tm1_load, lv[1].data, 1, (uint64_t)lv[1].w, (uint64_t)lv[1].h, (uint64_t)lv[1].pitch, B_P1W, B_P1H)) return 0;
        if (!make_tmap(&tm2, lv[2].data, 1, (uint64_t)lv[2].w, (uint64_t)lv[2].h, (uint64_t)lv[2].pitch, 2 * B_T3W, 2 * B_T3H)) return 0;
        const Image8& l3 = levels >= 4 ? lv[3] : lv[2];
        if (!make_tmap(&tm3, l3.data, 1, (uint64_t)l3.w, (uint64_t)l3.h, (uint64_t)l3.pitch, B_T3W, B_T3H)) return 0;
    }
    dim3 ga((lv[1].w + A_TW - 1) / A_TW, (lv[1].h + A_TH - 1) / A_TH);
    gray_l1_tma_kernel<<<ga, A_THREADS, 0, s>>>(tm_rgb, tm0, tm1, w, h);
    *launches = 1;
    if (!two_more) return 2;
    const int with_l3 = levels >= 4 ? 1 : 0;
    // the grid covers level 2 in 64 x 16 blocks (a level-3 tile each when level 3 exists)
    dim3 gb((lv[2].w + 2 * B_T3W - 1) / (2 * B_T3W), (lv[2].h + 2 * B_T3H - 1) / (2 * B_T3H));
    l2_l3_tma_kernel<<<gb, B_THREADS, 0, s>>>(tm1_load, tm2, tm3, lv[1].w, lv[1].h, lv[2].w, lv[2].h, with_l3);
    *launches = 2;
    return with_l3 ? 4 : 3;
}

}  // namespace pc
