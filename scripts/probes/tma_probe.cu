// Which TMA box shapes / coordinates does this device accept?  (diagnostic for pyramid_tma.cu)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe scripts/probes/tma_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, int x, int y, int bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         smem_u32(buf)),
                     "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(x), "r"(y)
                     : "memory");
    }
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
            smem_u32(&bar)),
        "r"(0)
        : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int W = 1920, H = 480;                 // bytes per row, rows
    uint8_t* d;
    cudaMalloc(&d, (size_t)W * H);
    std::vector<uint8_t> h((size_t)W * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 11));
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    uint8_t* out;
    cudaMalloc(&out, 65536);
    struct Case { const char* name; int elem; int bw, bh, x, y; };
    const Case cases[] = {
        // result on B200 (driver 580): every case whose inner start is a multiple of 16 bytes runs and matches (rows,
        // negative / past-the-end coordinates are free: zero fill); an inner start off a 16-byte boundary faults with
        // cudaErrorIllegalInstruction -- keep such a case LAST, the error is sticky
        {"u8 128x32 aligned", 1, 128, 32, 128, 32},   {"u8 160x43 x=-16 y=-6", 1, 160, 43, -16, -6},
        {"u8 160x43 x=1904 y=470", 1, 160, 43, 1904, 470}, {"u32 64x32 aligned", 4, 64, 32, 64, 32},
        {"u32 104x35 x=92 y=30", 4, 104, 35, 92, 30}, {"u32 104x35 x=-4 y=-2", 4, 104, 35, -4, -2},
        {"u32 104x35 x=476 y=460", 4, 104, 35, 476, 460}, {"u8 256x35 x=368", 1, 256, 35, 368, 30},
        {"u8 128x32 x=5 (unaligned)", 1, 128, 32, 5, 3},
    };
    for (const Case& c : cases) {
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)(W / c.elem), (cuuint64_t)H};
        const cuuint64_t strides[1] = {(cuuint64_t)W};
        const cuuint32_t box[2] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, c.elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box,
                         es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int bytes = c.bw * c.elem * c.bh;
        cudaError_t e = cudaSuccess;
        int bad = -1;
        if (r == CUDA_SUCCESS) {
            probe_kernel<<<1, 128, bytes>>>(tm, c.x, c.y, bytes, out);
            e = cudaDeviceSynchronize();
            if (e == cudaSuccess) {
                std::vector<uint8_t> o(bytes);
                cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
                bad = 0;
                for (int rr = 0; rr < c.bh; rr++)
                    for (int cc = 0; cc < c.bw * c.elem; cc++) {
                        const int gy = c.y + rr, gx = c.x * c.elem + cc;
                        const uint8_t want = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? h[(size_t)gy * W + gx] : 0;
                        if (o[rr * c.bw * c.elem + cc] != want) bad++;
                    }
            }
        }
        printf("%-32s encode=%d run=%s mismatches=%d\n", c.name, (int)r, cudaGetErrorName(e), bad);
        if (e != cudaSuccess) { printf("  (sticky error: stopping)\n"); return 0; }
    }
    return 0;
}
