// Device-side helpers shared by the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pc {

// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
__host__ __device__ __forceinline__ int reflect101(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        p = p < 0 ? -p : 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// Streaming 16-byte load: read-only path, do not allocate in L1 (each byte is used once).
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Total order on floats as signed ints (negative floats compare correctly).
__host__ __device__ __forceinline__ int float_to_ordered_int(float f) {
    int i;
#ifdef __CUDA_ARCH__
    i = __float_as_int(f);
#else
    union { float f; int i; } u; u.f = f; i = u.i;
#endif
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ordered_int_to_float(int i) {
    i = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    union { float f; int i; } u; u.i = i; return u.f;
#endif
}
// Same order, unsigned (for radix keys): larger float -> larger key.
__device__ __forceinline__ uint32_t float_to_ordered_uint(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_uint_to_float(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

}  // namespace pc
