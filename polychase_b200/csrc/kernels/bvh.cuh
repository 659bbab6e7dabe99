// Stack-based nearest-hit BVH traversal (K10), shared by the track and refine kernels.
// Stands in for Embree's rtcIntersect1 as the reference uses it
// (/root/reference/cpp/ray_casting.cc:65-121): nearest hit, tnear = 0, tfar = inf.
#pragma once

#include "geom.cuh"
#include "track_kernels.h"

namespace pc {

struct HitRec {
    float t, u, v;
    int prim;      // -1 = miss
};

__device__ __forceinline__ HitRec bvh_nearest_hit(const BvhView& bvh, V3 o, V3 d) {
    HitRec best{INFINITY, 0.f, 0.f, -1};
    if (bvh.num_nodes == 0) return best;
    // Embree's rcp_safe: a zero direction component becomes +-1e-18 so that a ray lying exactly
    // in a box face gives 0 * big = 0 instead of 0 * inf = NaN (which would cull both neighbours).
    auto rcp_safe = [](float x) { return 1.f / (fabsf(x) < 1e-18f ? copysignf(1e-18f, x) : x); };
    const float idx = rcp_safe(d.x), idy = rcp_safe(d.y), idz = rcp_safe(d.z);
    int stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const int ni = stack[--sp];
        const float4 a = __ldg(&bvh.nodes[2 * ni]);       // bmin.xyz, left/first
        const float4 b = __ldg(&bvh.nodes[2 * ni + 1]);   // bmax.xyz, count
        float t0 = (a.x - o.x) * idx, t1 = (b.x - o.x) * idx;
        float tmin = fminf(t0, t1), tmax = fmaxf(t0, t1);
        t0 = (a.y - o.y) * idy; t1 = (b.y - o.y) * idy;
        tmin = fmaxf(tmin, fminf(t0, t1)); tmax = fminf(tmax, fmaxf(t0, t1));
        t0 = (a.z - o.z) * idz; t1 = (b.z - o.z) * idz;
        tmin = fmaxf(tmin, fminf(t0, t1)); tmax = fminf(tmax, fmaxf(t0, t1));
        // conservative slab test (NaNs from 0*inf fall through as "hit")
        if (tmax < fmaxf(tmin, 0.f) * 0.9999f - 1e-6f || tmin > best.t) continue;
        const int first = __float_as_int(a.w), count = __float_as_int(b.w);
        if (count > 0) {
            for (int k = 0; k < count; k++) {
                const float4* tp = bvh.tris + 3 * (first + k);
                const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
                const V3 p1 = v3(q0.x, q0.y, q0.z), e1 = v3(q1.x, q1.y, q1.z) - p1, e2 = v3(q2.x, q2.y, q2.z) - p1;
                const V3 pv = cross(d, e2);
                const float det = dot(e1, pv);
                if (det == 0.f) continue;
                const float inv = 1.f / det;
                const V3 s = o - p1;
                const float u = inv * dot(s, pv);
                if (u < 0.f || u > 1.f) continue;
                const V3 qv = cross(s, e1);
                const float v = inv * dot(d, qv);
                if (v < 0.f || u + v > 1.f) continue;
                const float t = inv * dot(e2, qv);
                if (t < 0.f || t >= best.t) continue;
                best.t = t; best.u = u; best.v = v; best.prim = __float_as_int(q0.w);
            }
        } else {
            if (sp + 2 <= 64) { stack[sp++] = first; stack[sp++] = first + 1; }
        }
    }
    return best;
}


}  // namespace pc
