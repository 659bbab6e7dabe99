// K12 + K13 + K14: bundle-adjustment residuals / Jacobians / normal equations, total cost
// with the stateful primitive-id cache, and the block-banded Cholesky solve.
//
//   Evaluate / EvaluateWithJacobian      /root/reference/cpp/refiner.cc:274-506
//   BuildNormalEquations / TotalCost     /root/reference/cpp/pnp/lev_marq.h:653-824
//   ComputeStep (SimplicialLLT)          /root/reference/cpp/pnp/lev_marq.h:826-841
//   ray/plane, ray/triangle              /root/reference/cpp/ray_casting.h:76-179
//
// Layout: keypoints of all frames concatenated; one CTA per edge (directed flow) walks its
// matches.  The ray of a residual depends only on (source frame, source keypoint), so the
// cached-triangle test + ray-cast fallback + cache update of Evaluate runs once per keypoint
// (ba_refresh_points) and the per-edge cost kernel only projects into the target camera.
// Per-edge 2p x 2p normal-equation blocks are reduced in registers/shared memory, written
// once, and assembled per frame in a fixed order (deterministic, no float atomics) into a
// block-banded matrix (half-bandwidth 8 frames) that one CTA factors with a blocked Cholesky.
#include "ba_kernels.h"
#include "bvh.cuh"
#include "common.cuh"

namespace pc {

namespace {

__device__ __forceinline__ V3 hnorm_mul(const float* M, V3 p) {   // (M * [p,1]).hnormalized()
    const float x = M[0] * p.x + M[1] * p.y + M[2] * p.z + M[3];
    const float y = M[4] * p.x + M[5] * p.y + M[6] * p.z + M[7];
    const float z = M[8] * p.x + M[9] * p.y + M[10] * p.z + M[11];
    const float w = M[12] * p.x + M[13] * p.y + M[14] * p.z + M[15];
    return V3{x / w, y / w, z / w};
}
__device__ __forceinline__ V3 mul3x3(const float* M, V3 p) {      // M.block<3,3>(0,0) * p
    return V3{M[0] * p.x + M[1] * p.y + M[2] * p.z, M[4] * p.x + M[5] * p.y + M[6] * p.z,
              M[8] * p.x + M[9] * p.y + M[10] * p.z};
}
__device__ __forceinline__ V3 mul3x3_t(const float* M, V3 p) {    // M.block<3,3>(0,0)^T * p
    return V3{M[0] * p.x + M[4] * p.y + M[8] * p.z, M[1] * p.x + M[5] * p.y + M[9] * p.z,
              M[2] * p.x + M[6] * p.y + M[10] * p.z};
}

__device__ __forceinline__ void load_tri(const MeshView& mesh, uint32_t prim, V3& p1, V3& p2, V3& p3) {
    const uint32_t i0 = mesh.tris[3 * prim], i1 = mesh.tris[3 * prim + 1], i2 = mesh.tris[3 * prim + 2];
    p1 = v3(mesh.verts[3 * i0], mesh.verts[3 * i0 + 1], mesh.verts[3 * i0 + 2]);
    p2 = v3(mesh.verts[3 * i1], mesh.verts[3 * i1 + 1], mesh.verts[3 * i1 + 2]);
    p3 = v3(mesh.verts[3 * i2], mesh.verts[3 * i2 + 1], mesh.verts[3 * i2 + 2]);
}

// IntersectWithJac(ray, triangle) without Jacobians (ray_casting.h:125-179)
__device__ __forceinline__ bool intersect_tri(V3 o, V3 d, V3 p1, V3 p2, V3 p3, V3& out) {
    const float eps = 1e-10f;
    const V3 e1 = p2 - p1, e2 = p3 - p1;
    const V3 rxe2 = cross(d, e2);
    const float det = dot(e1, rxe2);
    if (det > -eps && det < eps) return false;
    const float inv = (float)(1.0 / (double)det);
    const V3 s = o - p1;
    const float u = inv * dot(s, rxe2);
    if (u < 0.f || u > 1.f) return false;
    const V3 sxe1 = cross(s, e1);
    const float v = inv * dot(d, sxe1);
    if (v < 0.f || u + v > 1.f) return false;
    const float t = inv * dot(e2, sxe1);
    if (t < 0.f) return false;
    out = o + d * t;
    return true;
}

}  // namespace

// ---- first half of Evaluate, once per referenced keypoint (refiner.cc:306-350) -----------
__device__ __forceinline__ bool gate_closed(const BALmState* st, int gate) {
    if (st == nullptr || gate == GATE_NONE) return false;
    if (st->done) return true;
    return gate == GATE_BUILD ? !st->rebuild : st->skip != 0;
}

__global__ void __launch_bounds__(128) ba_refresh_kernel(BAView v, MeshView mesh, const BALmState* st, int gate) {
    if (gate_closed(st, gate)) return;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= v.n_kps) return;
    if (!v.referenced[g]) { v.pt_valid[g] = 0; return; }
    const Cam c = make_cam(v.cams[v.kp_frame[g]]);
    const V3 dc = unproject(c, v.kps[2 * g], v.kps[2 * g + 1]);
    const V3 ow = cam_center(c);
    const V3 dw = mul_t(c.R, dc);
    const V3 oo = hnorm_mul(v.Minv, ow);
    const V3 dobj = mul3x3(v.Minv, dw);
    bool found = false;
    V3 P = v3(0, 0, 0);
    const uint32_t prim = v.cache[g];
    if (prim != kInvalidPrim) {                               // :326-334
        V3 p1, p2, p3;
        load_tri(mesh, prim, p1, p2, p3);
        found = intersect_tri(oo, dobj, p1, p2, p3, P);
    }
    if (!found) {                                             // :336-350
        const HitRec h = bvh_nearest_hit(mesh.bvh, oo, dobj);
        bool ok = h.prim >= 0;
        if (ok && mesh.mask != nullptr && ((mesh.mask[h.prim >> 5] >> (h.prim & 31)) & 1u)) ok = false;
        if (ok) {
            V3 p1, p2, p3;
            load_tri(mesh, (uint32_t)h.prim, p1, p2, p3);
            const float w = (float)(1.0 - (double)h.u - (double)h.v);
            P = p1 * w + p2 * h.u + p3 * h.v;
            v.cache[g] = (uint32_t)h.prim;
            found = true;
        } else {
            v.cache[g] = kInvalidPrim;
        }
    }
    v.pt_valid[g] = found ? 1 : 0;
    if (found) {
        const V3 Pw = hnorm_mul(v.M, P);                      // :352-353
        v.pts[3 * g] = Pw.x; v.pts[3 * g + 1] = Pw.y; v.pts[3 * g + 2] = Pw.z;
    }
}

void launch_ba_refresh_points(const BAView& v, const MeshView& mesh, const BALmState* st, int gate, cudaStream_t s) {
    if (v.n_kps <= 0) return;      // every keypoint fell outside the projected mesh bbox: nothing to evaluate
    ba_refresh_kernel<<<(v.n_kps + 127) / 128, 128, 0, s>>>(v, mesh, st, gate);
}

// ---- TotalCost: one CTA per edge (lev_marq.h:773-824, refiner.cc:354-360) -----------------
constexpr int BA_THREADS = 256;

__global__ void __launch_bounds__(BA_THREADS) ba_cost_kernel(BAView v, Loss loss, const BALmState* st, int gate) {
    __shared__ double red_sum[BA_THREADS / 32];
    __shared__ int red_cnt[BA_THREADS / 32];
    if (gate_closed(st, gate)) return;
    const int e = blockIdx.x;
    if (v.edge_mask != nullptr && !v.edge_mask[e]) {           // another rank's edge
        if (threadIdx.x == 0) v.edge_cost[e] = 0.f;
        return;
    }
    const pc_ba_edge ed = v.edges[e];
    const Cam ct = make_cam(v.cams[ed.tgt_frame_idx]);
    const int koff = v.kp_offsets[ed.src_frame_idx];
    float acc = 0.f;
    int cnt = 0;
    for (int r = threadIdx.x; r < ed.rows; r += BA_THREADS) {
        const int row = ed.first_row + r;
        const int g = koff + (int)v.src_idx[row];
        if (!v.pt_valid[g]) continue;
        const V3 Pw = v3(v.pts[3 * g], v.pts[3 * g + 1], v.pts[3 * g + 2]);
        const V3 Pc = mul(ct.R, Pw) + ct.t;
        if (is_behind(ct, Pc)) continue;
        const float rx = ct.fx * Pc.x / Pc.z + ct.cx - v.tgt[2 * row];
        const float ry = ct.fy * Pc.y / Pc.z + ct.cy - v.tgt[2 * row + 1];
        acc += loss_value(loss, rx * rx + ry * ry);          // ResidualWeight == 1 (refiner.cc:259-266)
        cnt++;
    }
    double a = (double)acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if ((threadIdx.x & 31) == 0) { red_sum[threadIdx.x >> 5] = a; red_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0;
        int n = 0;
        for (int k = 0; k < BA_THREADS / 32; k++) { sum += red_sum[k]; n += red_cnt[k]; }
        float ec = (float)sum;
        if (n > 0) ec = ec / (float)n;                       // kShouldNormalize (lev_marq.h:812-816)
        v.edge_cost[e] = v.edge_weight[e] * ec;              // :818
    }
}

// cost = sum of the per-edge costs in edge order (float32, like the reference's cost += in TotalCost with one
// thread, lev_marq.h:818): a warp walks the list 32 at a time, lane 0 adds each group's terms in order.
__global__ void __launch_bounds__(32) ba_sum_cost_kernel(BAView v, const BALmState* st, int gate, float* cost_out) {
    if (gate_closed(st, gate)) return;
    float c = 0.f;
    for (int e0 = 0; e0 < v.n_edges; e0 += 32) {
        const int e = e0 + threadIdx.x;
        const float x = e < v.n_edges ? v.edge_cost[e] : 0.f;
#pragma unroll
        for (int l = 0; l < 32; l++) c += __shfl_sync(0xffffffffu, x, l);
    }
    if (threadIdx.x == 0) *cost_out = c;
}

void launch_ba_cost_edges(const BAView& v, const Loss& loss, const BALmState* st, int gate, cudaStream_t s) {
    if (v.n_edges > 0) ba_cost_kernel<<<v.n_edges, BA_THREADS, 0, s>>>(v, loss, st, gate);
}

void launch_ba_cost_sum(const BAView& v, const BALmState* st, int gate, float* cost_out, cudaStream_t s) {
    ba_sum_cost_kernel<<<1, 32, 0, s>>>(v, st, gate, cost_out);
}

void launch_ba_cost(const BAView& v, const Loss& loss, const BALmState* st, int gate, float* cost_out, cudaStream_t s) {
    launch_ba_cost_edges(v, loss, st, gate, s);
    launch_ba_cost_sum(v, st, gate, cost_out, s);
}

// ---- BuildNormalEquations: one CTA per edge ---------------------------------------------------
// EvaluateWithJacobian (refiner.cc:363-506).  J = [J_src | J_tgt] (2 x 2P).  Returns false when
// the residual is dropped (no cached primitive, degenerate plane, point behind the target).
template <int P>
__device__ __forceinline__ bool ba_jacobian(const BAView& v, const MeshView& mesh, const Cam& cs, const Cam& ct,
                                            bool src_fixed, bool tgt_fixed, int g, int row, float* J0, float* J1,
                                            float& rx, float& ry) {
    const uint32_t prim = v.cache[g];
    if (prim == kInvalidPrim) return false;                   // :391-393
    const float sx = v.kps[2 * g], sy = v.kps[2 * g + 1];
    const V3 dirCam = unproject(cs, sx, sy);
    const V3 origin = cam_center(cs);
    const V3 dirW = mul_t(cs.R, dirCam);
    V3 p1, p2, p3;
    load_tri(mesh, prim, p1, p2, p3);
    const V3 plane_pt = v3(v.M[0] * p1.x + v.M[1] * p1.y + v.M[2] * p1.z + v.M[3],
                           v.M[4] * p1.x + v.M[5] * p1.y + v.M[6] * p1.z + v.M[7],
                           v.M[8] * p1.x + v.M[9] * p1.y + v.M[10] * p1.z + v.M[11]);       // :422-423
    const V3 n = mul3x3_t(v.Minv, cross(p2 - p1, p3 - p1));                                // :424-428
    const float d_dot_n = dot(dirW, n);                       // ray_casting.h:90
    if ((double)d_dot_n > -1e-10 && (double)d_dot_n < 1e-10) return false;   // reference CHECK(ok) would throw
    const float p0 = dot(plane_pt - origin, n);
    const float t = (float)((double)p0 / (double)d_dot_n);
    const V3 X = origin + dirW * t;
    const V3 XCam = mul(ct.R, X) + ct.t;
    if (is_behind(ct, XCam)) return false;                    // :444-446
    rx = ct.fx * XCam.x / XCam.z + ct.cx - v.tgt[2 * row];
    ry = ct.fy * XCam.y / XCam.z + ct.cy - v.tgt[2 * row + 1];
    // dp/dXCam (types.h:79-85)
    const float iz = 1.f / XCam.z;
    const V3 dp0 = v3(ct.fx * iz, 0.f, -ct.fx * XCam.x / (XCam.z * XCam.z));
    const V3 dp1 = v3(0.f, ct.fy * iz, -ct.fy * XCam.y / (XCam.z * XCam.z));
#pragma unroll
    for (int k = 0; k < 2 * P; k++) { J0[k] = 0.f; J1[k] = 0.f; }
    if (!src_fixed) {
        // dp_dX = dp_dXCam * R_t ; rows as vectors: R_t^T dp
        const V3 a0 = mul_t(ct.R, dp0), a1 = mul_t(ct.R, dp1);
        // G = dp_dX * (I - dirW n^T / d_dot_n)   (rows g0, g1)
        const float k0 = dot(a0, dirW) / d_dot_n, k1 = dot(a1, dirW) / d_dot_n;
        const V3 g0 = a0 - n * k0, g1 = a1 - n * k1;
        // J_src[:,0:3] = G (Skew(origin) + t Skew(dirW));  row * Skew(w) = cross(row, w)
        const V3 r0 = cross(g0, origin) + cross(g0, dirW) * t;
        const V3 r1 = cross(g1, origin) + cross(g1, dirW) * t;
        J0[0] = r0.x; J0[1] = r0.y; J0[2] = r0.z;
        J1[0] = r1.x; J1[1] = r1.y; J1[2] = r1.z;
        // J_src[:,3:6] = G * (-R_s^T): row * R_s^T = (R_s row)
        const V3 t0 = -mul(cs.R, g0), t1 = -mul(cs.R, g1);
        J0[3] = t0.x; J0[4] = t0.y; J0[5] = t0.z;
        J1[3] = t1.x; J1[4] = t1.y; J1[5] = t1.z;
        if (P == 9) {
            // J_src[:,6:9] = G * t * R_s^T * dDirCam_dIntrin  (types.h:116-123)
            const V3 h0 = mul(cs.R, g0) * t, h1 = mul(cs.R, g1) * t;
            const float i00 = cs.sgn * (cs.cx - sx) / (cs.fy * cs.fy * cs.aspect), i01 = -cs.sgn / cs.fx;
            const float i10 = cs.sgn * (cs.cy - sy) / (cs.fy * cs.fy), i12 = -cs.sgn / cs.fy;
            if (v.opt_f) { J0[6] = h0.x * i00 + h0.y * i10; J1[6] = h1.x * i00 + h1.y * i10; }
            if (v.opt_pp) {
                J0[7] = h0.x * i01; J1[7] = h1.x * i01;
                J0[8] = h0.y * i12; J1[8] = h1.y * i12;
            }
        }
    }
    if (!tgt_fixed) {
        // J_tgt[:,0:3] = dp * (R_t Skew(-X)): row*R_t = R_t^T row, then cross with -X
        const V3 b0 = mul_t(ct.R, dp0), b1 = mul_t(ct.R, dp1);
        const V3 q0 = cross(b0, -X), q1 = cross(b1, -X);
        J0[P + 0] = q0.x; J0[P + 1] = q0.y; J0[P + 2] = q0.z;
        J1[P + 0] = q1.x; J1[P + 1] = q1.y; J1[P + 2] = q1.z;
        J0[P + 3] = dp0.x; J0[P + 4] = dp0.y; J0[P + 5] = dp0.z;
        J1[P + 3] = dp1.x; J1[P + 4] = dp1.y; J1[P + 5] = dp1.z;
        if (P == 9) {
            if (v.opt_f) { J0[P + 6] = ct.aspect * XCam.x / XCam.z; J1[P + 6] = XCam.y / XCam.z; }
            if (v.opt_pp) { J0[P + 7] = 1.f; J1[P + 8] = 1.f; }
        }
    }
    return true;
}

// Accumulates rows [R0, R1) of the lower triangle of J^T W J and of J^T W r.
template <int P, int R0, int R1>
__device__ __forceinline__ void ba_edge_accumulate(const BAView& v, const MeshView& mesh, const Loss& loss,
                                                   const pc_ba_edge& ed, float ew, int tid_in_half, int half_threads,
                                                   float* pair_out_smem, int* cnt_out) {
    constexpr int NP = 2 * P;
    constexpr int NACC = (R1 * (R1 + 1) - R0 * (R0 + 1)) / 2 + (R1 - R0);
    const Cam cs = make_cam(v.cams[ed.src_frame_idx]);
    const Cam ct = make_cam(v.cams[ed.tgt_frame_idx]);
    const bool src_fixed = ed.src_frame_idx == 0 || ed.src_frame_idx == v.nf - 1;   // IsGroundTruth (refiner.cc:268-271)
    const bool tgt_fixed = ed.tgt_frame_idx == 0 || ed.tgt_frame_idx == v.nf - 1;
    const int koff = v.kp_offsets[ed.src_frame_idx];
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; k++) acc[k] = 0.f;
    int cnt = 0;
    for (int r = tid_in_half; r < ed.rows; r += half_threads) {
        const int row = ed.first_row + r;
        const int g = koff + (int)v.src_idx[row];
        float J0[NP], J1[NP], rx, ry;
        if (!ba_jacobian<P>(v, mesh, cs, ct, src_fixed, tgt_fixed, g, row, J0, J1, rx, ry)) continue;
        const float w = ew * 1.0f * loss_weight(loss, rx * rx + ry * ry);          // lev_marq.h:690-693
        cnt++;
        int k = 0;
#pragma unroll
        for (int a = R0; a < R1; a++)
#pragma unroll
            for (int b = 0; b <= a; b++) acc[k++] += (J0[a] * J0[b] + J1[a] * J1[b]) * w;
        const float wrx = w * rx, wry = w * ry;
#pragma unroll
        for (int a = R0; a < R1; a++) acc[k++] += J0[a] * wrx + J1[a] * wry;
    }
    // warp reduce, then one slot per warp in shared memory: [warp][NACC]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NACC; k++) {
        float a = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) pair_out_smem[wid * 128 + k] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) cnt_out[wid] = cnt;
}

// P == 6: one half (256 threads).  P == 9: two halves of 128 threads that each evaluate every
// residual and keep rows [0,12) resp. [12,18) of the 18x18 lower triangle (register budget).
template <int P>
__global__ void __launch_bounds__(256) ba_build_kernel(BAView v, MeshView mesh, Loss loss, const BALmState* st, int gate) {
    constexpr int NP = 2 * P;
    constexpr int HALVES = P == 6 ? 1 : 2;
    constexpr int HT = 256 / HALVES;          // threads per half
    constexpr int HW = HT / 32;               // warps per half
    constexpr int SPLIT = 12;
    __shared__ float s_acc[8 * 128];
    __shared__ int s_cnt[8];
    if (gate_closed(st, gate)) return;
    const int e = blockIdx.x;
    const pc_ba_edge ed = v.edges[e];
    const float ew = (v.edge_mask != nullptr && !v.edge_mask[e]) ? 0.f : v.edge_weight[e];   // another rank's edge: no contribution
    float* out = v.edge_pair + (size_t)e * v.pair_stride;
    const int half = threadIdx.x / HT;
    if (ew == 0.f) {                                            // lev_marq.h:670-673
        for (int k = threadIdx.x; k < v.pair_stride; k += blockDim.x) out[k] = 0.f;
        return;
    }
    if constexpr (HALVES == 1) {
        ba_edge_accumulate<P, 0, NP>(v, mesh, loss, ed, ew, threadIdx.x, HT, s_acc, s_cnt);
    } else {
        if (half == 0) ba_edge_accumulate<P, 0, SPLIT>(v, mesh, loss, ed, ew, threadIdx.x, HT, s_acc, s_cnt);
        else ba_edge_accumulate<P, SPLIT, NP>(v, mesh, loss, ed, ew, threadIdx.x - HT, HT, s_acc, s_cnt);
    }
    __syncthreads();
    int n = 0;
    for (int k = 0; k < HW; k++) n += s_cnt[k];                 // valid residuals (half 0 saw them all)
    const float inv_n = n > 0 ? 1.f / (float)n : 1.f;
    // output layout: lower triangle rows 0..NP-1 (row-major packed), then Jtr[NP], then count
    constexpr int TRI = NP * (NP + 1) / 2;
    for (int k = threadIdx.x; k < TRI + NP; k += blockDim.x) {
        // locate (half, local index) of packed entry k
        int h = 0, local = k;
        if (HALVES == 2) {
            constexpr int TRI0 = SPLIT * (SPLIT + 1) / 2;
            if (k < TRI) {
                if (k < TRI0) { h = 0; local = k; } else { h = 1; local = k - TRI0; }
            } else {
                const int a = k - TRI;
                if (a < SPLIT) { h = 0; local = TRI0 + a; } else { h = 1; local = (TRI - TRI0) + (a - SPLIT); }
            }
        }
        float sum = 0.f;
        for (int wq = 0; wq < HW; wq++) sum += s_acc[(h * HW + wq) * 128 + local];
        out[k] = n > 0 ? sum / (float)n : sum;                  // lev_marq.h:705-710
    }
    (void)inv_n;
    if (threadIdx.x == 0) out[TRI + NP] = (float)n;
}

void launch_ba_build(const BAView& v, const MeshView& mesh, const Loss& loss, const BALmState* st, int gate, cudaStream_t s) {
    if (v.n_edges <= 0) return;
    if (v.p == 6) ba_build_kernel<6><<<v.n_edges, 256, 0, s>>>(v, mesh, loss, st, gate);
    else ba_build_kernel<9><<<v.n_edges, 256, 0, s>>>(v, mesh, loss, st, gate);
}

// ---- assembly into the block-banded matrix: one CTA per frame -----------------------------------
// band[i][k] = block (row i, col i-k), p x p row-major; the diagonal block is stored full
// (symmetric).  AccumBlockInSparseMatrix / Jtr accumulation (lev_marq.h:712-766) in ascending
// edge order.
__device__ __forceinline__ float pair_at(const float* pr, int a, int b) {   // symmetric read of packed lower
    if (a < b) { const int t = a; a = b; b = t; }
    return pr[a * (a + 1) / 2 + b];
}

__global__ void __launch_bounds__(256) ba_assemble_kernel(BAView v, const BALmState* st, int gate) {
    if (gate_closed(st, gate)) return;
    const int i = blockIdx.x;
    const int p = v.p, pp = p * p, NP = 2 * p, TRI = NP * (NP + 1) / 2;
    float* band_i = v.band + (size_t)i * kBandBlocks * pp;
    const int e0 = v.inc_offsets[i], e1 = v.inc_offsets[i + 1];
    for (int idx = threadIdx.x; idx < kBandBlocks * pp + p; idx += blockDim.x) {
        float acc = 0.f;
        if (idx < kBandBlocks * pp) {
            const int k = idx / pp, a = (idx - k * pp) / p, b = idx - k * pp - a * p;
            for (int q = e0; q < e1; q++) {
                const int e = v.inc_edges[q];
                const pc_ba_edge ed = v.edges[e];
                const float* pr = v.edge_pair + (size_t)e * v.pair_stride;
                if (k == 0) {
                    if (ed.src_frame_idx == i) acc += pair_at(pr, a, b);
                    if (ed.tgt_frame_idx == i) acc += pair_at(pr, p + a, p + b);
                } else {
                    const int j = i - k;
                    if (ed.src_frame_idx == i && ed.tgt_frame_idx == j) acc += pair_at(pr, a, p + b);      // J1^T J2
                    else if (ed.tgt_frame_idx == i && ed.src_frame_idx == j) acc += pair_at(pr, p + a, b); // J2^T J1
                }
            }
            band_i[idx] = acc;
        } else {
            const int a = idx - kBandBlocks * pp;
            for (int q = e0; q < e1; q++) {
                const int e = v.inc_edges[q];
                const pc_ba_edge ed = v.edges[e];
                const float* pr = v.edge_pair + (size_t)e * v.pair_stride + TRI;
                if (ed.src_frame_idx == i) acc += pr[a];
                if (ed.tgt_frame_idx == i) acc += pr[p + a];
            }
            v.jtr[i * p + a] = acc;
        }
    }
    __syncthreads();
    if (threadIdx.x < p) {                                       // JtJ_diag clamp (lev_marq.h:770)
        const float d = band_i[threadIdx.x * p + threadIdx.x];
        v.diag[i * p + threadIdx.x] = fminf(fmaxf(d, 1e-6f), 1e32f);
    }
}

__global__ void __launch_bounds__(1024) ba_grad_norm_kernel(BAView v) {
    __shared__ double red[32];
    const int n = v.nf * v.p;
    double a = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) a += (double)v.jtr[k] * (double)v.jtr[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < 32; k++) s += red[k];
        v.scalars[1] = (float)sqrt(s);
    }
}

void launch_ba_assemble(const BAView& v, const BALmState* st, int gate, cudaStream_t s) {
    ba_assemble_kernel<<<v.nf, 256, 0, s>>>(v, st, gate);
    if (st == nullptr) ba_grad_norm_kernel<<<1, 1024, 0, s>>>(v);   // the LM loop forms |Jtr| in its solve kernel
}

}  // namespace pc
