// K8q: the 10x10 LK kernel of lk10.cu run as a work queue.
//
// lk10_kernel<CACHED> gives every pentad (five lanes = OpenCV's five float accumulation chains, see
// lk10.cu) one keypoint and walks the pyramid levels in lock step, so a warp iterates at every level
// until the slowest of its six keypoints has converged: ncu shows 19.9 of 32 lanes active
// (profiles/r1_v_ncu_lk10_with_templates.txt).  With the per-frame template cache a pentad can change
// (keypoint, level) for a few 16-byte loads, so here each pentad runs its own state machine:
//
//   NEED  -> (item finished: write next/status, take the next (pair, keypoint) from the launch's queue)
//            load the level's source template (15 x LDG.128 per lane, already unpacked) + A11/A12/A22/1/D,
//            apply the level's entry tests
//   ITER  -> one window pass per round together with every other iterating pentad of the warp
//   IDLE  -> queue empty
//
// and the warp only runs the (cheap) NEED phase in the rounds where one of its pentads asks for it.
// The queue is one global counter per launch; items are ordered slow pairs first (|skip| 8 .. 1), so the
// launch ends on the short ones.  The L1 error of the final position (level 0) does not depend on the
// iteration history, so it is a separate, perfectly balanced pass (lk10q_err_kernel) instead of a
// one-pentad-in-six phase of the queue kernel.
//
// Arithmetic per (pair, keypoint, level) is that of lk10_kernel -- same helpers (lk10_common.cuh), same
// explicit-rounding intrinsics, compiled with -fmad=false -- so results are bit-identical to the oracle
// (oracle/restate.c::orc_lk restating cv::calcOpticalFlowPyrLK, call site
// /root/reference/cpp/opticalflow.cc:119-125); only the schedule differs.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "lk10_common.cuh"

namespace pc {

namespace {

using namespace lk10;

constexpr int QUADS = kLkQueueQuadsPerLane;      // uint4 per lane and (level, keypoint)
static_assert(kLkQueueTemplateBytesPerPoint == QUADS * PENTAD * 16, "template stride");

// ---- template store (queue layout) -------------------------------------------------------------
// Per (level, keypoint): QUADS x PENTAD uint4, quad-major, so that the five lanes of a pentad read 80
// consecutive bytes per load.  Quads 0..4 hold the lane's Ival[0..19], 5..9 Ix, 10..14 Iy as plain
// ints (they go straight into the registers the window pass reads); quad 15 holds, for role 0,
// (A11, A12, A22, 1/D) and for role 1 (minEig, D, 0, 0) -- the level's entry tests and the 2x2 solve
// no longer cost the LK kernel a sqrt and two divisions per (keypoint, level, pair).
__global__ void __launch_bounds__(LK_WARPS * 32, 4) lk10q_template_kernel(PyramidView a, const float* __restrict__ pts,
                                                                          const int* __restrict__ n_pts, int cap, int nlev,
                                                                          uint4* __restrict__ words) {
    const int level = blockIdx.y;
    if (level >= min(a.levels, nlev)) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / PENTAD, role = lane - grp * PENTAD;
    const int base = grp < PTS_PER_WARP ? grp * PENTAD : 0;
    const int npts = min(*n_pts, cap);
    const int pi = (blockIdx.x * LK_WARPS + wib) * PTS_PER_WARP + grp;
    const bool valid = grp < PTS_PER_WARP && pi < npts;
    if (!__any_sync(FULL, valid)) return;
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
    // the level's 2x2 system, in the operation order of lk10_kernel
    const float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dA = __fsub_rn(A11, A22);
    const float rad = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * WIN * WIN));
    const float Dinv = __fdiv_rn(1.f, D);
    uint4* dst = words + ((size_t)level * cap + pi) * (QUADS * PENTAD) + role;
#pragma unroll
    for (int q = 0; q < 5; q++) {
        dst[q * PENTAD] = make_uint4((uint32_t)Ival[4 * q], (uint32_t)Ival[4 * q + 1], (uint32_t)Ival[4 * q + 2], (uint32_t)Ival[4 * q + 3]);
        dst[(5 + q) * PENTAD] = make_uint4((uint32_t)Ix[4 * q], (uint32_t)Ix[4 * q + 1], (uint32_t)Ix[4 * q + 2], (uint32_t)Ix[4 * q + 3]);
        dst[(10 + q) * PENTAD] = make_uint4((uint32_t)Iy[4 * q], (uint32_t)Iy[4 * q + 1], (uint32_t)Iy[4 * q + 2], (uint32_t)Iy[4 * q + 3]);
    }
    if (role == 0) dst[15 * PENTAD] = make_uint4(__float_as_uint(A11), __float_as_uint(A12), __float_as_uint(A22), __float_as_uint(Dinv));
    if (role == 1) dst[15 * PENTAD] = make_uint4(__float_as_uint(minEig), __float_as_uint(D), 0u, 0u);
}

// ---- the queue kernel ---------------------------------------------------------------------------
struct QPair {                     // what a pentad needs of its pair, staged in shared memory
    const float* pts;
    float* next;
    uint8_t* status;
    const uint4* tmpl;
    int tcap, nlev;
    const uint8_t* bimg[kMaxLevels];
    int bpitch[kMaxLevels], bw[kMaxLevels], bh[kMaxLevels], aw[kMaxLevels], ah[kMaxLevels];
};

enum { ST_NEED = 0, ST_ITER = 1, ST_IDLE = 2 };

// ---- shared-memory template slots, filled by the bulk-copy engine ---------------------------------
// The order in which a pentad needs templates is known in advance -- (item, top level) .. (item, 0), then
// the item it has already reserved from the queue -- so the next one is always in flight while the current
// level iterates: one cp.async.bulk (1280 B, global -> shared, completion on the pentad's mbarrier) issued
// by the pentad's first lane.  Entering a level then costs an mbarrier test and 16 LDS.128 instead of a
// DRAM/L2 round trip that stalls all six pentads of the warp (the first version of this kernel loaded the
// templates with LDG at the transition: 22 % fewer instructions than the lock-step kernel but 47 % issue
// utilisation against 62 %, long-scoreboard stalls 3.9 per issue -- profiles/r2_i_ncu_lk10q_ldg.txt).
constexpr int SLOT_BYTES = QUADS * PENTAD * 16 + 80;         // +80: consecutive slots start 20 banks apart
constexpr int SLOT_QUADS = SLOT_BYTES / 16;
constexpr int SLOTS = LK_WARPS * PTS_PER_WARP;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(uint4* dst, const uint4* src, uint64_t* bar) {
    constexpr uint32_t bytes = QUADS * PENTAD * 16;
    // the slot's previous contents were read through the generic proxy (LDS); order those reads before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    }
}

__global__ void __launch_bounds__(LK_WARPS * 32, 4) lk10q_kernel(LKBatch batch, LKParams prm) {
    __shared__ __align__(128) uint4 s_tmpl[SLOTS * SLOT_QUADS];
    __shared__ uint64_t s_bar[SLOTS];
    __shared__ QPair s_pair[kMaxPairsPerLaunch];
    __shared__ int s_off[kMaxPairsPerLaunch + 1];
    __shared__ int s_taken;
    const int np = batch.num_pairs;
    if (threadIdx.x < kMaxPairsPerLaunch) {
        const int k = threadIdx.x;                   // queue order: slow pairs (appended last) first
        int cnt = 0;
        if (k < np) {
            const LKPair& pr = batch.pair[np - 1 - k];
            QPair& q = s_pair[k];
            q.pts = pr.pts; q.next = pr.next; q.status = pr.status;
            q.tmpl = pr.tmpl.words; q.tcap = pr.tmpl.cap;
            q.nlev = min(min(pr.a.levels, pr.b.levels), prm.max_level + 1);
#pragma unroll
            for (int l = 0; l < kMaxLevels; l++) {
                q.bimg[l] = pr.b.data[l]; q.bpitch[l] = pr.b.pitch[l];
                q.bw[l] = pr.b.w[l]; q.bh[l] = pr.b.h[l];
                q.aw[l] = pr.a.w[l]; q.ah[l] = pr.a.h[l];
            }
            cnt = q.nlev > 0 ? min(*pr.n_pts, batch.cap) : 0;
        }
        // exclusive prefix over the eight counts (one warp's first lanes)
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < kMaxPairsPerLaunch; d <<= 1) {
            const int v = __shfl_up_sync(0xffu, incl, d, 8);
            if (k >= d) incl += v;
        }
        s_off[k] = incl - cnt;
        if (k == kMaxPairsPerLaunch - 1) s_off[kMaxPairsPerLaunch] = incl;
        if (k == 0) s_taken = 0;
    }
    if (threadIdx.x < SLOTS) bar_init(&s_bar[threadIdx.x]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int total = s_off[kMaxPairsPerLaunch];
    const int budget = batch.queue_budget;

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / PENTAD, role = lane - grp * PENTAD;
    const bool worker = grp < PTS_PER_WARP;
    const int base = worker ? grp * PENTAD : 0;       // first lane of this pentad (shuffle source)
    const int simd_flag = role < 4 ? 1 : 0;
    const int xa = simd_flag ? role : 8;              // first column of the lane's chain (second: +4, tail +1)
    constexpr unsigned kLeaders = 0x02108421u;        // role-0 lanes of the six pentads
    const unsigned my_leader = worker ? (1u << base) : 0u;
    const int slot = wib * PTS_PER_WARP + (worker ? grp : 0);
    uint4* const my_tmpl = s_tmpl + slot * SLOT_QUADS;
    uint64_t* const my_bar = &s_bar[slot];

    const float halfw = (WIN - 1) * 0.5f;
    const double eps2 = prm.eps * prm.eps;
    const float eps2_lo = (float)(eps2 * (1.0 - 1e-6)), eps2_hi = (float)(eps2 * (1.0 + 1e-6));
    const float FLT_SCALE = 1.f / (float)(1 << 20);

    int state = worker ? ST_NEED : ST_IDLE;
    int level = -1, qk = -1, pi = 0, status = 1, j = 0;
    int nxt_it = -1;                                  // the item this pentad has reserved (-1: none left)
    uint32_t ph = 0;                                  // parity of the slot's next completion
    bool pf_issued = false;                           // the slot holds / is receiving the template needed next
    bool pf_due = false;                              // the slot has been consumed: request its successor
    unsigned pend_mask = 0;                           // leaders whose reservation is in flight (warp-uniform)
    int pend_first = 0;                               // lane 0: the atomic's result, read when it is collected
    float ptx = 0.f, pty = 0.f, nextx = 0.f, nexty = 0.f, nx = 0.f, ny = 0.f, pdx = 0.f, pdy = 0.f;
    float A11 = 0.f, A12 = 0.f, A22 = 0.f, Dinv = 0.f;
    LevelRef B = {nullptr, 0, 0, 0};
    int Ival[20], Ix[20], Iy[20];
#pragma unroll
    for (int i = 0; i < 20; i++) { Ival[i] = 0; Ix[i] = 0; Iy[i] = 0; }

    // queue index -> (pair slot, keypoint)
    auto decode = [&](int it, int& k, int& p) {
        k = 0;
#pragma unroll
        for (int m = 1; m < kMaxPairsPerLaunch; m++) k += it >= s_off[m] ? 1 : 0;
        p = it - s_off[k];
    };
    // reserve one item for every pentad in `leaders` (warp-uniform); the result is collected later
    auto reserve = [&](unsigned leaders) {
        if (lane == 0) {
            int n = __popc(leaders);
            if (budget > 0 && atomicAdd(&s_taken, n) >= budget) n = 0;        // this block has had its share
            pend_first = n ? atomicAdd(batch.queue, n) : total;
        }
        pend_mask = leaders;
    };
    auto collect = [&]() {
        const int first = __shfl_sync(FULL, pend_first, 0);
        if (pend_mask & my_leader) {
            const int it = first + __popc(pend_mask & (my_leader - 1u));
            nxt_it = it < total ? it : -1;
        }
        pend_mask = 0;
    };
    // request the template that follows (qk, pi, level) in this pentad's sequence
    auto prefetch_next = [&]() {
        if (level > 0) {
            if (role == 0) bulk_load(my_tmpl, s_pair[qk].tmpl + ((size_t)(level - 1) * s_pair[qk].tcap + pi) * (QUADS * PENTAD), my_bar);
            pf_issued = true;
        } else if (nxt_it >= 0) {
            int k2, p2;
            decode(nxt_it, k2, p2);
            const QPair& q2 = s_pair[k2];
            if (role == 0) bulk_load(my_tmpl, q2.tmpl + ((size_t)(q2.nlev - 1) * q2.tcap + p2) * (QUADS * PENTAD), my_bar);
            pf_issued = true;
        }
    };

    // every pentad starts with one reserved item (blocking) and asks for it like for any other
    reserve(kLeaders);
    collect();

    for (;;) {
        bool skipped = false;
        // ---- NEED: finished items leave, reserved items / next levels enter ------------------------
        if (__any_sync(FULL, state == ST_NEED)) {
            const bool fin = state == ST_NEED && level < 0;
            const unsigned fin_mask = __ballot_sync(FULL, fin);
            if (fin_mask) {
                if (fin && qk >= 0 && role == 0) {
                    const QPair& q = s_pair[qk];
                    reinterpret_cast<float2*>(q.next)[pi] = make_float2(nextx, nexty);
                    q.status[pi] = (uint8_t)status;
                }
                if (pend_mask) collect();             // a reservation still in flight: its owner may be leaving now
                if (fin) {
                    if (nxt_it < 0) {
                        state = ST_IDLE; qk = -1;
                    } else {
                        decode(nxt_it, qk, pi);
                        const float2 p = __ldg(reinterpret_cast<const float2*>(s_pair[qk].pts) + pi);
                        ptx = p.x; pty = p.y;
                        level = s_pair[qk].nlev - 1;
                        status = 1;
                        nxt_it = -1;
                    }
                }
                const unsigned leaders = __ballot_sync(FULL, fin && state == ST_NEED) & kLeaders;
                if (leaders) reserve(leaders);        // collected when the item reaches level 0 (or at the next exit)
            }
            if (state == ST_NEED) {
                const QPair& q = s_pair[qk];
                const float scale = __int_as_float((127 - level) << 23);           // 1 / 2^level
                float prevx = __fmul_rn(ptx, scale), prevy = __fmul_rn(pty, scale);
                if (level == q.nlev - 1) { nextx = prevx; nexty = prevy; }
                else { nextx = __fmul_rn(nextx, 2.f); nexty = __fmul_rn(nexty, 2.f); }
                prevx = __fsub_rn(prevx, halfw); prevy = __fsub_rn(prevy, halfw);
                const int ipx = __float2int_rd(prevx), ipy = __float2int_rd(prevy);
                bool ok = !(ipx < -WIN || ipx >= q.aw[level] || ipy < -WIN || ipy >= q.ah[level]);
                // the first iteration's test of the start position against the target level (lk10_kernel makes it at
                // j = 0, after the template; all three entry tests only clear `status` at level 0 and leave the
                // level, so their order is immaterial) -- made here so that a level whose template is read into
                // registers always runs a window pass in the same round
                {
                    const int inx0 = __float2int_rd(__fsub_rn(nextx, halfw)), iny0 = __float2int_rd(__fsub_rn(nexty, halfw));
                    if (prm.iters > 0 &&
                        ((unsigned)(inx0 + WIN) >= (unsigned)(q.bw[level] + WIN) || (unsigned)(iny0 + WIN) >= (unsigned)(q.bh[level] + WIN)))
                        ok = false;
                }
                // the slot: this (item, level)'s template (requested while the previous level iterated)
                if (!pf_issued && role == 0)
                    bulk_load(my_tmpl, q.tmpl + ((size_t)level * q.tcap + pi) * (QUADS * PENTAD), my_bar);
                bar_wait(my_bar, ph);
                ph ^= 1u;
                pf_issued = false;
                if (ok) {
                    const uint4 s0 = my_tmpl[15 * PENTAD], s1 = my_tmpl[15 * PENTAD + 1];
                    const float minEig = __uint_as_float(s1.x), D = __uint_as_float(s1.y);
                    if ((double)minEig < prm.min_eig || D < 1.1920928955078125e-07f) {
                        ok = false;
                    } else {
                        A11 = __uint_as_float(s0.x); A12 = __uint_as_float(s0.y); A22 = __uint_as_float(s0.z);
                        Dinv = __uint_as_float(s0.w);
                        const uint4* tl = my_tmpl + role;
#pragma unroll
                        for (int k = 0; k < 5; k++) {
                            const uint4 v = tl[k * PENTAD];
                            Ival[4 * k] = (int)v.x; Ival[4 * k + 1] = (int)v.y; Ival[4 * k + 2] = (int)v.z; Ival[4 * k + 3] = (int)v.w;
                        }
#pragma unroll
                        for (int k = 0; k < 5; k++) {
                            const uint4 v = tl[(5 + k) * PENTAD];
                            Ix[4 * k] = (int)v.x; Ix[4 * k + 1] = (int)v.y; Ix[4 * k + 2] = (int)v.z; Ix[4 * k + 3] = (int)v.w;
                        }
#pragma unroll
                        for (int k = 0; k < 5; k++) {
                            const uint4 v = tl[(10 + k) * PENTAD];
                            Iy[4 * k] = (int)v.x; Iy[4 * k + 1] = (int)v.y; Iy[4 * k + 2] = (int)v.z; Iy[4 * k + 3] = (int)v.w;
                        }
                        B.img = q.bimg[level]; B.w = q.bw[level]; B.h = q.bh[level]; B.pitch = q.bpitch[level];
                        nx = __fsub_rn(nextx, halfw); ny = __fsub_rn(nexty, halfw);
                        pdx = 0.f; pdy = 0.f; j = 0;
                    }
                }
                pf_due = true;
                if (ok && prm.iters > 0) {
                    state = ST_ITER;
                } else {
                    skipped = true;                          // leaves this level after the slot's successor is requested
                    if (!ok && level == 0) status = 0;
                }
            }
        }
        // ---- ITER: one window pass for every iterating pentad -------------------------------------
        bool done = false;
        if (__any_sync(FULL, state == ST_ITER)) {
            bool iterating = state == ST_ITER;
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (iterating && ((unsigned)(inx + WIN) >= (unsigned)(B.w + WIN) || (unsigned)(iny + WIN) >= (unsigned)(B.h + WIN))) {
                if (level == 0) status = 0;                  // inx < -WIN || inx >= B.w || iny < -WIN || iny >= B.h
                iterating = false;
                done = true;
            }
            int w00, w01, w10, w11;
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            float bx = 0.f, by = 0.f;
            int unused = 0;
            if (iterating) window_pass<false>(B, inx, iny, xa, simd_flag, w00, w01, w10, w11, Ival, Ix, Iy, bx, by, unused);
            __syncwarp();
            const float b1 = __fmul_rn(pentad_total(bx, base, role), FLT_SCALE);
            const float b2 = __fmul_rn(pentad_total(by, base, role), FLT_SCALE);
            if (iterating) {
                const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dinv);
                const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dinv);
                nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
                nextx = __fadd_rn(nx, halfw); nexty = __fadd_rn(ny, halfw);
                // OpenCV tests dx*dx + dy*dy <= eps^2 in double; the float sum decides it except within
                // 1e-6 (relative) of the threshold, where the double expression is evaluated
                const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                bool small = d2 <= eps2_lo;
                if (!small && d2 < eps2_hi)
                    small = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps2;
                if (small) {
                    done = true;
                } else if (j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
                    nextx = __fsub_rn(nextx, __fmul_rn(dx, 0.5f));
                    nexty = __fsub_rn(nexty, __fmul_rn(dy, 0.5f));
                    done = true;
                }
                pdx = dx; pdy = dy;
                if (++j >= prm.iters) done = true;
            }
        } else if (!__any_sync(FULL, state == ST_NEED)) {
            break;
        }
        // ---- the slot has been read into registers (and the window pass above has used every one of them,
        // so no LDS of the slot is in flight): request this pentad's next template ---------------------
        if (__any_sync(FULL, pf_due)) {
            if (pend_mask && __any_sync(FULL, pf_due && level == 0)) collect();     // its reserved item is needed now
            if (pf_due) {
                prefetch_next();
                pf_due = false;
            }
        }
        if (skipped || (state == ST_ITER && done)) { level--; state = ST_NEED; }
    }
}

// ---- L1 error of the final position (level 0), one pentad per (pair, keypoint) ------------------
__global__ void __launch_bounds__(LK_WARPS * 32) lk10q_err_kernel(LKBatch batch) {
    const LKPair& pr = batch.pair[blockIdx.y];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / PENTAD, role = lane - grp * PENTAD;
    const int base = grp < PTS_PER_WARP ? grp * PENTAD : 0;
    const int npts = min(*pr.n_pts, batch.cap);
    const int pi = (blockIdx.x * LK_WARPS + wib) * PTS_PER_WARP + grp;
    const bool valid = grp < PTS_PER_WARP && pi < npts;
    if (!__any_sync(FULL, valid)) return;
    const int simd_flag = role < 4 ? 1 : 0;
    const int xa = simd_flag ? role : 8;
    const float halfw = (WIN - 1) * 0.5f;
    int status = 0;
    float nextx = 0.f, nexty = 0.f;
    if (valid) {
        status = pr.status[pi];
        const float2 p = reinterpret_cast<const float2*>(pr.next)[pi];
        nextx = p.x; nexty = p.y;
    }
    const LevelRef B = {pr.b.data[0], pr.b.w[0], pr.b.h[0], pr.b.pitch[0]};
    bool want = valid && status != 0;
    const float fx = __fsub_rn(nextx, halfw), fy = __fsub_rn(nexty, halfw);
    const int inx = __float2int_rd(fx), iny = __float2int_rd(fy);
    if (want && (inx < -WIN || inx >= B.w || iny < -WIN || iny >= B.h)) {
        status = 0;
        want = false;
    }
    int w00, w01, w10, w11;
    bilinear_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), w00, w01, w10, w11);
    int Ival[20];
    int esum = 0;
    float f0, f1;
    if (want) {
        const uint4* tl = pr.tmpl.words + (size_t)pi * (QUADS * PENTAD) + role;      // level 0
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint4 v = __ldg(tl + k * PENTAD);
            Ival[4 * k] = (int)v.x; Ival[4 * k + 1] = (int)v.y; Ival[4 * k + 2] = (int)v.z; Ival[4 * k + 3] = (int)v.w;
        }
        window_pass<true>(B, inx, iny, xa, simd_flag, w00, w01, w10, w11, Ival, Ival, Ival, f0, f1, esum);
    }
    __syncwarp();
    int tot = 0;
#pragma unroll
    for (int k = 0; k < PENTAD; k++) tot += __shfl_sync(FULL, esum, base + k);
    // every partial sum is an integer < 2^24: the float summation order of the CPU is exact
    if (valid && role == 0) {
        pr.err[pi] = want ? __fdiv_rn(__fmul_rn((float)tot, 1.f), (float)(32 * WIN * WIN)) : 0.f;
        pr.status[pi] = (uint8_t)status;
    }
}

int sm_count_of_current_device() {
    static int cached[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

}  // namespace

void launch_lk10q_templates(const PyramidView& a, const float* pts, const int* n_pts, int cap, const LKParams& p,
                            uint4* words, cudaStream_t s) {
    const int per_block = LK_WARPS * PTS_PER_WARP;
    const int nlev = std::min(a.levels, p.max_level + 1);
    dim3 grid((cap + per_block - 1) / per_block, nlev);
    lk10q_template_kernel<<<grid, LK_WARPS * 32, 0, s>>>(a, pts, n_pts, cap, p.max_level + 1, words);
}

void launch_lk10q(const LKBatch& batch, const LKParams& p, cudaStream_t s) {
    const int per_block = LK_WARPS * PTS_PER_WARP;
    const long long items = (long long)batch.cap * batch.num_pairs;
    const int resident = sm_count_of_current_device() * 4;
    long long blocks = (items + per_block - 1) / per_block;
    if (batch.queue_budget > 0) {
        // blocks leave after `queue_budget` items each, so that blocks of other streams get their slots;
        // enough blocks for every item plus one resident wave that may find the queue empty
        blocks = (items + batch.queue_budget - 1) / batch.queue_budget + resident;
    } else {
        blocks = std::min<long long>(blocks, resident);
    }
    cudaMemsetAsync(batch.queue, 0, sizeof(int), s);
    lk10q_kernel<<<(unsigned)std::max<long long>(blocks, 1), LK_WARPS * 32, 0, s>>>(batch, p);
    dim3 egrid((batch.cap + per_block - 1) / per_block, batch.num_pairs);
    lk10q_err_kernel<<<egrid, LK_WARPS * 32, 0, s>>>(batch);
}

}  // namespace pc
