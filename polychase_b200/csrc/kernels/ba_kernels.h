// Launcher interface of the refine (bundle adjustment) kernels K12-K14.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "track_kernels.h"

namespace pc {

constexpr int kBandBlocks = 9;          // block (i, i-k), k = 0..8: flows exist for |d| in {1,2,4,8}
constexpr uint32_t kInvalidPrim = 0xFFFFFFFFu;

struct BAView {
    int nf, p;                          // frames, params per camera (6 or 9)
    int n_kps, n_edges, n_rows;
    int opt_f, opt_pp;
    float M[16], Minv[16];              // model matrix and its inverse, row-major
    const float* kps;                   // n_kps x 2 (all frames concatenated)
    const int* kp_frame;                // n_kps
    const int* kp_offsets;              // nf + 1
    const uint8_t* referenced;          // n_kps: keypoint appears in some edge
    uint32_t* cache;                    // n_kps: cached primitive id (refiner.cc:547-559)
    float* pts;                         // n_kps x 3 world points of the current evaluation
    uint8_t* pt_valid;                  // n_kps
    const pc_ba_edge* edges;            // n_edges
    const float* edge_weight;           // n_edges
    const uint32_t* src_idx;            // n_rows
    const float* tgt;                   // n_rows x 2
    const pc_camera_state* cams;        // nf
    float* edge_cost;                   // n_edges
    float* edge_pair;                   // n_edges x pair_stride
    int pair_stride;                    // lower triangle of (2p x 2p) + 2p + 1 (valid count)
    // assembly
    const int* inc_offsets;             // nf + 1
    const int* inc_edges;               // incident edge ids per frame, ascending
    float* band;                        // nf x 9 x p x p
    float* jtr;                         // nf x p
    float* diag;                        // nf x p (clamped)
    float* lband;                       // factor workspace, same shape as band
    float* step;                        // nf x p
    float* tmp;                         // nf x p
    float* scalars;                     // [0] cost [1] grad_norm [2] step_norm [3] expected_change [4] llt_ok
    // edge-sharded refine (multi-GPU): this rank evaluates the edges whose mask byte is non-zero (nullptr: all).
    // band and jtr are contiguous (one allocation: band, then jtr) so that one all-reduce covers both.
    const uint8_t* edge_mask;
};

// Device-resident state of LevMarqSparseSolver::Solve (lev_marq.h:492-588); see ba_lm.cu.
struct BALmState {
    // options (constant during a solve)
    float gradient_tol, step_tol, min_lambda, max_lambda;
    unsigned long long max_iterations;
    // loop state
    float cost, lambda, v;
    int rebuild;                        // the next iteration rebuilds the normal equations
    int done;                           // the loop has ended (a `break`, or max_iterations)
    int skip;                           // this iteration's factorisation failed: the evaluation kernels do nothing
    int llt_ok;
    unsigned long long iterations, invalid_steps;
    float initial_cost, step_norm, grad_norm, cost_new, expected;
    // the stats the reference passes to the iteration callback after this iteration (valid iff snap_valid)
    pc_bundle_stats snap;
    int snap_valid;
};

// `gate` (nullable): the kernels of a stage return at once unless the LM state says the reference's loop would run
// that stage now -- GATE_BUILD: !done && rebuild;  GATE_EVAL: !done && !skip.
enum BAGate { GATE_NONE = 0, GATE_BUILD = 1, GATE_EVAL = 2 };
void launch_ba_refresh_points(const BAView& v, const MeshView& mesh, const BALmState* st, int gate, cudaStream_t s);
// cost -> *cost_out (device).  edge_mask (nullable, n_edges bytes): only edges with a non-zero byte are evaluated
// (edge-sharded multi-GPU refine); their costs are written to v.edge_cost, the others are left at zero.
void launch_ba_cost(const BAView& v, const Loss& loss, const BALmState* st, int gate, float* cost_out, cudaStream_t s);
void launch_ba_cost_edges(const BAView& v, const Loss& loss, const BALmState* st, int gate, cudaStream_t s);   // per-edge costs only
void launch_ba_cost_sum(const BAView& v, const BALmState* st, int gate, float* cost_out, cudaStream_t s);      // sum in edge order
void launch_ba_build(const BAView& v, const MeshView& mesh, const Loss& loss, const BALmState* st, int gate, cudaStream_t s);
void launch_ba_assemble(const BAView& v, const BALmState* st, int gate, cudaStream_t s);
// banded LLT + solve, step = -x.  st == nullptr: plain solve with `lambda`; else the LM state supplies lambda and
// receives the outcome (grad / step tolerance breaks, failed factorisation).
void launch_ba_solve(const BAView& v, BALmState* st, float lambda, cudaStream_t s);
void launch_ba_step(const BAView& v, BALmState* st, const pc_camera_state* params, pc_camera_state* params_new,
                    const Bounds& bounds, float* expected_part, cudaStream_t s);
void launch_ba_decide(const BAView& v, BALmState* st, pc_camera_state* params, const pc_camera_state* params_new,
                      const float* expected_part, const float* cost_new, cudaStream_t s);

}  // namespace pc
