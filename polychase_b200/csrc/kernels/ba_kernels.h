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
};

void launch_ba_refresh_points(const BAView& v, const MeshView& mesh, cudaStream_t s);
void launch_ba_cost(const BAView& v, const Loss& loss, cudaStream_t s);
void launch_ba_build(const BAView& v, const MeshView& mesh, const Loss& loss, cudaStream_t s);
void launch_ba_assemble(const BAView& v, cudaStream_t s);
void launch_ba_solve(const BAView& v, float lambda, cudaStream_t s);     // banded LLT + solve, step = -x
void launch_ba_expected_change(const BAView& v, cudaStream_t s);         // step^T (2 Jtr + A step)

}  // namespace pc
