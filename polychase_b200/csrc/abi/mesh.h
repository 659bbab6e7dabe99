// Device-resident mesh (BVH + raw buffers) and the track path's scratch buffers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../kernels/track_kernels.h"

struct pc_ctx;

namespace pc {

struct MeshData {
    int nv = 0, nt = 0, num_nodes = 0;
    float4* d_nodes = nullptr;
    float4* d_tris4 = nullptr;
    float* d_verts = nullptr;
    uint32_t* d_tris = nullptr;
    uint32_t* d_mask = nullptr;
    float bbox_min[3] = {0, 0, 0}, bbox_max[3] = {0, 0, 0};
    // track scratch
    RaySource* d_srcs = nullptr;
    float* d_X = nullptr; size_t cap_X = 0;
    float* d_x = nullptr; size_t cap_x = 0;
    float* d_w = nullptr; size_t cap_w = 0;
    uint8_t* d_valid = nullptr; size_t cap_valid = 0;
    float* d_kps = nullptr; size_t cap_kps = 0;
    uint32_t* d_idx = nullptr; size_t cap_idx = 0;
    float* d_tgt = nullptr; size_t cap_tgt = 0;
    uint32_t* d_prim = nullptr; size_t cap_prim = 0;
    float* d_uv = nullptr; size_t cap_uv = 0;
    float* d_t = nullptr; size_t cap_t = 0;
    float* d_pos = nullptr; size_t cap_pos = 0;
    pc_camera_state* d_cam = nullptr;
    PnpResult* d_result = nullptr;
};

MeshView mesh_view(const MeshData* m);
int validate_bundle_opts(pc_ctx* c, const pc_bundle_opts* o);

}  // namespace pc
