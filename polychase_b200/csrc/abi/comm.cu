// C ABI: the path's collectives over NCCL (NVLink / NVSwitch inside a node).
//
//   pc_traj_allgather   the one collective of the Analyze -> Track -> Refine path: stitches the per-GPU trajectory
//                       segments (packed pc_camera_state, 64 B per frame) before the global refine
//   comm_allgather_inplace   edge-sharded refine: per-edge normal-equation blocks / costs
//
// One process per GPU; rank 0 creates the id (pc_comm_unique_id), the launcher ships its 128 bytes to the other
// ranks (torch.distributed, MPI, a file: anything), every rank calls pc_comm_init.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "comm.h"
#include "context.h"

namespace pc {

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.error = std::string("libnccl.so.2 could not be opened: ") + (dlerror() ? dlerror() : "?");
            return;
        }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString)
            api.error = "libnccl.so.2 lacks one of ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllGather";
    });
    return &api;
}

int nccl_fail(pc_ctx* c, ncclResult_t r, const char* what) {
    NcclApi* a = nccl_api();
    return fail(c, PC_ERR_CUDA, std::string("NCCL error in ") + what + ": " + (a->GetErrorString ? a->GetErrorString(r) : "?"));
}

}  // namespace

struct CommData {
    ncclComm_t comm = nullptr;
    int world = 0, rank = 0;
    pc_camera_state* d_traj = nullptr;     // staging of pc_traj_allgather
    size_t traj_cap = 0;
};

void free_comm(CommData* m) {
    if (!m) return;
    if (m->comm && nccl_api()->CommDestroy) nccl_api()->CommDestroy(m->comm);
    cudaFree(m->d_traj);
    delete m;
}

int comm_world(const pc_ctx* c) { return c->comm ? c->comm->world : 0; }
int comm_rank(const pc_ctx* c) { return c->comm ? c->comm->rank : 0; }

int comm_allgather_inplace(pc_ctx* c, float* buf, size_t count_per_rank, cudaStream_t s) {
    CommData* m = c->comm;
    if (!m || m->world <= 1 || count_per_rank == 0) return PC_OK;
    const ncclResult_t r = nccl_api()->AllGather(buf + (size_t)m->rank * count_per_rank, buf, count_per_rank, ncclFloat,
                                                 m->comm, s);
    if (r != ncclSuccess) return nccl_fail(c, r, "ncclAllGather");
    return PC_OK;
}

}  // namespace pc

using namespace pc;

extern "C" {

int pc_comm_unique_id(uint8_t id_out[PC_COMM_ID_BYTES]) {
    NcclApi* a = nccl_api();
    if (!a->error.empty()) return fail(nullptr, PC_ERR_STATE, a->error);
    static_assert(sizeof(ncclUniqueId) <= PC_COMM_ID_BYTES, "ncclUniqueId must fit PC_COMM_ID_BYTES");
    ncclUniqueId id;
    const ncclResult_t r = a->GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(nullptr, r, "ncclGetUniqueId");
    memset(id_out, 0, PC_COMM_ID_BYTES);
    memcpy(id_out, &id, sizeof(id));
    return PC_OK;
}

int pc_comm_init(pc_ctx* c, int world, int rank, const uint8_t id_in[PC_COMM_ID_BYTES]) {
    PC_CUDA(c, cudaSetDevice(c->device));
    NcclApi* a = nccl_api();
    if (!a->error.empty()) return fail(c, PC_ERR_STATE, a->error);
    PC_CHECK(c, world >= 1 && rank >= 0 && rank < world && id_in != nullptr, "bad communicator arguments");
    if (c->comm) { free_comm(c->comm); c->comm = nullptr; }
    CommData* m = new CommData();
    m->world = world;
    m->rank = rank;
    ncclUniqueId id;
    memcpy(&id, id_in, sizeof(id));
    const ncclResult_t r = a->CommInitRank(&m->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete m;
        return nccl_fail(c, r, "ncclCommInitRank");
    }
    c->comm = m;
    return PC_OK;
}

int pc_comm_destroy(pc_ctx* c) {
    if (c->comm) {
        PC_CUDA(c, cudaSetDevice(c->device));
        PC_CUDA(c, cudaStreamSynchronize(c->compute));
        free_comm(c->comm);
        c->comm = nullptr;
    }
    return PC_OK;
}

int pc_traj_allgather(pc_ctx* c, const pc_camera_state* local, int n_local, const int* counts, pc_camera_state* all) {
    PC_CUDA(c, cudaSetDevice(c->device));
    CommData* m = c->comm;
    if (!m) return fail(c, PC_ERR_STATE, "pc_comm_init has not been called on this context");
    PC_CHECK(c, counts != nullptr && all != nullptr && n_local >= 0 && (n_local == 0 || local != nullptr), "bad arguments");
    PC_CHECK(c, counts[m->rank] == n_local, "counts[rank] must equal n_local");
    int pad = 0;
    for (int r = 0; r < m->world; r++) {
        PC_CHECK(c, counts[r] >= 0, "negative segment length");
        pad = std::max(pad, counts[r]);
    }
    if (pad == 0) return PC_OK;
    const size_t need = (size_t)pad * m->world;
    if (m->traj_cap < need) {
        cudaFree(m->d_traj);
        m->d_traj = nullptr;
        m->traj_cap = 0;
        PC_CUDA(c, cudaMalloc(&m->d_traj, need * sizeof(pc_camera_state)));
        m->traj_cap = need;
    }
    cudaStream_t s = c->compute;
    pc_camera_state* mine = m->d_traj + (size_t)m->rank * pad;
    PC_CUDA(c, cudaMemsetAsync(mine, 0, (size_t)pad * sizeof(pc_camera_state), s));
    if (n_local) PC_CUDA(c, cudaMemcpyAsync(mine, local, (size_t)n_local * sizeof(pc_camera_state), cudaMemcpyHostToDevice, s));
    int rc = comm_allgather_inplace(c, reinterpret_cast<float*>(m->d_traj), (size_t)pad * (sizeof(pc_camera_state) / sizeof(float)), s);
    if (rc) return rc;
    std::vector<pc_camera_state> host(need);
    PC_CUDA(c, cudaMemcpyAsync(host.data(), m->d_traj, need * sizeof(pc_camera_state), cudaMemcpyDeviceToHost, s));
    PC_CUDA(c, cudaStreamSynchronize(s));
    size_t o = 0;
    for (int r = 0; r < m->world; r++) {
        memcpy(all + o, host.data() + (size_t)r * pad, (size_t)counts[r] * sizeof(pc_camera_state));
        o += (size_t)counts[r];
    }
    return PC_OK;
}

}  // extern "C"
