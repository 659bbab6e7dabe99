// NCCL plumbing behind the C ABI's collectives (pc_comm_*, pc_traj_allgather, edge-sharded refine).
// libnccl.so.2 is opened at run time (dlopen): a single-GPU user never needs NCCL installed, and inside a
// torch process the already loaded copy is the one that gets used.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

struct pc_ctx;

namespace pc {

struct CommData;
void free_comm(CommData*);
int comm_world(const pc_ctx* c);      // 0 = no communicator
int comm_rank(const pc_ctx* c);
// All-gather of float chunks in place: rank r's chunk is buf[r * count_per_rank .. (r+1) * count_per_rank).
// A no-op without a communicator or with world == 1.
int comm_allgather_inplace(pc_ctx* c, float* buf, size_t count_per_rank, cudaStream_t s);

}  // namespace pc
