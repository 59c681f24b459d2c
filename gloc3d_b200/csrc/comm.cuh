// comm.cuh -- the library's communicator (internal).  See comm.cu.
#pragma once
#include "common.cuh"

struct gloc_comm {
  void* nccl = nullptr;   // ncclComm_t
  int rank = 0, size = 1, device = 0;
};

namespace gloc {
// collectives on byte buffers, enqueued on `s` (device memory of the communicator's device)
int comm_all_gather(gloc_comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
int comm_all_reduce_max_u64(gloc_comm* c, const void* send, void* recv, size_t count, cudaStream_t s);
int comm_all_to_all(gloc_comm* c, const void* send, void* recv, size_t bytes_per_block, cudaStream_t s);
}  // namespace gloc
