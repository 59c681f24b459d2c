// comm.cuh -- the library's communicator (internal).  See comm.cu.
#pragma once
#include "common.cuh"

#include <vector>

// A local device buffer and where every rank's counterpart is mapped in this process (peer
// memory over NVLink: direct pointers inside one process, CUDA IPC between processes).
struct gloc_peer_map {
  void* local = nullptr;
  std::vector<void*> ptrs;       // [size]; ptrs[rank] == local
  std::vector<char> opened;      // ptrs[i] came from cudaIpcOpenMemHandle
};

struct gloc_comm {
  void* nccl = nullptr;   // ncclComm_t
  int rank = 0, size = 1, device = 0;
  void* d_scratch = nullptr;     // small device buffer for host-side exchanges
  cudaStream_t xstream = nullptr;
  std::vector<gloc_peer_map> maps;
};

namespace gloc {
// collectives on byte buffers, enqueued on `s` (device memory of the communicator's device)
int comm_all_gather(gloc_comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
int comm_all_reduce_max_u64(gloc_comm* c, const void* send, void* recv, size_t count, cudaStream_t s);
int comm_all_to_all(gloc_comm* c, const void* send, void* recv, size_t bytes_per_block, cudaStream_t s);
// in place: rank r holds bytes [offsets[r], offsets[r + 1]) of buf; afterwards every rank holds all of them
int comm_all_gather_v(gloc_comm* c, void* buf, const size_t* offsets, cudaStream_t s);
// Collective: every rank passes its own cudaMalloc'ed buffer; out[i] = rank i's buffer as this
// process can address it (GLOC_ERR_CUDA when peers cannot reach each other).  Cached per pointer.
int comm_map_peers(gloc_comm* c, void* local, void*** out);
void comm_unmap_peers(gloc_comm* c, void* local);
// Collective: every rank's `bytes` bytes of HOST data to every rank's host (recv holds size * bytes).
int comm_host_all_gather(gloc_comm* c, const void* send, void* recv, size_t bytes);
}  // namespace gloc
