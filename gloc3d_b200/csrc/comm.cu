// comm.cu -- multi-GPU plumbing behind the C ABI: NCCL communicators owned by libgloc3d.so, so
// that a C++ host (the reference is a single C++ process) can shard the database over the GPUs of
// one box without Python.  The reference has no distributed code at all; this is
// the B200 design for its north-star scaling configs.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): a process that already carries an NCCL
// (PyTorch bundles its own) shares that copy, a plain C++ host gets the system library, and
// single-GPU users never load it.
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>   // types and prototypes only; every call goes through the table below

#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "comm.cuh"

namespace gloc {

namespace {

struct NcclApi {
  bool ok = false;
  std::string why;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

NcclApi& api() {
  static NcclApi A;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      A.why = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    bool all = true;
    auto sym = [&](const char* name) {
      void* p = dlsym(h, name);
      if (!p) {
        all = false;
        A.why = std::string("libnccl lacks ") + name;
      }
      return p;
    };
    A.GetVersion = (decltype(A.GetVersion))sym("ncclGetVersion");
    A.GetUniqueId = (decltype(A.GetUniqueId))sym("ncclGetUniqueId");
    A.CommInitRank = (decltype(A.CommInitRank))sym("ncclCommInitRank");
    A.CommInitAll = (decltype(A.CommInitAll))sym("ncclCommInitAll");
    A.CommDestroy = (decltype(A.CommDestroy))sym("ncclCommDestroy");
    A.GetErrorString = (decltype(A.GetErrorString))sym("ncclGetErrorString");
    A.AllGather = (decltype(A.AllGather))sym("ncclAllGather");
    A.AllReduce = (decltype(A.AllReduce))sym("ncclAllReduce");
    A.Send = (decltype(A.Send))sym("ncclSend");
    A.Recv = (decltype(A.Recv))sym("ncclRecv");
    A.Broadcast = (decltype(A.Broadcast))sym("ncclBroadcast");
    A.GroupStart = (decltype(A.GroupStart))sym("ncclGroupStart");
    A.GroupEnd = (decltype(A.GroupEnd))sym("ncclGroupEnd");
    A.ok = all;
  });
  return A;
}

int nccl_fail(const char* what, ncclResult_t r) {
  return fail(GLOC_ERR_CUDA, std::string(what) + ": " + (api().GetErrorString ? api().GetErrorString(r) : "NCCL error"));
}

#define GLOC_NCCL_TRY(expr)                                   \
  do {                                                        \
    ncclResult_t _r = (expr);                                 \
    if (_r != ncclSuccess) return nccl_fail(#expr, _r);       \
  } while (0)

}  // namespace

int comm_all_gather(gloc_comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s) {
  GLOC_NCCL_TRY(api().AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t)c->nccl, s));
  return GLOC_OK;
}

int comm_all_reduce_max_u64(gloc_comm* c, const void* send, void* recv, size_t count, cudaStream_t s) {
  GLOC_NCCL_TRY(api().AllReduce(send, recv, count, ncclUint64, ncclMax, (ncclComm_t)c->nccl, s));
  return GLOC_OK;
}

// In place: bytes [offsets[r], offsets[r + 1]) of `buf` are valid on rank r; afterwards all of
// [offsets[0], offsets[size]) is valid everywhere (an all-gather of unequal parts: one grouped
// broadcast per rank, over NVLink).
int comm_all_gather_v(gloc_comm* c, void* buf, const size_t* offsets, cudaStream_t s) {
  GLOC_NCCL_TRY(api().GroupStart());
  for (int r = 0; r < c->size; ++r) {
    const size_t n = offsets[r + 1] - offsets[r];
    if (n == 0) continue;
    char* p = (char*)buf + offsets[r];
    ncclResult_t a = api().Broadcast(p, p, n, ncclUint8, r, (ncclComm_t)c->nccl, s);
    if (a != ncclSuccess) {
      api().GroupEnd();
      return nccl_fail("ncclBroadcast", a);
    }
  }
  GLOC_NCCL_TRY(api().GroupEnd());
  return GLOC_OK;
}

// rank r receives block r of every rank's `send` ([size][bytes_per_block]) into recv[src]
int comm_all_to_all(gloc_comm* c, const void* send, void* recv, size_t bytes_per_block, cudaStream_t s) {
  GLOC_NCCL_TRY(api().GroupStart());
  for (int r = 0; r < c->size; ++r) {
    ncclResult_t a = api().Send((const char*)send + (size_t)r * bytes_per_block, bytes_per_block, ncclUint8, r,
                                (ncclComm_t)c->nccl, s);
    ncclResult_t b = api().Recv((char*)recv + (size_t)r * bytes_per_block, bytes_per_block, ncclUint8, r,
                                (ncclComm_t)c->nccl, s);
    if (a != ncclSuccess || b != ncclSuccess) {
      api().GroupEnd();
      return nccl_fail("ncclSend/ncclRecv", a != ncclSuccess ? a : b);
    }
  }
  GLOC_NCCL_TRY(api().GroupEnd());
  return GLOC_OK;
}

namespace {

struct PeerInfo {
  unsigned long long pid, ptr;
  int device, pad;
  cudaIpcMemHandle_t handle;
};

// every rank's `bytes` bytes of host data to every rank (through a small device buffer and NCCL)
int host_all_gather(gloc_comm* c, const void* send, void* recv, size_t bytes) {
  const size_t need = bytes * (size_t)(c->size + 1);
  if (need > 65536) return fail(GLOC_ERR_RANGE, "comm: host exchange too large");
  if (!c->d_scratch) GLOC_CUDA_TRY(cudaMalloc(&c->d_scratch, 65536));
  if (!c->xstream) GLOC_CUDA_TRY(cudaStreamCreateWithFlags(&c->xstream, cudaStreamNonBlocking));
  char* d = (char*)c->d_scratch;
  GLOC_CUDA_TRY(cudaMemcpyAsync(d, send, bytes, cudaMemcpyHostToDevice, c->xstream));
  GLOC_NCCL_TRY(api().AllGather(d, d + bytes, bytes, ncclUint8, (ncclComm_t)c->nccl, c->xstream));
  GLOC_CUDA_TRY(cudaMemcpyAsync(recv, d + bytes, bytes * (size_t)c->size, cudaMemcpyDeviceToHost, c->xstream));
  GLOC_CUDA_TRY(cudaStreamSynchronize(c->xstream));
  return GLOC_OK;
}

}  // namespace

int comm_host_all_gather(gloc_comm* c, const void* send, void* recv, size_t bytes) {
  if (bytes == 0) return GLOC_OK;
  if (bytes * (size_t)(c->size + 1) <= 65536) return host_all_gather(c, send, recv, bytes);
  if (!c->xstream) GLOC_CUDA_TRY(cudaStreamCreateWithFlags(&c->xstream, cudaStreamNonBlocking));
  char* d = nullptr;
  GLOC_CUDA_TRY(cudaMalloc((void**)&d, bytes * (size_t)(c->size + 1)));
  cudaError_t e = cudaMemcpyAsync(d, send, bytes, cudaMemcpyHostToDevice, c->xstream);
  ncclResult_t ne = ncclSuccess;
  if (e == cudaSuccess) ne = api().AllGather(d, d + bytes, bytes, ncclUint8, (ncclComm_t)c->nccl, c->xstream);
  if (e == cudaSuccess && ne == ncclSuccess)
    e = cudaMemcpyAsync(recv, d + bytes, bytes * (size_t)c->size, cudaMemcpyDeviceToHost, c->xstream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->xstream);
  cudaFree(d);
  if (ne != ncclSuccess) return nccl_fail("comm_host_all_gather: ncclAllGather", ne);
  GLOC_CUDA_TRY(e);
  return GLOC_OK;
}

int comm_map_peers(gloc_comm* c, void* local, void*** out) {
  for (auto& m : c->maps)
    if (m.local == local) {
      *out = m.ptrs.data();
      return GLOC_OK;
    }
  PeerInfo mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.pid = (unsigned long long)getpid();
  mine.ptr = (unsigned long long)(uintptr_t)local;
  mine.device = c->device;
  cudaError_t he = cudaIpcGetMemHandle(&mine.handle, local);
  if (he != cudaSuccess) (void)cudaGetLastError();   // still exchanged: peers in this process do not need it
  std::vector<PeerInfo> all((size_t)c->size);
  int rc = host_all_gather(c, &mine, all.data(), sizeof(PeerInfo));
  if (rc != GLOC_OK) return rc;
  gloc_peer_map m;
  m.local = local;
  m.ptrs.assign((size_t)c->size, nullptr);
  m.opened.assign((size_t)c->size, 0);
  int bad = 0;
  for (int i = 0; i < c->size; ++i) {
    if (i == c->rank) {
      m.ptrs[i] = local;
    } else if (all[i].pid == mine.pid) {            // same process: the pointer itself, once peer access is on
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, c->device, all[i].device) != cudaSuccess || !can) {
        bad = 1;
        continue;
      }
      cudaError_t e = cudaDeviceEnablePeerAccess(all[i].device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) bad = 1;
      (void)cudaGetLastError();
      m.ptrs[i] = (void*)(uintptr_t)all[i].ptr;
    } else {
      void* p = nullptr;
      cudaError_t e = he == cudaSuccess ? cudaIpcOpenMemHandle(&p, all[i].handle, cudaIpcMemLazyEnablePeerAccess)
                                        : he;
      if (e != cudaSuccess) {
        (void)cudaGetLastError();
        bad = 1;
        continue;
      }
      m.ptrs[i] = p;
      m.opened[i] = 1;
    }
  }
  // agree: either every rank mapped every peer, or nobody uses peer memory
  std::vector<int> flags((size_t)c->size);
  rc = host_all_gather(c, &bad, flags.data(), sizeof(int));
  if (rc != GLOC_OK) return rc;
  for (int f : flags) bad |= f;
  if (bad) {
    for (int i = 0; i < c->size; ++i)
      if (m.opened[i]) cudaIpcCloseMemHandle(m.ptrs[i]);
    return fail(GLOC_ERR_CUDA, "comm: the GPUs of this communicator cannot address each other's memory");
  }
  c->maps.push_back(std::move(m));
  *out = c->maps.back().ptrs.data();
  return GLOC_OK;
}

void comm_unmap_peers(gloc_comm* c, void* local) {
  for (size_t k = 0; k < c->maps.size(); ++k)
    if (c->maps[k].local == local) {
      for (int i = 0; i < c->size; ++i)
        if (c->maps[k].opened[i]) cudaIpcCloseMemHandle(c->maps[k].ptrs[i]);
      c->maps.erase(c->maps.begin() + k);
      return;
    }
}

}  // namespace gloc

using namespace gloc;

extern "C" {

int gloc_comm_unique_id(uint8_t* id, size_t capacity) {
  if (!id || capacity < GLOC_COMM_ID_BYTES) return fail(GLOC_ERR_INVALID, "gloc_comm_unique_id: need a 128-byte buffer");
  if (!api().ok) return fail(GLOC_ERR_CUDA, "gloc_comm_unique_id: " + api().why);
  static_assert(sizeof(ncclUniqueId) == GLOC_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  GLOC_NCCL_TRY(api().GetUniqueId(&u));
  std::memcpy(id, &u, sizeof(u));
  return GLOC_OK;
}

int gloc_comm_create(gloc_comm** out, const uint8_t* id, int n_ranks, int rank, int device) {
  if (!out) return fail(GLOC_ERR_INVALID, "gloc_comm_create: out is null");
  *out = nullptr;
  if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(GLOC_ERR_INVALID, "gloc_comm_create: bad argument");
  if (!api().ok) return fail(GLOC_ERR_CUDA, "gloc_comm_create: " + api().why);
  DeviceGuard g(device);
  if (!g.ok) return fail(GLOC_ERR_CUDA, "gloc_comm_create: cudaSetDevice failed");
  gloc_comm* c = new (std::nothrow) gloc_comm;
  if (!c) return fail(GLOC_ERR_NOMEM, "gloc_comm_create: out of host memory");
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  ncclResult_t r = api().CommInitRank(&comm, n_ranks, u, rank);
  if (r != ncclSuccess) {
    delete c;
    return nccl_fail("ncclCommInitRank", r);
  }
  c->nccl = comm;
  c->rank = rank;
  c->size = n_ranks;
  c->device = device;
  *out = c;
  return GLOC_OK;
}

int gloc_comm_create_local(gloc_comm** out, int n_devices, const int* devices) {
  if (!out || n_devices < 1) return fail(GLOC_ERR_INVALID, "gloc_comm_create_local: bad argument");
  for (int i = 0; i < n_devices; ++i) out[i] = nullptr;
  if (!api().ok) return fail(GLOC_ERR_CUDA, "gloc_comm_create_local: " + api().why);
  std::vector<int> devs((size_t)n_devices);
  for (int i = 0; i < n_devices; ++i) devs[i] = devices ? devices[i] : i;
  std::vector<ncclComm_t> comms((size_t)n_devices, nullptr);
  GLOC_NCCL_TRY(api().CommInitAll(comms.data(), n_devices, devs.data()));
  for (int i = 0; i < n_devices; ++i) {
    gloc_comm* c = new (std::nothrow) gloc_comm;
    if (!c) return fail(GLOC_ERR_NOMEM, "gloc_comm_create_local: out of host memory");
    c->nccl = comms[i];
    c->rank = i;
    c->size = n_devices;
    c->device = devs[i];
    out[i] = c;
  }
  return GLOC_OK;
}

void gloc_comm_destroy(gloc_comm* c) {
  if (!c) return;
  {
    DeviceGuard g(c->device);
    while (!c->maps.empty()) comm_unmap_peers(c, c->maps.back().local);
    if (c->d_scratch) cudaFree(c->d_scratch);
    if (c->xstream) cudaStreamDestroy(c->xstream);
    if (c->nccl && api().ok) api().CommDestroy((ncclComm_t)c->nccl);
  }
  delete c;
}

int gloc_comm_rank(const gloc_comm* c) { return c ? c->rank : -1; }
int gloc_comm_size(const gloc_comm* c) { return c ? c->size : 0; }

int gloc_comm_nccl_version(void) {
  int v = 0;
  if (!api().ok || api().GetVersion(&v) != ncclSuccess) return 0;
  return v;
}

}  // extern "C"
