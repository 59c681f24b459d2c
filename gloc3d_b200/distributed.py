"""Multi-GPU retrieval: the database shards by rows across the ranks of one box (one
process per GPU, torch.distributed / NCCL over NVLink for the plumbing).

Every rank holds rows [lo, hi) of the database, answers all queries against its shard
(local top-k with GLOBAL indices), the per-rank lists are exchanged with ONE all-gather of
nq*k*(8+4) bytes per rank, and every rank merges them with K4 (`gloc_knn_merge_topk_device`)
-- same (d2, idx) order, so the result is identical on every rank and identical to a
single-GPU search.  The reference has no distributed code at all (F1); this
is the B200 design for its north-star scaling config.

The host logic (shard bounds, gather layout, merge) is backend-agnostic so that it is
testable on CPU with gloo by injecting a local-search and a merge function.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable

import numpy as np


def shard_bounds(n_rows: int, world_size: int) -> list[int]:
    """Contiguous, balanced row ranges: rank r owns [b[r], b[r+1])."""
    return [n_rows * r // world_size for r in range(world_size + 1)]


class ShardedRetrieval:
    """Row-sharded exact top-k over `world_size` ranks.

    local_search(q, k) -> (idx, d2) for this rank's shard with global indices;
    merge(idx[g, nq, k], d2[g, nq, k]) -> (idx[nq, k], d2[nq, k]).
    With the defaults both run on the GPU through libgloc3d.so.
    """

    def __init__(self, rank: int, world_size: int, group=None,
                 local_search: Callable | None = None, merge: Callable | None = None):
        self.rank, self.world_size, self.group = rank, world_size, group
        self._local_search = local_search
        self._merge = merge
        self.index = None
        self.lo = self.hi = 0

    # -- GPU-backed construction ------------------------------------------
    def load_shard(self, shard_rows, lo: int, device: int, mode: int = 0):
        """shard_rows: this rank's rows [lo, lo+len) as numpy (host) or CUDA tensor."""
        from .retrieval import KnnIndex, merge_topk_device

        dim = shard_rows.shape[1]
        self.index = KnnIndex(dim, device)
        if isinstance(shard_rows, np.ndarray):
            self.index.set_db(shard_rows)
        else:
            self.index.set_db_device(shard_rows)
        self.index.set_index_offset(lo)
        self.index.set_mode(mode)
        self.lo, self.hi = lo, lo + shard_rows.shape[0]
        self._local_search = lambda q, k: self.index.query_device(q, k)
        self._merge = merge_topk_device
        return self

    # -- query --------------------------------------------------------------
    def query(self, q, k: int):
        """q: the full query batch, replicated on every rank (CUDA tensor on the GPU path,
        CPU tensor under gloo).  Returns the global top-k on every rank."""
        import torch
        import torch.distributed as dist

        idx, d2 = self._local_search(q, k)
        if self.world_size == 1:
            return idx, d2
        nq = idx.shape[0]
        all_idx = torch.empty((self.world_size, nq, k), dtype=idx.dtype, device=idx.device)
        all_d2 = torch.empty((self.world_size, nq, k), dtype=d2.dtype, device=d2.device)
        if idx.is_cuda:   # NCCL: one fused gather per array, straight into the merge layout
            dist.all_gather_into_tensor(all_idx, idx.contiguous(), group=self.group)
            dist.all_gather_into_tensor(all_d2, d2.contiguous(), group=self.group)
        else:             # gloo (CPU tests of the host logic)
            dist.all_gather(list(all_idx.unbind(0)), idx.contiguous(), group=self.group)
            dist.all_gather(list(all_d2.unbind(0)), d2.contiguous(), group=self.group)
        return self._merge(all_idx, all_d2)

    def query_host(self, q_pinned, k: int, out_idx_pinned=None, out_d2_pinned=None):
        """End-to-end call with HOST buffers: H2D of the queries, sharded search, all-gather,
        merge, D2H of the result (pinned torch CPU tensors in and out)."""
        import torch

        dev = torch.device("cuda", self.index.device)
        q = q_pinned.to(dev, non_blocking=True)
        idx, d2 = self.query(q, k)
        if out_idx_pinned is None:
            out_idx_pinned = torch.empty(idx.shape, dtype=idx.dtype, pin_memory=True)
            out_d2_pinned = torch.empty(d2.shape, dtype=d2.dtype, pin_memory=True)
        out_idx_pinned.copy_(idx, non_blocking=True)
        out_d2_pinned.copy_(d2, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out_idx_pinned, out_d2_pinned

    def close(self):
        if self.index is not None:
            self.index.close()
            self.index = None


class Comm:
    """`gloc_comm`: an NCCL communicator owned by libgloc3d.so (the multi-GPU entry points of the C
    ABI take it, so a C++ host needs no Python).  Under torchrun the 128-byte NCCL id travels from
    rank 0 to the other ranks through torch.distributed -- the only thing the process group is used
    for; every data-path collective runs inside the library."""

    def __init__(self, handle, rank: int, size: int):
        self._h, self.rank, self.size = handle, rank, size

    @classmethod
    def from_torch(cls, device: int, group=None) -> "Comm":
        import torch
        import torch.distributed as dist

        from . import _lib

        rank, size = dist.get_rank(group), dist.get_world_size(group)
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            _lib.check(_lib.lib().gloc_comm_unique_id(buf, 128))
        t = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = t.cuda(device)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        _lib.check(_lib.lib().gloc_comm_create(C.byref(h), ident, size, rank, device))
        return cls(h, rank, size)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            from . import _lib

            _lib.lib().gloc_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def query_sharded_device(index, comm: Comm, q, k: int, replicated: bool = False, out_idx=None, out_d2=None):
    """Row-sharded top-k through the C ABI (`gloc_knn_query_sharded_device`): q is this rank's slice
    of the batch (replicated=False) or the whole batch on every rank (replicated=True)."""
    import torch

    from . import _lib

    nq = q.shape[0]
    if out_idx is None:
        out_idx = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        out_d2 = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    stream = torch.cuda.current_stream(q.device).cuda_stream
    _lib.check(_lib.lib().gloc_knn_query_sharded_device(index._h, comm._h, q.data_ptr(), nq, k, out_idx.data_ptr(),
                                                        out_d2.data_ptr(), int(replicated), stream))
    return out_idx, out_d2


def query_sharded_host(index, comm: Comm, q_ptr: int, nq: int, k: int, idx_ptr: int, d2_ptr: int,
                       replicated: bool = False) -> None:
    """The same with host buffers (pinned tensors' pointers): H2D and D2H inside the call."""
    from . import _lib

    _lib.check(_lib.lib().gloc_knn_query_sharded(index._h, comm._h, q_ptr, nq, k, idx_ptr, d2_ptr, int(replicated)))


def partition_pairs(idx: np.ndarray, lo: int, hi: int):
    """Which (query, candidate) pairs a rank verifies in the sharded localizer: those whose
    retrieved row it owns.  Returns flat positions into the [nq, k] result (host logic mirrored
    from loc_api.cu, testable without a GPU)."""
    flat = np.asarray(idx).reshape(-1)
    return np.nonzero((flat >= lo) & (flat < hi))[0]


def assign_pairs(idx: np.ndarray, bounds) -> np.ndarray:
    """Which rank verifies which (query, candidate) pair once the ranks share their grid stores
    (gloc_loc_share_grids): `bounds` = the row ranges [bounds[r], bounds[r + 1]) of the ranks.  Calls the
    library's own rule (gloc_loc_assign_pairs, host only -- no GPU needed); -1 = a slot without a row."""
    from . import _lib
    flat = np.ascontiguousarray(np.asarray(idx).reshape(-1).astype(np.int64))
    b = np.asarray(bounds, np.int64)
    owner = (np.searchsorted(b, flat, side="right") - 1).astype(np.int32)
    owner[(flat < b[0]) | (flat >= b[-1])] = -1
    out = np.empty(flat.size, np.int32)
    _lib.check(_lib.lib().gloc_loc_assign_pairs(owner.ctypes.data, flat.size, len(b) - 1, out.ctypes.data))
    return out


def combine_pair_keys(keys_per_rank: list[np.ndarray]) -> np.ndarray:
    """The all-reduce(max) of the sharded localizer: every pair has one owner, the others hold 0."""
    return np.maximum.reduce([np.asarray(k, np.uint64) for k in keys_per_rank])
