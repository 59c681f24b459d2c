"""Stage 1 -- exhaustive exact top-k L2 retrieval on B200, behind the reference's
own adaptor interface.

`InvKeyTree` mirrors `KDTreeVectorOfVectorsAdaptor<KeyMat, float>`
(/root/reference/registration/KDTreeVectorOfVectorsAdaptor.h:52-102, aliased at
registration/loop_detector.h:31-32): same constructor arguments, same `query`
argument meaning, same error behaviour (empty data / dimension mismatch raise).
`KnnIndex` is the thin handle over the C ABI (batch queries, device buffers,
sharding); both call libgloc3d.so -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KnnStats, check


class KnnIndex:
    """One GPU's shard of the descriptor database (C ABI: gloc_knn_*)."""

    def __init__(self, dim: int, device: int = 0):
        self._h = C.c_void_p()
        check(_lib.lib().gloc_knn_create(C.byref(self._h), dim, device))
        self.dim = dim
        self.device = device

    # -- database ---------------------------------------------------------
    def set_db(self, rows: np.ndarray) -> None:
        rows = self._rows(rows)
        check(_lib.lib().gloc_knn_set_db(self._h, rows.ctypes.data, rows.shape[0]))

    def set_db_device(self, rows) -> None:
        """rows: CUDA torch.Tensor float32 [n, dim] on this index's device."""
        assert rows.is_cuda and rows.is_contiguous() and rows.dtype.is_floating_point
        assert rows.element_size() == 4 and rows.shape[1] == self.dim
        check(_lib.lib().gloc_knn_set_db_device(self._h, rows.data_ptr(), rows.shape[0]))

    def append(self, rows: np.ndarray) -> None:
        rows = self._rows(rows)
        check(_lib.lib().gloc_knn_append(self._h, rows.ctypes.data, rows.shape[0]))

    def __len__(self) -> int:
        return int(_lib.lib().gloc_knn_size(self._h))

    def set_search_limit(self, n_search: int | None) -> None:
        v = (1 << 64) - 1 if n_search is None else int(n_search)
        check(_lib.lib().gloc_knn_set_search_limit(self._h, v))

    def set_index_offset(self, offset: int) -> None:
        check(_lib.lib().gloc_knn_set_index_offset(self._h, int(offset)))

    def set_mode(self, mode: int) -> None:
        check(_lib.lib().gloc_knn_set_mode(self._h, int(mode)))

    def stats(self) -> KnnStats:
        s = KnnStats()
        check(_lib.lib().gloc_knn_get_stats(self._h, C.byref(s)))
        return s

    def set_profiling(self, enabled: bool) -> None:
        check(_lib.lib().gloc_knn_set_profiling(self._h, int(enabled)))

    def profile(self):
        """(summed ms, launches) of the dominant kernel since the last call."""
        p = _lib.Profile()
        check(_lib.lib().gloc_knn_get_profile(self._h, C.byref(p)))
        return p.dominant_ms, int(p.dominant_launches)

    # -- queries ----------------------------------------------------------
    def query(self, q: np.ndarray, k: int, out_idx: np.ndarray | None = None,
              out_d2: np.ndarray | None = None):
        """Host buffers in, host buffers out (H2D/D2H inside the call)."""
        q = np.ascontiguousarray(q, np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise ValueError(f"query dim {q.shape[1]} != index dim {self.dim}")
        nq = q.shape[0]
        if out_idx is None:
            out_idx = np.empty((nq, k), np.uint64)
        if out_d2 is None:
            out_d2 = np.empty((nq, k), np.float32)
        assert out_idx.dtype == np.uint64 and out_d2.dtype == np.float32
        assert out_idx.size >= nq * k and out_d2.size >= nq * k
        check(_lib.lib().gloc_knn_query(self._h, q.ctypes.data, nq, k, out_idx.ctypes.data,
                                        out_d2.ctypes.data))
        return out_idx, out_d2

    def query_ptr(self, q_ptr: int, nq: int, k: int, idx_ptr: int, d2_ptr: int) -> None:
        """Raw host pointers (e.g. pinned torch tensors): the plain C call."""
        check(_lib.lib().gloc_knn_query(self._h, q_ptr, nq, k, idx_ptr, d2_ptr))

    def query_device(self, q, k: int, out_idx=None, out_d2=None, stream: int | None = None):
        """q: CUDA torch.Tensor float32 [nq, dim].  Returns CUDA tensors (idx as int64
        holding the uint64 bit pattern).  Enqueued on the current torch stream."""
        import torch

        assert q.is_cuda and q.is_contiguous() and q.dtype == torch.float32
        nq = q.shape[0]
        if out_idx is None:
            out_idx = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        if out_d2 is None:
            out_d2 = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        if stream is None:
            stream = torch.cuda.current_stream(q.device).cuda_stream
        check(_lib.lib().gloc_knn_query_device(self._h, q.data_ptr(), nq, k, out_idx.data_ptr(),
                                               out_d2.data_ptr(), stream))
        return out_idx, out_d2

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().gloc_knn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _rows(self, rows) -> np.ndarray:
        rows = np.ascontiguousarray(rows, np.float32)
        if rows.ndim == 1:
            rows = rows.reshape(1, -1) if rows.size else rows.reshape(0, self.dim)
        if rows.shape[0] and rows.shape[1] != self.dim:
            raise ValueError(f"row dim {rows.shape[1]} != index dim {self.dim}")
        return rows


def merge_topk_device(idx, d2, out_idx=None, out_d2=None):
    """K4: idx/d2 CUDA tensors [g, nq, k] of per-shard lists with global indices."""
    import torch

    g, nq, k = idx.shape
    assert idx.is_contiguous() and d2.is_contiguous() and idx.dtype == torch.int64
    if out_idx is None:
        out_idx = torch.empty((nq, k), dtype=torch.int64, device=idx.device)
    if out_d2 is None:
        out_d2 = torch.empty((nq, k), dtype=torch.float32, device=idx.device)
    stream = torch.cuda.current_stream(idx.device).cuda_stream
    check(_lib.lib().gloc_knn_merge_topk_device(idx.data_ptr(), d2.data_ptr(), g, nq, k,
                                                out_idx.data_ptr(), out_d2.data_ptr(),
                                                idx.device.index or 0, stream))
    return out_idx, out_d2


class InvKeyTree:
    """Drop-in for the reference's `InvKeyTree`
    (= KDTreeVectorOfVectorsAdaptor<std::vector<std::vector<float>>, float>).

        tree = InvKeyTree(512, db_features, 10)       # loop_detector.cpp:36
        tree.query(feat, 20, ret_indexes, out_dists)  # loop_detector.cpp:45

    `leaf_max_size` is accepted and ignored (the GPU search is exhaustive; no tree).
    """

    def __init__(self, dimensionality: int, mat, leaf_max_size: int = 10, device: int = 0):
        mat = np.ascontiguousarray(mat, np.float32)
        # assert(mat.size() != 0 && mat[0].size() != 0)  KDTreeVectorOfVectorsAdaptor.h:75
        if mat.ndim != 2 or mat.shape[0] == 0 or mat.shape[1] == 0:
            raise AssertionError("mat.size() != 0 && mat[0].size() != 0")
        del dimensionality, leaf_max_size  # the adaptor ignores the first one too (:71,76)
        self.m_data = mat
        self.index = KnnIndex(mat.shape[1], device)
        self.index.set_db(mat)

    def kdtree_get_point_count(self) -> int:
        return len(self.index)

    def query(self, query_point, num_closest: int, out_indices=None, out_distances_sq=None):
        q = np.ascontiguousarray(query_point, np.float32).reshape(1, -1)
        idx, d2 = self.index.query(q, num_closest)
        if out_indices is not None:
            out_indices[:num_closest] = idx[0]
        if out_distances_sq is not None:
            out_distances_sq[:num_closest] = d2[0]
        return idx[0], d2[0]
