"""The whole query path in one call -- descriptors + scans -> located frames and poses.

Mirrors the evaluation loop of the reference driver
(/root/reference/registration/global_localization.cpp:482-574: `detect_all_query` then
`global_registraion`, candidates in retrieval order, first match wins) with
FastCorrelativeScanMatcher2D as the verifier.  Everything runs in libgloc3d.so
(`gloc_loc_*`); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import (LOC_FIRST_MATCH, LOC_VERIFY_ALL, CsmResult, LocParams, LocResult, LocStats,  # noqa: F401
                   check)
from .retrieval import KnnIndex
from .scan_matching import CsmStore


@dataclass
class LocalizeOutput:
    idx: np.ndarray          # [nq, k] uint64, the retrieval result
    d2: np.ndarray           # [nq, k] float32
    candidates: list         # nq * k CsmResult (reserved == -1: not evaluated), or None
    results: list            # nq LocResult


class Localizer:
    """`gloc_localizer`: borrows a KnnIndex and a CsmStore on the same device."""

    def __init__(self, index: KnnIndex, store: CsmStore):
        self._h = C.c_void_p()
        self.index, self.store = index, store      # keep them alive
        check(_lib.lib().gloc_loc_create(C.byref(self._h), index._h, store._h))

    def set_row_grids(self, grid_of_row: np.ndarray | None) -> None:
        if grid_of_row is None:
            check(_lib.lib().gloc_loc_set_row_grids(self._h, None, 0))
            return
        m = np.ascontiguousarray(grid_of_row, np.int32)
        check(_lib.lib().gloc_loc_set_row_grids(self._h, m.ctypes.data, m.shape[0]))

    @staticmethod
    def params(k: int, n_lin: int, n_ang: int, ang_step: float, depth: int, min_score: float,
               policy: int = LOC_VERIFY_ALL) -> LocParams:
        return LocParams(k, n_lin, n_ang, ang_step, depth, min_score, policy)

    @staticmethod
    def pack_scans(scans: list[np.ndarray]):
        scans = [np.ascontiguousarray(s, np.float32).reshape(-1, 3) for s in scans]
        offs = np.zeros(len(scans) + 1, np.int64)
        offs[1:] = np.cumsum([s.shape[0] for s in scans])
        pts = np.ascontiguousarray(np.concatenate(scans, axis=0), np.float32)
        return pts, offs

    def localize(self, queries: np.ndarray, scans, params: LocParams, inits=None,
                 per_candidate: bool = True) -> LocalizeOutput:
        """Host buffers in, host buffers out.  scans: list of [P, 3] clouds (one per query) or a
        (pts, offsets) pair from pack_scans."""
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        pts, offs = scans if isinstance(scans, tuple) else self.pack_scans(scans)
        assert offs.shape[0] == nq + 1
        init = None if inits is None else np.ascontiguousarray(inits, np.float64).reshape(nq, 3)
        idx = np.empty((nq, params.k), np.uint64)
        d2 = np.empty((nq, params.k), np.float32)
        cand = (CsmResult * (nq * params.k))() if per_candidate else None
        res = (LocResult * nq)()
        check(_lib.lib().gloc_loc_localize(self._h, q.ctypes.data, nq, pts.ctypes.data, offs.ctypes.data,
                                           None if init is None else init.ctypes.data, C.byref(params),
                                           idx.ctypes.data, d2.ctypes.data,
                                           C.cast(cand, C.c_void_p) if cand is not None else None,
                                           C.cast(res, C.c_void_p)))
        return LocalizeOutput(idx, d2, list(cand) if cand is not None else None, list(res))

    def localize_ptr(self, q_ptr: int, nq: int, pts_ptr: int, offs: np.ndarray, params: LocParams,
                     idx_ptr: int, d2_ptr: int, res, cand=None, device: bool = False) -> None:
        """Raw pointers (pinned host tensors, or device tensors with device=True): the plain C call."""
        fn = _lib.lib().gloc_loc_localize_device if device else _lib.lib().gloc_loc_localize
        check(fn(self._h, q_ptr, nq, pts_ptr, offs.ctypes.data, None, C.byref(params), idx_ptr, d2_ptr,
                 C.cast(cand, C.c_void_p) if cand is not None else None, C.cast(res, C.c_void_p)))

    def localize_sharded(self, comm, queries: np.ndarray, scans, params: LocParams, inits=None,
                         per_candidate: bool = True) -> LocalizeOutput:
        """Collective: the same batch on every rank of a row-sharded database (host buffers)."""
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        pts, offs = scans if isinstance(scans, tuple) else self.pack_scans(scans)
        init = None if inits is None else np.ascontiguousarray(inits, np.float64).reshape(nq, 3)
        idx = np.empty((nq, params.k), np.uint64)
        d2 = np.empty((nq, params.k), np.float32)
        cand = (CsmResult * (nq * params.k))() if per_candidate else None
        res = (LocResult * nq)()
        check(_lib.lib().gloc_loc_localize_sharded(self._h, comm._h, q.ctypes.data, nq, pts.ctypes.data,
                                                   offs.ctypes.data, None if init is None else init.ctypes.data,
                                                   C.byref(params), idx.ctypes.data, d2.ctypes.data,
                                                   C.cast(cand, C.c_void_p) if cand is not None else None,
                                                   C.cast(res, C.c_void_p), 0))
        return LocalizeOutput(idx, d2, list(cand) if cand is not None else None, list(res))

    def localize_sharded_ptr(self, comm, q_ptr: int, nq: int, pts_ptr: int, offs: np.ndarray, params: LocParams,
                             idx_ptr: int, d2_ptr: int, res, cand=None, device: bool = False) -> None:
        check(_lib.lib().gloc_loc_localize_sharded(self._h, comm._h, q_ptr, nq, pts_ptr, offs.ctypes.data, None,
                                                   C.byref(params), idx_ptr, d2_ptr,
                                                   C.cast(cand, C.c_void_p) if cand is not None else None,
                                                   C.cast(res, C.c_void_p), int(device)))

    def share_grids(self, comm) -> None:
        """Collective: make every rank's grid store readable by its peers (NVLink peer memory) so that
        localize_sharded can hand surplus (query, candidate) pairs to ranks with room."""
        check(_lib.lib().gloc_loc_share_grids(self._h, comm._h))

    def unshare_grids(self, comm) -> None:
        check(_lib.lib().gloc_loc_unshare_grids(self._h, comm._h))

    def set_profiling(self, enabled: bool) -> None:
        check(_lib.lib().gloc_loc_set_profiling(self._h, int(enabled)))

    def profile(self):
        """(total ms, retrieval ms, calls) measured on the device since the last call."""
        p = _lib.LocProfile()
        check(_lib.lib().gloc_loc_get_profile(self._h, C.byref(p)))
        return p.total_ms, p.retrieval_ms, int(p.calls)

    def stats(self) -> LocStats:
        s = LocStats()
        check(_lib.lib().gloc_loc_get_stats(self._h, C.byref(s)))
        return s

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().gloc_loc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
