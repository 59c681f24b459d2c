"""Stage 2 -- correlative / branch-and-bound scan matching of BEV probability grids
on B200, behind the reference's FastCorrelativeScanMatcher2D interface.

Mirrors /root/reference/registration/2d:
  MapLimits / CellLimits            map_limits.h:40-90, xy_index.h:34-45
  Grid2D / ProbabilityGrid          grid_2d.h:34-111, probability_grid.cpp:27-71
  FastCorrelativeScanMatcherOptions2D  fast_correlative_scan_matcher_2d.h:43-52
  FastCorrelativeScanMatcher2D      fast_correlative_scan_matcher_2d.h:137-200
All scoring runs in libgloc3d.so (no CPU path).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import CsmResult, CsmStats, check


@dataclass
class MapLimits:
    """resolution, max() (world corner) and cell limits -- map_limits.h:40-47."""
    resolution: float
    max_x: float
    max_y: float
    num_x_cells: int
    num_y_cells: int


@dataclass
class Rigid2d:
    """transform::Rigid2d as (translation, rotation angle) -- 3d/rigid_transform.h."""
    x: float = 0.0
    y: float = 0.0
    yaw: float = 0.0


class ProbabilityGrid:
    """Grid2D's uint16 correspondence-cost cells (grid_2d.h:100-101; 0 = unknown,
    flat index num_x_cells*y + x) plus the author's explicit origin (grid_2d.h:72-76)."""

    def __init__(self, limits: MapLimits, cells: np.ndarray | None = None,
                 ox: float = 0.0, oy: float = 0.0):
        self.limits = limits
        shape = (limits.num_y_cells, limits.num_x_cells)
        self.cells = (np.zeros(shape, np.uint16) if cells is None
                      else np.ascontiguousarray(cells, np.uint16).reshape(shape))
        self.ox, self.oy = ox, oy


@dataclass
class FastCorrelativeScanMatcherOptions2D:
    linear_search_window: float = 3.0
    angular_search_window: float = 3.0
    branch_and_bound_depth: int = 5


class CsmStore:
    """Many map grids on one GPU + batched matching (C ABI: gloc_csm_*)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(_lib.lib().gloc_csm_create(C.byref(self._h), device))
        self.device = device
        self.limits: list[MapLimits] = []

    def add_grid_cells(self, cells: np.ndarray, resolution: float, max_x: float, max_y: float) -> int:
        cells = np.ascontiguousarray(cells, np.uint16)
        ny, nx = cells.shape
        gid = C.c_int()
        check(_lib.lib().gloc_csm_add_grid_cells(self._h, cells.ctypes.data, nx, ny, resolution,
                                                 max_x, max_y, C.byref(gid)))
        self.limits.append(MapLimits(resolution, max_x, max_y, nx, ny))
        return gid.value

    def add_grid_u8(self, level1: np.ndarray, resolution: float, max_x: float, max_y: float) -> int:
        level1 = np.ascontiguousarray(level1, np.uint8)
        ny, nx = level1.shape
        gid = C.c_int()
        check(_lib.lib().gloc_csm_add_grid_u8(self._h, level1.ctypes.data, nx, ny, resolution,
                                              max_x, max_y, C.byref(gid)))
        self.limits.append(MapLimits(resolution, max_x, max_y, nx, ny))
        return gid.value

    def __len__(self) -> int:
        return int(_lib.lib().gloc_csm_num_grids(self._h))

    def store_bytes(self):
        """(bytes of the resident grids, bytes of the store's reusable work buffers)."""
        a, b = C.c_uint64(), C.c_uint64()
        check(_lib.lib().gloc_csm_store_bytes(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def grid_info(self, grid_id: int) -> MapLimits:
        info = _lib.GridInfo()
        check(_lib.lib().gloc_csm_get_grid_info(self._h, grid_id, C.byref(info)))
        return MapLimits(info.resolution, info.max_x, info.max_y, info.nx, info.ny)

    def save_grids(self, path: str) -> None:
        """Every grid of the store -> a grid store file (gloc_csm_save_grids)."""
        check(_lib.lib().gloc_csm_save_grids(self._h, path.encode()))

    def load_grids(self, path: str) -> list[int]:
        """Appends the grids of a grid store file; returns their ids (gloc_csm_load_grids)."""
        first, n = C.c_int(), C.c_int()
        check(_lib.lib().gloc_csm_load_grids(self._h, path.encode(), C.byref(first), C.byref(n)))
        ids = list(range(first.value, first.value + n.value))
        for gid in ids:
            self.limits.append(self.grid_info(gid))
        return ids

    def precomputation_grid(self, grid_id: int, width: int) -> np.ndarray:
        lim = self.limits[grid_id]
        out = np.empty((lim.num_y_cells + width - 1, lim.num_x_cells + width - 1), np.uint8)
        check(_lib.lib().gloc_csm_get_precomputation_grid(self._h, grid_id, width, out.ctypes.data))
        return out

    def discretize(self, pts, init, n_ang, ang_step, resolution, max_x, max_y) -> np.ndarray:
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty((2 * n_ang + 1, pts.shape[0], 2), np.int32)
        check(_lib.lib().gloc_csm_discretize(self._h, pts.ctypes.data, pts.shape[0], init[0],
                                             init[1], init[2], n_ang, ang_step, resolution, max_x,
                                             max_y, out.ctypes.data))
        return out

    def match_batch(self, scans: list[np.ndarray], grid_ids, scan_ids, inits, n_lin: int,
                    n_ang: int, ang_step: float, depth: int, min_score: float) -> list[CsmResult]:
        """scans: list of [P_i, 3] float32 clouds; per pair grid_ids[i], scan_ids[i], inits[i]."""
        scans = [np.ascontiguousarray(s, np.float32).reshape(-1, 3) for s in scans]
        offs = np.zeros(len(scans) + 1, np.int64)
        offs[1:] = np.cumsum([s.shape[0] for s in scans])
        pts = np.concatenate(scans, axis=0) if scans else np.zeros((0, 3), np.float32)
        pts = np.ascontiguousarray(pts, np.float32)
        gi = np.ascontiguousarray(grid_ids, np.int32)
        si = np.ascontiguousarray(scan_ids, np.int32)
        init = np.ascontiguousarray(inits, np.float64).reshape(-1, 3)
        n = gi.shape[0]
        assert si.shape[0] == n and init.shape[0] == n
        out = (CsmResult * max(n, 1))()
        check(_lib.lib().gloc_csm_match_batch(self._h, pts.ctypes.data, offs.ctypes.data,
                                              len(scans), gi.ctypes.data, si.ctypes.data,
                                              init.ctypes.data, n, n_lin, n_ang, ang_step, depth,
                                              min_score, out))
        return list(out)[:n]

    def set_profiling(self, enabled: bool) -> None:
        check(_lib.lib().gloc_csm_set_profiling(self._h, int(enabled)))

    def profile(self):
        p = _lib.Profile()
        check(_lib.lib().gloc_csm_get_profile(self._h, C.byref(p)))
        return p.dominant_ms, int(p.dominant_launches)

    def stats(self) -> CsmStats:
        s = CsmStats()
        check(_lib.lib().gloc_csm_get_stats(self._h, C.byref(s)))
        return s

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().gloc_csm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def search_parameters(linear_search_window: float, angular_search_window: float,
                      point_cloud: np.ndarray, resolution: float):
    """SearchParameters production ctor (correlative_scan_matcher_2d.cpp:27-55) ->
    (num_linear_perturbations, num_angular_perturbations, angular_perturbation_step_size)."""
    pts = np.ascontiguousarray(point_cloud, np.float32).reshape(-1, 3)
    nl, na, st = C.c_int(), C.c_int(), C.c_double()
    check(_lib.lib().gloc_csm_search_params(linear_search_window, angular_search_window,
                                            pts.ctypes.data, pts.shape[0], resolution,
                                            C.byref(nl), C.byref(na), C.byref(st)))
    return nl.value, na.value, st.value


def grid_to_virtual_point_cloud(grid: ProbabilityGrid) -> np.ndarray:
    """GridToVirtualPointCloud (fast_correlative_scan_matcher_2d.cpp:78-95)."""
    lim = grid.limits
    n = C.c_int()
    fn = _lib.lib().gloc_csm_grid_to_points
    check(fn(grid.cells.ctypes.data, lim.num_x_cells, lim.num_y_cells, lim.resolution, grid.ox,
             grid.oy, None, 0, C.byref(n)))
    pts = np.zeros((n.value, 3), np.float32)
    check(fn(grid.cells.ctypes.data, lim.num_x_cells, lim.num_y_cells, lim.resolution, grid.ox,
             grid.oy, pts.ctypes.data, n.value, C.byref(n)))
    return pts


class FastCorrelativeScanMatcher2D:
    """Same constructor and Match* methods as the reference class
    (fast_correlative_scan_matcher_2d.h:137-200).  Each Match* returns
    (ok, score, pose): ok is the reference's bool return value; score and pose are
    the values the reference writes through its out-pointers (None when ok is False,
    where the reference leaves them untouched)."""

    def __init__(self, grid: ProbabilityGrid, options: FastCorrelativeScanMatcherOptions2D,
                 device: int = 0):
        if options.branch_and_bound_depth < 1:  # CHECK_GE, fast_..._2d.cpp:195
            raise ValueError("Check failed: options.branch_and_bound_depth() >= 1")
        self.options_ = options
        self.limits_ = grid.limits
        self._store = CsmStore(device)
        self._gid = self._store.add_grid_cells(grid.cells, grid.limits.resolution,
                                               grid.limits.max_x, grid.limits.max_y)

    # PrecomputationGridStack2D::Get(index) -> width 2^index grid
    def precomputation_grid(self, index: int) -> np.ndarray:
        return self._store.precomputation_grid(self._gid, 1 << index)

    def Match(self, initial_pose_estimate: Rigid2d, point_cloud_or_grid, min_score: float):
        """fast_..._2d.cpp:219-238 (both overloads)."""
        cloud = self._cloud(point_cloud_or_grid)
        n_lin, n_ang, step = search_parameters(self.options_.linear_search_window,
                                               self.options_.angular_search_window, cloud,
                                               self.limits_.resolution)
        return self.MatchWithSearchParameters((n_lin, n_ang, step), initial_pose_estimate, cloud,
                                              min_score)

    def MatchFullSubmap(self, point_cloud_or_grid, min_score: float):
        """fast_..._2d.cpp:240-268: +-25 cells, +-pi around the grid centre."""
        cloud = self._cloud(point_cloud_or_grid)
        lim = self.limits_
        n_lin, n_ang, step = search_parameters(25 * lim.resolution, np.pi, cloud, lim.resolution)
        center = Rigid2d(lim.max_x - 0.5 * lim.resolution * lim.num_x_cells,
                         lim.max_y - 0.5 * lim.resolution * lim.num_y_cells, 0.0)
        return self.MatchWithSearchParameters((n_lin, n_ang, step), center, cloud, min_score)

    def MatchWithSearchParameters(self, search_parameters_, initial_pose_estimate: Rigid2d,
                                  point_cloud: np.ndarray, min_score: float):
        """fast_..._2d.cpp:270-320.  search_parameters_ = (n_lin, n_ang, step)."""
        n_lin, n_ang, step = search_parameters_
        init = (initial_pose_estimate.x, initial_pose_estimate.y, initial_pose_estimate.yaw)
        r = self._store.match_batch([point_cloud], [self._gid], [0], [init], n_lin, n_ang, step,
                                    self.options_.branch_and_bound_depth, min_score)[0]
        if not r.found:
            return False, None, None
        return True, r.score, Rigid2d(r.pose_x, r.pose_y, r.pose_yaw)

    @staticmethod
    def _cloud(x) -> np.ndarray:
        if isinstance(x, ProbabilityGrid):
            return grid_to_virtual_point_cloud(x)
        return np.ascontiguousarray(x, np.float32).reshape(-1, 3)
