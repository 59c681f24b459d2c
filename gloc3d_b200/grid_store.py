"""Grid store file (include/gloc3d.h "grid store file"): a map's BEV grids
on disk, bit-packed, so that a database is projected once instead of at every start-up as the
reference does (global_localization.cpp:419-449).  Thin ctypes wrappers; the file code is host
code inside libgloc3d.so and needs no device."""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import numpy as np

from . import _lib
from ._lib import check


class StoredGrid(NamedTuple):
    level1: np.ndarray      # [ny, nx] uint8 width-1 precomputation grid (0 .. 255)
    resolution: float
    max_x: float
    max_y: float


def write_grid_file(path: str, grids) -> None:
    """grids: iterable of (level1 [ny, nx] uint8, resolution, max_x, max_y)."""
    grids = [StoredGrid(np.ascontiguousarray(g[0], np.uint8), float(g[1]), float(g[2]), float(g[3]))
             for g in grids]
    n = len(grids)
    infos = (_lib.GridInfo * max(n, 1))()
    ptrs = (C.c_void_p * max(n, 1))()
    for i, g in enumerate(grids):
        if g.level1.ndim != 2:
            raise ValueError("level1 must be a 2-D uint8 array [ny, nx]")
        infos[i] = _lib.GridInfo(g.level1.shape[1], g.level1.shape[0], g.resolution, g.max_x, g.max_y)
        ptrs[i] = g.level1.ctypes.data
    check(_lib.lib().gloc_grid_file_write(path.encode(), infos, ptrs, n))


def read_grid_file(path: str) -> list[StoredGrid]:
    h = C.c_void_p()
    n = C.c_size_t()
    check(_lib.lib().gloc_grid_file_open(path.encode(), C.byref(h), C.byref(n)))
    out = []
    try:
        for _ in range(n.value):
            info = _lib.GridInfo()
            check(_lib.lib().gloc_grid_file_next(h, C.byref(info), None, 0))          # sizes only
            cells = np.empty((info.ny, info.nx), np.uint8)
            check(_lib.lib().gloc_grid_file_next(h, C.byref(info), cells.ctypes.data, cells.size))
            out.append(StoredGrid(cells, info.resolution, info.max_x, info.max_y))
    finally:
        _lib.lib().gloc_grid_file_close(h)
    return out
