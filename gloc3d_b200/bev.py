"""BEV projection of LiDAR scans on B200 -- the producer of both stages' inputs.

Mirrors RpyPCLoopDetector::get_projected_grid / crop_pad_occupancy
(/root/reference/registration/loop_detector.cpp:83-135) and the projection helpers of
3d/submap_3d.cpp:238-429.  All work runs in libgloc3d.so (no CPU path).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BevInfo, check


class BevProjector:
    """One scan -> the reference's BEV occupancy image (0 = occupied, 255 = free)."""

    def __init__(self, device: int = 0, resolution: float = 0.2, max_range: float = 100.0):
        self._h = C.c_void_p()
        check(_lib.lib().gloc_bev_create(C.byref(self._h), device, resolution, max_range))
        self.info: BevInfo | None = None

    def project(self, pts: np.ndarray) -> BevInfo:
        """pts: [n, >=3] float32 (x, y, z first; KITTI scans are [n, 4])."""
        pts = np.ascontiguousarray(pts, np.float32)
        if pts.ndim != 2 or pts.shape[1] < 3:
            raise ValueError("points must be [n, >= 3]")
        info = BevInfo()
        check(_lib.lib().gloc_bev_project(self._h, pts.ctypes.data, pts.shape[0], pts.shape[1],
                                          C.byref(info)))
        self.info = info
        return info

    def image(self) -> np.ndarray:
        """cv::Mat of ProjectToCvMat: uint8 [height, width]."""
        img = np.empty((self.info.height, self.info.width), np.uint8)
        check(_lib.lib().gloc_bev_get_image(self._h, img.ctypes.data, img.size))
        return img

    def xy_res(self):
        """(ox, oy, resolution) as get_projected_grid returns them (loop_detector.cpp:133)."""
        return self.info.ox, self.info.oy, self.info.resolution

    def cnn_input(self, width: int = 768, height: int = 768) -> np.ndarray:
        """crop_pad_occupancy (loop_detector.cpp:83-106), one channel."""
        out = np.empty((height, width), np.uint8)
        check(_lib.lib().gloc_bev_get_cnn_input(self._h, width, height, out.ctypes.data))
        return out

    def cnn_input_roi(self, width: int = 768, height: int = 768):
        """cnn_input plus roi_dst of crop_pad_occupancy (x0, y0, w, h): outside it lies padding."""
        out = np.empty((height, width), np.uint8)
        roi = np.zeros(4, np.int32)
        check(_lib.lib().gloc_bev_get_cnn_input_roi(self._h, width, height, out.ctypes.data, roi.ctypes.data))
        return out, roi

    def occupied_points(self) -> np.ndarray:
        """GridToVirtualPointCloud of the projected grid (fast_..._2d.cpp:78-95): [n, 3] float32."""
        n = C.c_size_t()
        check(_lib.lib().gloc_bev_get_occupied_points(self._h, None, 0, C.byref(n)))
        pts = np.zeros((n.value, 3), np.float32)
        if n.value:
            check(_lib.lib().gloc_bev_get_occupied_points(self._h, pts.ctypes.data, n.value,
                                                          C.byref(n)))
        return pts

    def add_to_store(self, store) -> int:
        """ProjectToGrid (3d/submap_3d.cpp:328-429) -> a map grid of a CsmStore, device to device."""
        from .scan_matching import MapLimits

        gid = C.c_int()
        check(_lib.lib().gloc_csm_add_grid_from_bev(store._h, self._h, C.byref(gid)))
        i = self.info
        store.limits.append(MapLimits(i.resolution, (i.min_ix + i.width - 1) * i.resolution,
                                      (i.min_iy + i.height - 1) * i.resolution, i.width, i.height))
        return gid.value

    def add_to_store_aligned(self, store) -> int:
        """The same occupancy as a MapLimits-consistent map grid (gloc_csm_add_grid_from_bev_aligned):
        a world point looks up the pixel of its own voxel, so rigid alignment is possible."""
        from .scan_matching import MapLimits

        gid = C.c_int()
        check(_lib.lib().gloc_csm_add_grid_from_bev_aligned(store._h, self._h, C.byref(gid)))
        i = self.info
        store.limits.append(MapLimits(i.resolution, (i.min_ix + i.width - 1 + 0.5) * i.resolution,
                                      (i.min_iy + i.height - 1 + 0.5) * i.resolution, i.height, i.width))
        return gid.value

    def kernel_launches(self) -> int:
        return int(_lib.lib().gloc_bev_kernel_launches(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().gloc_bev_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
