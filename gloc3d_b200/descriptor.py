"""NetVLAD_fc pooling head on the GPU (C ABI: gloc_vlad_*; reference: model/netvlad_fc.py:73-109
as run by RpyPCLoopDetector::get_place_feature, loop_detector.cpp:137-172): encoder feature maps
in, 512-d place descriptors out, batched; the device entry point feeds KnnIndex.query_device /
set_db_device without a host round trip.  The VGG16 encoder is not part of this package."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


class NetVladHead:
    def __init__(self, conv_w, centroids, hidden_w, conv_b=None, device: int = 0):
        """conv_w [K, C], centroids [K, C], hidden_w [K*C, D], conv_b [K] or None (numpy)."""
        conv_w = np.ascontiguousarray(conv_w, np.float32)
        centroids = np.ascontiguousarray(centroids, np.float32)
        hidden_w = np.ascontiguousarray(hidden_w, np.float32)
        if conv_w.ndim != 2 or centroids.shape != conv_w.shape or hidden_w.ndim != 2 or \
                hidden_w.shape[0] != conv_w.size:
            raise ValueError("conv_w [K, C], centroids [K, C], hidden_w [K*C, D] expected")
        if conv_b is not None:
            conv_b = np.ascontiguousarray(conv_b, np.float32)
            if conv_b.shape != (conv_w.shape[0],):
                raise ValueError("conv_b must be [K]")
        self.clusters, self.dim = conv_w.shape
        self.out_dim = hidden_w.shape[1]
        self.device = device
        self._h = C.c_void_p()
        check(_lib.lib().gloc_vlad_create(C.byref(self._h), device, self.dim, self.clusters, self.out_dim,
                                          conv_w.ctypes.data, None if conv_b is None else conv_b.ctypes.data,
                                          centroids.ctypes.data, hidden_w.ctypes.data))

    def forward(self, feat: np.ndarray) -> np.ndarray:
        """feat [B, C, S] or [B, C, H, W] float32 (host) -> [B, D] float32 (host)."""
        feat = np.ascontiguousarray(feat, np.float32)
        feat = feat.reshape(feat.shape[0], feat.shape[1], -1)
        if feat.shape[1] != self.dim:
            raise ValueError(f"feature maps must have {self.dim} channels")
        out = np.empty((feat.shape[0], self.out_dim), np.float32)
        check(_lib.lib().gloc_vlad_forward(self._h, feat.ctypes.data, feat.shape[0], feat.shape[2],
                                           out.ctypes.data))
        return out

    def forward_device(self, feat_ptr: int, batch: int, n_loc: int, out_ptr: int) -> None:
        """Device pointers: feat [batch][C][n_loc] float32 -> out [batch][D] float32."""
        check(_lib.lib().gloc_vlad_forward_device(self._h, C.c_void_p(feat_ptr), batch, n_loc,
                                                  C.c_void_p(out_ptr)))

    @property
    def kernel_launches(self) -> int:
        return int(_lib.lib().gloc_vlad_kernel_launches(self._h))

    def close(self) -> None:
        if self._h:
            _lib.lib().gloc_vlad_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
