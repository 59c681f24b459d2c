"""Descriptor extraction on the GPU: the VGG16 encoder (C ABI: gloc_enc_*; reference:
main.py:531-536 as run by RpyPCLoopDetector::get_place_feature, loop_detector.cpp:137-172) and the
NetVLAD_fc pooling head on the GPU (C ABI: gloc_vlad_*; reference: model/netvlad_fc.py:73-109
as run by RpyPCLoopDetector::get_place_feature, loop_detector.cpp:137-172): encoder feature maps
in, 512-d place descriptors out, batched; the device entry point feeds KnnIndex.query_device /
set_db_device without a host round trip."""
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


VGG16_COUT = (64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512)


class Encoder:
    """VGG16 features[:-2] on uint8 BEV images [B, H, W] -> feature maps [B, 512, H/16 * W/16]."""

    def __init__(self, conv_w, conv_b, height: int = 768, width: int = 768, device: int = 0):
        """conv_w: 13 arrays [Cout, Cin, 3, 3] float32 (torchvision layout), conv_b: 13 arrays [Cout]."""
        if len(conv_w) != 13 or len(conv_b) != 13:
            raise ValueError("13 convolution layers expected")
        cin = 3
        ws, bs = [], []
        for w, b, cout in zip(conv_w, conv_b, VGG16_COUT):
            w = np.ascontiguousarray(w, np.float32)
            b = np.ascontiguousarray(b, np.float32)
            if w.shape != (cout, cin, 3, 3) or b.shape != (cout,):
                raise ValueError(f"layer with {cout} outputs: weight {w.shape}, bias {b.shape}")
            ws.append(w)
            bs.append(b)
            cin = cout
        self._keep = (ws, bs)
        wp = (C.c_void_p * 13)(*[w.ctypes.data for w in ws])
        bp = (C.c_void_p * 13)(*[b.ctypes.data for b in bs])
        self.height, self.width, self.device = height, width, device
        self._h = C.c_void_p()
        check(_lib.lib().gloc_enc_create(C.byref(self._h), device, height, width, wp, bp))
        self.channels, self.n_loc = 512, (height // 16) * (width // 16)

    def forward(self, images: np.ndarray, rois=None) -> np.ndarray:
        """rois [B, 4] int32 (x0, y0, w, h of the BEV image inside the plane): the rest of the plane is
        the reference's (255, 0, 0) canvas padding; None: every pixel is image."""
        images = np.ascontiguousarray(images, np.uint8)
        if images.ndim != 3 or images.shape[1:] != (self.height, self.width):
            raise ValueError(f"images must be [B, {self.height}, {self.width}] uint8")
        out = np.empty((images.shape[0], self.channels, self.n_loc), np.float32)
        if rois is None:
            check(_lib.lib().gloc_enc_forward(self._h, images.ctypes.data, images.shape[0], out.ctypes.data))
        else:
            rois = np.ascontiguousarray(rois, np.int32).reshape(images.shape[0], 4)
            check(_lib.lib().gloc_enc_forward_padded(self._h, images.ctypes.data, rois.ctypes.data, images.shape[0],
                                                     out.ctypes.data))
        return out

    def forward_device(self, images_ptr: int, batch: int, feat_ptr: int) -> None:
        check(_lib.lib().gloc_enc_forward_device(self._h, C.c_void_p(images_ptr), batch, C.c_void_p(feat_ptr)))

    @property
    def kernel_launches(self) -> int:
        return int(_lib.lib().gloc_enc_kernel_launches(self._h))

    def close(self) -> None:
        if self._h:
            _lib.lib().gloc_enc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DescriptorExtractor:
    """BEV image plane(s) -> place descriptors: encoder + NetVLAD_fc head, device-resident in
    between (torch only provides the device buffers)."""

    def __init__(self, conv_w, conv_b, vlad_conv_w, centroids, hidden_w, vlad_conv_b=None,
                 height: int = 768, width: int = 768, device: int = 0):
        self.device = device
        self.enc = Encoder(conv_w, conv_b, height=height, width=width, device=device)
        self.head = NetVladHead(vlad_conv_w, centroids, hidden_w, conv_b=vlad_conv_b, device=device)
        self.out_dim = self.head.out_dim
        self._feat = None

    @classmethod
    def from_file(cls, path: str, **kw):
        from .weights import load_weights

        conv_w, conv_b, vw, vb, cent, hid = load_weights(path)
        return cls(conv_w, conv_b, vw, cent, hid, vlad_conv_b=vb, **kw)

    def describe_device(self, images_ptr: int, batch: int, desc_ptr: int) -> None:
        """uint8 planes [batch][H][W] (device) -> descriptors [batch][D] float32 (device)."""
        import torch

        dev = torch.device("cuda", self.device)
        if self._feat is None or self._feat.shape[0] < batch:
            self._feat = torch.empty((batch, self.enc.channels, self.enc.n_loc), dtype=torch.float32, device=dev)
            torch.cuda.synchronize(dev)
        self.enc.forward_device(images_ptr, batch, self._feat.data_ptr())
        self.head.forward_device(self._feat.data_ptr(), batch, self.enc.n_loc, desc_ptr)

    def describe(self, images: np.ndarray) -> np.ndarray:
        """uint8 planes [B, H, W] (host) -> descriptors [B, D] float32 (host)."""
        import torch

        dev = torch.device("cuda", self.device)
        images = np.ascontiguousarray(images, np.uint8)
        d_img = torch.from_numpy(images).to(dev)
        d_desc = torch.empty((images.shape[0], self.out_dim), dtype=torch.float32, device=dev)
        torch.cuda.synchronize(dev)
        self.describe_device(d_img.data_ptr(), images.shape[0], d_desc.data_ptr())
        return d_desc.cpu().numpy()

    def close(self) -> None:
        self.enc.close()
        self.head.close()
        self._feat = None
