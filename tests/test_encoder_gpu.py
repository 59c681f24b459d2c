"""VGG16 encoder on the GPU (gloc_enc_*) against the oracle (pinned to torchvision's feature
stack as the reference cuts it), and the whole descriptor path image -> encoder -> NetVLAD_fc
head -> retrieval on the device.  First passed on a B200 in round 1's driver run; the kernel source is also
checked on the host by tests/test_encoder_emulated.py."""
import ctypes as C
import os

import numpy as np
import pytest

import gloc3d_b200 as g
from gloc3d_b200 import _lib
from oracle import encoder_oracle as eo
from oracle import vlad_oracle as vo



def test_create_validates_arguments_and_fails_loudly_without_gpu():
    L = _lib.lib()
    ws, bs = eo.hashed_vgg_weights(1)
    wp = (C.c_void_p * 13)(*[w.ctypes.data for w in ws])
    bp = (C.c_void_p * 13)(*[b.ctypes.data for b in bs])
    h = C.c_void_p()
    assert L.gloc_enc_create(None, 0, 768, 768, wp, bp) == _lib.GLOC_ERR_INVALID
    assert L.gloc_enc_create(C.byref(h), 0, 700, 768, wp, bp) == _lib.GLOC_ERR_RANGE
    assert L.gloc_enc_create(C.byref(h), 0, 768, 640, wp, bp) == _lib.GLOC_ERR_RANGE
    bad = (C.c_void_p * 13)(*([w.ctypes.data for w in ws[:12]] + [None]))
    assert L.gloc_enc_create(C.byref(h), 0, 768, 768, bad, bp) == _lib.GLOC_ERR_INVALID
    assert L.gloc_enc_forward(None, None, 1, None) == _lib.GLOC_ERR_INVALID
    assert L.gloc_enc_kernel_launches(None) == 0
    L.gloc_enc_destroy(None)
    with pytest.raises(ValueError):
        g.Encoder(ws[:12], bs[:12])
    if L.gloc_device_count() == 0:
        with pytest.raises(g.GlocError) as e:
            g.Encoder(ws, bs)
        assert e.value.code == _lib.GLOC_ERR_CUDA


def bev_like(B, H, W, seed):
    rng = np.random.default_rng(seed)
    img = np.full((B, H, W), 255, np.uint8)                    # free = 255, occupied = 0 (the JPG convention)
    for b in range(B):
        for _ in range(40):
            y, x = rng.integers(0, H), rng.integers(0, W)
            if rng.random() < 0.5:
                img[b, y, x:x + rng.integers(5, 60)] = 0
            else:
                img[b, y:y + rng.integers(5, 60), x] = 0
    return img


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,B", [(128, 256, 3), (256, 256, 1)])
def test_encoder_against_oracle(H, W, B):
    ws, bs = eo.hashed_vgg_weights(11)
    img = bev_like(B, H, W, 5)
    enc = g.Encoder(ws, bs, height=H, width=W)
    out = enc.forward(img)
    ref = eo.vgg16_features(img, ws, bs).reshape(B, 512, -1)
    # FP16 operands through 13 layers against float32: a few 1e-3 of the largest activation
    assert np.abs(out - ref).max() <= 2e-2 * np.abs(ref).max(), np.abs(out - ref).max() / np.abs(ref).max()
    assert enc.kernel_launches == 13 + 4
    assert np.array_equal(enc.forward(img[B - 1:]), out[B - 1:])      # frames are independent
    enc.close()


@pytest.mark.gpu
def test_reference_canvas_padding():
    """Small BEV images are padded with the reference's (255, 0, 0) canvas (loop_detector.cpp:84:
    cv::Mat::ones sets channel 0 only): conv1_1 must treat padding and image differently."""
    H, W, B = 256, 256, 3
    ws, bs = eo.hashed_vgg_weights(13)
    img = bev_like(B, H, W, 6)
    rois = np.array([[40, 30, 150, 200], [0, 0, 256, 256], [100, 0, 156, 97]], np.int32)
    for b, (x0, y0, w, h) in enumerate(rois):       # outside the rectangle: the canvas value
        keep = np.zeros((H, W), bool)
        keep[y0:y0 + h, x0:x0 + w] = True
        img[b][~keep] = 255
    enc = g.Encoder(ws, bs, height=H, width=W)
    out = enc.forward(img, rois)
    ref = eo.vgg16_features(img, ws, bs, rois).reshape(B, 512, -1)
    scale = np.abs(ref).max()
    assert np.abs(out - ref).max() <= 2e-2 * scale
    plain = eo.vgg16_features(img, ws, bs).reshape(B, 512, -1)        # all-255 padding: a different network input
    assert np.abs(ref[0] - plain[0]).max() > 0.1 * scale              # the distinction matters ...
    assert np.abs(out[0] - plain[0]).max() > 0.05 * scale             # ... and the GPU follows the reference
    assert np.array_equal(enc.forward(img[1:2]), enc.forward(img[1:2], rois[1:2]))   # full rectangle == no padding
    with pytest.raises(g.GlocError):
        enc.forward(img[:1], np.array([[200, 0, 100, 10]], np.int32))
    enc.close()


@pytest.mark.gpu
def test_reference_input_size_batch_16():
    """768 x 768 (loop_detector.cpp:144-147), 16 frames per call: two frames against the float32
    oracle, and every frame bit-identical to the same frame encoded alone."""
    H = W = 768
    ws, bs = eo.hashed_vgg_weights(17)
    img = bev_like(16, H, W, 8)
    enc = g.Encoder(ws, bs, height=H, width=W)
    out = enc.forward(img)
    for b in (0, 15):
        ref = eo.vgg16_features(img[b:b + 1], ws, bs).reshape(512, -1)
        err = np.abs(out[b] - ref).max() / np.abs(ref).max()
        assert err <= 1e-2, err
    for b in (3, 9):
        assert np.array_equal(enc.forward(img[b:b + 1])[0], out[b])
    enc.close()


@pytest.mark.gpu
def test_descriptor_path_on_the_device():
    import torch

    H, W, B = 256, 256, 6
    ws, bs = eo.hashed_vgg_weights(21)
    conv_w, cent, hid = vo.hashed_weights(64, 512, 512, 31)
    img = bev_like(B, H, W, 9)
    enc, head = g.Encoder(ws, bs, height=H, width=W), g.NetVladHead(conv_w, cent, hid)
    d_img = torch.from_numpy(img).cuda()
    d_feat = torch.empty((B, 512, enc.n_loc), dtype=torch.float32, device="cuda")
    d_desc = torch.empty((B, 512), dtype=torch.float32, device="cuda")
    enc.forward_device(d_img.data_ptr(), B, d_feat.data_ptr())
    head.forward_device(d_feat.data_ptr(), B, enc.n_loc, d_desc.data_ptr())
    ref = vo.netvlad_fc(eo.vgg16_features(img, ws, bs).reshape(B, 512, -1), conv_w, cent, hid)
    got = d_desc.cpu().numpy()
    assert np.abs(got - ref).max() <= 3e-2 * np.abs(ref).max()
    # nearest neighbour of every descriptor among the reference descriptors is itself
    d = ((got[:, None, :] - ref[None, :, :]) ** 2).sum(2)
    assert np.array_equal(d.argmin(1), np.arange(B))
    enc.close()
    head.close()
