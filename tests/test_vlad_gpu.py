"""NetVLAD_fc head on the GPU (gloc_vlad_*) against the oracle, which is pinned to the
reference's own module (tests/golden/vlad_*.npz).  The kernel source is also checked on the host by
tests/test_vlad_emulated.py."""
import os

import numpy as np
import pytest

import gloc3d_b200 as g
from gloc3d_b200 import _lib
from oracle import vlad_oracle as vo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_create_fails_loudly_without_gpu_and_validates_arguments():
    L = _lib.lib()
    import ctypes as C

    h = C.c_void_p()
    w = np.zeros((8, 32), np.float32)
    hid = np.zeros((256, 32), np.float32)
    assert L.gloc_vlad_create(None, 0, 32, 8, 32, w.ctypes.data, None, w.ctypes.data, hid.ctypes.data) == _lib.GLOC_ERR_INVALID
    assert L.gloc_vlad_create(C.byref(h), 0, 33, 8, 32, w.ctypes.data, None, w.ctypes.data, hid.ctypes.data) == _lib.GLOC_ERR_RANGE
    assert L.gloc_vlad_create(C.byref(h), 0, 32, 65, 32, w.ctypes.data, None, w.ctypes.data, hid.ctypes.data) == _lib.GLOC_ERR_RANGE
    assert L.gloc_vlad_forward(None, None, 1, 1, None) == _lib.GLOC_ERR_INVALID
    assert L.gloc_vlad_kernel_launches(None) == 0
    L.gloc_vlad_destroy(None)
    if L.gloc_device_count() == 0:
        with pytest.raises(g.GlocError) as e:
            g.NetVladHead(w, w, hid)
        assert e.value.code == _lib.GLOC_ERR_CUDA


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vlad_small.npz", "vlad_full.npz"])
def test_reference_goldens(name):
    z = np.load(os.path.join(GOLD, name))
    K, C, H, W, B, seed = (int(z[k]) for k in ("K", "C", "H", "W", "B", "seed"))
    conv_w, cent, hid = vo.hashed_weights(K, C, C, seed)
    x = vo.hashed_features(B, C, H * W, seed + 10)
    head = g.NetVladHead(conv_w, cent, hid)
    out = head.forward(x)
    assert np.abs(out - z["out"]).max() <= 1e-5 * np.abs(z["out"]).max() + 1e-7
    assert head.kernel_launches == 5
    head.close()


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,S,K,D,bias", [(9, 64, 70, 64, 130, True), (2, 96, 129, 5, 17, False),
                                            (33, 512, 2304, 64, 512, False)])
def test_against_oracle_and_batch_independence(B, C, S, K, D, bias):
    conv_w, cent, hid = vo.hashed_weights(K, C, D, 100 + K)
    x = vo.hashed_features(B, C, S, 200 + S)
    conv_b = np.linspace(-0.5, 0.5, K).astype(np.float32) if bias else None
    head = g.NetVladHead(conv_w, cent, hid, conv_b=conv_b)
    out = head.forward(x)
    n_ref = min(B, 3)                      # the numpy oracle is slow at full size
    ref = vo.netvlad_fc(x[:n_ref], conv_w, cent, hid, conv_b=conv_b)
    assert np.abs(out[:n_ref] - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-7
    assert np.array_equal(head.forward(x[B - 1:]), out[B - 1:])      # fixed summation order
    head.close()


@pytest.mark.gpu
def test_descriptors_feed_retrieval_on_the_device():
    import torch

    K, C, D, S = 64, 512, 512, 144
    conv_w, cent, hid = vo.hashed_weights(K, C, D, 5)
    feats = torch.from_numpy(vo.hashed_features(80, C, S, 6)).cuda()
    desc = torch.empty((80, D), dtype=torch.float32, device="cuda")
    head = g.NetVladHead(conv_w, cent, hid)
    head.forward_device(feats.data_ptr(), 80, S, desc.data_ptr())
    ix = g.KnnIndex(D, 0)
    ix.set_db(desc.cpu().numpy())
    idx, d2 = ix.query(desc[:5].cpu().numpy(), 3)
    assert np.array_equal(idx[:, 0], np.arange(5)) and np.all(d2[:, 0] == 0)
    ix.close()
    head.close()
