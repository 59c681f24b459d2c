"""The NetVLAD_fc restatement (oracle/vlad_oracle.py) against outputs of the reference's own
module (tests/golden/vlad_*.npz, minted by tests/golden/make_golden.py from
/root/reference/model/netvlad_fc.py): this part of the oracle is PINNED."""
import os

import numpy as np
import pytest

from oracle import vlad_oracle as vo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name))
    K, C, H, W, B, seed = (int(z[k]) for k in ("K", "C", "H", "W", "B", "seed"))
    conv_w, cent, hid = vo.hashed_weights(K, C, C, seed)
    x = vo.hashed_features(B, C, H * W, seed + 10)
    # the hash generators reproduce the tensors the fixture was minted with
    assert np.array_equal(np.array([conv_w[0, 0], cent[-1, -1], hid[-1, -1], x[-1, -1, -1]]), z["probe"])
    return x, conv_w, cent, hid, z["out"]


@pytest.mark.parametrize("name", ["vlad_small.npz", "vlad_full.npz"])
def test_restatement_matches_the_reference_module(name):
    x, conv_w, cent, hid, ref = load(name)
    out = vo.netvlad_fc(x, conv_w, cent, hid)
    assert out.shape == ref.shape and out.dtype == np.float32
    # the reference sums in float32 (torch CPU), the restatement in float64
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-7, np.abs(out - ref).max()


def test_properties_of_the_head():
    x, conv_w, cent, hid, _ = load("vlad_small.npz")
    base = vo.netvlad_fc(x, conv_w, cent, hid)
    # the input is L2-normalised per location: positive rescaling of any location changes nothing
    scaled = x * np.linspace(0.5, 4.0, x.shape[2], dtype=np.float32)[None, None, :]
    assert np.allclose(vo.netvlad_fc(scaled, conv_w, cent, hid), base, atol=2e-6)
    # the sum over locations does not depend on their order
    perm = np.random.default_rng(0).permutation(x.shape[2])
    assert np.allclose(vo.netvlad_fc(x[:, :, perm], conv_w, cent, hid), base, atol=2e-6)
    # the output is linear in hidden_w, and frames are independent
    assert np.allclose(vo.netvlad_fc(x, conv_w, cent, 2 * hid), 2 * base, atol=2e-6)
    assert np.array_equal(vo.netvlad_fc(x[1:2], conv_w, cent, hid), base[1:2])
    # an all-zero location (eps clamp of F.normalize) is tolerated: it only moves mass to -a*c
    z = x.copy()
    z[:, :, 0] = 0
    assert np.all(np.isfinite(vo.netvlad_fc(z, conv_w, cent, hid)))
    # bias shifts the logits (vladv2 layout)
    b = np.linspace(-1, 1, conv_w.shape[0]).astype(np.float32)
    assert not np.allclose(vo.netvlad_fc(x, conv_w, cent, hid, conv_b=b), base, atol=1e-4)
