"""CPU restatement of the reference's NetVLAD_fc pooling head -- TEST INFRASTRUCTURE ONLY
(imported by tests/ and tests/golden/make_golden.py; the product never imports oracle/).

Follows /root/reference/model/netvlad_fc.py:73-109 (NetVLAD.forward, vladv2 = False, gating off),
numpy, one frame at a time, float32 inputs with float64 accumulation:

    x      [C, S]   feature map of the encoder (C = 512 channels, S = H*W locations)
    x^     = x / max(||x[:, s]||_2, 1e-12)                 F.normalize(x, p=2, dim=1)      :76-77
    a      = softmax_k(conv_w @ x^ (+ conv_b))             1x1 conv + softmax over clusters :80-81
    V[k]   = sum_s a[k, s] (x^[:, s] - c_k)                residuals to each centroid       :88-96
    V[k]   = V[k] / max(||V[k]||_2, 1e-12)                 intra-normalisation              :99
    v      = flatten(V) / max(||flatten(V)||_2, 1e-12)     [K*C]                            :101-102
    out    = v @ hidden1_weights                           [K*C] -> [out_dim]               :105

PINNED: tests/golden/vlad_*.npz hold the outputs of the reference's own NetVLAD module
(imported from /root/reference/model/netvlad_fc.py by tests/golden/make_golden.py, torch CPU
float32) for seeded weights and inputs; tests/test_oracle_vlad.py checks this restatement
against them.
"""
import numpy as np

EPS = 1e-12   # F.normalize's default eps


def netvlad_fc(x, conv_w, centroids, hidden_w, conv_b=None):
    """x [B, C, S] (or [B, C, H, W]); conv_w [K, C]; centroids [K, C]; hidden_w [K*C, D].
    Returns [B, D] float32."""
    x = np.asarray(x, np.float32)
    x = x.reshape(x.shape[0], x.shape[1], -1).astype(np.float64)
    w = np.asarray(conv_w, np.float32).astype(np.float64)
    c = np.asarray(centroids, np.float32).astype(np.float64)
    h = np.asarray(hidden_w, np.float32).astype(np.float64)
    out = np.empty((x.shape[0], h.shape[1]), np.float32)
    for b in range(x.shape[0]):
        xb = x[b]
        xn = xb / np.maximum(np.sqrt((xb * xb).sum(axis=0, keepdims=True)), EPS)
        logits = w @ xn
        if conv_b is not None:
            logits = logits + np.asarray(conv_b, np.float64)[:, None]
        logits -= logits.max(axis=0, keepdims=True)
        a = np.exp(logits)
        a /= a.sum(axis=0, keepdims=True)
        V = a @ xn.T - a.sum(axis=1, keepdims=True) * c
        V = V / np.maximum(np.sqrt((V * V).sum(axis=1, keepdims=True)), EPS)
        v = V.reshape(-1)
        v = v / max(np.sqrt((v * v).sum()), EPS)
        out[b] = (v @ h).astype(np.float32)
    return out


# synthetic weights / features: the integer-hash generators live with the other synthetic inputs
from gloc3d_b200.synth import hash_uniform as _hash_uniform  # noqa: E402,F401
from gloc3d_b200.synth import hashed_features, hashed_vlad_weights as hashed_weights  # noqa: E402,F401
