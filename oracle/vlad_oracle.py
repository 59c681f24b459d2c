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


def _hash_uniform(shape, seed):
    """Deterministic pseudo-random float32 values in [-1, 1) from integer arithmetic only (a
    splitmix64 finaliser of the element index): identical on every platform and numpy version,
    so fixtures need not store large weight tensors."""
    n = int(np.prod(shape))
    off = np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) + off
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u24 = (z >> np.uint64(40)).astype(np.float32)          # 24 random bits: exact in float32
    return (u24 / np.float32(1 << 23) - np.float32(1.0)).reshape(shape)


def hashed_weights(K, C, D, seed):
    """conv_w [K, C] ~ U(-1, 1)/sqrt(C) (the Conv2d default init range, netvlad_fc.py:34),
    centroids [K, C] ~ U(0, 1) (:35), hidden_w [K*C, D] ~ U(-1, 1) sqrt(3/C) (same variance as
    the reference's randn/sqrt(dim), :37-38)."""
    conv_w = _hash_uniform((K, C), seed) / np.float32(np.sqrt(C))
    centroids = (_hash_uniform((K, C), seed + 1) + np.float32(1.0)) * np.float32(0.5)
    hidden_w = _hash_uniform((K * C, D), seed + 2) * np.float32(np.sqrt(3.0 / C))
    return conv_w.astype(np.float32), centroids.astype(np.float32), hidden_w.astype(np.float32)


def hashed_features(B, C, S, seed):
    """A feature map [B, C, S] with conv5_3-like statistics (no ReLU: both signs)."""
    return (_hash_uniform((B, C, S), seed) * np.float32(3.0)).astype(np.float32)
