"""Weight file of the descriptor network (encoder + NetVLAD_fc head) for gloc_enc_* / gloc_vlad_*.

The reference ships its network as a TorchScript module (MODEL argument of global_localization,
loop_detector.cpp:157-163), which needs libtorch to read.  tools/export_weights.py turns such a
module into this plain container once; everything else reads the container without torch.

Layout (little-endian): b"GLOCW001", u32 n_arrays, then per array
    u16 name_len | name (utf-8) | u32 ndim | u64 shape[ndim] | float32 data (C order)
Names: enc.<l>.weight [Cout, Cin, 3, 3], enc.<l>.bias [Cout] for l = 0..12;
       vlad.conv.weight [K, C], vlad.conv.bias [K] (optional), vlad.centroids [K, C],
       vlad.hidden [K*C, D].
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"GLOCW001"


def save_arrays(path: str, arrays: dict) -> None:
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<I", len(arrays)))
        for name, a in arrays.items():
            a = np.ascontiguousarray(a, np.float32)
            nb = name.encode()
            f.write(struct.pack("<H", len(nb)) + nb + struct.pack("<I", a.ndim))
            f.write(struct.pack(f"<{a.ndim}Q", *a.shape))
            f.write(a.tobytes())


def load_arrays(path: str) -> dict:
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a GLOCW001 weight file")
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            (ln,) = struct.unpack("<H", f.read(2))
            name = f.read(ln).decode()
            (nd,) = struct.unpack("<I", f.read(4))
            shape = struct.unpack(f"<{nd}Q", f.read(8 * nd))
            cnt = int(np.prod(shape)) if nd else 1
            buf = f.read(4 * cnt)
            if len(buf) != 4 * cnt:
                raise ValueError(f"{path}: truncated at {name}")
            out[name] = np.frombuffer(buf, np.float32).reshape(shape).copy()
    return out


def save_weights(path: str, conv_w, conv_b, vlad_conv_w, centroids, hidden_w, vlad_conv_b=None) -> None:
    arrays = {}
    for l, (w, b) in enumerate(zip(conv_w, conv_b)):
        arrays[f"enc.{l}.weight"] = w
        arrays[f"enc.{l}.bias"] = b
    arrays["vlad.conv.weight"] = np.asarray(vlad_conv_w, np.float32).reshape(np.shape(vlad_conv_w)[0], -1)
    if vlad_conv_b is not None:
        arrays["vlad.conv.bias"] = vlad_conv_b
    arrays["vlad.centroids"] = centroids
    arrays["vlad.hidden"] = hidden_w
    save_arrays(path, arrays)


def load_weights(path: str):
    """-> (conv_w[13], conv_b[13], vlad_conv_w, vlad_conv_b or None, centroids, hidden_w)."""
    a = load_arrays(path)
    conv_w = [a[f"enc.{l}.weight"] for l in range(13)]
    conv_b = [a[f"enc.{l}.bias"] for l in range(13)]
    return conv_w, conv_b, a["vlad.conv.weight"], a.get("vlad.conv.bias"), a["vlad.centroids"], a["vlad.hidden"]
