#!/usr/bin/env python
"""Turn the reference's network file into the plain weight container of gloc3d_b200/weights.py.

    python tools/export_weights.py MODEL.pt weights.glocw

MODEL.pt: the TorchScript module the reference's driver loads (`--mode save_pt` of main.py: a
traced VGGVLAD with submodules `encoder` = VGG16 features[:-2] and `pool` = NetVLAD_fc), or a
checkpoint / state_dict with the same parameter names.  Needs torch; nothing else in this
repository does at run time.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gloc3d_b200 import weights as W  # noqa: E402


def state_dict_of(path):
    import torch

    try:
        return {k: v.detach().cpu() for k, v in torch.jit.load(path, map_location="cpu").state_dict().items()}
    except RuntimeError:
        obj = torch.load(path, map_location="cpu")
        if isinstance(obj, dict) and "state_dict" in obj:
            obj = obj["state_dict"]
        return {k: v.detach().cpu() for k, v in obj.items()}


def export(model_path, out_path):
    sd = state_dict_of(model_path)
    sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
    idx = sorted({int(k.split(".")[1]) for k in sd if k.startswith("encoder.") and k.endswith(".weight")})
    if len(idx) != 13:
        raise SystemExit(f"expected the 13 convolutions of VGG16 under encoder.*, found {len(idx)}")
    conv_w = [sd[f"encoder.{i}.weight"].numpy().astype(np.float32) for i in idx]
    conv_b = [sd[f"encoder.{i}.bias"].numpy().astype(np.float32) for i in idx]
    vw = sd["pool.conv.weight"].numpy().astype(np.float32)
    vb = sd["pool.conv.bias"].numpy().astype(np.float32) if "pool.conv.bias" in sd else None
    W.save_weights(out_path, conv_w, conv_b, vw.reshape(vw.shape[0], -1), sd["pool.centroids"].numpy(),
                   sd["pool.hidden1_weights"].numpy(), vlad_conv_b=vb)
    return len(idx)


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    n = export(sys.argv[1], sys.argv[2])
    print(f"wrote {sys.argv[2]}: {n} convolutions + NetVLAD_fc head")
