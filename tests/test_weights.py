"""Weight container (gloc3d_b200/weights.py) and the TorchScript exporter (tools/export_weights.py).
The second test assembles the reference's network the way main.py does -- torchvision's VGG16
features[:-2] as `encoder`, the reference's own NetVLAD module as `pool` -- traces it like
`--mode save_pt`, exports it, and checks the oracle chain (encoder_oracle + vlad_oracle) on the
exported weights against the traced module: the descriptor oracles are pinned end to end."""
import os
import sys

import numpy as np
import pytest

from gloc3d_b200 import weights as W
from oracle import encoder_oracle as eo
from oracle import vlad_oracle as vo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MODEL = "/root/reference/model"


def test_container_round_trip(tmp_path):
    ws, bs = eo.hashed_vgg_weights(1)
    cw, cent, hid = vo.hashed_weights(8, 512, 16, 2)
    p = str(tmp_path / "w.glocw")
    W.save_weights(p, ws, bs, cw, cent, hid, vlad_conv_b=np.arange(8, dtype=np.float32))
    conv_w, conv_b, vw, vb, c2, h2 = W.load_weights(p)
    assert all(np.array_equal(a, b) for a, b in zip(conv_w, ws)) and all(np.array_equal(a, b) for a, b in zip(conv_b, bs))
    assert np.array_equal(vw, cw) and np.array_equal(vb, np.arange(8)) and np.array_equal(c2, cent) and np.array_equal(h2, hid)
    W.save_weights(p, ws, bs, cw, cent, hid)
    assert W.load_weights(p)[3] is None
    open(p, "wb").write(b"NOTAFILE")
    with pytest.raises(ValueError):
        W.load_arrays(p)


@pytest.mark.skipif(not os.path.isdir(REF_MODEL), reason="needs the reference's model/ directory (build container only)")
def test_exported_torchscript_matches_the_oracle_chain(tmp_path):
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")
    sys.path.insert(0, REF_MODEL)
    import netvlad_fc
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import export_weights

    class Net(torch.nn.Module):          # main.py's VGGVLAD: encoder, then pool
        def __init__(self, encoder, pool):
            super().__init__()
            self.encoder, self.pool = encoder, pool

        def forward(self, x):
            return self.pool(self.encoder(x))

    torch.manual_seed(0)
    enc = torch.nn.Sequential(*list(tv.models.vgg16(weights=None).features.children())[:-2])
    net = Net(enc, netvlad_fc.NetVLAD(num_clusters=64, dim=512)).eval()
    rng = np.random.default_rng(0)
    img = (rng.random((2, 64, 96)) < 0.1).astype(np.uint8) * 255
    x = torch.from_numpy(img).float().div(255.0)[:, None].expand(-1, 3, -1, -1).contiguous()
    with torch.no_grad():
        traced = torch.jit.trace(net, x)
        ref = traced(x).numpy()
    pt, out = str(tmp_path / "model.pt"), str(tmp_path / "model.glocw")
    traced.save(pt)
    assert export_weights.export(pt, out) == 13
    conv_w, conv_b, vw, vb, cent, hid = W.load_weights(out)
    assert vb is None and vw.shape == (64, 512) and hid.shape == (64 * 512, 512)
    feat = eo.vgg16_features(img, conv_w, conv_b)
    desc = vo.netvlad_fc(feat, vw, cent, hid)
    assert desc.shape == ref.shape == (2, 512)
    assert np.abs(desc - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-7, np.abs(desc - ref).max()
