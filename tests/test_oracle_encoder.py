"""The encoder restatement (oracle/encoder_oracle.py) against torchvision's own VGG16 feature
stack cut the way the reference cuts it (main.py:531-536): PINNED."""
import numpy as np
import pytest

from oracle import encoder_oracle as eo


def test_restatement_is_torchvisions_vgg16_features_without_the_last_two_children():
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")
    ws, bs = eo.hashed_vgg_weights(3)
    net = tv.models.vgg16(weights=None)
    layers = list(net.features.children())[:-2]                 # the reference's cut
    convs = [m for m in layers if isinstance(m, torch.nn.Conv2d)]
    assert len(convs) == 13 and tuple(c.out_channels for c in convs) == eo.VGG16_COUT
    assert not isinstance(layers[-1], (torch.nn.ReLU, torch.nn.MaxPool2d))
    with torch.no_grad():
        for c, w, b in zip(convs, ws, bs):
            c.weight.copy_(torch.from_numpy(w))
            c.bias.copy_(torch.from_numpy(b))
    rng = np.random.default_rng(0)
    img = (rng.random((2, 32, 48)) < 0.1).astype(np.uint8) * 255
    x = torch.from_numpy(img).float().div(255.0)[:, None].expand(-1, 3, -1, -1).contiguous()
    with torch.no_grad():
        ref = torch.nn.Sequential(*layers)(x).numpy()
    out = eo.vgg16_features(img, ws, bs)
    assert out.shape == (2, 512, 2, 3) and np.array_equal(out, ref)
    assert np.abs(out).max() > 1e-3                             # the hashed weights keep the signal alive
