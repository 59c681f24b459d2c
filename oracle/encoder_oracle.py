"""CPU restatement of the reference's encoder -- TEST INFRASTRUCTURE ONLY.

The reference builds it from torchvision (/root/reference/main.py:531-536, :563):
    encoder = models.vgg16(...);  layers = list(encoder.features.children())[:-2]
i.e. the 13 3x3 convolutions (padding 1) with ReLU, 2x2 max-pools after convolutions 2, 4, 7
and 10, and -- the last two children dropped -- neither ReLU nor pool after convolution 13.
Its input is the BEV image of loop_detector.cpp:137-172: uint8, three identical channels,
scaled by 1/255.  Float32 torch on the CPU (a floating-point kernel's reference, as the task
allows); PINNED by tests/test_oracle_encoder.py against torchvision's own vgg16().features[:-2]
carrying the same weights.
"""
import numpy as np

VGG16_COUT = (64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512)
POOL_AFTER = (1, 3, 6, 9)          # 0-based convolution indices followed by a 2x2 max-pool


from gloc3d_b200.synth import hashed_vgg_weights  # noqa: E402,F401


def vgg16_features(images_u8, conv_w, conv_b, rois=None):
    """images_u8 [B, H, W] uint8 -> [B, 512, H/16, W/16] float32.  rois [B, 4] (x0, y0, w, h): the
    rectangle of the BEV image inside the plane; everything outside is the reference's canvas
    padding, cv::Mat::ones(h, w, CV_8UC3) * 255 = (255, 0, 0) per pixel (loop_detector.cpp:84;
    Mat::ones sets channel 0 only), so channels 1 and 2 are zero there."""
    import torch
    import torch.nn.functional as F

    x = torch.from_numpy(np.asarray(images_u8, np.uint8)).float().div(255.0)
    x = x[:, None, :, :].expand(-1, 3, -1, -1).contiguous()
    if rois is not None:
        for b, (x0, y0, w, h) in enumerate(np.asarray(rois).reshape(-1, 4)):
            keep = torch.zeros(x.shape[2:], dtype=torch.bool)
            keep[y0:y0 + h, x0:x0 + w] = True
            x[b, 1:][:, ~keep] = 0.0
    with torch.no_grad():
        for l in range(13):
            x = F.conv2d(x, torch.from_numpy(conv_w[l]), torch.from_numpy(conv_b[l]), padding=1)
            if l != 12:
                x = F.relu(x)
            if l in POOL_AFTER:
                x = F.max_pool2d(x, 2)
    return x.numpy()
