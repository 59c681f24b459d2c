"""The encoder kernels (gloc3d_b200/csrc/encoder.cu) executed on the HOST: the tcgen05
convolution runs from its own source against the functional mbarrier / TMA / tcgen05 model of
tests/cpp/tc_emu.hpp (4-D TMA boxes whose out-of-bounds zero fill IS the padding), the folded
first layer and the max-pool one OS thread per CUDA thread; all against numpy.  The GPU parity
test proper is tests/test_encoder_gpu.py."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
EXE = os.path.join(CPP, "_encoder_emu_test")


@pytest.fixture(scope="module")
def exe():
    src = open(os.path.join(ROOT, "gloc3d_b200", "csrc", "encoder.cu")).read()
    m = re.search(r"// \[enc-kernels-begin\].*?\n(.*?)// \[enc-kernels-end\]", src, re.S)
    assert m, "kernel markers missing in encoder.cu"
    text = m.group(1).replace("extern __shared__ unsigned char enc_smem_raw[];",
                              "unsigned char* enc_smem_raw = t_smem_raw;")
    assert "extern" not in text and "asm" not in text
    open(os.path.join(CPP, "_encoder_kernels.inc"), "w").write(text)
    r = subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-ffp-contract=off", "-fno-strict-aliasing",
                        os.path.join(CPP, "encoder_emu_test.cpp"), "-o", EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-6000:]
    return EXE


def call(exe, mode, tmp_path, header, arrays, async_seed=None):
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.array(header, np.int32).tobytes())
        for a in arrays:
            f.write(np.ascontiguousarray(a).tobytes())
    env = dict(os.environ)
    env.pop("GLOC_EMU_ASYNC", None)
    if async_seed is not None:
        env["GLOC_EMU_ASYNC"] = str(async_seed)
    r = subprocess.run([exe, mode, inp, outp], capture_output=True, text=True, timeout=1500, env=env)
    assert r.returncode == 0, "emulated kernel failed\n" + r.stderr[-4000:]
    return outp


def conv_ref(act, w_oihw, bias, relu):
    """float64 3x3 convolution, zero padding 1, on NHWC input; w [Cout][Cin][3][3] (PyTorch: correlation)."""
    B, H, W, C = act.shape
    pad = np.zeros((B, H + 2, W + 2, C))
    pad[:, 1:-1, 1:-1] = act
    out = np.zeros((B, H, W, w_oihw.shape[0]))
    for ky in range(3):
        for kx in range(3):
            out += pad[:, ky:ky + H, kx:kx + W, :] @ w_oihw[:, :, ky, kx].astype(np.float64).T
    out += bias.astype(np.float64)
    return np.maximum(out, 0) if relu else out


@pytest.mark.parametrize("B,H,W,Cin,Cout,last,workers,async_seed", [
    (2, 16, 32, 64, 64, 0, 3, None),     # BN = 64: 4 patches per image, workers run several tiles each
    (1, 8, 32, 128, 256, 0, 1, 1),       # BN = 256, two k-blocks per tap, one CTA runs both patches
    (3, 8, 16, 64, 128, 1, 2, 2),        # the last layer's form: FP32 NCHW, no ReLU
    (1, 16, 16, 64, 512, 0, 2, 3),       # two channel blocks of 256 per patch
])
def test_emulated_convolution(exe, tmp_path, B, H, W, Cin, Cout, last, workers, async_seed):
    rng = np.random.default_rng(B * 1000 + Cout)
    act = np.maximum(rng.standard_normal((B, H, W, Cin)), 0).astype(np.float16)      # post-ReLU input
    w = (rng.standard_normal((Cout, Cin, 3, 3)) / np.sqrt(9 * Cin)).astype(np.float32)
    bias = rng.standard_normal(Cout).astype(np.float32) * 0.1
    w_h = w.astype(np.float16)
    packed = np.ascontiguousarray(w_h.transpose(0, 2, 3, 1)).reshape(Cout, 9 * Cin)   # [Cout][tap][Cin]
    outp = call(exe, "conv", tmp_path, [B, H, W, Cin, Cout, last, workers], [act, packed, bias], async_seed)
    ref = conv_ref(act.astype(np.float64), w_h.astype(np.float64), bias, relu=not last)
    if last:
        got = np.fromfile(outp, np.float32).reshape(B, Cout, H * W).transpose(0, 2, 1).reshape(B, H, W, Cout)
        assert np.allclose(got, ref, rtol=1e-4, atol=1e-4), np.abs(got - ref).max()
    else:
        got = np.fromfile(outp, np.float16).reshape(B, H, W, Cout).astype(np.float64)
        assert np.allclose(got, ref, rtol=2e-3, atol=2e-3), np.abs(got - ref).max()
        assert np.all(got >= 0)


def test_emulated_first_layer(exe, tmp_path):
    rng = np.random.default_rng(5)
    B, H, W = 2, 12, 40
    img = (rng.random((B, H, W)) < 0.2).astype(np.uint8) * 255
    img[0, 0, :] = 255
    img[1, :, -1] = 255                                     # activity on the borders: the padding matters
    w = (rng.standard_normal((64, 3, 3, 3)) / 5).astype(np.float32)
    bias = (rng.standard_normal(64) * 0.1).astype(np.float32)
    w1 = (w.sum(axis=1).reshape(64, 9) / np.float32(255)).astype(np.float32)          # the library's folding
    outp = call(exe, "conv1", tmp_path, [B, H, W], [img, w1, bias])
    got = np.fromfile(outp, np.float16).reshape(B, H, W, 64).astype(np.float64)
    x3 = np.repeat((img.astype(np.float64) / 255.0)[..., None], 3, axis=3)           # GRAY2BGR, 1/255
    ref = conv_ref(x3, w.astype(np.float64), bias, relu=True)
    assert np.allclose(got, ref, rtol=2e-3, atol=2e-3), np.abs(got - ref).max()


def test_emulated_max_pool(exe, tmp_path):
    rng = np.random.default_rng(6)
    B, H, W, C = 2, 6, 10, 72
    x = np.maximum(rng.standard_normal((B, H, W, C)), 0).astype(np.float16)
    x[0, 0, 0, :8] = np.float16(-0.0)                       # what fmaxf(-0, 0) may leave behind
    outp = call(exe, "pool", tmp_path, [B, H, W, C], [x])
    got = np.fromfile(outp, np.float16).reshape(B, H // 2, W // 2, C)
    ref = x.reshape(B, H // 2, 2, W // 2, 2, C).max(axis=(2, 4))
    assert np.array_equal(got.astype(np.float32), ref.astype(np.float32))
