"""The kernels of gloc3d_b200/csrc/vlad.cu executed on the HOST through a small CUDA emulator
(tests/cpp/cuda_emu.hpp: one OS thread per CUDA thread, real barriers and shuffles) with the
launch geometry of the library, against the oracle.  This checks indexing, barrier placement
and arithmetic of the kernel source without a GPU; the parity test proper is
tests/test_vlad_gpu.py."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import vlad_oracle as vo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
EXE = os.path.join(CPP, "_vlad_emu_test")


@pytest.fixture(scope="module")
def exe():
    src = open(os.path.join(ROOT, "gloc3d_b200", "csrc", "vlad.cu")).read()
    m = re.search(r"// \[kernels-begin\].*?\n(.*)// \[kernels-end\]", src, re.S)
    assert m, "kernel markers missing in vlad.cu"
    text = m.group(1).replace("extern __shared__", "extern")
    assert "__shared__" in text and "__syncthreads" in text
    open(os.path.join(CPP, "_vlad_kernels.inc"), "w").write(text)
    r = subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-fno-strict-aliasing",
                        os.path.join(CPP, "vlad_emu_test.cpp"), "-o", EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return EXE


def run(exe, tmp_path, x, conv_w, conv_b, cent, hid):
    B, C, S = x.shape
    K, D = conv_w.shape[0], hid.shape[1]
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.array([B, C, S, K, D, 0 if conv_b is None else 1], np.int32).tobytes())
        for arr in (x, conv_w, np.zeros(K, np.float32) if conv_b is None else conv_b, cent, hid):
            f.write(np.ascontiguousarray(arr, np.float32).tobytes())
    r = subprocess.run([exe, inp, outp], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    raw = np.fromfile(outp, np.float32)
    return raw[:B * D].reshape(B, D), raw[B * D:].reshape(B, K * C)


@pytest.mark.parametrize("B,C,S,K,D,bias", [
    (3, 32, 30, 8, 32, False),      # the small golden's shape: single tiles, ragged S
    (9, 64, 70, 64, 130, True),     # two FC passes (9 > 8 frames), D not a multiple of 128, S spans 3 tiles, bias
    (2, 96, 129, 5, 17, False),     # S one past a CTA of 128 locations, K not a multiple of 8
])
def test_emulated_kernels_match_the_oracle(exe, tmp_path, B, C, S, K, D, bias):
    conv_w, cent, hid = vo.hashed_weights(K, C, D, 100 + K)
    x = vo.hashed_features(B, C, S, 200 + S)
    conv_b = np.linspace(-0.5, 0.5, K).astype(np.float32) if bias else None
    out, V = run(exe, tmp_path, x, conv_w, conv_b, cent, hid)
    ref = vo.netvlad_fc(x, conv_w, cent, hid, conv_b=conv_b)
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-7, np.abs(out - ref).max()
    # the normalised VLAD vector has unit length and unit-length (or empty) cluster rows
    assert np.allclose(np.linalg.norm(V, axis=1), 1.0, atol=1e-5)
    rows = np.linalg.norm(V.reshape(B, K, C), axis=2)
    assert np.allclose(rows, rows[:, :1], rtol=1e-4)


def test_emulated_kernels_reproduce_the_reference_golden(exe, tmp_path):
    z = np.load(os.path.join(ROOT, "tests", "golden", "vlad_small.npz"))
    K, C, H, W, B, seed = (int(z[k]) for k in ("K", "C", "H", "W", "B", "seed"))
    conv_w, cent, hid = vo.hashed_weights(K, C, C, seed)
    x = vo.hashed_features(B, C, H * W, seed + 10)
    out, _ = run(exe, tmp_path, x, conv_w, None, cent, hid)
    assert np.abs(out - z["out"]).max() <= 1e-5 * np.abs(z["out"]).max() + 1e-7


def test_a_frame_does_not_depend_on_its_batch(exe, tmp_path):
    conv_w, cent, hid = vo.hashed_weights(8, 32, 32, 7)
    x = vo.hashed_features(10, 32, 40, 8)
    full, _ = run(exe, tmp_path, x, conv_w, None, cent, hid)
    part, _ = run(exe, tmp_path, x[8:], conv_w, None, cent, hid)
    assert np.array_equal(full[8:], part)           # bit for bit: fixed summation order
