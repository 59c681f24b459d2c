"""The C++ host mirrors (gloc3d_b200/host/*.hpp) compile against the C ABI (CPU) and
reproduce the oracle through the reference's own interfaces (GPU)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_host_mirror_test")


def build():
    from oracle import pyoracle

    pyoracle.build()
    cmd = ["g++", "-O2", "-std=c++14", "-ffp-contract=off", os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
           "-o", BIN, f"-L{ROOT}/gloc3d_b200", f"-L{ROOT}/oracle", "-lgloc3d", "-lgloc_oracle",
           f"-Wl,-rpath,{ROOT}/gloc3d_b200", f"-Wl,-rpath,{ROOT}/oracle", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_host_mirrors_compile_and_link():
    build()
    assert os.path.exists(BIN)


def test_host_only_logic_without_gpu():
    """Guards, SearchParameters, GridToVirtualPointCloud, option defaults: no device needed."""
    from oracle import pyoracle

    pyoracle.build()
    exe = os.path.join(ROOT, "tests", "cpp", "_host_cpu_test")
    cmd = ["g++", "-O2", "-std=c++14", "-ffp-contract=off", os.path.join(ROOT, "tests", "cpp", "host_cpu_test.cpp"),
           "-o", exe, f"-L{ROOT}/gloc3d_b200", f"-L{ROOT}/oracle", "-lgloc3d", "-lgloc_oracle",
           f"-Wl,-rpath,{ROOT}/gloc3d_b200", f"-Wl,-rpath,{ROOT}/oracle", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASS host cpu" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_host_mirrors_match_oracle():
    build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
