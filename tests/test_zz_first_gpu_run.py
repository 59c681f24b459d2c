"""First contact with a GPU for the code written after round 1's GPU minutes were spent: the
NetVLAD head, the encoder, the grid-store round trip and the CTA-pair GEMM.  Each group runs its
own (opt-in) parity tests in a SEPARATE process -- a faulting kernel cannot poison the CUDA
context of the main suite -- and is a non-strict xfail: the suite stays green either way, and
the outcome (xpassed / xfailed) is on record.  This file sorts last on purpose.  Once a group
has passed on a B200, drop its GLOC_TEST_UNVERIFIED guard and its entry here."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_group(files, extra_env, timeout_s):
    env = dict(os.environ, GLOC_TEST_UNVERIFIED="1", **extra_env)
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + files
    try:
        r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout_s)
    except subprocess.TimeoutExpired:
        pytest.fail(f"timed out after {timeout_s} s")
    tail = (r.stdout + r.stderr)[-1500:]
    print(tail)
    assert r.returncode == 0, tail


@pytest.mark.gpu
@pytest.mark.xfail(reason="NetVLAD head / grid store round trip: not yet run on a GPU", strict=False)
def test_first_run_vlad_head_and_grid_store():
    run_group(["tests/test_vlad_gpu.py", "tests/test_grid_store.py"], {}, 600)


@pytest.mark.gpu
@pytest.mark.xfail(reason="tcgen05 encoder: not yet run on a GPU", strict=False)
def test_first_run_encoder():
    run_group(["tests/test_encoder_gpu.py", "tests/test_driver_network_gpu.py"], {}, 1200)


@pytest.mark.gpu
@pytest.mark.xfail(reason="CTA-pair shortlist GEMM (GLOC_KNN_PAIR=1): not yet run on a GPU", strict=False)
def test_first_run_pair_gemm():
    run_group(["tests/test_knn_gpu.py"], {"GLOC_KNN_PAIR": "1"}, 900)
