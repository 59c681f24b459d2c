"""Row-sharded retrieval and localisation through the C ABI (gloc_comm_*, gloc_knn_query_sharded*,
gloc_loc_localize_sharded: NCCL inside libgloc3d.so) on 2 GPUs against the single-process oracles.
Needs a box with at least two GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_one_process_per_gpu_under_torchrun():
    if n_gpus() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert tail.count("sharded paths ok") == 2, tail


def test_one_process_driving_two_gpus():
    """The C++ host's shape: one process, gloc_comm_create_local, one thread per device."""
    if n_gpus() < 2:
        pytest.skip("needs two GPUs")
    import ctypes as C

    import gloc3d_b200 as g
    from gloc3d_b200 import _lib, synth
    from oracle import pyoracle as po

    L = _lib.lib()
    comms = (C.c_void_p * 2)()
    _lib.check(L.gloc_comm_create_local(comms, 2, None))
    assert L.gloc_comm_size(comms[0]) == 2 and L.gloc_comm_rank(comms[1]) == 1
    n, k, nq = 3001, 20, 5
    db = synth.make_descriptors(n, seed=12, dup_run=8)
    q = synth.make_queries(db, nq, seed=13, sigma=0.01)
    ref_idx, ref_d2 = po.knn(db, q, k)
    bounds = [0, n // 2, n]
    shards = []
    for d in range(2):
        ix = g.KnnIndex(512, d)
        ix.set_db(db[bounds[d]:bounds[d + 1]])
        ix.set_index_offset(bounds[d])
        shards.append(ix)
    outs = [(np.empty((nq, k), np.uint64), np.empty((nq, k), np.float32)) for _ in range(2)]
    errs = []

    def work(d):
        try:
            _lib.check(L.gloc_knn_query_sharded(shards[d]._h, comms[d], q.ctypes.data, nq, k, outs[d][0].ctypes.data,
                                                outs[d][1].ctypes.data, 1))
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(d,)) for d in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not errs, errs
    for d in range(2):
        assert np.array_equal(outs[d][0], ref_idx) and np.array_equal(outs[d][1].view(np.uint32), ref_d2.view(np.uint32))
    for ix in shards:
        ix.close()
    for c in comms:
        L.gloc_comm_destroy(c)
