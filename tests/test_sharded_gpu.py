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
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
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


def test_one_process_two_gpus_balanced_localize():
    """gloc_loc_share_grids inside ONE process (peer pointers, no IPC): all queries revisit shard 0, so
    device 0 hands pairs to device 1, which reads device 0's grids over NVLink; results = the oracles."""
    if n_gpus() < 2:
        pytest.skip("needs two GPUs")
    import ctypes as C

    import gloc3d_b200 as g
    from gloc3d_b200 import _lib, synth
    from oracle import pyoracle as po

    class LocalComm:
        def __init__(self, h):
            self._h = h

    L = _lib.lib()
    comms = (C.c_void_p * 2)()
    _lib.check(L.gloc_comm_create_local(comms, 2, None))
    res, nx, ny, rows_per, k, nq = 0.2, 96, 80, 40, 5, 6
    mx, my = synth.centered_limits(nx, ny, res)
    db = synth.make_descriptors(2 * rows_per, seed=21, dup_run=4)
    grids = [synth.make_bev_grid(nx, ny, seed=300 + i, n_segments=10, n_blobs=6) for i in range(2 * rows_per)]
    rng = np.random.default_rng(5)
    qrows = rng.integers(0, rows_per, nq)                     # shard 0 only
    q = (db[qrows] + rng.standard_normal((nq, 512)).astype(np.float32) * 0.01).astype(np.float32)
    scans = [synth.planted_scan(grids[int(r)], res, mx, my, rng.uniform(-0.3, 0.3), rng.uniform(-1, 1), rng.uniform(-1, 1),
                                dropout=0.2, seed=int(r)) for r in qrows]
    n_lin, n_ang, step, depth, min_score = 16, 12, 2 * np.pi / 360, 4, 0.4
    ref_idx, _ = po.knn(db, q, k)
    ref = [[po.csm_match(grids[int(ref_idx[qi, c])], res, mx, my, depth, scans[qi], (0, 0, 0), n_lin, n_ang, step,
                         min_score, 0) for c in range(k)] for qi in range(nq)]
    parts = []
    for d in range(2):
        ix = g.KnnIndex(512, d)
        ix.set_db(db[d * rows_per:(d + 1) * rows_per])
        ix.set_index_offset(d * rows_per)
        st = g.CsmStore(d)
        for i in range(rows_per):                              # identity row -> grid table
            st.add_grid_u8(grids[d * rows_per + i], res, mx, my)
        parts.append((ix, st, g.Localizer(ix, st)))
    outs, errs = [None, None], []

    def work(d):
        try:
            loc = parts[d][2]
            loc.share_grids(LocalComm(comms[d]))
            prm = loc.params(k, n_lin, n_ang, step, depth, min_score, g.LOC_VERIFY_ALL)
            outs[d] = (loc.localize_sharded(LocalComm(comms[d]), q, scans, prm), loc.stats())
            loc.unshare_grids(LocalComm(comms[d]))
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(d,)) for d in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=180)
    assert not errs, errs
    quota = -(-nq * k // 2)
    for d in range(2):
        o, s = outs[d]
        assert np.array_equal(o.idx, ref_idx)
        for qi in range(nq):
            for c in range(k):
                r, e = o.candidates[qi * k + c], ref[qi][c]
                assert r.found == e.found and np.float32(r.score).view(np.uint32) == np.float32(e.score).view(np.uint32)
                if e.found:
                    assert (r.scan_index, r.x_offset, r.y_offset) == (e.scan_index, e.x_offset, e.y_offset)
        assert s.pairs_verified <= quota
    assert outs[0][1].pairs_migrated == 0 and outs[1][1].pairs_migrated > 0
    assert outs[0][1].pairs_verified + outs[1][1].pairs_verified == nq * k
    for ix, st, loc in parts:
        loc.close(); st.close(); ix.close()
    for c in comms:
        L.gloc_comm_destroy(c)
