"""CPU: the stage-1 oracle against the reference's own nanoflann (oracle/_ref) and the
golden vectors minted from it."""
import glob
import os

import numpy as np
import pytest

from gloc3d_b200 import synth
from helpers import canonical_ties


def np_l2_reference_order(q, x):
    """evalMetric (nanoflann.hpp:453-487) in numpy float32, one op at a time."""
    q = q.astype(np.float32)
    x = x.astype(np.float32)
    r = np.float32(0)
    n4 = (len(q) // 4) * 4
    for g in range(0, n4, 4):
        d = (q[g:g + 4] - x[g:g + 4]).astype(np.float32)
        s = np.float32(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2]))
        s = np.float32(s + np.float32(d[3] * d[3]))
        r = np.float32(r + s)
    for i in range(n4, len(q)):
        d = np.float32(q[i] - x[i])
        r = np.float32(r + np.float32(d * d))
    return r


@pytest.mark.parametrize("dim", [1, 3, 4, 7, 30, 512])
def test_l2_matches_numpy_restatement(oracle, dim):
    rng = np.random.default_rng(dim)
    for _ in range(20):
        q = rng.standard_normal(dim).astype(np.float32)
        x = rng.standard_normal(dim).astype(np.float32)
        assert oracle.l2(q, x).view(np.uint32) == np_l2_reference_order(q, x).view(np.uint32)


def test_golden_vectors_from_reference(oracle, golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "knn_*.npz")))
    assert len(files) >= 3
    for f in files:
        z = np.load(f)
        k = int(z["k"])
        idx, d2 = oracle.knn(z["db"], z["q"], k)
        assert np.array_equal(d2.view(np.uint32), z["d2"].view(np.uint32)), f
        assert np.array_equal(idx, canonical_ties(z["idx"], z["d2"])), f


@pytest.mark.parametrize("n,k,sigma,dup", [(4541, 20, None, 0), (4541, 25, 0.002, 8), (3000, 25, 0.01, 16)])
def test_oracle_equals_compiled_reference(oracle, n, k, sigma, dup):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    db = synth.make_descriptors(n, seed=1234, dup_run=dup)
    q = synth.make_queries(db, 150, seed=5678, sigma=sigma)
    tree = oracle.RefTree(db, 10)
    ridx, rd2 = tree.query(q, k, nthreads=8)
    idx, d2 = oracle.knn(db, q, k, nthreads=8)
    assert np.array_equal(d2.view(np.uint32), rd2.view(np.uint32))
    # indices: equal up to the order inside runs of bit-equal distances; a tie that
    # straddles the k boundary may pick either row, so compare those rows as distances only
    idx1, d21 = oracle.knn(db, q, k + 1, nthreads=8)
    straddle = d21[:, k - 1] == d21[:, k]
    ok = ~straddle
    assert np.array_equal(idx[ok], canonical_ties(ridx, rd2)[ok])


def test_short_database_and_merge(oracle):
    db = synth.make_descriptors(10, 16, seed=3)
    q = synth.make_queries(db, 4, seed=4)
    idx, d2 = oracle.knn(db, q, 16)
    assert (idx[:, 10:] == np.iinfo(np.uint64).max).all()
    assert (d2[:, 10:] == np.finfo(np.float32).max).all()
    assert (np.diff(d2[:, :10].astype(np.float64), axis=1) >= 0).all()
    # sharded search + merge == unsharded search, for 1, 2, 4, 8 shards
    db = synth.make_descriptors(1000, 32, seed=5, dup_run=4)
    q = synth.make_queries(db, 33, seed=6, sigma=0.001)
    full_idx, full_d2 = oracle.knn(db, q, 25)
    for g in (1, 2, 4, 8):
        bounds = [1000 * i // g for i in range(g + 1)]
        parts = [oracle.knn(db[bounds[i]:bounds[i + 1]], q, 25) for i in range(g)]
        pidx = np.stack([np.where(p[0] == np.iinfo(np.uint64).max, p[0], p[0] + np.uint64(bounds[i]))
                         for i, p in enumerate(parts)])
        pd2 = np.stack([p[1] for p in parts])
        midx, md2 = oracle.topk_merge(pidx, pd2)
        assert np.array_equal(midx, full_idx) and np.array_equal(md2, full_d2)
