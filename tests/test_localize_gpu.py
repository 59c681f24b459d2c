"""GPU parity tests of the whole query path (gloc_loc_*: retrieval -> gather of the candidates'
grids -> verification -> located frame and pose) against the two oracles stage by stage, and of the
pooled grid store behind it: every batch meets grids it has never seen, nothing is cached per grid.

Reference flow: global_localization.cpp:482-574 (detect_all_query, global_registraion: candidates in
retrieval order, first match wins)."""
import numpy as np
import pytest

import gloc3d_b200 as g
from gloc3d_b200 import synth

pytestmark = pytest.mark.gpu


def same(r, o):
    assert r.found == o.found
    if o.found:
        assert (r.scan_index, r.x_offset, r.y_offset) == (o.scan_index, o.x_offset, o.y_offset)
        assert np.float32(r.score).view(np.uint32) == np.float32(o.score).view(np.uint32)
        assert (r.pose_x, r.pose_y, r.pose_yaw) == (o.pose_x, o.pose_y, o.pose_yaw)
    else:
        assert np.float32(r.score) == np.float32(o.score)


def make_world(n_rows, n_grids, nq, nx, ny, seed, graded_every=0):
    """A database whose row r was taken at place r % n_grids, and queries that revisit the place of
    a random row: perturbed descriptor + a scan planted in that place's grid."""
    res = 0.2
    db = synth.make_descriptors(n_rows, seed=seed, dup_run=4)
    mx, my = synth.centered_limits(nx, ny, res)
    grids = [synth.make_bev_grid(nx, ny, seed=seed * 1000 + i, n_segments=14, n_blobs=8,
                                 graded=bool(graded_every and i % graded_every == 0)) for i in range(n_grids)]
    rng = np.random.default_rng(seed + 7)
    rows = rng.integers(0, n_rows, nq)
    q = (db[rows] + rng.standard_normal((nq, 512)).astype(np.float32) * 0.01).astype(np.float32)
    scans, poses = [], []
    for i, r in enumerate(rows):
        yaw, dx, dy = rng.uniform(-0.4, 0.4), rng.uniform(-2, 2), rng.uniform(-2, 2)
        scans.append(synth.planted_scan(grids[r % n_grids], res, mx, my, yaw, dx, dy, dropout=0.2,
                                        seed=seed + i))
        poses.append((yaw, dx, dy))
    return db, grids, (res, mx, my), q, scans, rows, poses


@pytest.mark.parametrize("graded_every", [0, 5])
def test_localize_equals_the_two_oracles(oracle, graded_every):
    n_rows, n_grids, nq, k = 700, 60, 9, 6
    n_lin, n_ang, step, depth, min_score = 24, 30, 2 * np.pi / 360, 4, 0.45
    db, grids, (res, mx, my), q, scans, rows, _ = make_world(n_rows, n_grids, nq, 150, 120, 11, graded_every)
    ix = g.KnnIndex(512, 0)
    ix.set_db(db)
    st = g.CsmStore(0)
    for gr in grids:
        st.add_grid_u8(gr, res, mx, my)
    loc = g.Localizer(ix, st)
    loc.set_row_grids(np.arange(n_rows, dtype=np.int32) % n_grids)
    inits = np.zeros((nq, 3))
    inits[:, 2] = np.linspace(-0.1, 0.1, nq)      # a non-trivial initial yaw per query
    inits[:, 0] = 0.2
    prm = loc.params(k, n_lin, n_ang, step, depth, min_score, g.LOC_VERIFY_ALL)
    out = loc.localize(q, scans, prm, inits)
    ref_idx, ref_d2 = oracle.knn(db, q, k, nthreads=4)
    assert np.array_equal(out.idx, ref_idx)
    assert np.array_equal(out.d2.view(np.uint32), ref_d2.view(np.uint32))
    n_located = 0
    for qi in range(nq):
        first = -1
        for c in range(k):
            gid = int(ref_idx[qi, c]) % n_grids
            o = oracle.csm_match(grids[gid], res, mx, my, depth, scans[qi], tuple(inits[qi]), n_lin, n_ang,
                                 step, min_score, 0)
            same(out.candidates[qi * k + c], o)
            if o.found and first < 0:
                first = c
        R = out.results[qi]
        assert R.n_verified == k
        assert R.located == int(first >= 0) and R.candidate == first
        if first >= 0:
            n_located += 1
            assert R.db_index == ref_idx[qi, first]
            same(R.match, out.candidates[qi * k + first])
            best = max(range(k), key=lambda c: (out.candidates[qi * k + c].found,
                                                out.candidates[qi * k + c].score, -c))
            assert R.best_candidate == best
    assert n_located >= nq - 2          # the planted place is among the candidates and matches

    # the reference's evaluation order: same located frame and pose, fewer verifications
    prm.policy = g.LOC_FIRST_MATCH
    out2 = loc.localize(q, scans, prm, inits)
    assert np.array_equal(out2.idx, out.idx)
    for qi in range(nq):
        A, B = out.results[qi], out2.results[qi]
        assert (A.located, A.candidate, A.db_index) == (B.located, B.candidate, B.db_index)
        same(B.match, A.match)
        assert B.n_verified <= A.n_verified
        for c in range(k):
            r2 = out2.candidates[qi * k + c]
            if r2.reserved == -1:
                assert A.located and c > A.candidate
            else:
                same(r2, out.candidates[qi * k + c])
    assert loc.stats().pairs_verified < 2 * nq * k
    loc.close()
    st.close()
    ix.close()


def test_identity_row_grid_table_and_errors(oracle):
    db, grids, (res, mx, my), q, scans, rows, _ = make_world(40, 40, 3, 90, 70, 5)
    ix = g.KnnIndex(512, 0)
    ix.set_db(db)
    st = g.CsmStore(0)
    loc = g.Localizer(ix, st)
    prm = loc.params(4, 12, 10, 0.02, 3, 0.4)
    with pytest.raises(g.GlocError):          # empty store
        loc.localize(q, scans, prm)
    for gr in grids[:-1]:
        st.add_grid_u8(gr, res, mx, my)
    with pytest.raises(g.GlocError):          # a database row without a grid
        loc.localize(q, scans, prm)
    st.add_grid_u8(grids[-1], res, mx, my)
    out = loc.localize(q, scans, prm)         # db_grids_[db_idx]: grid r belongs to row r
    ref_idx, _ = oracle.knn(db, q, 4)
    assert np.array_equal(out.idx, ref_idx)
    for qi in range(3):
        for c in range(4):
            o = oracle.csm_match(grids[int(ref_idx[qi, c])], res, mx, my, 3, scans[qi], (0, 0, 0), 12, 10,
                                 0.02, 0.4, 0)
            same(out.candidates[qi * 4 + c], o)
    with pytest.raises(g.GlocError):
        loc.set_row_grids(np.array([0, 1, 99], np.int32))
    prm.k = 41
    with pytest.raises(g.GlocError):          # fewer than k rows
        loc.localize(q, scans, prm)
    loc.close()
    st.close()
    ix.close()


def test_store_keeps_bits_only_and_rebuilds_per_batch(oracle):
    """SURVEY a-11: the precomputation stack must be on-the-fly on the GPU.  200 grids of 400 x 400 cost
    about their bit-packed size; matching a batch in which every pair meets a different grid, in
    several sub-batches, equals the oracle; the work buffers do not grow with the number of grids."""
    res, nx, ny, n = 0.2, 400, 400, 200
    mx, my = synth.centered_limits(nx, ny, res)
    st = g.CsmStore(0)
    grids = [synth.make_bev_grid(nx, ny, seed=9000 + i, n_segments=30, n_blobs=20) for i in range(n)]
    for gr in grids:
        st.add_grid_u8(gr, res, mx, my)
    grid_bytes, _ = st.store_bytes()
    per_grid = ny * ((((nx + 31) // 32) + 2) & ~1) * 4     # csm_bit_stride: even, one zero word after the row
    assert grid_bytes <= n * (per_grid + 256)
    assert grid_bytes < n * nx * ny // 6          # far below one byte per cell
    rng = np.random.default_rng(3)
    scans = [synth.planted_scan(grids[i], res, mx, my, rng.uniform(-0.5, 0.5), rng.uniform(-3, 3),
                                rng.uniform(-3, 3), dropout=0.3, seed=i) for i in range(0, n, 8)]
    gi = list(range(n))
    si = [(i // 8) for i in range(n)]             # grid i vs the scan planted in grid 8 * (i // 8)
    inits = [(0.0, 0.0, 0.0)] * n
    import os
    os.environ["GLOC_CSM_SUB"] = "64"             # four sub-batches: slots are reused
    try:
        out = st.match_batch(scans, gi, si, inits, 40, 30, 2 * np.pi / 360, 5, 0.35)
    finally:
        del os.environ["GLOC_CSM_SUB"]
    _, ws1 = st.store_bytes()
    probe = list(range(0, n, 8)) + [1, 2, 3, 77, 199]
    for i in probe:
        o = oracle.csm_match(grids[gi[i]], res, mx, my, 5, scans[si[i]], (0, 0, 0), 40, 30, 2 * np.pi / 360,
                             0.35, 0)
        same(out[i], o)
    assert sum(out[i].found for i in range(0, n, 8)) >= n // 8 - 1
    # the same batch again, and the same grids shared by many pairs (deduplicated slots)
    out2 = st.match_batch(scans, gi, si, inits, 40, 30, 2 * np.pi / 360, 5, 0.35)
    assert [r.as_tuple() for r in out2] == [r.as_tuple() for r in out]
    shared = st.match_batch(scans, [gi[8]] * 5 + [gi[0]] * 5, [1, 0, 1, 2, 1, 0, 1, 0, 3, 0], inits[:10], 40, 30,
                            2 * np.pi / 360, 5, 0.35)
    same(shared[0], out[8])
    same(shared[2], out[8])
    same(shared[5], out[0])
    same(shared[9], out[0])
    _, ws2 = st.store_bytes()
    assert ws2 <= max(ws1, 1) * 4                 # work buffers are per batch, not per grid
    st.close()
