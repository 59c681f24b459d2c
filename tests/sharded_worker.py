"""Worker of tests/test_sharded_gpu.py: run under torchrun with one rank per GPU.  Every rank checks
the row-sharded C-ABI paths (NCCL inside libgloc3d.so) against the single-process oracles."""
import datetime
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gloc3d_b200 as g  # noqa: E402
from gloc3d_b200 import synth  # noqa: E402
from gloc3d_b200.distributed import (Comm, query_sharded_device, query_sharded_host,  # noqa: E402
                                     shard_bounds)
from oracle import pyoracle as po  # noqa: E402


def same(r, o):
    assert r.found == o.found
    if o.found:
        assert (r.scan_index, r.x_offset, r.y_offset) == (o.scan_index, o.x_offset, o.y_offset)
        assert np.float32(r.score).view(np.uint32) == np.float32(o.score).view(np.uint32)
        assert (r.pose_x, r.pose_y, r.pose_yaw) == (o.pose_x, o.pose_y, o.pose_yaw)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    comm = Comm.from_torch(local)
    assert (comm.rank, comm.size) == (rank, world)
    dev = torch.device("cuda", local)

    # ---- retrieval: sliced batch (all-gather queries, all-to-all lists, merge at the owner)
    n, k = 6000 + 37, 25
    db = synth.make_descriptors(n, seed=3, dup_run=8)
    per = 160
    q_all = np.concatenate([synth.make_queries(db, per * world // 2, seed=4),
                            synth.make_queries(db, per * world - per * world // 2, seed=5, sigma=0.01)])
    b = shard_bounds(n, world)
    ix = g.KnnIndex(512, local)
    ix.set_db(db[b[rank]:b[rank + 1]])
    ix.set_index_offset(b[rank])
    ref_idx, ref_d2 = po.knn(db, q_all, k, nthreads=4)
    mine = slice(rank * per, (rank + 1) * per)
    for mode in (g.KNN_AUTO, g.KNN_EXACT_SCAN):
        ix.set_mode(mode)
        oi, od = query_sharded_device(ix, comm, torch.from_numpy(q_all[mine]).to(dev), k, False)
        torch.cuda.synchronize()
        assert np.array_equal(oi.cpu().numpy().view(np.uint64), ref_idx[mine]), f"rank {rank}: sliced indices (mode {mode})"
        assert np.array_equal(od.cpu().numpy().view(np.uint32), ref_d2[mine].view(np.uint32))
    ix.set_mode(g.KNN_AUTO)
    # host buffers
    hi_, hd_ = np.empty((per, k), np.uint64), np.empty((per, k), np.float32)
    qh = np.ascontiguousarray(q_all[mine])
    query_sharded_host(ix, comm, qh.ctypes.data, per, k, hi_.ctypes.data, hd_.ctypes.data, False)
    assert np.array_equal(hi_, ref_idx[mine]) and np.array_equal(hd_.view(np.uint32), ref_d2[mine].view(np.uint32))
    # ---- retrieval: the same query on every rank (online localisation), 1 and 3 queries
    for nq in (1, 3):
        oi, od = query_sharded_device(ix, comm, torch.from_numpy(q_all[:nq]).to(dev), k, True)
        torch.cuda.synchronize()
        assert np.array_equal(oi.cpu().numpy().view(np.uint64), ref_idx[:nq]), f"rank {rank}: replicated indices"
        assert np.array_equal(od.cpu().numpy().view(np.uint32), ref_d2[:nq].view(np.uint32))
    # a shard with fewer rows than k still merges correctly
    tiny = g.KnnIndex(512, local)
    tb = [0, 7, 60] if world == 2 else shard_bounds(60, world)
    tiny.set_db(db[tb[rank]:tb[rank + 1]])
    tiny.set_index_offset(tb[rank])
    tiny.set_mode(g.KNN_EXACT_SCAN)
    oi, od = query_sharded_device(tiny, comm, torch.from_numpy(q_all[:2]).to(dev), k, True)
    torch.cuda.synchronize()
    t_idx, t_d2 = po.knn(db[:tb[-1]], q_all[:2], k)
    assert np.array_equal(oi.cpu().numpy().view(np.uint64), t_idx)
    tiny.close()

    # ---- whole path over the sharded database
    res, nx, ny, G = 0.2, 150, 120, 12            # G places per shard
    mx, my = synth.centered_limits(nx, ny, res)
    rows = 400 * world
    ldb = synth.make_descriptors(rows, seed=8, dup_run=4)
    lb = shard_bounds(rows, world)

    def place(r):
        s = int(np.searchsorted(lb, r, side="right") - 1)
        return s * G + (r - lb[s]) % G

    grids = [synth.make_bev_grid(nx, ny, seed=900 + i, n_segments=14, n_blobs=8) for i in range(G * world)]
    lix = g.KnnIndex(512, local)
    lix.set_db(ldb[lb[rank]:lb[rank + 1]])
    lix.set_index_offset(lb[rank])
    st = g.CsmStore(local)
    for i in range(G):
        st.add_grid_u8(grids[rank * G + i], res, mx, my)
    loc = g.Localizer(lix, st)
    loc.set_row_grids((np.arange(lb[rank + 1] - lb[rank]) % G).astype(np.int32))
    rng = np.random.default_rng(17)
    nq, kk = 7, 6
    qrows = rng.integers(0, rows, nq)
    lq = (ldb[qrows] + rng.standard_normal((nq, 512)).astype(np.float32) * 0.01).astype(np.float32)
    scans = [synth.planted_scan(grids[place(int(r))], res, mx, my, rng.uniform(-0.4, 0.4), rng.uniform(-2, 2),
                                rng.uniform(-2, 2), dropout=0.2, seed=int(r)) for r in qrows]
    n_lin, n_ang, step, depth, min_score = 24, 30, 2 * np.pi / 360, 4, 0.45
    prm = loc.params(kk, n_lin, n_ang, step, depth, min_score, g.LOC_VERIFY_ALL)
    out = loc.localize_sharded(comm, lq, scans, prm)
    lref_idx, lref_d2 = po.knn(ldb, lq, kk, nthreads=4)
    assert np.array_equal(out.idx, lref_idx) and np.array_equal(out.d2.view(np.uint32), lref_d2.view(np.uint32))
    located = 0
    for qi in range(nq):
        first = -1
        for c in range(kk):
            o = po.csm_match(grids[place(int(lref_idx[qi, c]))], res, mx, my, depth, scans[qi], (0, 0, 0), n_lin, n_ang,
                             step, min_score, 0)
            same(out.candidates[qi * kk + c], o)
            if o.found and first < 0:
                first = c
        R = out.results[qi]
        assert (R.located, R.candidate) == (int(first >= 0), first) and R.n_verified == kk
        located += R.located
    assert located >= nq - 2
    prm.policy = g.LOC_FIRST_MATCH
    out2 = loc.localize_sharded(comm, lq, scans, prm)
    for qi in range(nq):
        A, B = out.results[qi], out2.results[qi]
        assert (A.located, A.candidate, A.db_index) == (B.located, B.candidate, B.db_index)
        same(B.match, A.match)
    # every rank verified only the pairs it owns
    owned = int(((lref_idx >= lb[rank]) & (lref_idx < lb[rank + 1])).sum())
    assert loc.stats().pairs_verified <= 2 * owned

    # ---- balanced verification over peer memory (gloc_loc_share_grids): queries that all revisit rank
    # 0's rows, so rank 0 owns most candidates and must hand pairs to its peers, which read rank 0's
    # bit-packed grids over NVLink.  Same results as the oracle, per-rank work within the quota.
    nq2 = 9
    qrows2 = rng.integers(lb[0], lb[1], nq2)
    lq2 = (ldb[qrows2] + rng.standard_normal((nq2, 512)).astype(np.float32) * 0.01).astype(np.float32)
    scans2 = [synth.planted_scan(grids[place(int(r))], res, mx, my, rng.uniform(-0.4, 0.4), rng.uniform(-2, 2),
                                 rng.uniform(-2, 2), dropout=0.2, seed=int(r) + 5) for r in qrows2]
    ref2_idx, _ = po.knn(ldb, lq2, kk, nthreads=4)
    ref2 = [[po.csm_match(grids[place(int(ref2_idx[qi, c]))], res, mx, my, depth, scans2[qi], (0, 0, 0), n_lin, n_ang,
                          step, min_score, 0) for c in range(kk)] for qi in range(nq2)]
    owned2 = [int(((ref2_idx >= lb[r]) & (ref2_idx < lb[r + 1])).sum()) for r in range(world)]
    quota = -(-nq2 * kk // world)
    assert owned2[0] > quota                       # the scenario is skewed
    loc.share_grids(comm)

    def check_balanced(tag):
        prm.policy = g.LOC_VERIFY_ALL
        s0 = loc.stats()
        o = loc.localize_sharded(comm, lq2, scans2, prm)
        s1 = loc.stats()
        assert np.array_equal(o.idx, ref2_idx), tag
        for qi in range(nq2):
            for c in range(kk):
                same(o.candidates[qi * kk + c], ref2[qi][c])
        mine_now = s1.pairs_verified - s0.pairs_verified
        assert mine_now <= quota, (tag, rank, mine_now, quota)
        moved = s1.pairs_migrated - s0.pairs_migrated
        assert (moved == 0) if rank == 0 else (moved >= 0)
        tot = torch.tensor([mine_now, moved], device=dev, dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot[0]) == nq2 * kk and int(tot[1]) >= owned2[0] - quota, (tag, tot.tolist())
        prm.policy = g.LOC_FIRST_MATCH
        o2 = loc.localize_sharded(comm, lq2, scans2, prm)
        for qi in range(nq2):
            A, B = o.results[qi], o2.results[qi]
            assert (A.located, A.candidate, A.db_index) == (B.located, B.candidate, B.db_index), tag
            same(B.match, A.match)

    check_balanced("bits over peer memory")
    # the first batch again: balanced and owner-only verification agree
    prm.policy = g.LOC_VERIFY_ALL
    out3 = loc.localize_sharded(comm, lq, scans, prm)
    assert [c.as_tuple() for c in out3.candidates] == [c.as_tuple() for c in out.candidates]
    # a graded grid anywhere in the job switches every rank to the uint8 kernels; the tables must be
    # shared again after a store changed (the stale state is an error, not a silent fallback)
    if rank == world - 1:
        st.add_grid_u8(synth.make_bev_grid(nx, ny, seed=77, n_segments=14, n_blobs=8, graded=True), res, mx, my)
        try:
            loc.localize_sharded(comm, lq2, scans2, prm)
            raise AssertionError("a changed shard must be refused")
        except g.GlocError:
            pass
    dist.barrier()
    loc.share_grids(comm)
    check_balanced("uint8 kernels over peer memory")
    loc.unshare_grids(comm)
    prm.policy = g.LOC_VERIFY_ALL
    out4 = loc.localize_sharded(comm, lq2, scans2, prm)          # owner-only again
    for qi in range(nq2):
        for c in range(kk):
            same(out4.candidates[qi * kk + c], ref2[qi][c])
    loc.close(); st.close(); lix.close(); ix.close(); comm.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: sharded paths ok", flush=True)


if __name__ == "__main__":
    main()
