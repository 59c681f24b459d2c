"""CPU: the multi-GPU host logic (row sharding, all-gather layout, merge) with world_size 2
over gloo.  The GPU kernels are replaced by injected functions (the oracle as the local
searcher and as the merge) -- only the plumbing of gloc3d_b200.distributed is under test."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gloc3d_b200 import synth
    from gloc3d_b200.distributed import ShardedRetrieval, shard_bounds
    from oracle import pyoracle as po

    n, nq, k = 3001, 37, 25
    db = synth.make_descriptors(n, 64, seed=7, dup_run=4)
    q = synth.make_queries(db, nq, seed=8, sigma=0.002)
    b = shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]

    def local_search(qt, kk):
        idx, d2 = po.knn(db[lo:hi], qt.numpy(), kk)
        idx = np.where(idx == np.iinfo(np.uint64).max, idx, idx + np.uint64(lo))
        return torch.from_numpy(idx.view(np.int64)), torch.from_numpy(d2)

    def merge(all_idx, all_d2):
        mi, md = po.topk_merge(all_idx.numpy().view(np.uint64), all_d2.numpy())
        return torch.from_numpy(mi.view(np.int64)), torch.from_numpy(md)

    sr = ShardedRetrieval(rank, world, local_search=local_search, merge=merge)
    idx, d2 = sr.query(torch.from_numpy(q), k)
    ref_idx, ref_d2 = po.knn(db, q, k)
    ok = np.array_equal(idx.numpy().view(np.uint64), ref_idx) and np.array_equal(d2.numpy(), ref_d2)
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_shard_bounds():
    from gloc3d_b200.distributed import shard_bounds

    for n in (0, 1, 7, 100_000, 1_000_003):
        for g in (1, 2, 4, 8):
            b = shard_bounds(n, g)
            assert b[0] == 0 and b[-1] == n and len(b) == g + 1
            sizes = np.diff(b)
            assert (sizes >= 0).all() and sizes.max() - sizes.min() <= 1


@pytest.mark.parametrize("world", [2])
def test_sharded_query_gloo(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"


def test_bench_steps_are_rank_uniform():
    """bench.py: a step may contain a collective (row-sharded protocol: all-gather of the local
    top-k), so no step may run under a rank-dependent condition -- a rank-0-only loop of steps
    dead-locks every N > 1 run.  Static check of the control flow."""
    import ast
    import os

    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    tree = ast.parse(src)

    def mentions_rank(node):
        return any(isinstance(n, ast.Name) and n.id == "rank" for n in ast.walk(node))

    def step_calls(node):
        out = []
        for n in ast.walk(node):
            if isinstance(n, ast.Call):
                f = n.func
                name = f.attr if isinstance(f, ast.Attribute) else getattr(f, "id", "")
                if name in ("query", "query_host", "query_device", "query_ptr", "step", "e2e_step", "match_batch",
                            "all_reduce", "all_gather", "all_gather_into_tensor", "barrier", "measure"):
                    out.append(name)
        return out

    bad = []
    for node in ast.walk(tree):
        if isinstance(node, (ast.If, ast.While)) and mentions_rank(node.test):
            for stmt in node.body:
                bad += step_calls(stmt)
    assert not bad, f"steps / collectives under a rank-dependent condition: {bad}"


def test_hold_steps_is_a_pure_function():
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.hold_steps(1.0) == 250 and bench.hold_steps(5.2) == 48 and bench.hold_steps(1e-9) == 4000
    assert bench.hold_steps(1e9) == 1


def test_sharded_localizer_host_logic():
    """partition_pairs / combine_pair_keys: every (query, candidate) pair has exactly one owner and
    the all-reduce(max) of the per-rank key arrays (0 = not mine) reassembles all results."""
    from gloc3d_b200.distributed import combine_pair_keys, partition_pairs, shard_bounds

    rng = np.random.default_rng(0)
    rows, world, nq, k = 1000, 4, 9, 25
    idx = rng.integers(0, rows, (nq, k)).astype(np.uint64)
    keys = rng.integers(1, 1 << 62, nq * k).astype(np.uint64)
    b = shard_bounds(rows, world)
    seen = np.zeros(nq * k, int)
    per_rank = []
    for r in range(world):
        mine = partition_pairs(idx, b[r], b[r + 1])
        seen[mine] += 1
        kr = np.zeros(nq * k, np.uint64)
        kr[mine] = keys[mine]
        per_rank.append(kr)
    assert (seen == 1).all()
    assert np.array_equal(combine_pair_keys(per_rank), keys)


def test_balanced_pair_assignment_rule():
    """gloc_loc_assign_pairs -- the library's own rule (host only) for which rank verifies which pair once the
    ranks share their grid stores: every pair exactly once, nobody above the quota, an owner keeps its first
    quota pairs, surplus goes in order to the ranks with room; owner-only when nothing is skewed."""
    from gloc3d_b200.distributed import assign_pairs, partition_pairs, shard_bounds

    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 8):
        rows, nq, k = 4000, 37, 25
        b = np.asarray(shard_bounds(rows, world))
        # skewed: most candidates of a query come from one shard (the place it revisits)
        home = rng.integers(0, world, nq)
        idx = np.where(rng.random((nq, k)) < 0.8,
                       b[home][:, None] + rng.integers(0, rows // world, (nq, k)),
                       rng.integers(0, rows, (nq, k))).astype(np.uint64)
        idx[3, 7] = rows + 5                        # a slot without a row (fewer than k rows in the job)
        ver = assign_pairs(idx, b)
        flat = idx.reshape(-1)
        n = int((flat < rows).sum())
        quota = -(-n // world)
        assert ver[3 * k + 7] == -1 and (ver >= 0).sum() == n
        counts = np.bincount(ver[ver >= 0], minlength=world)
        assert counts.sum() == n and counts.max() <= quota
        for r in range(world):
            own = partition_pairs(idx, b[r], b[r + 1])
            kept = own[ver[own] == r]
            # a rank keeps a PREFIX of its own pairs, and takes foreign pairs only if it owns less than the quota
            assert np.array_equal(kept, own[:len(kept)]) and len(kept) == min(len(own), quota)
            foreign = np.nonzero(ver == r)[0]
            foreign = foreign[~np.isin(foreign, own)]
            assert len(foreign) == 0 or len(own) < quota
        if world == 1:
            assert (ver[ver >= 0] == 0).all()
    # nothing to balance: exactly quota pairs per owner -> everybody verifies its own
    b = np.asarray(shard_bounds(800, 4))
    idx = np.concatenate([np.full(10, b[r] + 1) for r in range(4)]).astype(np.uint64)
    assert np.array_equal(assign_pairs(idx, b), np.repeat(np.arange(4), 10))
    # deterministic
    assert np.array_equal(assign_pairs(idx[::-1].copy(), b), np.repeat(np.arange(4)[::-1], 10))
