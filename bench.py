#!/usr/bin/env python
"""bench.py -- headline benchmark of the global-localization query path.

    python bench.py --gpus N --steps K --warmup W [--workload retrieval|verify|stream] [--impl reference]

Metric (BASELINE.json): global-localization queries/s.  One "step" = one pass of the hot
path over one batch of synthetic input.

Workload "retrieval" (default; BASELINE.json configs[1], the configuration the metric is
quoted on that fits one GPU): 100k-descriptor database (512-d float32), 10k-query batch per
GPU, top-25 exact L2.  N > 1, per-GPU work (10k queries x 100k rows) fixed as N grows
("scaling": "weak"; value = all N x 10k queries answered by the job / max-over-ranks device time):
  --sharding queries (headline): the 205 MB database is replicated, every rank answers its
      own 10k-query slice; the path partitions into independent queries, no collective;
  --sharding db: the north-star protocol for databases that do not fit one GPU -- rows
      sharded over the ranks, every rank answers all N x 10k queries on its shard, ONE NCCL
      all-gather of the local top-k lists, K4 merge on every rank.
Both are measured in every N > 1 run; the non-headline one is reported under "other_sharding".

Workload "verify" (configs[2]): 25 candidate grids per query, 361 yaw bins, +-100 cells at
0.2 m, depth 5, 800x800 BEV grids; pairs are split across ranks, no collective.

Workload "stream": the reference's own call pattern -- ONE query per call -- against a resident
1M-descriptor database; HBM-bound (every call streams the 2 GB database once).

`value`  : inputs resident in HBM, CUDA-event timed.  `e2e`: the same metric through the
C ABI / public API with HOST (pinned) buffers, H2D + D2H inside the timed region.
`--impl reference`: the reference's own CPU implementation (its nanoflann compiled from
/root/reference into oracle/_ref, or the oracle port for verify) on all host threads.
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DIM, K_NN = 512, 25
DB_ROWS, Q_PER_GPU = 100_000, 10_000
METRIC, UNIT = "global-loc queries/s", "queries/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (20 ms period; the
    timed regions here last tens of milliseconds, so the sampler keeps running while a
    trailing burst of the same steps keeps the GPU under the same load -- see `hold`)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


DUP_RUN = 8   # database rows form runs of near-duplicates (consecutive KITTI frames); --dup-run 0: iid rows


def hold_steps(ms_per_step: float) -> int:
    """Extra steps that keep the measured load up for ~0.25 s after the timed region so that
    the 20 ms clock sampler sees it.  A pure function of a value that is identical on every
    rank (the max-reduced step time): all ranks run the SAME number of steps, which matters
    when a step contains a collective."""
    return int(min(4000, max(1, round(250.0 / max(ms_per_step, 0.05)))))


def make_retrieval_inputs(world: int, rows: int = 0, queries: int = 0):
    from gloc3d_b200 import synth

    rows = rows or DB_ROWS
    db = (synth.make_descriptors(rows, DIM, seed=1234, dup_run=DUP_RUN) if rows <= 200_000
          else synth.make_descriptors_mt(rows, DIM, seed=1234, dup_run=max(DUP_RUN, 1)))
    nq = queries or Q_PER_GPU * world
    qa = synth.make_queries(db, nq // 2, seed=5678)                    # set A: independent
    qb = synth.make_queries(db, nq - nq // 2, seed=5679, sigma=0.01)   # set B: perturbed copies
    return db, np.ascontiguousarray(np.concatenate([qa, qb]))


# ------------------------------------------------------------------ retrieval
def run_retrieval(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import gloc3d_b200 as g
    from gloc3d_b200.distributed import shard_bounds

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    db, q_all = make_retrieval_inputs(world, args.rows, args.queries)
    mode = {"auto": g.KNN_AUTO, "exact": g.KNN_EXACT_SCAN, "shortlist": g.KNN_SHORTLIST}[args.mode]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def reduce_ranks(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    rows_total = args.rows or DB_ROWS
    strong = bool(args.queries)             # a fixed total batch (configs[3]) instead of 10k per GPU
    comm = None
    if world > 1:
        from gloc3d_b200.distributed import Comm

        comm = Comm.from_torch(local_rank)

    def measure(sharding: str, steps: int, warmup: int):
        """sharding "queries": every rank holds the whole DB and answers its own query slice (no
        data-path collective).  "db": the north-star protocol through the C ABI
        (gloc_knn_query_sharded*): rows sharded; every rank uploads ITS slice of the batch, the
        queries are all-gathered over NVLink, every rank searches the whole batch on its shard, the
        local top-k lists go to the rank that owns the query (all-to-all) and are merged there."""
        from gloc3d_b200.distributed import query_sharded_device, query_sharded_host

        nq_job = args.queries if strong else Q_PER_GPU * world
        nq = nq_job // world                  # queries this rank uploads / downloads per step
        q = np.ascontiguousarray(q_all[rank * nq:(rank + 1) * nq])
        db_sharded = sharding == "db" and world > 1
        if db_sharded:
            b = shard_bounds(rows_total, world)
            lo, hi = b[rank], b[rank + 1]
        else:
            lo, hi = 0, rows_total
        ix = g.KnnIndex(DIM, local_rank)
        ix.set_db(db[lo:hi])
        ix.set_index_offset(lo)
        ix.set_mode(mode)
        q_dev = torch.from_numpy(q).to(dev)
        q_pin = torch.from_numpy(q).pin_memory()
        oi_pin = torch.empty((nq, K_NN), dtype=torch.int64).pin_memory()
        od_pin = torch.empty((nq, K_NN), dtype=torch.float32).pin_memory()
        oi_dev = torch.empty((nq, K_NN), dtype=torch.int64, device=dev)
        od_dev = torch.empty((nq, K_NN), dtype=torch.float32, device=dev)

        def dev_step():
            if db_sharded:
                return query_sharded_device(ix, comm, q_dev, K_NN, False, oi_dev, od_dev)
            return ix.query_device(q_dev, K_NN, oi_dev, od_dev)

        # ---- device-resident timing (value)
        for _ in range(warmup):
            out = dev_step()
        barrier()
        ix.set_profiling(True)
        launches0 = ix.stats().kernel_launches
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            out = dev_step()
        e1.record()
        barrier()
        dom_ms, dom_n = ix.profile()
        ix.set_profiling(False)
        st = ix.stats()
        ms_total = reduce_ranks(e0.elapsed_time(e1), dist.ReduceOp.MAX)
        for _ in range(hold_steps(ms_total / steps)):   # every rank: the same count (see hold_steps)
            dev_step()
        barrier()
        clk = clocks.stop() if rank == 0 else None
        launches = int(reduce_ranks(float(st.kernel_launches - launches0), dist.ReduceOp.SUM))
        ms_per_step = ms_total / steps

        # ---- end to end with host buffers (e2e): pinned H2D of the queries, D2H of the result
        def e2e_step():
            if db_sharded:
                query_sharded_host(ix, comm, q_pin.data_ptr(), nq, K_NN, oi_pin.data_ptr(), od_pin.data_ptr())
            else:
                ix.query_ptr(q_pin.data_ptr(), nq, K_NN, oi_pin.data_ptr(), od_pin.data_ptr())

        for _ in range(max(1, min(warmup, 3))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        torch.cuda.synchronize(dev)
        e2e_ms = reduce_ranks((time.perf_counter() - t0) * 1e3, dist.ReduceOp.MAX) / steps
        barrier()
        # the timed path and the host-buffer path return the same bits
        assert np.array_equal(out[0].cpu().numpy(), oi_pin.numpy()) and \
            np.array_equal(out[1].cpu().numpy(), od_pin.numpy())
        res = dict(sharding=sharding, nq=nq, nq_job=nq_job, rows=hi - lo, ms_per_step=ms_per_step,
                   value=nq_job / (ms_per_step * 1e-3), e2e_ms=e2e_ms, e2e_value=nq_job / (e2e_ms * 1e-3),
                   h2d=int(q.nbytes) * world, d2h=int(nq * K_NN * 12) * world, launches=launches, clk=clk,
                   dom_ms=dom_ms, dom_n=dom_n, st=st, steps=steps, db_sharded=db_sharded,
                   nq_searched=nq * world if db_sharded else nq,
                   sample=(q, out[0].cpu().numpy().view(np.uint64), out[1].cpu().numpy()))
        ix.close()
        return res

    primary = args.sharding if world > 1 else "queries"
    r = measure(primary, args.steps, args.warmup)
    other = None
    if world > 1:   # the other protocol, measured in the same run for the record
        other = measure("db" if primary == "queries" else "queries", min(args.steps, 5), 3)
    if rank != 0:
        return None
    st = r["st"]
    last_mode = int(st.last_mode)
    # the library may split a batch into several launches: algorithmic work per LAUNCH
    launches_per_step = max(r["dom_n"], 1) / r["steps"]
    flops = 2.0 * r["nq_searched"] * r["rows"] * DIM / launches_per_step
    alg_bytes = r["rows"] * DIM * 4 + (r["nq_searched"] * DIM * 4 + r["nq_searched"] * K_NN * 12) / launches_per_step
    avg_ms = r["dom_ms"] / max(r["dom_n"], 1)
    peak_tf = peaks["bf16_tflops"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp) and world == 1:
        traffic = json.load(open(tp)).get({1: "exact_scan", 2: "shortlist_gemm"}.get(last_mode, ""), None)
    nq_job = r["nq_job"]
    shard_txt = {"queries": f"queries/{world} (each rank: whole DB replicated, its own query slice; no collective)",
                 "db": f"rows/{world}: queries all-gathered over NVLink, local top-k per shard, all-to-all of the lists to "
                       f"the query's owner, K4 merge there (gloc_knn_query_sharded, NCCL inside libgloc3d.so)"}
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32" if last_mode == g.KNN_EXACT_SCAN else "fp16 tensor shortlist + f32 exact re-rank",
        "data": "synthetic",
        "config": {"workload": (f"configs[3]: {rows_total} x 512-d f32 descriptor DB, {nq_job}-query batch (whole job), "
                                if strong else
                                f"configs[1]: {rows_total} x 512-d f32 descriptor DB, 10k-query batch per GPU, ") +
                               "top-25 exact L2 retrieval (bit-exact vs nanoflann)",
                   "db_rows": rows_total, "queries_per_step": nq_job, "k": K_NN, "dim": DIM,
                   "sharding": shard_txt[r["sharding"]] if world > 1 else "none",
                   "strategy": {1: "exact_scan", 2: "tensor_shortlist"}.get(last_mode, str(last_mode)),
                   "gemm_variant": "cta_pair (GLOC_KNN_PAIR)" if os.environ.get("GLOC_KNN_PAIR", "0") not in ("", "0")
                                   else "one CTA per SM",
                   "l2": "inputs larger than L2 (DB + query batch > 126 MB per rank); no flush"},
        "e2e": {"value": r["e2e_value"], "unit": UNIT, "ms_per_step": r["e2e_ms"],
                "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
        "gpu_launches": r["launches"],
        "clocks": r["clk"],
        "roofline": {"bound": "tensor", "kernel": {1: "knn_exact_scan_kernel", 2: "knn_shortlist_gemm_kernel"}.get(last_mode),
                     "achieved": flops / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else None,
                     "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (flops / (avg_ms * 1e-3) / 1e12) / peak_tf if avg_ms > 0 else None,
                     "traffic": traffic, "peak_source": peaks["source"] + " (burst bf16)",
                     "kernel_ms": avg_ms, "kernel_launches_timed": r["dom_n"],
                     "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": alg_bytes,
                     "hbm_frac_of_algorithmic_bytes": (alg_bytes / (avg_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if avg_ms > 0 else None},
        "stats": {"fallback_queries": int(st.fallback_queries), "shortlist_rows_per_query":
                  (st.shortlist_rows / max(st.shortlist_queries, 1))},
    }
    # a sample of the timed step's answers against the brute-force oracle (bit-exact indices and distances)
    from oracle import pyoracle as po
    sq, sidx, sd2 = r["sample"]
    n_chk = min(32, sq.shape[0])
    ref_idx, ref_d2 = po.knn(db, sq[:n_chk], K_NN, nthreads=os.cpu_count() or 1)
    assert np.array_equal(sidx[:n_chk], ref_idx) and np.array_equal(sd2[:n_chk].view(np.uint32), ref_d2.view(np.uint32)), \
        "retrieval: result differs from the oracle"
    line["stats"]["checked_against_oracle"] = f"{n_chk} queries of rank 0's slice, indices and distances bit-equal"
    if other is not None:
        line["other_sharding"] = {"sharding": shard_txt[other["sharding"]], "value": other["value"],
                                  "ms_per_step": other["ms_per_step"], "e2e_value": other["e2e_value"],
                                  "steps": other["steps"]}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_retrieval(db, q_all, budget_s=args.cpu_budget)
    return line


def cpu_baseline_retrieval(db, q, budget_s: float):
    """The reference's nanoflann (oracle/_ref) on all host threads, bounded sample."""
    from oracle import pyoracle as po

    cores = os.cpu_count() or 1
    if po.have_ref():
        t0 = time.perf_counter()
        tree = po.RefTree(db, 10)
        build_s = time.perf_counter() - t0
        run = lambda qs: tree.query(qs, K_NN, nthreads=cores)  # noqa: E731
        kind = "reference"
    else:
        build_s = 0.0
        run = lambda qs: po.knn(db, qs, K_NN, nthreads=cores)  # noqa: E731
        kind = "port"
    n0 = min(q.shape[0], cores)
    t0 = time.perf_counter()
    run(q[:n0])
    per_q = (time.perf_counter() - t0) / n0
    n = int(max(n0, min(q.shape[0], budget_s / max(per_q, 1e-9))))
    n = max(cores, n // cores * cores)
    t0 = time.perf_counter()
    run(q[:n])
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n} of the step's queries against the full 100k DB, nanoflann KD-tree "
                      f"(leaf 10) built once in {build_s:.2f} s (not counted), {cores} threads"}


def run_reference_retrieval(args):
    from oracle import pyoracle as po

    db, q = make_retrieval_inputs(1)
    cores = os.cpu_count() or 1
    if po.have_ref():
        tree = po.RefTree(db, 10)
        run = lambda qs: tree.query(qs, K_NN, nthreads=cores)  # noqa: E731
        kind = "reference"
    else:
        run = lambda qs: po.knn(db, qs, K_NN, nthreads=cores)  # noqa: E731
        kind = "port"
    n0 = min(q.shape[0], cores)
    t0 = time.perf_counter()
    run(q[:n0])
    per_q = (time.perf_counter() - t0) / n0
    total_steps = args.steps + args.warmup
    n = int(min(q.shape[0], max(cores, (args.cpu_budget * 4 / total_steps) / max(per_q, 1e-9))))
    n = max(cores, n // cores * cores)
    for _ in range(args.warmup):
        run(q[:n])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(q[:n])
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    v = n / (ms * 1e-3)
    sample = (f"{n} queries/step of the 10k batch against the full 100k DB; nanoflann KD-tree (leaf 10), "
              f"{cores} host threads")
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 100k x 512-d f32 descriptor DB, 10k-query batch per GPU, "
                                   "top-25 exact L2 retrieval (bit-exact vs nanoflann)",
                       "db_rows": DB_ROWS, "k": K_NN, "dim": DIM},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}



# --------------------------------------------------------------------- stream
STREAM_ROWS = 1_000_000


def run_stream(args, rank, world, local_rank):
    """Online localisation as the reference issues it: ONE query per call against a resident
    1M-descriptor database (configs[3] size; --rows 5000000 for configs[4]).  HBM-bound: every
    step streams the float32 database once.  N > 1: rows sharded over the ranks, the same query
    on every rank, per-GPU streaming scan of the shard, NCCL all-gather of the N x 25-entry lists,
    merge (gloc_knn_query_sharded, replicated) -- strong scaling of a single query's latency."""
    import torch
    import torch.distributed as dist

    import gloc3d_b200 as g
    from gloc3d_b200 import synth
    from gloc3d_b200.distributed import Comm, query_sharded_device, query_sharded_host, shard_bounds

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    nq = max(1, min(4, args.stream_queries))
    rows_total = args.rows or STREAM_ROWS
    b = shard_bounds(rows_total, world)
    lo, hi = b[rank], b[rank + 1]
    # every rank generates its own shard only (block-seeded streams) plus the first block, where the
    # queries' source rows live
    db = synth.make_descriptors_mt(rows_total, DIM, seed=1234, dup_run=max(DUP_RUN, 1), rows=(lo, hi))
    src = synth.make_descriptors_mt(rows_total, DIM, seed=1234, dup_run=max(DUP_RUN, 1),
                                    rows=(0, min(synth.MT_BLOCK, rows_total)))
    qs = synth.make_queries(src, 64, seed=5678, sigma=0.01)
    comm = Comm.from_torch(local_rank) if world > 1 else None
    ix = g.KnnIndex(DIM, local_rank)
    ix.set_db(db)
    ix.set_index_offset(lo)
    q_dev = torch.from_numpy(qs).to(dev)
    q_pin = torch.from_numpy(qs).pin_memory()
    oi_pin = torch.empty((nq, K_NN), dtype=torch.int64).pin_memory()
    od_pin = torch.empty((nq, K_NN), dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def step(i):
        j = (i * nq) % (64 - nq + 1)
        if comm is not None:
            return query_sharded_device(ix, comm, q_dev[j:j + nq], K_NN, True)
        return ix.query_device(q_dev[j:j + nq], K_NN)

    for i in range(args.warmup):
        out = step(i)
    barrier()
    ix.set_profiling(True)
    l0 = ix.stats().kernel_launches
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        out = step(i)
    e1.record()
    barrier()
    dom_ms, dom_n = ix.profile()
    ix.set_profiling(False)
    st = ix.stats()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    for _ in range(hold_steps(ms_total / args.steps)):   # every rank: the same count (see hold_steps)
        step(0)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = (st.kernel_launches - l0) * world
    ms = ms_total / args.steps

    def e2e_step(i):
        j = (i * nq) % (64 - nq + 1)
        if comm is not None:
            query_sharded_host(ix, comm, q_pin[j:j + nq].data_ptr(), nq, K_NN, oi_pin.data_ptr(), od_pin.data_ptr(), True)
        else:
            ix.query_ptr(q_pin[j:j + nq].data_ptr(), nq, K_NN, oi_pin.data_ptr(), od_pin.data_ptr())

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    barrier()
    assert np.array_equal(out[0].cpu().numpy(), oi_pin.numpy()) and \
        np.array_equal(out[1].cpu().numpy(), od_pin.numpy())
    # the sharded answer is the single-GPU answer: checked against the brute-force oracle on rank 0
    if rank != 0:
        ix.close()
        return None
    avg_ms = dom_ms / max(dom_n, 1)
    alg_bytes = (hi - lo) * DIM * 4 + nq * DIM * 4 + nq * K_NN * 12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp) and rows_total == STREAM_ROWS and world == 1:
        traffic = json.load(open(tp)).get("stream_scan", None)
    line = {
        "metric": METRIC, "value": nq / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"online localisation: {nq} query per call (the reference's call pattern, "
                               f"loop_detector.cpp:42-45) against a resident {rows_total} x 512-d f32 database, "
                               "top-25 exact L2 (bit-exact vs nanoflann)",
                   "db_rows": rows_total, "queries_per_step": nq, "k": K_NN, "dim": DIM,
                   "sharding": "none" if world == 1 else
                               f"rows/{world}: per-GPU streaming scan of {hi - lo} rows, all-gather of the {world} x {K_NN}-entry "
                               "lists, merge (gloc_knn_query_sharded, NCCL inside libgloc3d.so)",
                   "strategy": "stream_scan",
                   "l2": f"every step streams the {(hi - lo) * DIM * 4 / 1e9:.2f} GB shard (> L2); no flush needed"},
        "e2e": {"value": nq / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": nq * DIM * 4 * world, "d2h_bytes_per_step": nq * K_NN * 12 * world},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "knn_stream_kernel (per GPU, its shard)",
                     "achieved": alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms else None,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": (alg_bytes / (avg_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if avg_ms else None,
                     "traffic": traffic, "peak_source": peaks["source"] + " (copy, read+write)",
                     "kernel_ms": avg_ms, "kernel_launches_timed": dom_n,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "whole_step_hbm_frac": (alg_bytes / (ms * 1e-3) / 1e9) / peaks["hbm_gbs"],
                     "aggregate_gbs_all_gpus": rows_total * DIM * 4 / (ms * 1e-3) / 1e9},
    }
    from oracle import pyoracle as po
    j = ((args.steps - 1) * nq) % (64 - nq + 1)
    if rows_total <= 2_000_000:
        full = db if world == 1 else synth.make_descriptors_mt(rows_total, DIM, seed=1234, dup_run=max(DUP_RUN, 1))
        ref_idx, ref_d2 = po.knn(full, qs[j:j + nq], K_NN, nthreads=os.cpu_count() or 1)
        assert np.array_equal(oi_pin.numpy().view(np.uint64), ref_idx) and \
            np.array_equal(od_pin.numpy().view(np.uint32), ref_d2.view(np.uint32)), "stream: result differs from the oracle"
        line["stats"] = {"checked_against_oracle": f"{nq} queries of the last step, indices and distances bit-equal"}
    else:
        # too large to regenerate on one host: the nearest row of the source block must lead the global answer
        s_idx, s_d2 = po.knn(src, qs[j:j + nq], 1, nthreads=os.cpu_count() or 1)
        assert np.array_equal(oi_pin.numpy().view(np.uint64)[:, 0], s_idx[:, 0]) and \
            np.array_equal(od_pin.numpy().view(np.uint32)[:, 0], s_d2[:, 0].view(np.uint32)), "stream: top-1 differs"
        line["stats"] = {"checked": f"top-1 of {nq} queries equals the brute-force nearest row of their source block"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_stream(db, qs, args.cpu_budget)
    ix.close()
    return line


def cpu_baseline_stream(db, q, budget_s: float):
    """The reference's nanoflann on the 1M-row database, all host threads, bounded sample."""
    from oracle import pyoracle as po

    cores = os.cpu_count() or 1
    if po.have_ref():
        t0 = time.perf_counter()
        tree = po.RefTree(db, 10)
        build_s = time.perf_counter() - t0
        run = lambda qs: tree.query(qs, K_NN, nthreads=cores)  # noqa: E731
        kind = "reference"
    else:
        build_s = 0.0
        run = lambda qs: po.knn(db, qs, K_NN, nthreads=cores)  # noqa: E731
        kind = "port"
    n = min(q.shape[0], cores)
    t0 = time.perf_counter()
    run(q[:n])
    dt = time.perf_counter() - t0
    reps = int(max(1, min(8, budget_s / max(dt, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(reps):
        run(q[:n])
    dt = (time.perf_counter() - t0) / reps
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n} queries in parallel (one per thread) against the full 1M DB, nanoflann KD-tree "
                      f"(leaf 10) built once in {build_s:.1f} s (not counted), {cores} threads; a single "
                      f"reference call answers one query on one core: {n / dt / cores:.2f} q/s"}

def run_reference_stream(args):
    from gloc3d_b200 import synth

    db = synth.make_descriptors_mt(args.rows or STREAM_ROWS, DIM, seed=1234, dup_run=max(DUP_RUN, 1))
    q = synth.make_queries(db, 64, seed=5678, sigma=0.01)
    cb = None
    for _ in range(max(1, args.warmup)):
        cb = cpu_baseline_stream(db, q, 0.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cb = cpu_baseline_stream(db, q, 0.0)
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "online localisation: 1 query per call against a resident 1M x 512-d f32 "
                                   "database, top-25 exact L2", "db_rows": STREAM_ROWS, "k": K_NN, "dim": DIM},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                        "d2h_bytes_per_step": 0}, "gpu_launches": 0}


# --------------------------------------------------------------------- verify
VER = dict(nx=800, ny=800, res=0.2, n_lin=100, n_ang=180, step=2 * np.pi / 360, depth=5,
           min_score=0.3, cands=25)


def make_verify_inputs(n_queries: int, n_maps: int = 32):
    from gloc3d_b200 import synth

    mx, my = synth.centered_limits(VER["nx"], VER["ny"], VER["res"])
    maps = [synth.make_bev_grid(VER["nx"], VER["ny"], seed=2222 + i) for i in range(n_maps)]
    rng = np.random.default_rng(3333)
    scans, pairs = [], []
    for qi in range(n_queries):
        m = qi % n_maps
        yaw, dx, dy = rng.uniform(-np.pi, np.pi), rng.uniform(-18, 18), rng.uniform(-18, 18)
        scans.append(synth.planted_scan(maps[m], VER["res"], mx, my, yaw, dx, dy, dropout=0.2,
                                        jitter_cells=1.0, seed=3333 + qi))
        for c in range(VER["cands"]):          # candidate 0 is the planted map, 24 are wrong places
            pairs.append(((m + c) % n_maps, qi))
    return maps, mx, my, scans, pairs


def verify_traffic(key="csm_coarse_bits"):
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    return json.load(open(tp)).get(key, None) if os.path.exists(tp) else None


def run_verify(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import gloc3d_b200 as g

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    nq = args.verify_queries * world
    maps, mx, my, scans, pairs = make_verify_inputs(nq)
    mine = pairs[rank::world]                      # pairs are independent: no collective
    st = g.CsmStore(local_rank)
    gids = [st.add_grid_u8(m, VER["res"], mx, my) for m in maps]
    gi = [gids[p[0]] for p in mine]
    si = [p[1] for p in mine]
    inits = [(0.0, 0.0, 0.0)] * len(mine)

    gpu_depth = args.verify_depth or VER["depth"]

    def step():
        return st.match_batch(scans, gi, si, inits, VER["n_lin"], VER["n_ang"], VER["step"],
                              gpu_depth, VER["min_score"])

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        out = step()
    st.set_profiling(True)
    l0 = st.stats().kernel_launches
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step()
    dt = (time.perf_counter() - t0) * 1e3
    barrier()
    dom_ms, dom_n = st.profile()
    st.set_profiling(False)
    launches = st.stats().kernel_launches - l0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    for _ in range(hold_steps(dt / args.steps)):   # every rank: the same count (see hold_steps)
        step()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    if rank != 0:
        return None
    ms = dt / args.steps
    found = sum(r.found for r in out)
    P = float(np.mean([s.shape[0] for s in scans]))
    side = (2 * VER["n_lin"]) // (1 << (VER["depth"] - 1)) + 1
    lookups = (2 * VER["n_ang"] + 1) * side * side * P * len(mine)      # coarse level, per launch
    peaks = load_peaks()
    # the scorer's binding unit: random 8-byte shared-memory loads, one per (rotation, point, candidate
    # row); ceiling measured live by the library's micro-benchmark
    lsu = None
    try:
        import ctypes

        from gloc3d_b200 import _lib
        peak = ctypes.c_double()
        if _lib.lib().gloc_bench_smem_gather(local_rank, ctypes.byref(peak)) == 0 and peak.value > 0:
            lsu = {"peak_random_lds64_per_s": peak.value}
    except Exception as e:   # the ceiling is an annotation: never fail the bench over it
        lsu = {"error": str(e)}
    hbm_bytes = len(set(gi)) * VER["nx"] * VER["ny"] + sum(s.shape[0] for s in scans) * 12 + len(mine) * 24
    avg_ms = dom_ms / max(dom_n, 1)
    line = {
        "metric": METRIC, "value": nq / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 sums -> f32 score", "data": "synthetic",
        "config": {"workload": "configs[2]: verification sweep, 25 candidate grids/query, 361 yaw bins, "
                               "+-100 cells @0.2 m, depth 5, 800x800 BEV grids",
                   "queries_per_step": nq, "pairs_per_step": len(pairs),
                   "timing": "host wall clock around the synchronous C-ABI call (H2D of scans, D2H of "
                             "results inside); value == e2e for this workload",
                   "l2": "grid stacks (3.2 MB each, 32 maps) are L2-resident by design"},
        "e2e": {"value": nq / (ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(sum(s.nbytes for s in scans)) * world,
                "d2h_bytes_per_step": len(pairs) * 8},
        "gpu_launches": int(launches) * world, "clocks": clk,
        "roofline": None,
        "stats": {"found": int(found), "pairs": len(mine)},
    }
    hbm = {"algorithmic_bytes_per_launch": hbm_bytes,
           "achieved_gbs": hbm_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms else None,
           "frac_of_peak": (hbm_bytes / (avg_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if avg_ms else None}
    if lsu and "peak_random_lds64_per_s" in lsu and avg_ms:
        # as the localize workload: the scorer is bound by shared-memory gathers, one 8-byte load per (rotation,
        # point, candidate row) -- per PAIR of candidate rows in the paired plane layout 800-cell grids get
        wc = 1 << (VER["depth"] - 1)
        c_lo, c_hi = VER["n_lin"] // wc, (VER["nx"] + wc - 2 + VER["n_lin"]) // wc
        paired = side <= 14 and max(c_lo, c_hi - 31) - c_lo <= 33 - side and not os.environ.get("GLOC_CSM_NO_PAIRED")
        loads = lookups / side / side * ((side + 2) // 2 if paired else side)
        ach = loads / (avg_ms * 1e-3)
        line["roofline"] = {"bound": "lsu", "kernel": "csm_coarse_bits_kernel", "achieved": ach / 1e9,
                            "peak": lsu["peak_random_lds64_per_s"] / 1e9, "unit": "G LDS.64/s",
                            "frac": ach / lsu["peak_random_lds64_per_s"], "traffic": None, "kernel_ms": avg_ms,
                            "peak_source": "measured live: gloc_bench_smem_gather (random 8-byte shared-memory loads, chip-wide)",
                            "algorithmic_lookups_per_launch": lookups, "algorithmic_lds64_per_launch": loads,
                            "note": "200 pairs = 400 CTAs do not fill the chip for long: see the localize workload "
                                    "(3200 pairs per launch) for the scorer's steady state", "hbm": hbm}
    else:
        line["roofline"] = {"bound": "hbm", "kernel": "csm_coarse_bits_kernel", "achieved": hbm["achieved_gbs"],
                            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm["frac_of_peak"], "traffic": verify_traffic(),
                            "kernel_ms": avg_ms, "lsu_gather": lsu,
                            "note": "the LSU ceiling could not be measured; compulsory HBM bytes are tiny"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_verify(maps, mx, my, scans, pairs, args.cpu_budget)
    st.close()
    return line


def cpu_baseline_verify(maps, mx, my, scans, pairs, budget_s):
    from oracle import pyoracle as po

    cores = os.cpu_count() or 1
    n = min(len(pairs), max(cores, int(budget_s / 0.7) * cores))
    sel = pairs[:n]
    t0 = time.perf_counter()
    po.csm_match_batch([maps[p[0]] for p in sel], VER["res"], mx, my, VER["depth"],
                       [scans[p[1]] for p in sel], [(0, 0, 0)] * n, VER["n_lin"], VER["n_ang"],
                       VER["step"], VER["min_score"], 0, cores)
    dt = time.perf_counter() - t0
    return {"value": (n / VER["cands"]) / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} (grid, scan) pairs = {n / VER['cands']:.1f} queries, branch-and-bound restatement "
                      f"of registration/2d (oracle/csm_oracle.c), {cores} threads"}


def run_reference_verify(args):
    cores = os.cpu_count() or 1
    maps, mx, my, scans, pairs = make_verify_inputs(max(1, cores // VER["cands"] + 1))
    from oracle import pyoracle as po

    n = min(len(pairs), cores)
    sel = pairs[:n]

    def run():
        po.csm_match_batch([maps[p[0]] for p in sel], VER["res"], mx, my, VER["depth"],
                           [scans[p[1]] for p in sel], [(0, 0, 0)] * n, VER["n_lin"], VER["n_ang"],
                           VER["step"], VER["min_score"], 0, cores)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    v = (n / VER["cands"]) / (ms * 1e-3)
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 sums -> f32 score", "data": "synthetic",
            "config": {"workload": "configs[2]: verification sweep, 25 candidate grids/query, 361 yaw bins, "
                                   "+-100 cells @0.2 m, depth 5, 800x800 BEV grids"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} (grid, scan) pairs per step, {cores} threads"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}



# ------------------------------------------------------------------- localize
# The metric itself (BASELINE.json): retrieval + pose verification as ONE measured path,
# global_localization.cpp:482-574 (detect_all_query -> global_registraion -> match per candidate).
LOC = dict(rows=1_000_000, grids=8192, base_grids=1024, q_per_gpu=128, nx=800, ny=800, res=0.2,
           n_lin=100, n_ang=180, step=2 * np.pi / 360, depth=5, min_score=0.3, k=K_NN, batches=24)


class LocWorld:
    """1M-descriptor database (--loc-rows; runs of 8 near-duplicate rows = consecutive frames), 8192 distinct
    800 x 800 BEV grids (1024 seeded wall/blob layouts x the 8 dihedral variants).  The rows form 8
    contiguous super-blocks (sessions); block b's rows cycle through places [1024 b, 1024 (b + 1)),
    so a contiguous row shard of 1/N of the database (N in 1, 2, 4, 8) owns whole place ranges and
    the world is the same for every N.  A query revisits the place of a random row: its descriptor
    is that row plus N(0, 0.01^2) noise, its scan is the place's grid seen from a planted pose (yaw
    in [-pi, pi), |dx|, |dy| <= 18 m, 20 % dropout, +-1 cell jitter)."""

    BLOCKS = 8

    def __init__(self, rows=None, grids=None, rank=0, world=1):
        from gloc3d_b200 import synth

        self.synth = synth
        self.rows = rows or LOC["rows"]
        self.n_grids = grids or LOC["grids"]
        assert self.rows % self.BLOCKS == 0 and self.n_grids % self.BLOCKS == 0
        self.n_base = self.n_grids // self.BLOCKS
        self.rpb = self.rows // self.BLOCKS
        self.mx, self.my = synth.centered_limits(LOC["nx"], LOC["ny"], LOC["res"])
        # a rank generates its own shard of the database only (block-seeded streams: identical to the
        # slice of the whole) plus the head of every super-block, where the queries' source rows live
        self.lo, self.hi = self.shard(rank, world)[:2]
        self.db = synth.make_descriptors_mt(self.rows, DIM, seed=1234, dup_run=8, rows=(self.lo, self.hi))
        self.src_n = min(synth.MT_BLOCK, self.rpb)
        self.src = [synth.make_descriptors_mt(self.rows, DIM, seed=1234, dup_run=8,
                                              rows=(b * self.rpb, b * self.rpb + self.src_n))
                    for b in range(self.BLOCKS)]
        self.base = [synth.make_bev_grid(LOC["nx"], LOC["ny"], seed=2222 + i) for i in range(self.n_base)]

    def grid(self, gid: int) -> np.ndarray:
        b, v = gid % self.n_base, gid // self.n_base
        g = self.base[b]
        if v & 1:
            g = g[::-1, :]
        if v & 2:
            g = g[:, ::-1]
        if v & 4:
            g = g.T
        return np.ascontiguousarray(g)

    def place_of_rows(self, r) -> np.ndarray:
        r = np.asarray(r, np.int64)
        blk = r // self.rpb
        return (blk * self.n_base + (r - blk * self.rpb) % self.n_base).astype(np.int64)

    def grid_of_row(self, lo=0, hi=None) -> np.ndarray:
        """Row -> grid table of the shard [lo, hi) in LOCAL grid ids (the shard's first place = 0)."""
        hi = self.rows if hi is None else hi
        return (self.place_of_rows(np.arange(lo, hi)) - (lo // self.rpb) * self.n_base).astype(np.int32)

    def shard(self, rank, world):
        """(first row, end row, first place, end place) of rank's contiguous shard."""
        assert self.BLOCKS % world == 0, "the sharded bench wants 1, 2, 4 or 8 GPUs"
        b0, b1 = rank * self.BLOCKS // world, (rank + 1) * self.BLOCKS // world
        return b0 * self.rpb, b1 * self.rpb, b0 * self.n_base, b1 * self.n_base

    def batch(self, b: int, nq: int):
        """Query batch b: (descriptors [nq, 512], scans list, rows)."""
        rng = np.random.default_rng([5678, b])
        blk, off = rng.integers(0, self.BLOCKS, nq), rng.integers(0, self.src_n, nq)
        rows = blk * self.rpb + off
        src = np.stack([self.src[int(bb)][int(oo)] for bb, oo in zip(blk, off)])
        q = (src + rng.standard_normal((nq, DIM)).astype(np.float32) * np.float32(0.01)).astype(np.float32)
        scans = []
        for i, r in enumerate(rows):
            yaw, dx, dy = rng.uniform(-np.pi, np.pi), rng.uniform(-18, 18), rng.uniform(-18, 18)
            scans.append(self.synth.planted_scan(self.grid(int(self.place_of_rows(r))), LOC["res"], self.mx, self.my,
                                                 yaw, dx, dy, dropout=0.2, jitter_cells=1.0, seed=3333 + 1000 * b + i))
        return q, scans, rows


def loc_config(W, nq_job, sharding):
    return {"workload": f"configs[0]+[2] at configs[3] scale: {W.rows} x 512-d f32 descriptor DB, top-25 exact L2 retrieval "
                        "+ scan-match verification of all 25 candidates per query (361 yaw bins, +-100 cells "
                        "@0.2 m, branch and bound, 800x800 BEV grids, min_score 0.3) -> located frame + pose",
            "db_rows": W.rows, "distinct_grids": W.n_grids, "queries_per_step": nq_job, "k": LOC["k"],
            "candidates_verified_per_query": LOC["k"], "dim": DIM, "sharding": sharding,
            "grid_store": "bit-packed width-1 grids only; coarser levels rebuilt on the device per batch",
            "l2": f"every step is a new query batch whose ~{int(0.88 * nq_job * LOC['k'])} distinct candidate grids (82 KB each) and "
                  "2 GB database exceed L2; no flush"}


def run_localize(args, rank, world, local_rank):
    import ctypes

    import torch
    import torch.distributed as dist

    import gloc3d_b200 as g
    from gloc3d_b200 import _lib

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    W = LocWorld(args.loc_rows, args.loc_grids, rank, world)
    nq = args.loc_queries * world        # queries of the whole job per step (weak scaling)
    lo, hi, p0, p1 = W.shard(rank, world)
    comm = None
    if world > 1:
        from gloc3d_b200.distributed import Comm

        comm = Comm.from_torch(local_rank)
    ix = g.KnnIndex(DIM, local_rank)
    ix.set_db(W.db)
    ix.set_index_offset(lo)
    st = g.CsmStore(local_rank)
    t0 = time.perf_counter()
    for gid in range(p0, p1):            # this rank's rows and their places' grids only
        st.add_grid_u8(W.grid(gid), LOC["res"], W.mx, W.my)
    t_add = time.perf_counter() - t0
    loc = g.Localizer(ix, st)
    loc.set_row_grids(W.grid_of_row(lo, hi))
    # N > 1: the ranks can read each other's grid stores over NVLink, surplus pairs move to ranks with room
    shared = world > 1 and not os.environ.get("GLOC_BENCH_NO_SHARE")
    if shared:
        try:
            loc.share_grids(comm)
        except g.GlocError as e:     # peers cannot address each other (the failure is collective): owner-only
            shared = False
            if rank == 0:
                print(f"[bench] gloc_loc_share_grids unavailable, owner-only verification: {e}", file=sys.stderr, flush=True)
    prm = loc.params(LOC["k"], LOC["n_lin"], LOC["n_ang"], LOC["step"], args.verify_depth or LOC["depth"],
                     LOC["min_score"], g.LOC_FIRST_MATCH if args.loc_policy == "first" else g.LOC_VERIFY_ALL)
    n_batches = min(LOC["batches"] if world == 1 else 6, args.steps + args.warmup)
    batches = []
    for b in range(n_batches):                      # N > 1: the same batches on every rank (collective call)
        q, scans, rows = W.batch(b, nq)
        pts, offs = g.Localizer.pack_scans(scans)
        batches.append(dict(q=q, pts=pts, offs=offs, rows=rows, scans=scans,
                            q_dev=torch.from_numpy(q).to(dev), pts_dev=torch.from_numpy(pts).to(dev),
                            q_pin=torch.from_numpy(q).pin_memory(), pts_pin=torch.from_numpy(pts).pin_memory()))
    oi = torch.empty((nq, LOC["k"]), dtype=torch.int64).pin_memory()
    od = torch.empty((nq, LOC["k"]), dtype=torch.float32).pin_memory()
    res = (_lib.LocResult * nq)()
    cand = (_lib.CsmResult * (nq * LOC["k"]))()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def step(i, host=False, keep=False):
        B = batches[i % n_batches]
        if comm is not None:
            if host:
                loc.localize_sharded_ptr(comm, B["q_pin"].data_ptr(), nq, B["pts_pin"].data_ptr(), B["offs"], prm,
                                         oi.data_ptr(), od.data_ptr(), res, cand if keep else None)
            else:
                loc.localize_sharded_ptr(comm, B["q_dev"].data_ptr(), nq, B["pts_dev"].data_ptr(), B["offs"], prm,
                                         oi.data_ptr(), od.data_ptr(), res, cand if keep else None, device=True)
        elif host:
            loc.localize_ptr(B["q_pin"].data_ptr(), nq, B["pts_pin"].data_ptr(), B["offs"], prm, oi.data_ptr(),
                             od.data_ptr(), res, cand if keep else None)
        else:
            loc.localize_ptr(B["q_dev"].data_ptr(), nq, B["pts_dev"].data_ptr(), B["offs"], prm, oi.data_ptr(),
                             od.data_ptr(), res, cand if keep else None, device=True)

    # ---- device-resident timing (value): CUDA events on the library's stream + host clock around it
    for i in range(args.warmup):
        step(i)
    barrier()
    loc.set_profiling(True)
    st.set_profiling(True)
    ix.set_profiling(True)
    l0 = st.stats().kernel_launches + ix.stats().kernel_launches + loc.stats().kernel_launches
    pv0 = loc.stats().pairs_verified
    pm0 = loc.stats().pairs_migrated
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    torch.cuda.synchronize(dev)
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    total_ms, retr_ms, calls = loc.profile()
    coarse_ms, coarse_n = st.profile()
    gemm_ms, gemm_n = ix.profile()
    loc.set_profiling(False)
    st.set_profiling(False)
    ix.set_profiling(False)
    launches = st.stats().kernel_launches + ix.stats().kernel_launches + loc.stats().kernel_launches - l0
    pairs_verified = loc.stats().pairs_verified - pv0
    pairs_max = reduce_ranks(float(pairs_verified), dist.ReduceOp.MAX)
    pairs_migrated = reduce_ranks(float(loc.stats().pairs_migrated - pm0), dist.ReduceOp.SUM)
    ms_total = reduce_ranks(wall_ms, dist.ReduceOp.MAX)        # host clock >= device events (conservative)
    dev_ms = reduce_ranks(total_ms, dist.ReduceOp.MAX)
    for i in range(hold_steps(ms_total / args.steps)):
        step(i)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    # ---- end to end with host (pinned) buffers: H2D of descriptors + scans, D2H of the results
    for i in range(2):
        step(i, host=True)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i, host=True, keep=(i == args.steps - 1))
    torch.cuda.synchronize(dev)
    e2e_ms = reduce_ranks((time.perf_counter() - t0) * 1e3, dist.ReduceOp.MAX) / args.steps
    barrier()
    last = batches[(args.warmup + args.steps - 1) % n_batches]
    got_idx = oi.numpy().copy().view(np.uint64)
    got_res = [(r.located, r.candidate, r.db_index) for r in res]
    got_cand = [c.as_tuple() for c in cand]
    launches = int(reduce_ranks(float(launches), dist.ReduceOp.SUM))
    grid_bytes, ws_bytes = st.store_bytes()
    if rank != 0:
        return None
    nq_job = nq
    ms = ms_total / args.steps
    located = sum(r[0] for r in got_res)
    # did the query find its own place?
    right_place = sum(1 for r, row in zip(got_res, last["rows"])
                      if r[0] and int(W.place_of_rows(int(r[2]))) == int(W.place_of_rows(int(row))))
    P = float(np.mean([s.shape[0] for b in batches for s in b["scans"]]))
    side = (2 * LOC["n_lin"]) // (1 << (LOC["depth"] - 1)) + 1
    S = 2 * LOC["n_ang"] + 1
    pairs_per_launch = pairs_verified / max(coarse_n, 1)
    # one LDS.64 per (rotation, point, candidate row) -- or per PAIR of candidate rows when the grids admit the
    # paired plane layout (csm_make_plan: the data columns fit two overlapping 32-column halves)
    wc = 1 << (LOC["depth"] - 1)
    c_lo, c_hi = LOC["n_lin"] // wc, (LOC["nx"] + wc - 2 + LOC["n_lin"]) // wc
    paired = side <= 14 and max(c_lo, c_hi - 31) - c_lo <= 33 - side and not os.environ.get("GLOC_CSM_NO_PAIRED")
    loads_per_point = (side + 2) // 2 if paired else side       # 13 rows, either parity of the first: 7 pairs
    lds_per_launch = S * loads_per_point * P * pairs_per_launch
    avg_coarse = coarse_ms / max(coarse_n, 1)
    lsu = {}
    try:
        peak = ctypes.c_double()
        if _lib.lib().gloc_bench_smem_gather(local_rank, ctypes.byref(peak)) == 0 and peak.value > 0:
            lsu = {"peak_random_lds64_per_s": peak.value}
    except Exception as e:   # the ceiling is an annotation: never fail the bench over it
        lsu = {"error": str(e)}
    ach = lds_per_launch / (avg_coarse * 1e-3) if avg_coarse else None
    hbm_bytes = pairs_per_launch * (LOC["nx"] * LOC["ny"] / 8 + 12 * P + 24)
    roof = {"bound": "lsu", "kernel": "csm_coarse_bits_kernel",
            "achieved": ach / 1e9 if ach else None,
            "peak": lsu.get("peak_random_lds64_per_s", 0) / 1e9 or None, "unit": "G LDS.64/s",
            "frac": (ach / lsu["peak_random_lds64_per_s"]) if ach and lsu.get("peak_random_lds64_per_s") else None,
            "traffic": verify_traffic("csm_coarse_bits_localize") if (world == 1 and nq == LOC["q_per_gpu"]) else None,
            "peak_source": "measured live: gloc_bench_smem_gather (random 8-byte shared-memory loads, chip-wide)",
            "kernel_ms": avg_coarse, "kernel_launches_timed": coarse_n,
            "algorithmic_lookups_per_launch": S * side * side * P * pairs_per_launch,
            "algorithmic_lds64_per_launch": lds_per_launch,
            "plane_layout": ("paired rows: one 8-byte load = 2 candidate rows x 32 columns; %d loads per (rotation, point)"
                             % loads_per_point) if paired else "one 8-byte load = 1 candidate row x 64 columns",
            "note": "after the paired layout the scorer is co-limited by shared-memory wavefronts and instruction issue "
                    "(profiles/r02_ncu_coarse_paired_summary.md); frac is the share of the measured random-LDS.64 ceiling",
            "hbm": {"algorithmic_bytes_per_launch": hbm_bytes,
                    "achieved_gbs": hbm_bytes / (avg_coarse * 1e-3) / 1e9 if avg_coarse else None,
                    "frac_of_peak": hbm_bytes / (avg_coarse * 1e-3) / 1e9 / peaks["hbm_gbs"] if avg_coarse else None},
            "stages_ms_per_step": {"device_total": dev_ms / args.steps, "retrieval": retr_ms / args.steps,
                                   "verification": (total_ms - retr_ms) / args.steps,
                                   "coarse_scorer": coarse_ms / args.steps},
            "retrieval_gemm": {"bound": "tensor", "kernel": "knn_shortlist_gemm_kernel",
                               "kernel_ms": gemm_ms / max(gemm_n, 1), "launches": gemm_n,
                               "achieved_tflops": (2.0 * nq * (hi - lo) * DIM * args.steps / max(gemm_n, 1)) /
                                                  (gemm_ms / max(gemm_n, 1) * 1e-3) / 1e12 if gemm_ms else None,
                               "peak_tflops": peaks["bf16_tflops"],
                               "note": f"{nq} queries are one 128-row query tile at most: this launch is latency-, "
                                       "not tensor-bound; see --workload retrieval for the GEMM at batch size"}}
    line = {
        "metric": METRIC, "value": nq_job / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 tensor shortlist + f32 exact re-rank; u8/bit sums -> f32 score", "data": "synthetic",
        "config": loc_config(W, nq_job, "none" if world == 1 else
                             f"rows/{world}: every rank holds {hi - lo} descriptor rows and the {p1 - p0} map grids of their "
                             f"places; local top-k -> fused peer-memory gather + merge; a (query, candidate) pair is verified by the "
                             f"rank that owns the candidate" + ("; ranks owning more than 1/N of a step's pairs hand the surplus to "
                             "ranks with room, which read the owner's bit-packed grid over NVLink (gloc_loc_share_grids)" if shared else "") +
                             "; all-reduce of the 8-byte pair results (gloc_loc_localize_sharded)"),
        "e2e": {"value": nq_job / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(last["q"].nbytes + last["pts"].nbytes),
                "d2h_bytes_per_step": int(nq * LOC["k"] * 12 + nq * LOC["k"] * 8) * world,
                "note": "N > 1: every rank uploads the descriptors (2 KB) and scans (~52 KB) of 1/N of the queries, the parts "
                        "travel to the peers over NVLink; every rank downloads all results"},
        "gpu_launches": launches, "clocks": clk, "roofline": roof,
        "timing": "value: host clock around K synchronous C-ABI calls between device synchronisations "
                  f"(device events on the library's stream: {dev_ms / args.steps:.3f} ms/step)",
        "stats": {"policy": args.loc_policy, "located": int(located), "located_at_the_right_place": int(right_place),
                  "queries_last_step": nq, "grid_store_bytes": grid_bytes, "grid_store_workspace_bytes": ws_bytes,
                  "pairs_per_step_busiest_rank": pairs_max / args.steps, "pairs_per_step_mean": nq_job * LOC["k"] / world,
                  "pairs_migrated_per_step": pairs_migrated / args.steps,
                  "grids_added_s": t_add},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_localize(W, last, got_idx, got_cand, args.cpu_budget, check=True)[0]
    loc.close()
    st.close()
    ix.close()
    if comm is not None:
        comm.close()
    return line


def cpu_baseline_localize(W, B, gpu_idx, gpu_cand, budget_s, check, tree=None):
    """The reference's CPU path on all host threads, bounded sample of one query batch: nanoflann
    (compiled from /root/reference) against the full 1M-row database for n_r queries, the
    branch-and-bound restatement of registration/2d for n_v of their (query, candidate) pairs.  A
    query costs one retrieval + 25 verifications, all cores busy in both stages:
    q/s = 1 / (t_retrieval / n_r + 25 * t_verify / n_v).  With check=True the sample's indices and
    poses are asserted equal to the GPU's."""
    from oracle import pyoracle as po

    cores = os.cpu_count() or 1
    k = LOC["k"]
    build_s = 0.0
    if tree is None and po.have_ref():
        t0 = time.perf_counter()
        tree = po.RefTree(W.db, 10)
        build_s = time.perf_counter() - t0
    kind = "reference" if tree is not None else "port"
    n_r = min(B["q"].shape[0], cores)
    t0 = time.perf_counter()
    if tree is not None:
        ref_idx, ref_d2 = tree.query(B["q"][:n_r], k, nthreads=cores)
    else:
        ref_idx, ref_d2 = po.knn(W.db, B["q"][:n_r], k, nthreads=cores)
    t_r = time.perf_counter() - t0
    # verification sample: the first candidates of the first queries (the planted place is usually among them)
    n_v = int(max(cores, min(n_r * k, (budget_s / 1.5) * cores)))
    n_v = min(n_r * k, n_v // cores * cores)
    per_q = max(1, n_v // n_r)
    sel = [(qi, c) for qi in range(n_r) for c in range(per_q)][:n_v]
    grids = [W.grid(int(W.place_of_rows(int(ref_idx[qi, c])))) for qi, c in sel]
    scans_sel = [B["scans"][qi] for qi, _ in sel]
    if po.have_csm_ref():
        # the reference's own FastCorrelativeScanMatcher2D (registration/2d compiled unmodified), one
        # matcher -- hence one PrecomputationGridStack2D -- per (query, candidate) pair
        cells = [W.synth.level1_to_cells(gr) for gr in grids]
        t0 = time.perf_counter()
        out = po.ref_csm_match_batch(cells, LOC["res"], W.mx, W.my, LOC["depth"], scans_sel, [(0, 0, 0)] * len(sel),
                                     LOC["n_lin"], LOC["n_ang"], LOC["step"], LOC["min_score"], cores)
        t_v = time.perf_counter() - t0
        vkind = "reference"
    else:
        t0 = time.perf_counter()
        out = po.csm_match_batch(grids, LOC["res"], W.mx, W.my, LOC["depth"], scans_sel, [(0, 0, 0)] * len(sel),
                                 LOC["n_lin"], LOC["n_ang"], LOC["step"], LOC["min_score"], 0, cores)
        t_v = time.perf_counter() - t0
        vkind = "port"
    checked = None
    same_pose = 0
    if check:
        # nanoflann's order among exactly equal distances is traversal order; ours is (d2, idx)
        exact = po.knn(W.db, B["q"][:n_r], k, nthreads=cores)[0] if tree is not None else ref_idx
        assert np.array_equal(gpu_idx[:n_r], exact), "retrieval indices differ from the oracle"
        assert np.array_equal(np.sort(ref_idx, 1), np.sort(exact, 1)), "nanoflann and the brute-force oracle disagree"
        for (qi, c), o in zip(sel, out):
            r = gpu_cand[qi * k + c]
            assert r[0] == o.found, f"found flag differs at query {qi} candidate {c}"
            if o.found:
                assert np.float32(r[1]) == np.float32(o.score), f"score differs at query {qi} candidate {c}"
                same = r[5:8] == (o.pose_x, o.pose_y, o.pose_yaw)
                # the reference's std::sort leaves the order among equal scores unspecified: against it only a
                # tie may differ in pose; against the port (same tie rule as the GPU) nothing may
                assert same or vkind == "reference", f"pose differs at query {qi} candidate {c}"
                same_pose += int(same)
        checked = {"queries": n_r, "pairs": len(sel), "matched_pairs": int(sum(o.found for o in out)),
                   "matched_pairs_with_identical_pose": same_pose}
    v = 1.0 / (t_r / n_r + k * t_v / len(sel))
    return {"value": v, "unit": UNIT, "cores": cores,
            "kind": "reference" if (kind, vkind) == ("reference", "reference") else f"{kind} (retrieval) + {vkind} (verification)",
            "sample": f"{n_r} queries of the step against the full {W.rows}-row DB through nanoflann (KD-tree, leaf 10, "
                      f"built once in {build_s:.1f} s, not counted): {t_r:.2f} s; {len(sel)} of their (query, candidate) "
                      f"pairs through {'the reference FastCorrelativeScanMatcher2D (oracle/_ref/libcsm_ref.so)' if vkind == 'reference' else 'oracle/csm_oracle.c'}: {t_v:.2f} s; {cores} threads in both stages; "
                      f"q/s = 1 / (t_r/{n_r} + {k} * t_v/{len(sel)})",
            "retrieval_qps": n_r / t_r, "verification_pairs_per_s": len(sel) / t_v,
            "parity_checked_against_gpu": checked}, tree


def run_reference_localize(args):
    W = LocWorld(args.loc_rows, args.loc_grids)
    B = dict(zip(("q", "scans", "rows"), W.batch(0, min(args.loc_queries, os.cpu_count() or 1))))
    total = args.steps + args.warmup
    cb, tree = cpu_baseline_localize(W, B, None, None, 0.0, check=False)       # builds the tree once
    vals = []
    t0 = time.perf_counter()
    for i in range(total):
        cb, tree = cpu_baseline_localize(W, B, None, None, 0.0, check=False, tree=tree)
        if i >= args.warmup:
            vals.append(cb["value"])
        if i == args.warmup - 1:
            t0 = time.perf_counter()
    ms = (time.perf_counter() - t0) * 1e3 / max(args.steps, 1)
    v = float(len(vals) / sum(1.0 / x for x in vals))           # total queries / total time over the steps
    cb["value"] = v
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (nanoflann) + u8 sums -> f32 score", "data": "synthetic",
            "config": loc_config(W, args.loc_queries, "none"),
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


# ------------------------------------------------------------------- describe
DESC_HW = 768          # the reference's CNN input (loop_detector.cpp:144-147)
DESC_BATCH = 16        # frames per step and GPU


def make_describe_inputs(n_frames: int, seed: int = 77):
    """BEV-like planes (255 = free, 0 = occupied strokes) and hashed network weights."""
    from gloc3d_b200 import synth

    rng = np.random.default_rng(seed)
    img = np.full((n_frames, DESC_HW, DESC_HW), 255, np.uint8)
    for b in range(n_frames):
        for _ in range(300):
            y, x = int(rng.integers(0, DESC_HW)), int(rng.integers(0, DESC_HW))
            if rng.random() < 0.5:
                img[b, y, x:x + int(rng.integers(5, 80))] = 0
            else:
                img[b, y:y + int(rng.integers(5, 80)), x] = 0
    ws, bs = synth.hashed_vgg_weights(11)
    cw, cent, hid = synth.hashed_vlad_weights(64, 512, 512, 31)
    return img, ws, bs, cw, cent, hid


def describe_flops_per_frame() -> float:
    cout = (64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512)
    pool_after = (1, 3, 6, 9)
    h, cin, fl = DESC_HW, 3, 0.0
    for l, co in enumerate(cout):
        fl += 2.0 * 9 * cin * co * h * h
        cin = co
        if l in pool_after:
            h //= 2
    return fl


def cpu_describe(img, ws, bs, cw, cent, hid):
    from oracle import encoder_oracle as eo
    from oracle import vlad_oracle as vo

    return vo.netvlad_fc(eo.vgg16_features(img, ws, bs).reshape(img.shape[0], 512, -1), cw, cent, hid)


def run_describe(args, rank, world, local_rank):
    """BEV plane -> VGG16 encoder -> NetVLAD_fc head -> 512-d descriptor, batched
    and device-resident.  Frames are independent: N > 1 runs replicas on their own frames."""
    import torch
    import torch.distributed as dist

    import gloc3d_b200 as g

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    img, ws, bs, cw, cent, hid = make_describe_inputs(DESC_BATCH, seed=77 + rank)
    ex = g.DescriptorExtractor(ws, bs, cw, cent, hid, height=DESC_HW, width=DESC_HW, device=local_rank)
    d_img = torch.from_numpy(img).to(dev)
    d_desc = torch.empty((DESC_BATCH, 512), dtype=torch.float32, device=dev)
    img_pin = torch.from_numpy(img).pin_memory()
    desc_pin = torch.empty((DESC_BATCH, 512), dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def step():
        ex.describe_device(d_img.data_ptr(), DESC_BATCH, d_desc.data_ptr())

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    l0 = ex.enc.kernel_launches + ex.head.kernel_launches
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    t0 = time.perf_counter()          # the C ABI calls are synchronous: host wall clock brackets device work
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize(dev)
    ms_total = reduce_max((time.perf_counter() - t0) * 1e3)
    barrier()
    launches = (ex.enc.kernel_launches + ex.head.kernel_launches - l0) * world
    for _ in range(hold_steps(ms_total / args.steps)):   # every rank: the same count (see hold_steps)
        step()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms = ms_total / args.steps

    def e2e_step():
        d_img.copy_(img_pin, non_blocking=True)
        torch.cuda.synchronize(dev)
        step()
        desc_pin.copy_(d_desc, non_blocking=True)
        torch.cuda.synchronize(dev)

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_ms = reduce_max((time.perf_counter() - t0) * 1e3 / args.steps)
    barrier()
    if rank != 0:
        ex.close()
        return None
    frames = DESC_BATCH * world
    fl = describe_flops_per_frame() * DESC_BATCH
    line = {
        "metric": "global-loc descriptor frames/s", "value": frames / (ms * 1e-3), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16 operands, f32 accumulation", "data": "synthetic",
        "config": {"workload": f"descriptor extraction: {DESC_BATCH} BEV planes of {DESC_HW} x {DESC_HW} per GPU and step "
                               "-> VGG16 features[:-2] -> NetVLAD_fc (64 clusters) -> 512-d; hashed weights",
                   "frames_per_step": frames,
                   "timing": "host wall clock around the synchronous C-ABI calls, inputs resident in HBM",
                   "l2": "activations of one step (1.2 GB at conv1) exceed L2 many times over; no flush"},
        "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(img.nbytes) * world, "d2h_bytes_per_step": DESC_BATCH * 512 * 4 * world},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "enc_conv3x3_kernel (all 12 launches of a step together)",
                     "achieved": fl / (ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": fl / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"], "traffic": None,
                     "note": "whole-step time (encoder + head), not the convolution kernels alone: an upper "
                             "bound on their time", "algorithmic_flops_per_step": fl},
    }
    if not args.no_cpu_baseline:
        t0 = time.perf_counter()
        ref = cpu_describe(img[:1], ws, bs, cw, cent, hid)
        dt = time.perf_counter() - t0
        got = d_desc[:1].cpu().numpy()
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "1 frame through float32 torch on the host (oracle/encoder_oracle.py + "
                                          "vlad_oracle.py)",
                                "max_abs_diff_vs_gpu": float(np.abs(got - ref).max()),
                                "max_abs_ref": float(np.abs(ref).max())}
    ex.close()
    return line


def run_reference_describe(args):
    img, ws, bs, cw, cent, hid = make_describe_inputs(1)
    for _ in range(min(args.warmup, 1)):
        cpu_describe(img, ws, bs, cw, cent, hid)
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_describe(img, ws, bs, cw, cent, hid)
    ms = (time.perf_counter() - t0) * 1e3 / steps
    v = 1.0 / (ms * 1e-3)
    return {"impl": "reference", "metric": "global-loc descriptor frames/s", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"descriptor extraction: 1 BEV plane of {DESC_HW} x {DESC_HW} per step "
                                   "-> VGG16 features[:-2] -> NetVLAD_fc -> 512-d; float32 torch on the host"},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": "1 frame per step"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="localize", choices=["localize", "retrieval", "verify", "stream", "describe"])
    ap.add_argument("--loc-queries", type=int, default=LOC["q_per_gpu"], help="queries per GPU per step (localize)")
    ap.add_argument("--loc-rows", type=int, default=LOC["rows"], help="database rows (localize)")
    ap.add_argument("--loc-grids", type=int, default=LOC["grids"], help="distinct map grids (localize; multiple of 8)")
    ap.add_argument("--loc-policy", default="all", choices=["all", "first"],
                    help="all: verify every one of the 25 candidates (headline); first: the reference's order "
                         "of evaluation, stop at the first candidate that matches")
    ap.add_argument("--stream-queries", type=int, default=1, help="queries per call of the stream workload (1..4)")
    ap.add_argument("--mode", default="auto", choices=["auto", "exact", "shortlist"])
    ap.add_argument("--sharding", default="db", choices=["queries", "db"],
                    help="N > 1: 'queries' = DB replicated, queries split (no collective); 'db' = rows "
                         "sharded + all-gather top-k + merge (north-star protocol).  Both are measured; "
                         "this picks which one is the headline value.")
    ap.add_argument("--rows", type=int, default=0, help="retrieval: database rows (default 100k; configs[3]: 1000000)")
    ap.add_argument("--queries", type=int, default=0,
                    help="retrieval: total queries per step for the whole job (strong scaling; configs[3]: 100000); "
                         "default 10k per GPU (weak scaling)")
    ap.add_argument("--verify-queries", type=int, default=8, help="queries per GPU per step (verify)")
    ap.add_argument("--verify-depth", type=int, default=0,
                    help="internal branch-and-bound depth of the GPU verifier (0 = the reference's 5); "
                         "the result does not depend on it")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU baseline work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dup-run", type=int, default=8,
                    help="length of the near-duplicate runs in the synthetic database (0 = iid rows)")
    args = ap.parse_args()
    global DUP_RUN
    DUP_RUN = args.dup_run
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        line = {"localize": run_reference_localize, "retrieval": run_reference_retrieval,
                "verify": run_reference_verify, "stream": run_reference_stream,
                "describe": run_reference_describe}[args.workload](args)
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the product has no CPU path", file=sys.stderr)
        return 2
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a rank that dies must not hold the others (and the GPU box) for NCCL's default 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=300))
    try:
        fn = {"localize": run_localize, "retrieval": run_retrieval, "verify": run_verify,
              "stream": run_stream, "describe": run_describe}[args.workload]
        line = fn(args, rank, world, local_rank)
        if rank == 0:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
