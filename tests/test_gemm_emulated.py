"""The shortlist GEMM kernel (gloc3d_b200/csrc/knn_shortlist.cu) executed on the HOST: its own
source text runs with one OS thread per CUDA thread against a functional model of the hardware
it talks to (tests/cpp/gemm_emu_test.cpp: mbarriers incl. cluster-remote arrives, TMA tile loads
with the 128B swizzle, tcgen05.mma / commit / ld, tensor memory).  Both instantiations run:
<false> (one CTA per SM, the shipped and GPU-tested kernel -- the model has to reproduce its
known-good behaviour) and <true> (the CTA-pair variant, not yet run on a GPU).  Checked: the
barrier protocol terminates, every emitted group sits where its base row says (range, chunk
parity, scores = ||x||^2 - 2 q.x of exactly those rows), and the lists contain every true top-k
row."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
EXE = os.path.join(CPP, "_gemm_emu_test")
BM, BN = 128, 256


@pytest.fixture(scope="module")
def exe():
    src = open(os.path.join(ROOT, "gloc3d_b200", "csrc", "knn_shortlist.cu")).read()
    for tag in ("a", "b", "c", "k1", "k3"):
        m = re.search(r"// \[emu-%s-begin\].*?\n(.*?)// \[emu-%s-end\]" % (tag, tag), src, re.S)
        assert m, f"marker emu-{tag} missing"
        text = m.group(1)
        if tag == "c":
            text = text.replace("extern __shared__ unsigned char smem_raw[];", "unsigned char* smem_raw = t_smem_raw;")

            def asm_sub(mm):
                return "*tmem_slot = 0;" if "tcgen05.alloc" in mm.group(0) else "(void)0;"
            text, n = re.subn(r"asm volatile\(.*?\);", asm_sub, text, flags=re.S)
            assert n == 8 and "asm" not in text, n
        if tag == "k3":
            text = text.replace("extern __shared__ __align__(16) unsigned char sm_raw[];",
                                "unsigned char* sm_raw = t_smem_raw;")
            assert "extern" not in text
        open(os.path.join(CPP, f"_gemm_{tag}.inc"), "w").write(text)
    r = subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-ffp-contract=off", "-fno-strict-aliasing",
                        os.path.join(CPP, "gemm_emu_test.cpp"), "-o", EXE],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-6000:]
    return EXE


def pow2_scale_for(max_abs):
    _, e = np.frexp(np.float32(max_abs))
    return np.float32(np.ldexp(1.0, 14 - int(e)))


def prepare(nq, n_rows, dim, seed, dup=False):
    """K1 in numpy: FP16 copies after power-of-two scaling, norms, residual norms."""
    rng = np.random.default_rng(seed)
    db = (rng.standard_normal((n_rows, dim)) / np.sqrt(dim)).astype(np.float32)
    if dup:
        db[1::2] = db[0::2] + (1e-3 * rng.standard_normal(db[0::2].shape) / np.sqrt(dim)).astype(np.float32)
    q = db[rng.integers(0, n_rows, nq)] + (0.05 * rng.standard_normal((nq, dim)) / np.sqrt(dim)).astype(np.float32)
    q = q.astype(np.float32)
    sx = pow2_scale_for(np.abs(db).max())
    n_pad = (n_rows + BN - 1) // BN * BN + BN
    db_h = np.zeros((n_pad, dim), np.float16)
    db_h[:n_rows] = (db * sx).astype(np.float16)
    xn = np.full(n_pad, np.inf, np.float32)
    xn[:n_rows] = (db.astype(np.float64) ** 2).sum(1).astype(np.float32)
    dx2 = ((db.astype(np.float64) - db_h[:n_rows].astype(np.float64) / sx) ** 2).sum(1).max()
    sq = np.array([pow2_scale_for(np.abs(r).max()) for r in q], np.float32)
    q_h = (q * sq[:, None]).astype(np.float16)
    qn = (q.astype(np.float64) ** 2).sum(1).astype(np.float32)
    qe = ((q.astype(np.float64) - q_h.astype(np.float64) / sq[:, None]) ** 2).sum(1).astype(np.float32)
    stats = np.array([np.float32(xn[:n_rows].max()), 0, np.float32(dx2)], np.float32).view(np.uint32)
    return dict(db=db, q=q, db_h=db_h, q_h=q_h, xn=xn, qn=qn, qe=qe, qinv=(1 / sq).astype(np.float32),
                inv_sx=np.float32(1 / sx), stats=stats, n_pad=n_pad, sx=sx, sq=sq)


def run(exe, tmp_path, P, nq, n_rows, dim, k, cap, n_ranges, tiles_per_range, pair, workers, async_seed=None):
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.array([nq, n_rows, P["n_pad"], dim, k, cap, n_ranges, tiles_per_range, int(pair), workers, 0],
                         np.int32).tobytes())
        f.write(np.array([P["inv_sx"]], np.float32).tobytes())
        f.write(P["stats"].tobytes())
        for key in ("q_h", "db_h", "xn", "qn", "qe", "qinv"):
            f.write(np.ascontiguousarray(P[key]).tobytes())
    env = dict(os.environ)
    env.pop("GLOC_EMU_ASYNC", None)
    if async_seed is not None:
        env["GLOC_EMU_ASYNC"] = str(async_seed)
    r = subprocess.run([exe, inp, outp], capture_output=True, text=True, timeout=1500, env=env)
    assert r.returncode == 0, "emulated kernel failed (deadlock report below if any)\n" + r.stderr[-4000:]
    raw = np.fromfile(outp, np.uint32)
    lists = nq * n_ranges * 2
    o = 0
    thr = raw[o:o + nq]; o += nq
    eps2 = raw[o:o + nq].view(np.float32); o += nq
    cnt = raw[o:o + lists].reshape(nq, n_ranges * 2); o += lists
    cg = raw[o:o + lists * cap].reshape(nq, n_ranges * 2, cap); o += lists * cap
    cv = raw[o:o + lists * cap * 8].view(np.float32).reshape(nq, n_ranges * 2, cap, 8)
    return thr, eps2, cnt, cg, cv


def check(P, out, nq, n_rows, dim, k, cap, n_ranges, tiles_per_range):
    thr, eps2, cnt, cg, cv = out
    assert np.all(eps2 > 0) and np.all(cnt != 0xFFFFFFFF), "a unit did not report"
    assert np.all(cnt <= cap), "synthetic case is not meant to overflow"
    # scores the kernel should have seen: ||x||^2 + cm * <fp16 q, fp16 x>
    dots = P["q_h"].astype(np.float64) @ P["db_h"].astype(np.float64).T
    cm = (-2.0 * np.float64(P["inv_sx"]) * P["qinv"].astype(np.float64))[:, None]
    with np.errstate(invalid="ignore"):
        s_ref = P["xn"].astype(np.float64)[None, :] + cm * dots
    exact = ((P["q"][:, None, :].astype(np.float64) - P["db"][None, :, :].astype(np.float64)) ** 2).sum(2)
    n_emitted = 0
    for q in range(nq):
        rows_listed = set()
        for li in range(n_ranges * 2):
            rg, wg = li // 2, li % 2
            for i in range(int(cnt[q, li])):
                base = int(cg[q, li, i])
                assert base % 8 == 0 and rg * tiles_per_range * BN <= base < min((rg + 1) * tiles_per_range * BN,
                                                                                  P["n_pad"]), (q, li, base)
                assert ((base % BN) // 32) % 2 == wg, "chunk parity of the list's warpgroup"
                sc = cv[q, li, i].astype(np.float64)
                ref = s_ref[q, base:base + 8]
                live = np.arange(base, base + 8) < n_rows
                assert np.all(np.isinf(sc[~live])), "rows beyond the search limit must carry +inf"
                assert np.allclose(sc[live], ref[live], rtol=0, atol=2e-4 * (1 + np.abs(ref[live]).max())), \
                    (q, li, base, sc, ref)
                rows_listed.update(int(r) for r in np.arange(base, base + 8)[live])
                n_emitted += 1
        topk = np.argsort(exact[q], kind="stable")[:k]
        assert set(int(r) for r in topk) <= rows_listed, (q, sorted(set(topk) - rows_listed))
    return n_emitted


CASES = [
    # nq, n_rows, dim, k, n_ranges, tiles_per_range
    (200, 1000, 128, 10, 2, 2),     # two query tiles (second ragged), 4 database tiles (last ragged)
    (300, 700, 64, 25, 1, 3),       # three query tiles: an odd count leaves a pair with a phantom tile
    (130, 600, 512, 5, 1, 3),       # the reference's dimension: 8 k-blocks per tile, rings wrap several times
]


@pytest.mark.parametrize("async_seed", [None, 1, 2])
@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("case", CASES)
def test_emulated_gemm_epilogue(exe, tmp_path, case, pair, async_seed):
    nq, n_rows, dim, k, n_ranges, tpr = case
    P = prepare(nq, n_rows, dim, seed=nq + dim, dup=True)
    cap = 512
    # fewer workers than units, so that every worker runs several units back to back
    out = run(exe, tmp_path, P, nq, n_rows, dim, k, cap, n_ranges, tpr, pair, workers=1 if pair else 2,
              async_seed=async_seed)
    n = check(P, out, nq, n_rows, dim, k, cap, n_ranges, tpr)
    assert n > nq          # something was emitted for every query


FULL_CASES = [
    # nq, n_rows, dim, k, n_ranges, tiles_per_range
    (200, 1000, 128, 25, 2, 2),
    (64, 1024, 64, 32, 1, 4),        # the smallest batch the shortlist takes, k at its maximum, one k-block
    (129, 1500, 512, 1, 3, 2),       # k = 1, the reference's dimension, one query in the last tile
    (257, 2100, 192, 20, 2, 5),      # three query tiles (odd: a phantom tile in pair mode), ragged last range
]


@pytest.mark.parametrize("pair,case", [(False, FULL_CASES[0]), (False, FULL_CASES[1]), (False, FULL_CASES[2]),
                                       (True, FULL_CASES[0]), (True, FULL_CASES[1]), (True, FULL_CASES[3])])
def test_emulated_shortlist_path_is_bit_exact(exe, tmp_path, pair, case):
    """K1 -> K2 -> K3, all from the kernels' own source, against the nanoflann-order oracle:
    indices and distances bit for bit, no query falls back."""
    from oracle import pyoracle as po

    nq, n_rows, dim, k, n_ranges, tpr = case
    P = prepare(nq, n_rows, dim, seed=99 + nq, dup=True)
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.array([nq, n_rows, P["n_pad"], dim, k, 512, n_ranges, tpr, int(pair), 1 if pair else 2, 1],
                         np.int32).tobytes())
        f.write(np.ascontiguousarray(P["q"]).tobytes())
        f.write(np.ascontiguousarray(P["db"]).tobytes())
    env = dict(os.environ, GLOC_EMU_ASYNC="5")
    r = subprocess.run([exe, inp, outp], capture_output=True, text=True, timeout=2400, env=env)
    assert r.returncode == 0, r.stderr[-4000:]
    raw = open(outp, "rb").read()
    idx = np.frombuffer(raw, np.uint64, nq * k).reshape(nq, k)
    d2 = np.frombuffer(raw, np.float32, nq * k, offset=nq * k * 8).reshape(nq, k)
    n_ovf = np.frombuffer(raw, np.int32, 1, offset=nq * k * 12)[0]
    rows = np.frombuffer(raw, np.uint64, 2, offset=nq * k * 12 + 4)
    ref_idx, ref_d2 = po.knn(P["db"], P["q"], k, nthreads=4)
    assert n_ovf == 0
    assert np.array_equal(idx, ref_idx.astype(np.uint64))
    assert np.array_equal(d2.view(np.uint32), ref_d2.view(np.uint32))
    assert k * nq <= rows[0] <= max(12 * k, 40) * nq          # the shortlist is short


def test_emulated_shortlist_flags_what_it_cannot_answer(exe, tmp_path):
    """Thousands of identical rows pass every bound together: K3 must hand those queries to the
    exact scan (overflow list) instead of answering from a truncated shortlist, and still answer
    the other queries exactly."""
    from oracle import pyoracle as po

    nq, n_rows, dim, k, n_ranges, tpr = 128, 3072, 64, 10, 1, 12
    P = prepare(nq, n_rows, dim, seed=5)
    P["db"][:2600] = P["db"][0]                        # 2600 copies of one row
    P["q"][:40] = P["db"][0] + np.float32(1e-3)        # 40 queries right next to them
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.array([nq, n_rows, P["n_pad"], dim, k, 512, n_ranges, tpr, 0, 1, 1], np.int32).tobytes())
        f.write(np.ascontiguousarray(P["q"]).tobytes())
        f.write(np.ascontiguousarray(P["db"]).tobytes())
    r = subprocess.run([exe, inp, outp], capture_output=True, text=True, timeout=2400)
    assert r.returncode == 0, r.stderr[-4000:]
    raw = open(outp, "rb").read()
    idx = np.frombuffer(raw, np.uint64, nq * k).reshape(nq, k)
    d2 = np.frombuffer(raw, np.float32, nq * k, offset=nq * k * 8).reshape(nq, k)
    n_ovf = int(np.frombuffer(raw, np.int32, 1, offset=nq * k * 12)[0])
    ovf = set(np.frombuffer(raw, np.int32, nq, offset=nq * k * 12 + 4 + 16)[:n_ovf].tolist())
    assert set(range(40)) <= ovf, "the tie-heavy queries must be flagged"
    ref_idx, ref_d2 = po.knn(P["db"], P["q"], k, nthreads=4)
    ok = [q for q in range(nq) if q not in ovf]
    assert len(ok) >= 40
    assert np.array_equal(idx[ok], ref_idx[ok].astype(np.uint64))
    assert np.array_equal(d2[ok].view(np.uint32), ref_d2[ok].view(np.uint32))
