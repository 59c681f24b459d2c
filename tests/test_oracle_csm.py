"""CPU: self-consistency of the stage-2 restatement and its golden vectors (minted from the
reference's own registration/2d compiled into oracle/_ref/libcsm_ref.so; the direct comparison
with that library is tests/test_oracle_csm_ref.py)."""
import glob
import os

import numpy as np
import pytest

from gloc3d_b200 import synth


def test_value_codec(oracle):
    lo, hi = oracle.lib().gloc_oracle_min_cost(), oracle.lib().gloc_oracle_max_cost()
    assert np.float32(lo) == np.float32(1) - (np.float32(1) - np.float32(0.1))
    assert np.float32(hi) == np.float32(1) - np.float32(0.1)
    assert oracle.value_to_cost(0) == np.float32(hi)               # unknown -> kMax
    assert abs(oracle.value_to_cost(1) - lo) < 1e-6
    assert abs(oracle.value_to_cost(32767) - hi) < 1e-6
    assert oracle.value_to_cost(32768 + 5) == oracle.value_to_cost(5)  # update marker repeats
    for v in (1, 2, 100, 16384, 32766, 32767):                     # round trip
        assert oracle.cost_to_value(oracle.value_to_cost(v)) == v
    assert oracle.cell_value(np.float32(1) - np.float32(lo)) == 255  # occupied
    assert oracle.cell_value(np.float32(1) - np.float32(hi)) == 0    # free / unknown


@pytest.mark.parametrize("nx,ny,w", [(37, 23, 1), (37, 23, 2), (37, 23, 4), (37, 23, 16), (5, 3, 8),
                                     (1, 1, 4), (16, 16, 16), (15, 40, 32)])
def test_precomputation_float_path_equals_u8_path(oracle, nx, ny, w):
    rng = np.random.default_rng(nx * 100 + ny + w)
    cells = rng.integers(0, 32768, size=(ny, nx)).astype(np.uint16)
    cells[rng.random((ny, nx)) < 0.3] = 0
    a = oracle.precomp_from_cells(cells, w)
    b = oracle.precomp_from_level1(oracle.level1_from_cells(cells), w)
    assert a.shape == (ny + w - 1, nx + w - 1)
    assert np.array_equal(a, b)


def test_precomputation_is_windowed_max(oracle):
    rng = np.random.default_rng(7)
    l1 = rng.integers(0, 256, size=(20, 31)).astype(np.uint8)
    for w in (2, 4, 8):
        p = oracle.precomp_from_level1(l1, w)
        pad = np.zeros((20 + 2 * (w - 1), 31 + 2 * (w - 1)), np.uint8)
        pad[w - 1:w - 1 + 20, w - 1:w - 1 + 31] = l1
        for ly in range(p.shape[0]):
            for lx in range(p.shape[1]):
                assert p[ly, lx] == pad[ly:ly + w, lx:lx + w].max()


def test_search_params_and_grid_to_points(oracle):
    pts = np.array([[3, 4, 0], [0.1, 0.2, 5], [-60, 45, 1]], np.float32)
    n_lin, n_ang, step = oracle.search_params(3.0, 3.0, pts, 0.2)
    r = np.float32(np.sqrt(np.float32(np.float32(60 * 60) + np.float32(45 * 45))))
    want = (1 - 1e-3) * np.arccos(1 - 0.2 * 0.2 / (2.0 * float(np.float32(r * r))))
    assert n_lin == 15 and abs(step - want) < 1e-15 and n_ang == int(np.ceil(3.0 / step))
    # default depth-5 options (fast_..._2d.h:49-51) on an empty cloud: r_max = 3*res
    n_lin, n_ang, step = oracle.search_params(3.0, 3.0, np.zeros((0, 3), np.float32), 0.2)
    assert n_lin == 15 and n_ang == int(np.ceil(3.0 / step))
    cells = np.zeros((4, 6), np.uint16)
    cells[1, 2] = 1        # occupied (cost kMin < 0.11)
    cells[3, 5] = 400      # cost 0.1097 < 0.11
    cells[0, 0] = 500      # cost 0.1122 > 0.11
    cells[2, 2] = 32767    # free
    p = oracle.grid_to_points(cells, 0.2, 10.0, -5.0)
    # iteration order: i (x) outer, j (y) inner; point = (ox + i*res, oy + j*res, 0)
    assert np.allclose(p, [[10.4, -4.8, 0], [11.0, -4.4, 0]], atol=1e-6) and p.shape == (2, 3)


def test_discretize_axis_convention(oracle):
    # MapLimits::GetCellIndex swaps and flips: cell x <- world y, cell y <- world x
    res, mx, my = 0.5, 10.0, 20.0
    pts = np.array([[9.75, 19.75, 0], [9.75, 18.75, 0], [8.25, 19.75, 0]], np.float32)
    c = oracle.discretize(pts, (0, 0, 0), 0, 0.1, res, mx, my)
    assert c.shape == (1, 3, 2)
    assert c[0].tolist() == [[0, 0], [2, 0], [0, 3]]
    # pure translation of the initial pose shifts cells
    c2 = oracle.discretize(pts, (-1.0, -0.5, 0), 0, 0.1, res, mx, my)
    assert (c2[0] - c[0]).tolist() == [[1, 2]] * 3


@pytest.mark.parametrize("seed", range(8))
def test_branch_and_bound_equals_exhaustive(oracle, seed):
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(60, 140)), int(rng.integers(60, 140))
    res = 0.2
    g = synth.make_bev_grid(nx, ny, seed=100 + seed, n_segments=10, n_blobs=6, graded=bool(seed % 2))
    mx, my = synth.centered_limits(nx, ny, res)
    scan = synth.planted_scan(g, res, mx, my, yaw=rng.uniform(-0.5, 0.5), dx=rng.uniform(-1.5, 1.5),
                              dy=rng.uniform(-1.5, 1.5), dropout=0.2, jitter_cells=0.5, seed=seed)
    depth = int(rng.integers(1, 6))
    n_lin, n_ang = int(rng.integers(3, 14)), int(rng.integers(0, 18))
    a = oracle.csm_match(g, res, mx, my, depth, scan, (0.03, -0.02, 0.01), n_lin, n_ang, np.pi / 80, 0.15, 0)
    b = oracle.csm_match(g, res, mx, my, depth, scan, (0.03, -0.02, 0.01), n_lin, n_ang, np.pi / 80, 0.15, 1)
    assert a.found == b.found == 1
    assert np.float32(a.score) == np.float32(b.score)   # B&B max == exhaustive max
    assert a.n_scored > 0 and b.n_scored > 0


def test_planted_pose_recovered_and_wrong_map_rejected(oracle):
    res = 0.2
    g = synth.make_bev_grid(300, 300, seed=2222, n_segments=24, n_blobs=14)
    other = synth.make_bev_grid(300, 300, seed=77, n_segments=24, n_blobs=14)
    mx, my = synth.centered_limits(300, 300, res)
    scan = synth.planted_scan(g, res, mx, my, yaw=0.7, dx=3.4, dy=-2.0, dropout=0.2)
    r = oracle.csm_match(g, res, mx, my, 5, scan, (0, 0, 0), 30, 60, 2 * np.pi / 360, 0.3, 0)
    assert r.found and abs(r.pose_x - 3.4) <= 0.2 and abs(r.pose_y + 2.0) <= 0.2
    assert abs(r.pose_yaw - 0.7) <= 2 * np.pi / 360
    w = oracle.csm_match(other, res, mx, my, 5, scan, (0, 0, 0), 30, 60, 2 * np.pi / 360, 0.45, 0)
    assert not w.found and np.float32(w.score) == np.float32(0.45)


def test_binary_score_formula(oracle):
    # binary grids: score = min_s + 0.8 * hits / P up to float rounding
    res = 0.2
    g = synth.make_bev_grid(120, 120, seed=5, n_segments=10, n_blobs=5)
    mx, my = synth.centered_limits(120, 120, res)
    scan = synth.grid_points_world(g, res, mx, my)
    r = oracle.csm_match(g, res, mx, my, 3, scan, (0, 0, 0), 4, 2, 0.01, 0.2, 0)
    assert r.found and (r.scan_index, r.x_offset, r.y_offset) == (2, 0, 0)
    assert abs(r.score - 0.9) < 1e-6   # every point lands on an occupied cell


def test_full_submap_uses_25_cell_window(oracle):
    res = 0.2
    g = synth.make_bev_grid(100, 100, seed=9, n_segments=8, n_blobs=4)
    mx, my = 30.0, 50.0
    # scan expressed relative to the grid centre with a small offset
    pts = synth.grid_points_world(g, res, mx, my).astype(np.float64)
    cx, cy = mx - 0.5 * res * 100, my - 0.5 * res * 100
    pts[:, 0] -= cx + 0.6
    pts[:, 1] -= cy - 0.4
    r = oracle.csm_match_full_submap(g, res, mx, my, 4, pts.astype(np.float32), 0.5, 0)
    assert r.found and abs(r.pose_x - (cx + 0.6)) < 0.11 and abs(r.pose_y - (cy - 0.4)) < 0.11


def test_csm_golden_regression(oracle, golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "csm_*.npz")))
    assert len(files) >= 2
    for f in files:
        z = np.load(f)
        r = oracle.csm_match(z["grid"], float(z["res"]), float(z["max_x"]), float(z["max_y"]),
                             int(z["depth"]), z["scan"], tuple(z["init"]), int(z["n_lin"]),
                             int(z["n_ang"]), float(z["step"]), float(z["min_score"]), 0)
        assert np.array_equal(np.array(r.as_tuple(), np.float64), z["result"]), f
        c = oracle.discretize(z["scan"], tuple(z["init"]), int(z["n_ang"]), float(z["step"]),
                              float(z["res"]), float(z["max_x"]), float(z["max_y"]))
        assert np.array_equal(c[0], z["cells_first"]) and np.array_equal(c[-1], z["cells_last"])
        for w in (2, 4, 16):
            assert np.array_equal(oracle.precomp_from_level1(z["grid"], w), z[f"level{w}"])


def _numpy_matcher(level1, res, max_x, max_y, pts, init, n_lin, n_ang, step, min_score):
    """A second, independently written restatement of MatchWithSearchParameters (numpy, brute
    force) from SURVEY.md Appendix B: Eigen's quaternion rotation in float32, GetCellIndex in
    double with lround, ShrinkToFit, integer sums, float32 score, ties to the smallest
    (scan, x, y).  A third opinion next to oracle/csm_oracle.c and the compiled reference
    (tests/test_oracle_csm_ref.py)."""
    f32 = np.float32
    ny, nx = level1.shape
    P = pts.shape[0]

    def rot(p, theta):          # Quaternionf(AngleAxisf(theta, Z)) * p, fast_..._2d.cpp:278-283
        th = f32(theta)
        w, z = f32(np.cos(f32(0.5) * th, dtype=f32)), f32(np.sin(f32(0.5) * th, dtype=f32))
        x, y = p[:, 0].astype(f32), p[:, 1].astype(f32)
        ux, uy = -(z * y), z * x
        ux, uy = ux + ux, uy + uy
        return np.stack([(x + w * ux) + (-(z * uy)), (y + w * uy) + (z * ux)], axis=1).astype(f32)

    def lround(v):              # half away from zero (port.h:41-43)
        return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5)).astype(np.int64)

    p0 = rot(pts, init[2])
    min_s, max_s = f32(1) - (f32(1) - f32(0.1)), f32(1) - (f32(1) - (f32(1) - f32(0.1)))
    coef = (max_s - min_s) / f32(255)
    best = None
    theta = -n_ang * step       # accumulated in double, correlative_..._2d.cpp:99-107
    for s in range(2 * n_ang + 1):
        sc = rot(p0, theta)
        theta += step
        wx, wy = sc[:, 0] + f32(init[0]), sc[:, 1] + f32(init[1])
        cx = lround((max_y - wy.astype(np.float64)) / res - 0.5)      # map_limits.h:69-76
        cy = lround((max_x - wx.astype(np.float64)) / res - 0.5)
        lo_x, hi_x = max(-n_lin, min(0, int((-cx).min()))), min(n_lin, max(0, int((nx - 1 - cx).max())))
        lo_y, hi_y = max(-n_lin, min(0, int((-cy).min()))), min(n_lin, max(0, int((ny - 1 - cy).max())))
        for xo in range(lo_x, hi_x + 1):
            X = cx + xo
            okx = (X >= 0) & (X < nx)
            for yo in range(lo_y, hi_y + 1):
                Y = cy + yo
                ok = okx & (Y >= 0) & (Y < ny)
                total = int(level1[Y[ok], X[ok]].astype(np.int64).sum())
                score = f32(min_s + f32(f32(total) / f32(P)) * coef)
                if best is None or score > best[0]:
                    best = (score, s, xo, yo)
    if best is None or not best[0] > f32(min_score):
        return None
    return best


@pytest.mark.parametrize("seed", range(4))
def test_oracle_agrees_with_an_independent_numpy_restatement(oracle, seed):
    rng = np.random.default_rng(40 + seed)
    nx, ny = int(rng.integers(40, 70)), int(rng.integers(40, 70))
    res = 0.2
    g = synth.make_bev_grid(nx, ny, seed=300 + seed, n_segments=8, n_blobs=5, graded=bool(seed % 2))
    mx, my = synth.centered_limits(nx, ny, res)
    mx, my = mx + 0.37, my - 1.13                       # limits that do not centre the grid
    scan = synth.planted_scan(g, res, mx, my, yaw=rng.uniform(-0.3, 0.3), dx=rng.uniform(-0.8, 0.8),
                              dy=rng.uniform(-0.8, 0.8), dropout=0.3, jitter_cells=0.5, seed=seed)[:120]
    init = (0.05, -0.03, 0.02)
    n_lin, n_ang, step = 6, 4, np.pi / 70
    ref = _numpy_matcher(g, res, mx, my, scan, init, n_lin, n_ang, step, 0.12)
    for mode in (0, 1):
        o = oracle.csm_match(g, res, mx, my, 3, scan, init, n_lin, n_ang, step, 0.12, mode)
        assert ref is not None and o.found
        assert np.float32(o.score) == ref[0]
        if mode == 1:   # the exhaustive scan resolves ties to the smallest (scan, x, y), like the restatement
            assert (o.scan_index, o.x_offset, o.y_offset) == ref[1:]
    # the discretisation on its own (GenerateRotatedScans + DiscretizeScans), cell for cell
    c = oracle.discretize(scan, init, n_ang, step, res, mx, my)
    assert c.shape == (2 * n_ang + 1, scan.shape[0], 2)
    f32 = np.float32

    def rot(p, theta):
        th = f32(theta)
        w, z = f32(np.cos(f32(0.5) * th, dtype=f32)), f32(np.sin(f32(0.5) * th, dtype=f32))
        x, y = p[:, 0].astype(f32), p[:, 1].astype(f32)
        ux, uy = -(z * y), z * x
        ux, uy = ux + ux, uy + uy
        return np.stack([(x + w * ux) + (-(z * uy)), (y + w * uy) + (z * ux)], axis=1).astype(f32)

    p0, theta = rot(scan, init[2]), -n_ang * step
    for s_i in range(2 * n_ang + 1):
        sc = rot(p0, theta)
        theta += step
        wx, wy = sc[:, 0] + f32(init[0]), sc[:, 1] + f32(init[1])
        vx = (my - wy.astype(np.float64)) / res - 0.5
        vy = (mx - wx.astype(np.float64)) / res - 0.5
        cx = np.where(vx >= 0, np.floor(vx + 0.5), np.ceil(vx - 0.5)).astype(np.int64)
        cy = np.where(vy >= 0, np.floor(vy + 0.5), np.ceil(vy - 0.5)).astype(np.int64)
        assert np.array_equal(c[s_i, :, 0], cx) and np.array_equal(c[s_i, :, 1], cy)
