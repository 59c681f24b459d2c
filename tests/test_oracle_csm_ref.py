"""Pins the stage-2 restatement (oracle/csm_oracle.c) against the REFERENCE'S OWN matcher:
registration/2d/{fast_,}correlative_scan_matcher_2d.cpp, grid_2d.cpp, probability_grid.cpp,
3d/probability_values.cpp and 3d/point_cloud.cpp compiled UNMODIFIED from /root/reference into
oracle/_ref/libcsm_ref.so (oracle/csm_ref.cpp, oracle/shim/ for the absent Eigen / glog / OpenCV /
boost / ceres headers).  CPU only.  The reference's std::sort leaves the order among equal scores
unspecified, so where several candidates tie for the best score only the score must agree."""
import numpy as np
import pytest

from gloc3d_b200 import synth
from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not po.have_csm_ref(), reason="oracle/_ref/libcsm_ref.so not built "
                                "(needs /root/reference; the GPU box gets the prebuilt file)")

STEP = 2 * np.pi / 360


def random_cells(rng, nx, ny, graded):
    """Grid2D cells: walls/blobs, occupied cells at the minimum cost (binary) or anywhere in
    [1, 32767] (graded), the rest unknown (0) with a sprinkle of known free cells."""
    g = synth.make_bev_grid(nx, ny, seed=int(rng.integers(1 << 30)), n_segments=max(4, nx // 12), n_blobs=6)
    cells = np.zeros((ny, nx), np.uint16)
    occ = g > 0
    cells[occ] = rng.integers(1, 32768, int(occ.sum())).astype(np.uint16) if graded else 1
    free = (~occ) & (rng.random((ny, nx)) < 0.1)
    cells[free] = rng.integers(30000, 32768, int(free.sum())).astype(np.uint16) if graded else 32767
    return cells


def agree(r, o, tie_ok=True):
    """r: reference result (RefMatchResult), o: restatement (MatchResult, mode 0)."""
    assert r.found == o.found
    assert np.float32(r.score).view(np.uint32) == np.float32(o.score).view(np.uint32)
    if not o.found:
        return True
    same = (r.scan_index, r.x_offset, r.y_offset) == (o.scan_index, o.x_offset, o.y_offset)
    if same:
        assert (r.pose_x, r.pose_y, r.pose_yaw) == (o.pose_x, o.pose_y, o.pose_yaw)
        assert np.float32(r.cand_score) == np.float32(o.score)
    else:
        assert tie_ok, "different candidate"
    return same


def test_value_codec_table():
    R = po.csm_ref()
    for v in list(range(0, 65536, 7)) + [1, 32766, 32767, 32768, 32769, 65535]:
        assert np.float32(R.gloc_ref_value_to_cost(v)).view(np.uint32) == \
            np.float32(po.value_to_cost(v)).view(np.uint32), v
    for c in np.linspace(0.05, 0.95, 301, dtype=np.float32):
        assert R.gloc_ref_cost_to_value(float(c)) == po.cost_to_value(float(c))


@pytest.mark.parametrize("nx,ny,graded", [(37, 23, True), (1, 1, True), (5, 3, False), (16, 16, True),
                                          (64, 33, False), (130, 97, True)])
def test_precomputation_grids(nx, ny, graded):
    rng = np.random.default_rng(nx * 1000 + ny)
    if nx < 30 or ny < 30:
        cells = rng.integers(0, 32768, (ny, nx)).astype(np.uint16)
        cells[rng.random((ny, nx)) < 0.4] = 0
    else:
        cells = random_cells(rng, nx, ny, graded)
    for index in range(6):
        want = po.ref_precomp(cells, 6, index)            # the reference's sliding-window float path
        assert np.array_equal(po.precomp_from_cells(cells, 1 << index), want), index
        l1 = po.level1_from_cells(cells)
        assert np.array_equal(po.precomp_from_level1(l1, 1 << index), want), index   # what the GPU implements


def test_rotate_and_discretise_and_search_parameters():
    rng = np.random.default_rng(1)
    pts = np.concatenate([rng.uniform(-75, 75, (3000, 2)), rng.uniform(-2, 2, (3000, 1))], 1).astype(np.float32)
    for init, n_ang, step in [((0, 0, 0), 180, STEP), ((12.3, -4.56, 1.234), 25, 0.0026), ((-0.7, 0.3, -3.0), 0, 0.1),
                              ((100.05, -99.95, 6.5), 7, 0.3)]:
        assert np.array_equal(po.ref_discretize(pts, init, n_ang, step, 0.2, 80.0, 80.0),
                              po.discretize(pts, init, n_ang, step, 0.2, 80.0, 80.0))
    for lin, ang, res in [(3.0, 3.0, 0.2), (5.0, np.pi, 0.2), (0.5, 0.1, 0.05), (7.0, 1.0, 1.0)]:
        assert po.ref_search_params(lin, ang, pts, res) == po.search_params(lin, ang, pts, res)
    assert po.ref_search_params(3.0, 1.0, pts[:0], 0.2) == po.search_params(3.0, 1.0, pts[:0], 0.2)


@pytest.mark.parametrize("seed", range(10))
def test_match_with_search_parameters(seed):
    rng = np.random.default_rng(100 + seed)
    nx, ny = int(rng.integers(60, 170)), int(rng.integers(60, 170))
    res = 0.2
    graded = bool(seed % 2)
    cells = random_cells(rng, nx, ny, graded)
    level1 = po.level1_from_cells(cells)
    mx, my = synth.centered_limits(nx, ny, res)
    mx, my = mx + float(rng.uniform(-3, 3)), my + float(rng.uniform(-3, 3))
    yaw, dx, dy = rng.uniform(-0.5, 0.5), rng.uniform(-2, 2), rng.uniform(-2, 2)
    scan = synth.planted_scan(np.where(level1 > 128, 255, 0).astype(np.uint8), res, mx, my, yaw, dx, dy,
                              dropout=0.2, jitter_cells=0.5, seed=seed)
    init = (float(rng.uniform(-0.5, 0.5)), float(rng.uniform(-0.5, 0.5)), float(rng.uniform(-0.2, 0.2)))
    n_lin, n_ang, depth = int(rng.integers(8, 30)), int(rng.integers(10, 40)), int(rng.integers(1, 7))
    min_score = float(rng.choice([0.0, 0.2, 0.35, 0.95]))
    r = po.ref_csm_match(cells, res, mx, my, depth, scan, init, n_lin, n_ang, STEP, min_score)
    o = po.csm_match(level1, res, mx, my, depth, scan, init, n_lin, n_ang, STEP, min_score, 0)
    agree(r, o)
    # the canonical (exhaustive, smallest (scan, x, y)) answer has the same score
    e = po.csm_match(level1, res, mx, my, depth, scan, init, n_lin, n_ang, STEP, min_score, 1)
    assert e.found == r.found and np.float32(e.score) == np.float32(r.score)


def test_scan_off_the_map_and_tiny_inputs():
    cells = np.zeros((20, 30), np.uint16)
    cells[5:9, 4:20] = 1
    far = np.array([[500.0, 500.0, 0.0], [501.0, 499.0, 0.0]], np.float32)
    one = np.array([[0.3, -0.2, 0.0]], np.float32)
    for scan in (far, one):
        for depth in (1, 3, 5):
            r = po.ref_csm_match(cells, 0.2, 3.0, 2.0, depth, scan, (0, 0, 0), 6, 4, 0.1, 0.3)
            o = po.csm_match(po.level1_from_cells(cells), 0.2, 3.0, 2.0, depth, scan, (0, 0, 0), 6, 4, 0.1, 0.3, 0)
            agree(r, o)


def test_match_full_submap():
    rng = np.random.default_rng(7)
    cells = random_cells(rng, 120, 90, False)
    level1 = po.level1_from_cells(cells)
    mx, my = 14.0, 11.0
    scan = synth.planted_scan(level1, 0.2, mx, my, 0.4, 2.0, 1.0, dropout=0.1, seed=2)[::6]   # 25 * res window
    r = po.ref_csm_match_full_submap(cells, 0.2, mx, my, 4, scan, 0.3)
    o = po.csm_match_full_submap(level1, 0.2, mx, my, 4, scan, 0.3, 0)
    assert r.found == o.found and np.float32(r.score) == np.float32(o.score)
    if o.found and (r.pose_x, r.pose_y) == (o.pose_x, o.pose_y):
        assert r.pose_yaw == o.pose_yaw


def test_batch_threads_equal_single_calls():
    rng = np.random.default_rng(9)
    cells = [random_cells(rng, 80, 70, bool(i % 2)) for i in range(5)]
    mx, my = synth.centered_limits(80, 70, 0.2)
    scans = [synth.planted_scan(po.level1_from_cells(c), 0.2, mx, my, 0.1 * i, 0.5, -0.5, seed=i) for i, c in enumerate(cells)]
    inits = [(0.0, 0.1 * i, 0.0) for i in range(5)]
    out = po.ref_csm_match_batch(cells, 0.2, mx, my, 4, scans, inits, 12, 15, STEP, 0.3, nthreads=3)
    for i in range(5):
        r = po.ref_csm_match(cells[i], 0.2, mx, my, 4, scans[i], inits[i], 12, 15, STEP, 0.3)
        assert (out[i].found, out[i].score, out[i].pose_x, out[i].pose_y, out[i].pose_yaw) == \
            (r.found, r.score, r.pose_x, r.pose_y, r.pose_yaw)
