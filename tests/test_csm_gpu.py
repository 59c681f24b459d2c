"""GPU parity tests of stage 2 through the C ABI against the CPU restatement: integer
candidate sums -> bit-exact score, discrete pose and double pose; precomputation grids
and discretised scans byte-exact."""
import glob
import os

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


@pytest.mark.parametrize("nx,ny", [(37, 23), (1, 1), (5, 3), (16, 16), (781, 504)])
def test_precomputation_stack(oracle, nx, ny):
    rng = np.random.default_rng(nx + ny)
    cells = rng.integers(0, 65536, size=(ny, nx)).astype(np.uint16)
    cells[rng.random((ny, nx)) < 0.5] = 0
    st = g.CsmStore(0)
    gid = st.add_grid_cells(cells, 0.2, 10.0, 12.0)
    l1 = oracle.level1_from_cells(cells)
    gid2 = st.add_grid_u8(l1, 0.2, 10.0, 12.0)
    for w in (1, 2, 4, 8, 16, 32):
        want = oracle.precomp_from_cells(cells, w)      # the reference's float path
        assert np.array_equal(st.precomputation_grid(gid, w), want), w
        assert np.array_equal(st.precomputation_grid(gid2, w), want), w
    with pytest.raises(g.GlocError):
        st.precomputation_grid(gid, 3)
    st.close()


def test_rotate_and_discretise(oracle):
    rng = np.random.default_rng(1)
    pts = np.concatenate([rng.uniform(-75, 75, (3000, 2)), rng.uniform(-2, 2, (3000, 1))], 1).astype(np.float32)
    st = g.CsmStore(0)
    for init, n_ang, step in [((0, 0, 0), 180, 2 * np.pi / 360), ((12.3, -4.56, 1.234), 25, 0.0026),
                              ((-0.7, 0.3, -3.0), 0, 0.1)]:
        got = st.discretize(pts, init, n_ang, step, 0.2, 80.0, 80.0)
        assert np.array_equal(got, oracle.discretize(pts, init, n_ang, step, 0.2, 80.0, 80.0))
    st.close()


@pytest.mark.parametrize("seed", range(6))
def test_random_small_cases_vs_oracle(oracle, seed):
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(60, 160)), int(rng.integers(60, 160))
    res = 0.2
    grid = synth.make_bev_grid(nx, ny, seed=100 + seed, n_segments=10, n_blobs=6, graded=bool(seed % 2))
    mx, my = synth.centered_limits(nx, ny, res)
    scan = synth.planted_scan(grid, res, mx, my, yaw=rng.uniform(-0.5, 0.5), dx=rng.uniform(-1.5, 1.5),
                              dy=rng.uniform(-1.5, 1.5), dropout=0.2, jitter_cells=0.5, seed=seed)
    init = (0.03, -0.02, 0.01)
    n_lin, n_ang = int(rng.integers(3, 14)), int(rng.integers(0, 18))
    st = g.CsmStore(0)
    gid = st.add_grid_u8(grid, res, mx, my)
    exh = oracle.csm_match(grid, res, mx, my, 1, scan, init, n_lin, n_ang, np.pi / 80, 0.15, 1)
    for depth in (1, 2, 3, 5, 7):
        r = st.match_batch([scan], [gid], [0], [init], n_lin, n_ang, np.pi / 80, depth, 0.15)[0]
        same(r, exh)            # canonical argmax == exhaustive scan, whatever the depth
        bnb = oracle.csm_match(grid, res, mx, my, depth, scan, init, n_lin, n_ang, np.pi / 80, 0.15, 0)
        assert np.float32(r.score) == np.float32(bnb.score)
    st.close()


def test_ties_resolve_to_smallest_scan_x_y(oracle):
    # an empty-ish map with one occupied block: many poses tie at the maximum
    grid = np.zeros((40, 40), np.uint8)
    grid[10:30, 10:30] = 255
    res = 0.2
    mx, my = synth.centered_limits(40, 40, res)
    scan = np.array([[0.1, 0.1, 0], [-0.1, 0.1, 0], [0.1, -0.1, 0]], np.float32)
    st = g.CsmStore(0)
    gid = st.add_grid_u8(grid, res, mx, my)
    exh = oracle.csm_match(grid, res, mx, my, 1, scan, (0, 0, 0), 12, 3, 0.05, 0.2, 1)
    for depth in (1, 3, 5):
        same(st.match_batch([scan], [gid], [0], [(0, 0, 0)], 12, 3, 0.05, depth, 0.2)[0], exh)
    st.close()


def test_kitti_shape_pairs_full_window(oracle):
    # config 2: 360 yaw bins (361 scans), +-20 m at 0.2 m (+-100 cells), depth 5, 800x800 grid
    res = 0.2
    step = 2 * np.pi / 360
    grid = synth.make_bev_grid(800, 800, seed=2222)
    other = synth.make_bev_grid(800, 800, seed=99)
    mx, my = synth.centered_limits(800, 800, res)
    scan = synth.planted_scan(grid, res, mx, my, yaw=2.0, dx=11.0, dy=-7.4, dropout=0.2, jitter_cells=0.0)
    scan2 = synth.planted_scan(grid, res, mx, my, yaw=-1.1, dx=-15.5, dy=17.9, dropout=0.2, jitter_cells=1.0,
                               seed=4)
    st = g.CsmStore(0)
    g0 = st.add_grid_u8(grid, res, mx, my)
    g1 = st.add_grid_u8(other, res, mx, my)
    out = st.match_batch([scan, scan2], [g0, g1, g0], [0, 0, 1], [(0, 0, 0)] * 3, 100, 180, step, 5, 0.3)
    refs = [oracle.csm_match(gr, res, mx, my, 5, sc, (0, 0, 0), 100, 180, step, 0.3, 0)
            for gr, sc in ((grid, scan), (other, scan), (grid, scan2))]
    for r, o in zip(out, refs):
        same(r, o)
    assert out[0].found and abs(out[0].pose_x - 11.0) <= 0.21 and abs(out[0].pose_y + 7.4) <= 0.21
    assert abs(out[0].pose_yaw - 2.0) <= step and not out[1].found
    assert out[2].found and abs(out[2].pose_x + 15.5) <= 0.41 and abs(out[2].pose_yaw + 1.1) <= 2 * step
    s = st.stats()
    assert s.matches == 3 and s.kernel_launches >= 4
    st.close()


def test_large_scan_chunked_points(oracle):
    # more points than one shared-memory chunk (4096)
    res = 0.2
    grid = synth.make_bev_grid(500, 500, seed=8, n_segments=120, n_blobs=80)
    mx, my = synth.centered_limits(500, 500, res)
    scan = synth.planted_scan(grid, res, mx, my, yaw=0.2, dx=2.0, dy=1.0, dropout=0.0)
    assert scan.shape[0] > 4096
    st = g.CsmStore(0)
    gid = st.add_grid_u8(grid, res, mx, my)
    r = st.match_batch([scan], [gid], [0], [(0.5, 0.5, 0.1)], 30, 20, np.pi / 180, 5, 0.3)[0]
    same(r, oracle.csm_match(grid, res, mx, my, 5, scan, (0.5, 0.5, 0.1), 30, 20, np.pi / 180, 0.3, 0))
    st.close()


def test_interface_mirror(oracle):
    # FastCorrelativeScanMatcher2D(grid, options).Match / MatchFullSubmap with a ProbabilityGrid
    res = 0.2
    l1 = synth.make_bev_grid(200, 160, seed=5, n_segments=16, n_blobs=8)
    cells = synth.level1_to_cells(l1)
    mx, my = 25.0, 31.0
    grid = g.ProbabilityGrid(g.MapLimits(res, mx, my, 200, 160), cells)
    opt = g.FastCorrelativeScanMatcherOptions2D()
    assert (opt.linear_search_window, opt.angular_search_window, opt.branch_and_bound_depth) == (3.0, 3.0, 5)
    m = g.FastCorrelativeScanMatcher2D(grid, opt)
    assert np.array_equal(m.precomputation_grid(4), oracle.precomp_from_cells(cells, 16))
    level1 = oracle.level1_from_cells(cells)
    cx, cy = mx - 0.5 * res * 200, my - 0.5 * res * 160
    cloud = synth.grid_points_world(level1, res, mx, my).astype(np.float64)
    cloud[:, 0] -= cx + 0.4
    cloud[:, 1] -= cy + 0.2
    cloud = cloud.astype(np.float32)[::3]
    ok, score, pose = m.MatchFullSubmap(cloud, 0.4)
    o = oracle.csm_match_full_submap(level1, res, mx, my, 5, cloud, 0.4, 0)
    assert ok and o.found and np.float32(score) == np.float32(o.score)
    assert (pose.x, pose.y, pose.yaw) == (o.pose_x, o.pose_y, o.pose_yaw)
    ok2, _, _ = m.MatchFullSubmap(cloud, 0.95)          # score <= min_score -> false, outputs untouched
    assert not ok2
    # Match(initial pose, Grid2D): the query grid goes through GridToVirtualPointCloud
    qgrid = g.ProbabilityGrid(g.MapLimits(res, mx, my, 200, 160), cells, ox=-20.0, oy=-16.0)
    vpc = g.grid_to_virtual_point_cloud(qgrid)
    assert np.array_equal(vpc, oracle.grid_to_points(cells, res, -20.0, -16.0))
    n_lin, n_ang, st_ = g.search_parameters(3.0, 3.0, vpc, res)
    assert (n_lin, n_ang, st_) == oracle.search_params(3.0, 3.0, vpc, res)
    ok3, score3, pose3 = m.Match(g.Rigid2d(cx, cy, 0.3), qgrid, 0.1)
    o3 = oracle.csm_match(level1, res, mx, my, 5, vpc, (cx, cy, 0.3), n_lin, n_ang, st_, 0.1, 0)
    assert ok3 == bool(o3.found)
    if ok3:
        assert np.float32(score3) == np.float32(o3.score)


def test_csm_golden(golden_dir):
    for f in sorted(glob.glob(os.path.join(golden_dir, "csm_*.npz"))):
        z = np.load(f)
        st = g.CsmStore(0)
        gid = st.add_grid_u8(z["grid"], float(z["res"]), float(z["max_x"]), float(z["max_y"]))
        r = st.match_batch([z["scan"]], [gid], [0], [tuple(z["init"])], int(z["n_lin"]), int(z["n_ang"]),
                           float(z["step"]), int(z["depth"]), float(z["min_score"]))[0]
        assert np.array_equal(np.array(r.as_tuple(), np.float64), z["result"]), f
        c = st.discretize(z["scan"], tuple(z["init"]), int(z["n_ang"]), float(z["step"]), float(z["res"]),
                          float(z["max_x"]), float(z["max_y"]))
        assert np.array_equal(c[0], z["cells_first"]) and np.array_equal(c[-1], z["cells_last"])
        for w in (2, 4, 16):
            assert np.array_equal(st.precomputation_grid(gid, w), z[f"level{w}"])
        st.close()


def test_errors():
    st = g.CsmStore(0)
    with pytest.raises(g.GlocError):
        st.add_grid_u8(np.zeros((0, 5), np.uint8), 0.2, 1, 1)
    gid = st.add_grid_u8(np.zeros((8, 8), np.uint8), 0.2, 1, 1)
    scan = np.zeros((3, 3), np.float32)
    with pytest.raises(g.GlocError):
        st.match_batch([scan], [gid], [0], [(0, 0, 0)], 5, 5, 0.1, 0, 0.3)       # depth >= 1
    with pytest.raises(g.GlocError):
        st.match_batch([scan], [gid + 1], [0], [(0, 0, 0)], 5, 5, 0.1, 3, 0.3)   # bad grid id
    r = st.match_batch([scan], [gid], [0], [(0, 0, 0)], 5, 5, 0.1, 3, 0.3)[0]    # empty map: no match
    assert not r.found
    st.close()


def test_kitti_fixture_shape_off_centre_limits(oracle):
    # the shape of the reference's only real BEV (781 x 504) with limits that do not
    # centre the grid, full +-100-cell window: different plane widths / heights in x and y,
    # border points on one side only, and an initial pose away from the origin
    res = 0.2
    step = 2 * np.pi / 360
    grid = synth.make_bev_grid(781, 504, seed=77, n_segments=50, n_blobs=30)
    mx0, my0 = synth.centered_limits(781, 504, res)
    mx, my = mx0 + 13.4, my0 - 7.8
    scan = synth.planted_scan(grid, res, mx, my, yaw=-2.6, dx=9.0, dy=4.2, dropout=0.25, jitter_cells=0.5, seed=9)
    st = g.CsmStore(0)
    gid = st.add_grid_u8(grid, res, mx, my)
    init = (1.3, -0.7, 0.2)
    r = st.match_batch([scan], [gid], [0], [init], 100, 180, step, 5, 0.3)[0]
    o = oracle.csm_match(grid, res, mx, my, 5, scan, init, 100, 180, step, 0.3, 0)
    same(r, o)
    assert r.found and abs(r.pose_x - 9.0) <= 0.31 and abs(r.pose_y - 4.2) <= 0.31
    st.close()


def test_against_compiled_reference(oracle):
    """GPU vs the REFERENCE'S OWN matcher (registration/2d compiled unmodified into
    oracle/_ref/libcsm_ref.so): score always, candidate and pose unless several candidates tie for
    the best score (std::sort's order among equal scores is unspecified in the reference)."""
    if not oracle.have_csm_ref():
        pytest.skip("oracle/_ref/libcsm_ref.so not present")
    res, step = 0.2, 2 * np.pi / 360
    st = g.CsmStore(0)
    cases = []
    for seed in range(6):
        rng = np.random.default_rng(500 + seed)
        nx, ny = int(rng.integers(90, 260)), int(rng.integers(90, 260))
        grid = synth.make_bev_grid(nx, ny, seed=600 + seed, n_segments=16, n_blobs=8)
        cells = synth.level1_to_cells(grid)
        if seed % 2:     # graded costs at the occupied cells
            occ = cells > 0
            cells[occ] = rng.integers(1, 32768, int(occ.sum())).astype(np.uint16)
        mx, my = synth.centered_limits(nx, ny, res)
        scan = synth.planted_scan(grid, res, mx, my, rng.uniform(-0.6, 0.6), rng.uniform(-3, 3), rng.uniform(-3, 3),
                                  dropout=0.2, jitter_cells=0.5, seed=seed)
        init = (float(rng.uniform(-0.4, 0.4)), float(rng.uniform(-0.4, 0.4)), float(rng.uniform(-0.1, 0.1)))
        cases.append((cells, mx, my, scan, init, st.add_grid_cells(cells, res, mx, my)))
    n_lin, n_ang, depth, min_score = 30, 40, 5, 0.3
    out = st.match_batch([c[3] for c in cases], [c[5] for c in cases], list(range(6)), [c[4] for c in cases],
                         n_lin, n_ang, step, depth, min_score)
    exact = 0
    for (cells, mx, my, scan, init, gid), r in zip(cases, out):
        ref = oracle.ref_csm_match(cells, res, mx, my, depth, scan, init, n_lin, n_ang, step, min_score)
        assert r.found == ref.found
        assert np.float32(r.score).view(np.uint32) == np.float32(ref.score).view(np.uint32)
        for index in (0, 2, 4):
            assert np.array_equal(st.precomputation_grid(gid, 1 << index), oracle.ref_precomp(cells, 5, index))
        if ref.found and (r.scan_index, r.x_offset, r.y_offset) == (ref.scan_index, ref.x_offset, ref.y_offset):
            assert (r.pose_x, r.pose_y, r.pose_yaw) == (ref.pose_x, ref.pose_y, ref.pose_yaw)
            exact += 1
    assert exact >= 5        # ties for the maximum are the exception
    st.close()


@pytest.mark.parametrize("nx,ny,n_lin", [(800, 800, 100), (813, 500, 100), (500, 813, 104), (640, 700, 60)])
def test_bit_planes_at_the_borders_both_layouts(oracle, nx, ny, n_lin):
    """The paired plane layout of the bit-sliced scorer has no hit tests: rows and columns outside the
    grid are reached through clamped addresses and shifts that pull zeros in.  Grids with occupied
    borders at the widest sizes the layout admits (52 plane columns), a sensor near a corner, scan
    points all around and far beyond the map, a shifted initial pose: both layouts must return the
    oracle's canonical argmax."""
    res, step, n_ang, depth = 0.2, 2 * np.pi / 360, 8, 5
    rng = np.random.default_rng(nx * 7 + ny)
    grid = synth.make_bev_grid(nx, ny, seed=nx + ny, n_segments=30, n_blobs=16)
    for k in (0, 1, 5):                       # occupied frame: the outermost rows / columns matter
        grid[k, :] = grid[-1 - k, :] = 255
        grid[:, k] = grid[:, -1 - k] = 255
    mx, my = synth.centered_limits(nx, ny, res)
    ext_x, ext_y = nx * res / 2, ny * res / 2
    st = g.CsmStore(0)
    gid = st.add_grid_u8(grid, res, mx, my)
    import os
    for corner in ((1, 1), (-1, 1), (1, -1), (-1, -1)):
        dx, dy = corner[0] * (ext_y - 9.0), corner[1] * (ext_x - 11.0)
        yaw = float(rng.uniform(-3, 3))
        scan = synth.planted_scan(grid, res, mx, my, yaw, dx, dy, dropout=0.6, jitter_cells=0.5, seed=nx + corner[0])
        clutter = np.zeros((1500, 3), np.float32)
        clutter[:, :2] = rng.uniform(-1.6 * max(ext_x, ext_y), 1.6 * max(ext_x, ext_y), (1500, 2))
        scan = np.concatenate([scan, clutter]).astype(np.float32)
        init = (dx + float(rng.uniform(-6, 6)), dy + float(rng.uniform(-6, 6)), yaw + 0.03)
        o = oracle.csm_match(grid, res, mx, my, depth, scan, init, n_lin, n_ang, step, 0.05, 0)
        for env in (None, "1"):
            if env:
                os.environ["GLOC_CSM_NO_PAIRED"] = env
            try:
                r = st.match_batch([scan], [gid], [0], [init], n_lin, n_ang, step, depth, 0.05)[0]
            finally:
                os.environ.pop("GLOC_CSM_NO_PAIRED", None)
            same(r, o)
        assert o.found
    st.close()
