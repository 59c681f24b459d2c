"""GPU parity tests of the BEV projection through the C ABI, against
oracle/bev_oracle.c: images bit for bit, geometry, CNN input, occupied points, and the
projected grid as a scan-matcher map."""
import os

import numpy as np
import pytest

import gloc3d_b200 as g
from gloc3d_b200 import synth

pytestmark = pytest.mark.gpu


def check_scan(oracle, bev, scan):
    info = bev.project(scan)
    img, (ox, oy, res), (mx, my), nv, no = oracle.bev_project(scan)
    assert (info.height, info.width) == img.shape and (info.min_ix, info.min_iy) == (mx, my)
    assert (info.ox, info.oy, info.resolution) == (ox, oy, res) and info.n_occupied == no
    if img.size:
        assert np.array_equal(bev.image(), img)
    return info, img


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_synthetic_scans(oracle, seed):
    bev = g.BevProjector(0)
    scan = synth.make_lidar_scan(seed=seed, n_walls=30 + 10 * seed)
    info, img = check_scan(oracle, bev, scan)
    assert info.n_occupied > 500
    check_scan(oracle, bev, np.ascontiguousarray(scan[:, :3]))       # stride 3
    check_scan(oracle, bev, scan[:17])                               # a smaller scan on the same handle
    bev.close()


def test_edge_cases(oracle):
    bev = g.BevProjector(0)
    info = bev.project(np.zeros((0, 4), np.float32))
    assert info.width == 0 and info.height == 0 and info.n_occupied == 0
    info = bev.project(np.array([[150, 0, 0, 0], [np.nan, 1, 1, 0]], np.float32))   # only misses
    assert info.width == 0 and info.n_points_in_range == 0
    pts = np.array([[60, 80, 0, 0], [60, 80, 0.2, 0], [60.01, 80, 0, 0], [-100, 0, 0, 0], [-100, 0, -0.2, 0],
                    [0.1, -0.1, 0, 0], [0.1, -0.1, 0.4, 0], [0.3, 0.5, 0, 0]], np.float32)
    check_scan(oracle, bev, pts)
    with pytest.raises(g.GlocError):
        g.BevProjector(0, resolution=0.0)
    bev.close()


def test_golden_and_cnn_input(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "bev_kitti_subsample.npz"))
    bev = g.BevProjector(0)
    info = bev.project(z["pts"])
    img = bev.image()
    assert np.array_equal(np.packbits(img == 0), z["occupied_bits"]) and img.shape == tuple(z["shape"])
    assert (info.min_ix, info.min_iy, info.n_occupied) == tuple(int(v) for v in z["min_and_count"])
    for w, h in ((768, 768), (300, 200), (1000, 900)):     # crop both, crop, pad both
        assert np.array_equal(bev.cnn_input(w, h), oracle.crop_pad(img, w, h))
    bev.close()


def test_occupied_points_and_projected_grid(oracle):
    scan = synth.make_lidar_scan(seed=11)
    bev = g.BevProjector(0)
    info = bev.project(scan)
    img = bev.image()
    # ProjectToGrid: occupied -> probability 0.9 -> cost value 1, free -> cost value 32767
    cells = np.where(img == 0, np.uint16(1), np.uint16(32767)).astype(np.uint16)
    ref_pts = oracle.grid_to_points(cells, info.resolution, info.ox, info.oy)
    pts = bev.occupied_points()
    assert pts.shape == ref_pts.shape and np.array_equal(pts.view(np.uint32), ref_pts.view(np.uint32))
    # the projected grid inside a scan-match store == the grid built from the same cells
    st = g.CsmStore(0)
    gid = bev.add_to_store(st)
    max_x = (info.min_ix + info.width - 1) * info.resolution
    max_y = (info.min_iy + info.height - 1) * info.resolution
    gid2 = st.add_grid_cells(cells, info.resolution, max_x, max_y)
    for w in (1, 4, 16):
        assert np.array_equal(st.precomputation_grid(gid, w), st.precomputation_grid(gid2, w))
    assert np.array_equal(st.precomputation_grid(gid, 1), np.where(img == 0, 255, 0).astype(np.uint8))
    # a scan matches its own projected grid at the identity with the maximum score
    r = st.match_batch([pts], [gid], [0], [(0.0, 0.0, 0.0)], 8, 4, np.pi / 180, 3, 0.3)[0]
    ro = oracle.csm_match(np.where(img == 0, 255, 0).astype(np.uint8), info.resolution, max_x, max_y, 3,
                          pts, (0.0, 0.0, 0.0), 8, 4, np.pi / 180, 0.3, 0)
    assert r.as_tuple() == ro.as_tuple()
    st.close()
    bev.close()


def test_against_compiled_reference(oracle):
    """GPU vs the REFERENCE'S OWN projection (3d/submap_3d.cpp, range_data_inserter_3d.cpp, hybrid_grid.h
    compiled unmodified into oracle/_ref/libbev_ref.so): image and double-precision origin identical."""
    if not oracle.have_bev_ref():
        pytest.skip("oracle/_ref/libbev_ref.so not present")
    bev = g.BevProjector(0)
    rng = np.random.default_rng(2)
    scans = [synth.make_lidar_scan(seed=40 + s, n_walls=20 + 15 * s) for s in range(3)]
    k = rng.integers(-450, 450, (3000, 3)).astype(np.float32)
    half = (k + np.float32(0.5)) * np.float32(0.2)          # coordinates on voxel boundaries
    half[:, 2] = np.clip(half[:, 2], -3, 3)
    scans.append(np.concatenate([half, np.zeros((3000, 1), np.float32)], 1))
    for scan in scans:
        info = bev.project(scan)
        rimg, (rox, roy, rres) = oracle.ref_bev_project(scan)
        assert (info.height, info.width) == rimg.shape
        assert (info.ox, info.oy, info.resolution) == (rox, roy, rres)
        assert np.array_equal(bev.image(), rimg)
    bev.close()
