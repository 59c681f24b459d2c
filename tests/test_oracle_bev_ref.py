"""Pins the BEV-projection oracle (oracle/bev_oracle.c) to the REFERENCE'S OWN code: 3d/submap_3d.cpp
(Submap3D::InsertRangeData, ProjectToCvMat), 3d/range_data_inserter_3d.cpp, 3d/range_data.cpp and
3d/hybrid_grid.h compiled unmodified from /root/reference into oracle/_ref/libbev_ref.so
(oracle/bev_ref.cpp, `make -C oracle ref`).  Image, shape and the double-precision origin must be
identical.  (An empty scan is not compared: the reference's bounding box stays at INT_MAX / INT_MIN
and the image size overflows; NaN coordinates reach lround() there.)"""
import os

import numpy as np
import pytest

from gloc3d_b200 import synth


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.have_bev_ref():
        pytest.skip("oracle/_ref/libbev_ref.so not present")
    return oracle.ref_bev_project


def same_projection(oracle, ref, pts):
    img, (ox, oy, res), _, _, _ = oracle.bev_project(pts)
    rimg, (rox, roy, rres) = ref(pts)
    assert img.shape == rimg.shape
    assert np.array_equal(img, rimg)
    assert (ox, oy, res) == (rox, roy, rres)        # doubles, bit for bit
    return img


def test_known_answers_against_the_reference(oracle, ref):
    pts = np.array([
        [1.0, 2.0, 0.0, 0], [1.0, 2.0, 0.2, 0], [1.0, 2.0, 0.21, 0],
        [-3.0, 0.4, 1.0, 0], [-3.0, 0.4, 1.05, 0],
        [0.1, -0.1, 0.0, 0], [0.1, -0.1, 0.4, 0],   # float32(0.1)/float32(0.2) = 0.5 -> 1 ; -0.5 -> -1
        [60.0, 80.0, 0.0, 0],                        # range exactly 100: a return
        [60.0, 80.0, 5.0, 0],                        # beyond 100: a miss
    ], np.float32)
    img = same_projection(oracle, ref, pts)
    assert (img == 0).sum() == 2


@pytest.mark.parametrize("seed", range(6))
def test_synthetic_scans_against_the_reference(oracle, ref, seed):
    scan = synth.make_lidar_scan(seed=seed, n_walls=25 + 12 * seed)
    img = same_projection(oracle, ref, scan)
    assert (img == 0).sum() > 300
    same_projection(oracle, ref, np.ascontiguousarray(scan[:, :3]))     # stride 3


def test_rounding_boundaries_and_grid_growth(oracle, ref):
    """Coordinates on voxel boundaries (k + 0.5 voxels: RoundToInt rounds half away from zero), returns at
    the far corners (the hybrid grid grows several times), stacks of voxels per column."""
    rng = np.random.default_rng(5)
    r = np.float32(0.2)
    k = rng.integers(-480, 480, (4000, 3)).astype(np.float32)
    half = (k + np.float32(0.5)) * r
    half[:, 2] = np.clip(half[:, 2], -3, 3)
    cols = rng.integers(-450, 450, (600, 2)).astype(np.float32) * r
    stacks = np.concatenate([np.concatenate([cols, np.full((600, 1), z, np.float32)], 1) for z in (0.0, 0.2, 0.4)])
    corners = np.array([[70, 70, 1], [70, 70, 1.3], [-70, 70, 0], [-70, 70, 0.25], [70, -70, 2], [-70, -70, -2],
                        [-70, -70, -1.7], [99.9, 0, 0], [99.9, 0, 0.2], [0, -99.9, 0.2], [0, -99.9, 0.0]], np.float32)
    pts = np.concatenate([half, stacks, corners]).astype(np.float32)
    pts = pts[np.linalg.norm(pts.astype(np.float64), axis=1) < 140]
    same_projection(oracle, ref, pts)


def test_golden_scan_against_the_reference(oracle, ref, golden_dir):
    z = np.load(os.path.join(golden_dir, "bev_kitti_subsample.npz"))
    img = same_projection(oracle, ref, z["pts"])
    assert np.array_equal(np.packbits(img == 0), z["occupied_bits"])


def test_reference_kitti_scan(oracle, ref):
    path = "/root/reference/s2s_libtorch/000000.bin"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    pts = np.fromfile(path, np.float32).reshape(-1, 4)
    img = same_projection(oracle, ref, pts)
    assert img.shape == (504, 781) and (img == 0).sum() == 4698
