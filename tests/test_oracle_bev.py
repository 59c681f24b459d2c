"""CPU tests of the BEV-projection oracle (oracle/bev_oracle.c): hand-built known answers,
the committed golden scan, and -- in the build container only -- the reference's own KITTI
scan with the figures SURVEY.md 8c quotes for it."""
import os

import numpy as np
import pytest

from gloc3d_b200 import synth


def test_known_answers(oracle):
    r = np.float32(0.2)
    pts = np.array([
        [1.0, 2.0, 0.0, 0], [1.0, 2.0, 0.2, 0],     # two voxels in column (5, 10): occupied
        [1.0, 2.0, 0.21, 0],                         # same voxel as the previous point
        [-3.0, 0.4, 1.0, 0],                         # one voxel in column (-15, 2): free
        [-3.0, 0.4, 1.05, 0],                        # still voxel z = 5
        [0.1, -0.1, 0.0, 0], [0.1, -0.1, 0.4, 0],   # float32(0.1)/float32(0.2) = 0.5 -> 1 ; -0.5 -> -1
        [60.0, 80.0, 0.0, 0],                        # range exactly 100: a return
        [60.0, 80.0, 5.0, 0],                        # beyond 100: a miss
        [np.nan, 0.0, 0.0, 0],
    ], np.float32)
    img, (ox, oy, res), (mx, my), nv, no = oracle.bev_project(pts)
    assert (mx, my) == (-15, -1) and img.shape == (400 - (-1) + 1, 300 - (-15) + 1)
    assert res == float(r) and ox == -15 * float(r) and oy == -1 * float(r)
    occ = {(int(x) + mx, int(y) + my) for y, x in zip(*np.nonzero(img == 0))}
    assert occ == {(5, 10), (1, -1)} and no == 2 and nv == 6
    assert img[10 - my, 5 - mx] == 0 and img[2 - my, -15 - mx] == 255 and img[400 - my, 300 - mx] == 255


def test_empty_and_all_out_of_range(oracle):
    img, geo, mn, nv, no = oracle.bev_project(np.zeros((0, 4), np.float32))
    assert img.shape == (0, 0) and nv == 0 and no == 0
    img, geo, mn, nv, no = oracle.bev_project(np.array([[200, 0, 0, 0]], np.float32))
    assert img.shape == (0, 0)


def test_crop_pad(oracle):
    rng = np.random.default_rng(0)
    big = rng.integers(0, 2, (900, 1000), dtype=np.uint8) * 255
    out = oracle.crop_pad(big, 768, 768)
    assert np.array_equal(out, big[66:66 + 768, 116:116 + 768])
    small = rng.integers(0, 2, (504, 781), dtype=np.uint8) * 255   # wider than 768, shorter
    out = oracle.crop_pad(small, 768, 768)
    assert (out[:132] == 255).all() and (out[132 + 504:] == 255).all()
    assert np.array_equal(out[132:132 + 504], small[:, 6:6 + 768])


def test_synthetic_scan_properties(oracle):
    scan = synth.make_lidar_scan(seed=7)
    img, (ox, oy, res), (mx, my), nv, no = oracle.bev_project(scan)
    assert no == (img == 0).sum() and 500 < no < nv
    # the ground disc alone (one voxel per column) projects to an all-free image
    g = scan[np.abs(scan[:, 2] + 1.73) < 1e-3]
    img_g, *_ , no_g = oracle.bev_project(g)
    assert no_g == 0 and (img_g == 255).all()
    # stride 3 == stride 4
    img3, *_ = oracle.bev_project(np.ascontiguousarray(scan[:, :3]))
    assert np.array_equal(img3, img)


def test_golden_scan(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "bev_kitti_subsample.npz"))
    img, (ox, oy, res), (mx, my), nv, no = oracle.bev_project(z["pts"])
    assert np.array_equal(np.packbits(img == 0), z["occupied_bits"]) and img.shape == tuple(z["shape"])
    assert (mx, my, no) == tuple(int(v) for v in z["min_and_count"])


def test_reference_scan_matches_survey():
    # /root/reference exists only in the build container (781 x 504, 4 698 occupied)
    from oracle import pyoracle as po

    f = "/root/reference/s2s_libtorch/000000.bin"
    if not os.path.exists(f):
        pytest.skip("reference tree not present")
    pts = np.fromfile(f, np.float32).reshape(-1, 4)
    img, geo, mn, nv, no = po.bev_project(pts)
    assert img.shape == (504, 781) and no == 4698


def test_oracle_agrees_with_an_independent_numpy_restatement(oracle):
    # a second restatement of get_projected_grid, written from the sources' description
    #  with numpy set operations instead of a sort
    f32 = np.float32
    for seed in (21, 22):
        scan = synth.make_lidar_scan(seed=seed, n_walls=25)
        xyz = scan[:, :3].astype(f32)
        rng = np.sqrt((xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1]) + xyz[:, 2] * xyz[:, 2], dtype=f32)
        hits = xyz[rng <= f32(100.0)]
        q = hits / f32(0.2)                                                       # float division
        vox = np.where(q >= 0, np.floor(q + f32(0.5)), np.ceil(q - f32(0.5))).astype(np.int64)   # lround
        vox = np.unique(vox, axis=0)                                              # distinct hit voxels
        cols, counts = np.unique(vox[:, :2], axis=0, return_counts=True)
        mnx, mny = vox[:, 0].min(), vox[:, 1].min()
        w, h = vox[:, 0].max() - mnx + 1, vox[:, 1].max() - mny + 1
        img = np.full((h, w), 255, np.uint8)
        occ = cols[counts >= 2]                                                   # 2 x 0.55 > 0.9
        img[occ[:, 1] - mny, occ[:, 0] - mnx] = 0
        o_img, (ox, oy, res), (mx, my), nv, no = oracle.bev_project(scan)
        assert np.array_equal(o_img, img) and (mx, my) == (mnx, mny) and nv == len(vox) and no == len(occ)
        assert ox == mnx * float(f32(0.2)) and oy == mny * float(f32(0.2))
