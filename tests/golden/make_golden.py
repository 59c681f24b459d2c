"""Mint the golden fixtures under tests/golden/.  Run from the repo root in the build
container (needs /root/reference for the nanoflann build under oracle/_ref):

    python tests/golden/make_golden.py

knn_*.npz   outputs of the REFERENCE ITSELF (its nanoflann + adaptor compiled unmodified
            from /root/reference/registration, oracle/nanoflann_ref.cpp) on seeded inputs;
            inputs are stored too so the fixtures do not depend on numpy's RNG stream.
vlad_*.npz  outputs of the REFERENCE ITSELF (model/netvlad_fc.py, torch CPU) on hashed weights.
csm_*.npz   outputs of the REFERENCE ITSELF: registration/2d/*.cpp compiled unmodified into
            oracle/_ref/libcsm_ref.so (oracle/csm_ref.cpp + oracle/shim/): match result, rotated
            and discretised scans, precomputation grids.  The restatement (oracle/csm_oracle.c) is
            asserted equal at mint time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gloc3d_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def knn_case(name, n, dim, nq, k, seed, dup_run=0, sigma=None):
    db = synth.make_descriptors(n, dim, seed=seed, dup_run=dup_run)
    q = synth.make_queries(db, nq, seed=seed + 1, sigma=sigma)
    tree = po.RefTree(db, 10)
    idx, d2 = tree.query(q, k)
    np.savez_compressed(os.path.join(OUT, name), db=db, q=q, k=k, idx=idx, d2=d2)
    print(name, db.shape, q.shape, k)


def csm_case(name, nx, ny, seed, yaw, dx, dy, n_lin, n_ang, step, depth, min_score, graded=False):
    res = 0.2
    g = synth.make_bev_grid(nx, ny, seed=seed, n_segments=14, n_blobs=8, graded=graded)
    cells = synth.level1_to_cells(g)                      # Grid2D's uint16 cells
    g = po.level1_from_cells(cells)                       # the width-1 grid the reference derives
    assert np.array_equal(g, po.ref_precomp(cells, depth, 0))
    mx, my = synth.centered_limits(nx, ny, res)
    scan = synth.planted_scan(np.where(g > 0, 255, 0).astype(np.uint8), res, mx, my, yaw, dx, dy,
                              dropout=0.2, seed=seed + 1)
    init = (0.1, -0.05, 0.02)
    r = po.ref_csm_match(cells, res, mx, my, depth, scan, init, n_lin, n_ang, step, min_score)
    o = po.csm_match(g, res, mx, my, depth, scan, init, n_lin, n_ang, step, min_score, 0)
    result = (r.found, r.score, r.scan_index, r.x_offset, r.y_offset, r.pose_x, r.pose_y, r.pose_yaw)
    assert result == o.as_tuple(), (result, o.as_tuple())
    disc = po.ref_discretize(scan, init, n_ang, step, res, mx, my)
    assert np.array_equal(disc, po.discretize(scan, init, n_ang, step, res, mx, my))
    levels = {}
    for w in (2, 4, 16):
        levels[f"level{w}"] = po.ref_precomp(cells, 5, w.bit_length() - 1)
        assert np.array_equal(levels[f"level{w}"], po.precomp_from_level1(g, w))
    np.savez_compressed(os.path.join(OUT, name), grid=g, cells=cells, scan=scan, res=res, max_x=mx, max_y=my,
                        init=np.array(init), n_lin=n_lin, n_ang=n_ang, step=step,
                        depth=depth, min_score=min_score, result=np.array(result, np.float64),
                        cells_first=disc[0], cells_last=disc[-1], **levels)
    print(name, result)


def bev_case(name, every=4):
    """Every `every`-th point of the reference's only real scan (s2s_libtorch/000000.bin, a
    KITTI frame) and the BEV image the REFERENCE's own projection code computes for it (needs /root/reference)."""
    pts = np.fromfile("/root/reference/s2s_libtorch/000000.bin", np.float32).reshape(-1, 4)[::every].copy()
    img, (ox, oy, res), (mx, my), nv, no = po.bev_project(pts)
    # minted from the reference's own projection (oracle/_ref/libbev_ref.so): the oracle must agree
    rimg, rgeo = po.ref_bev_project(pts)
    assert np.array_equal(rimg, img) and rgeo == (ox, oy, res)
    img = rimg
    np.savez_compressed(os.path.join(OUT, name), pts=pts, occupied_bits=np.packbits(img == 0),
                        shape=np.array(img.shape), min_and_count=np.array([mx, my, no]))
    print(name, pts.shape, img.shape, no)


def vlad_case(name, K, C, H, W, B, seed):
    """Outputs of the REFERENCE's own NetVLAD_fc module (model/netvlad_fc.py imported from
    /root/reference, torch CPU float32) for hashed weights and features (oracle/vlad_oracle.py:
    integer-hash generators, so only the outputs need storing)."""
    import torch

    sys.path.insert(0, "/root/reference/model")
    import netvlad_fc
    from oracle import vlad_oracle as vo

    conv_w, cent, hid = vo.hashed_weights(K, C, C, seed)
    x = vo.hashed_features(B, C, H * W, seed + 10).reshape(B, C, H, W)
    m = netvlad_fc.NetVLAD(num_clusters=K, dim=C)
    with torch.no_grad():
        m.conv.weight.copy_(torch.from_numpy(conv_w)[:, :, None, None])
        m.centroids.copy_(torch.from_numpy(cent))
        m.hidden1_weights.copy_(torch.from_numpy(hid))
        out = m(torch.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(OUT, name), K=K, C=C, H=H, W=W, B=B, seed=seed, out=out,
                        probe=np.array([conv_w[0, 0], cent[-1, -1], hid[-1, -1], x[-1, -1, -1, -1]]))
    print(name, out.shape, float(np.abs(out).max()))


if __name__ == "__main__":
    bev_case("bev_kitti_subsample.npz")
    knn_case("knn_d512_k20.npz", 160, 512, 6, 20, 11)                      # reference k (loop_detector.h:98)
    knn_case("knn_d512_k25_dups.npz", 160, 512, 6, 25, 21, dup_run=8, sigma=0.002)
    knn_case("knn_d30_tail.npz", 200, 30, 5, 7, 31)                        # dim % 4 != 0 tail path
    csm_case("csm_binary.npz", 150, 110, 41, 0.35, 1.2, -0.8, 14, 24, np.pi / 90, 4, 0.3)
    csm_case("csm_graded.npz", 120, 140, 51, -0.6, -1.0, 0.6, 10, 20, np.pi / 60, 3, 0.2, graded=True)
    vlad_case("vlad_small.npz", 8, 32, 6, 5, 3, 61)
    vlad_case("vlad_full.npz", 64, 512, 48, 48, 2, 71)                     # the reference's sizes (768^2 input)
