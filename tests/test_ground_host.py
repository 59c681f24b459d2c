"""Host-side ground alignment and 6-DoF pose composition (gloc3d_b200/host/gloc_ground.hpp;
reference: registration/ground_estimator.cpp, global_localization.cpp:511-574) against
independent numpy/scipy restatements.  CPU only -- this is caller logic around the hot path."""
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_ground_cpu_test")


@pytest.fixture(scope="module")
def exe():
    r = subprocess.run(["g++", "-O2", "-std=c++14", "-ffp-contract=off", "-Wall", "-Werror",
                        os.path.join(ROOT, "tests", "cpp", "ground_cpu_test.cpp"), "-o", EXE],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return EXE


def run(exe, mode, n, values):
    text = f"{n}\n" + " ".join(repr(float(v)) for v in np.asarray(values, np.float64).ravel()) + "\n"
    r = subprocess.run([exe, mode], input=text, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def rows(out):
    return np.array([[float(v) for v in ln.split()] for ln in out.strip().splitlines()])


def eigen_euler_zyx(R):
    """Eigen's eulerAngles(2, 1, 0): R = Rz(a) Ry(b) Rx(c), a in [0, pi] -- built from scipy's
    canonical decomposition (yaw in [-pi, pi], pitch in [-pi/2, pi/2]) and the identity
    Rz(y) Ry(p) Rx(r) = Rz(y + pi) Ry(pi - p) Rx(r + pi)."""
    y, p, r = Rotation.from_matrix(R).as_euler("ZYX")
    if y < 0:
        y, p, r = y + np.pi, np.pi - p, r + np.pi
    wrap = lambda a: (a + np.pi) % (2 * np.pi) - np.pi
    return np.array([y, wrap(p), wrap(r)])


def rot(a, b, c):
    return Rotation.from_euler("ZYX", [a, b, c]).as_matrix()


def ang_close(a, b, tol):
    d = (np.asarray(a) - np.asarray(b) + np.pi) % (2 * np.pi) - np.pi
    return np.all(np.abs(d) < tol)


def test_euler_angles_follow_eigens_convention(exe):
    rng = np.random.default_rng(1)
    Rs = Rotation.random(300, random_state=2).as_matrix()
    # plus small tilts around the identity with both yaw signs: the case the ground transform hits
    small = [rot(y, p, r) for y, p, r in rng.uniform(-0.05, 0.05, (100, 3))]
    Rs = np.concatenate([Rs, np.array(small)])
    got = rows(run(exe, "euler", len(Rs), Rs.astype(np.float32)))
    for R, e in zip(Rs, got):
        assert 0.0 <= e[0] <= np.pi + 1e-6 and abs(e[1]) <= np.pi + 1e-6 and abs(e[2]) <= np.pi + 1e-6
        assert np.allclose(rot(*e), R, atol=2e-6)          # it is a decomposition of R
        assert ang_close(e, eigen_euler_zyx(R), 2e-4)      # and the one Eigen documents


def test_roll_pitch_yaw_quaternion(exe):
    rng = np.random.default_rng(3)
    rpy = rng.uniform(-np.pi, np.pi, (200, 3))
    got = rows(run(exe, "rpy", len(rpy), rpy))
    for (r, p, y), q in zip(rpy, got):
        ref = Rotation.from_euler("ZYX", [y, p, r]).as_quat()           # x y z w
        ref = np.array([ref[3], ref[0], ref[1], ref[2]])
        assert np.allclose(q, ref, atol=1e-12) or np.allclose(q, -ref, atol=1e-12)


def np_transform_to_ground(coeff):
    """ground_estimator.cpp:167-192 in numpy (float64)."""
    n = np.array(coeff[:3], np.float64)
    d = abs(coeff[3]) / np.linalg.norm(n)
    if coeff[2] < 0:
        n = -n
    n /= np.linalg.norm(n)
    z = np.array([0.0, 0.0, 1.0])
    axis = np.cross(n, z)
    s = np.sqrt((1 + n @ z) * 2)
    q = np.array([*(axis / s), s * 0.5])                                 # x y z w
    R = Rotation.from_quat(q / np.linalg.norm(q)).as_matrix()
    ypr = eigen_euler_zyx(R)
    T = np.eye(4)
    T[:3, :3] = rot(0.0, ypr[1], ypr[2])
    T[2, 3] = d
    return T


def test_transform_points_to_ground(exe):
    rng = np.random.default_rng(4)
    cases = []
    for _ in range(200):
        n = np.array([rng.normal(0, 0.08), rng.normal(0, 0.08), rng.choice([-1.0, 1.0])])
        n *= rng.uniform(0.5, 2.0)
        cases.append([*n, rng.uniform(-2.5, 2.5), *rng.uniform(-30, 30, 3), 0.25])
    out = [np.array([float(v) for v in ln.split()])
           for ln in run(exe, "transform", len(cases), np.array(cases, np.float32)).strip().splitlines()]
    flipped = 0
    for c, T, p in zip(np.array(cases, np.float32).astype(np.float64), out[0::2], out[1::2]):
        T = T.reshape(4, 4)
        ref = np_transform_to_ground(c[:4])
        assert np.allclose(T, ref, atol=3e-6), (T, ref)
        assert np.allclose(p[:3], ref[:3, :3] @ c[4:7] + ref[:3, 3], atol=2e-4) and p[3] == np.float32(0.25)
        # the transformed ground normal is +z: the plane becomes z = 0
        n = c[:3] / np.linalg.norm(c[:3]) * (1 if c[2] >= 0 else -1)
        assert np.allclose(T[:3, :3] @ n, [0, 0, 1], atol=1e-5)
        flipped += T[0, 0] < 0
    # Eigen's [0, pi] yaw range: about half of the tilts come back with a half turn about z
    assert 40 < flipped < 160


def np_compose(align, xy_yaw, Tq, Tdb):
    """global_localization.cpp:524-569 in numpy."""
    if align:
        inv = np.linalg.inv(Tdb)
        rpz = inv @ Tq
        e_rpz = eigen_euler_zyx(rpz[:3, :3])
        H = np.eye(4)
        H[:3, :3] = rot(xy_yaw[2], 0, 0)
        H[0, 3], H[1, 3] = xy_yaw[0], xy_yaw[1]
        yawxy = inv @ H @ Tq
        e_yawxy = eigen_euler_zyx(yawxy[:3, :3])
        roll, pitch, yaw = e_rpz[2], e_rpz[1], e_yawxy[0]
        t = [yawxy[0, 3], yawxy[1, 3], rpz[2, 3]]
    else:
        roll, pitch, yaw = 0.0, 0.0, xy_yaw[2]
        t = [xy_yaw[0], xy_yaw[1], 0.0]
    P = np.eye(4)
    P[:3, :3] = rot(yaw, pitch, roll)
    P[:3, 3] = t
    return P


def test_pose_composition(exe):
    rng = np.random.default_rng(5)
    cases, refs = [], []
    for i in range(200):
        align = i % 4 != 0
        xy_yaw = np.array([*rng.uniform(-15, 15, 2), rng.uniform(-np.pi, np.pi)], np.float32).astype(np.float64)
        Ts = []
        for _ in range(2):
            c = [rng.normal(0, 0.05), rng.normal(0, 0.05), 1.0, rng.uniform(1.5, 2.0)]
            Ts.append(np_transform_to_ground(c).astype(np.float32).astype(np.float64))
        cases.append(np.concatenate([[float(align)], xy_yaw, Ts[0].ravel(), Ts[1].ravel()]))
        refs.append(np_compose(align, xy_yaw, Ts[0], Ts[1]))
    out = rows(run(exe, "compose", len(cases), np.array(cases)))
    for c, P, ref in zip(cases, out, refs):
        P = P.reshape(4, 4)
        assert np.allclose(P[:3, 3], ref[:3, 3], atol=2e-4)
        assert np.allclose(P[:3, :3], ref[:3, :3], atol=5e-4), (c[0], P, ref)
        assert np.allclose(P[:3, :3] @ P[:3, :3].T, np.eye(3), atol=1e-5)


def test_composed_pose_recovers_a_planted_6dof_motion(exe):
    """Two sensors above one flat ground: perfect ground transforms plus the perfect planar match
    give back the planted relative pose under the reference's own error metric (which forgives a
    half turn, see TransformPointsToGround).  The reference's composition is first order in the
    tilt (roll, pitch and dz come from T_db^-1 T_q, which ignores the yaw and the offset between
    the frames), so the check uses tilts of a fraction of a degree."""
    rng = np.random.default_rng(6)
    cases, truth = [], []
    for _ in range(100):
        W = []
        for _ in range(2):   # world (ground frame, z up) <- sensor
            T = np.eye(4)
            T[:3, :3] = rot(rng.uniform(-np.pi, np.pi), rng.normal(0, 0.003), rng.normal(0, 0.003))
            T[:3, 3] = [*rng.uniform(-5, 5, 2), rng.uniform(1.6, 1.9)]
            W.append(T)
        Wdb, Wq = W
        Tl2g = []
        for Wx in (Wq, Wdb):
            # ground plane z_world = 0 in sensor coordinates: n = R^T z, d = height
            n = Wx[:3, :3].T @ np.array([0, 0, 1.0])
            Tl2g.append(np_transform_to_ground([*n, Wx[2, 3]]))
        Tq, Tdb = Tl2g
        # planar match between the two ground-aligned frames (what the scan matcher returns)
        H = Tdb @ np.linalg.inv(Wdb) @ Wq @ np.linalg.inv(Tq)
        assert np.allclose(H[2, :3], [0, 0, 1], atol=1e-9) and abs(H[2, 3]) < 1e-9
        xy_yaw = [H[0, 3], H[1, 3], np.arctan2(H[1, 0], H[0, 0])]
        cases.append(np.concatenate([[1.0], xy_yaw, Tq.ravel(), Tdb.ravel()]))
        truth.append((Wdb, Wq))
    located = rows(run(exe, "compose", len(cases), np.array(cases)))
    err_in = [np.concatenate([Wdb.ravel(), Wq.ravel(), L]) for (Wdb, Wq), L in zip(truth, located)]
    errs = rows(run(exe, "error", len(err_in), np.array(err_in)))
    assert np.all(errs[:, 0] < 1.0) and np.all(errs[:, 1] < 0.1), (errs[:, 0].max(), errs[:, 1].max())


def test_registration_error_metric(exe):
    rng = np.random.default_rng(7)
    cases, refs = [], []
    for i in range(100):
        M = []
        for _ in range(3):
            T = np.eye(4)
            T[:3, :3] = Rotation.random(random_state=int(rng.integers(1 << 30))).as_matrix()
            T[:3, 3] = rng.uniform(-20, 20, 3)
            M.append(T.astype(np.float32).astype(np.float64))
        if i % 3 == 0:   # a located pose half a turn away from the truth: forgiven by the metric
            q2db = np.linalg.inv(M[0]) @ M[1]
            M[2] = q2db.copy()
            M[2][:3, :3] = q2db[:3, :3] @ rot(np.pi - 0.02, 0, 0)
        q2db = np.linalg.inv(M[0]) @ M[1]
        tr = np.trace(q2db[:3, :3].T @ M[2][:3, :3])
        a = np.degrees(abs(np.arccos(np.clip(0.5 * (tr - 1), -0.999999, 0.999999))))
        if abs(a - 180) < 5:
            a = abs(a - 180)
        refs.append([a, np.linalg.norm(q2db[:3, 3] - M[2][:3, 3])])
        cases.append(np.concatenate([m.ravel() for m in M]))
    out = rows(run(exe, "error", len(cases), np.array(cases)))
    assert np.allclose(out, np.array(refs), atol=2e-2, rtol=1e-4)


def make_scene(seed, roll, pitch, height, n_ground=30000, n_wall=12000):
    """A tilted sensor over flat ground plus vertical walls, in sensor coordinates (x y z i)."""
    rng = np.random.default_rng(seed)
    r = np.sqrt(rng.uniform(2.0 ** 2, 30.0 ** 2, n_ground))
    a = rng.uniform(0, 2 * np.pi, n_ground)
    g = np.stack([r * np.cos(a), r * np.sin(a), rng.normal(0, 0.015, n_ground)], 1)
    walls = []
    for _ in range(12):
        p0 = rng.uniform(-18, 18, 2)
        d = rng.uniform(0, 2 * np.pi)
        L = rng.uniform(4, 12)
        t = rng.uniform(0, L, n_wall // 12)
        w = np.stack([p0[0] + t * np.cos(d), p0[1] + t * np.sin(d), rng.uniform(0, 2.5, len(t))], 1)
        walls.append(w + rng.normal(0, 0.01, w.shape))
    world = np.concatenate([g] + walls)
    R = rot(0.3, pitch, roll)                      # world <- sensor
    pts = (world - np.array([0, 0, height])) @ R   # = R^T (p - t)
    n_true = R.T @ np.array([0, 0, 1.0])
    return np.concatenate([pts, np.full((len(pts), 1), 0.5)], 1).astype(np.float32), n_true


def test_normals_match_an_independent_knn_pca(exe):
    pts, _ = make_scene(11, 0.02, -0.03, 1.73, n_ground=3000, n_wall=1200)
    xyz = pts[:, :3].astype(np.float64)
    got = rows(run(exe, "normals", len(xyz), xyz.astype(np.float32)))
    _, nn = cKDTree(xyz).query(xyz, k=10)
    agree = 0
    for i in range(len(xyz)):
        nb = xyz[nn[i]]
        w, v = np.linalg.eigh(np.cov(nb.T))
        n = v[:, 0]
        if -(xyz[i] @ n) < 0:
            n = -n
        # skip near-isotropic neighbourhoods where the smallest eigenvector is ill-conditioned
        if w[1] - w[0] > 1e-3 * w[2]:
            assert abs(got[i] @ n) > 0.999, (i, got[i], n)
            assert got[i] @ n > 0
            agree += 1
    assert agree > 0.8 * len(xyz)


def test_ground_is_found_and_levelled(exe):
    for seed, roll, pitch, h in ((21, 0.03, -0.02, 1.73), (22, -0.04, 0.05, 1.9), (23, 0.0, 0.0, 1.6)):
        pts, n_true = make_scene(seed, roll, pitch, h)
        out = run(exe, "ground", len(pts), pts).strip().splitlines()
        ok, n_near, n_ground = (int(v) for v in out[0].split())
        coeff = np.array([float(v) for v in out[1].split()])
        T = np.array([float(v) for v in out[2].split()]).reshape(4, 4)
        n_out = int(out[3])
        cloud = np.array([[float(v) for v in ln.split()] for ln in out[4:]])
        assert ok == 1 and n_ground > 0.5 * 0.4 * n_near          # most near points are ground
        n = coeff[:3] * np.sign(coeff[2])
        # RANSAC on three ground points with 1.5 cm noise: within ~1 degree and 10 cm of the truth
        assert np.degrees(np.arccos(np.clip(n @ n_true, -1, 1))) < 1.5
        assert abs(abs(coeff[3]) - h) < 0.1 and abs(T[2, 3] - abs(coeff[3])) < 1e-5
        assert n_out == pts.size and cloud.shape == (len(pts), 4)
        assert np.allclose(cloud[:, :3], pts[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3], atol=1e-3)
        assert np.array_equal(cloud[:, 3].astype(np.float32), pts[:, 3])
        ground_z = cloud[:30000, 2]                                   # the scene's ground points come first
        assert abs(np.median(ground_z)) < 0.1 and np.percentile(np.abs(ground_z - np.median(ground_z)), 90) < 0.4


def test_no_ground_returns_identity_and_an_empty_cloud(exe):
    rng = np.random.default_rng(31)
    t = rng.uniform(0, 10, 4000)
    wall = np.stack([np.full_like(t, 5.0) + rng.normal(0, 0.01, len(t)), t - 5, rng.uniform(-1, 2, len(t)),
                     np.zeros_like(t)], 1).astype(np.float32)   # a single vertical wall: normals horizontal
    out = run(exe, "ground", len(wall), wall).strip().splitlines()
    assert out[0].split()[0] == "0"
    assert np.allclose(np.array([float(v) for v in out[2].split()]).reshape(4, 4), np.eye(4))
    assert int(out[3]) == 0


KITTI_SCAN = "/root/reference/s2s_libtorch/000000.bin"


@pytest.mark.skipif(not os.path.exists(KITTI_SCAN), reason="the reference's one real scan (build container only)")
def test_ground_of_the_references_kitti_scan(exe):
    """KITTI's Velodyne sits 1.73 m above the road: the estimator has to find that ground."""
    pts = np.fromfile(KITTI_SCAN, np.float32).reshape(-1, 4)
    out = run(exe, "ground", len(pts), pts).strip().splitlines()
    ok, n_near, n_ground = (int(v) for v in out[0].split())
    coeff = np.array([float(v) for v in out[1].split()])
    assert ok == 1 and n_ground > 10000
    n = coeff[:3] * np.sign(coeff[2])
    assert abs(abs(coeff[3]) - 1.73) < 0.15
    assert np.degrees(np.arccos(n[2] / np.linalg.norm(n))) < 5.0
