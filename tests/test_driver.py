"""The evaluation driver (tools/global_localization.cpp): same command line, input formats,
log lines and result files as /root/reference/registration/global_localization.cpp.  CPU:
it compiles against the C ABI and reads the reference's file formats.  GPU: on a synthetic
drive (valset + poses + raw scans + descriptor table) its recalls and success counts equal
the same pipeline run through the Python mirrors of the C ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "_global_localization")


def build():
    cmd = ["g++", "-O2", "-std=c++14", os.path.join(ROOT, "tools", "global_localization.cpp"), "-o", BIN,
           f"-L{ROOT}/gloc3d_b200", "-lgloc3d", f"-Wl,-rpath,{ROOT}/gloc3d_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def make_drive(tmp, n_db=60, n_q=6, seed=5):
    """A synthetic drive through a field of walls: database keyframes every ~3 m along a
    path, queries near some of them with their own heading."""
    rng = np.random.default_rng(seed)
    walls = []
    for _ in range(200):
        x0, y0 = rng.uniform(-40, 240), rng.uniform(-90, 90)
        ang, length = rng.uniform(0, np.pi), rng.uniform(4, 30)
        n = int(length * 60)      # ~12 points per 0.2 m cell: every wall column holds >= 2 voxels
        t = rng.uniform(0, length, n)
        z = rng.uniform(-1.5, 2.0, n)
        walls.append(np.stack([x0 + t * np.cos(ang), y0 + t * np.sin(ang), z], axis=1))
    world = np.concatenate(walls)

    def scan_at(x, y, yaw):
        d = world[:, :2] - np.array([x, y])
        near = (d ** 2).sum(1) < 70.0 ** 2
        c, s = np.cos(-yaw), np.sin(-yaw)
        p = d[near]
        pts = np.stack([c * p[:, 0] - s * p[:, 1], s * p[:, 0] + c * p[:, 1], world[near, 2],
                        np.full(near.sum(), 0.3)], axis=1)
        return pts.astype(np.float32)

    db_pose = [(3.0 * i, 8.0 * np.sin(0.05 * i), 0.04 * i) for i in range(n_db)]
    q_src = rng.choice(n_db, n_q, replace=False)
    q_pose = [(db_pose[j][0] + rng.uniform(-1.2, 1.2), db_pose[j][1] + rng.uniform(-1.2, 1.2),
               db_pose[j][2] + rng.uniform(-0.6, 0.6)) for j in q_src]
    files = []
    for i, (x, y, yaw) in enumerate(db_pose + q_pose):
        f = os.path.join(tmp, f"{i:06d}.bin")
        scan_at(x, y, yaw).tofile(f)
        files.append(f)
    # descriptor table: a query's descriptor is its source keyframe's plus noise
    feats = (rng.standard_normal((n_db + n_q, 512)) / np.sqrt(512)).astype(np.float32)
    feats[n_db:] = feats[q_src] + (rng.standard_normal((n_q, 512)) * 0.01).astype(np.float32)
    model = os.path.join(tmp, "descriptors.bin")
    feats.tofile(model)
    # valset (dataset/kitti_i2i.py:76-104) and poses (:108-120, "qx qy qz qw x y z")
    valset = os.path.join(tmp, "valset.txt")
    with open(valset, "w") as f:
        f.write(f"{n_db} {n_q}\n")
        for p in files:
            f.write(p + "\n")
        for qi, (x, y, _) in enumerate(q_pose):
            pos = [i for i, (dx, dy, _) in enumerate(db_pose) if (dx - x) ** 2 + (dy - y) ** 2 < 16.0]
            f.write(f"{qi}:" + " ".join(str(i) for i in pos) + "\n")
    poses = os.path.join(tmp, "poses.txt")
    with open(poses, "w") as f:
        for (x, y, yaw) in db_pose + q_pose:
            f.write(f"0 0 {np.sin(yaw / 2):.9f} {np.cos(yaw / 2):.9f} {x:.6f} {y:.6f} 0\n")
    return valset, poses, model, files, feats, db_pose, q_pose


def test_driver_compiles_and_reads_formats(tmp_path):
    build()
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 1 and "usage: global_localization VALSET GT_POSE MODEL" in r.stderr
    r = subprocess.run([BIN, str(tmp_path / "missing.txt"), "x", "y"], capture_output=True, text=True)
    assert r.returncode == 1 and "failed to open file" in r.stdout


def test_driver_reads_a_whole_drive_without_gpu(tmp_path):
    # valset (dataset/kitti_i2i.py:76-104), poses (:108-120), raw scans, descriptor table
    build()
    tmp = str(tmp_path)
    valset, poses, model, files, feats, db_pose, q_pose = make_drive(tmp, n_db=12, n_q=3)
    n_pos = sum(1 for x, y, _ in q_pose for dx, dy, _ in db_pose if (dx - x) ** 2 + (dy - y) ** 2 < 16.0)
    n_pts = sum(os.path.getsize(f) // 16 for f in files)
    env = dict(os.environ, GLOC_DRIVER_PARSE_ONLY="1")
    r = subprocess.run([BIN, valset, poses, model], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert (f"inputs: 12 db scans, 3 query scans, 15 poses, {n_pos} positives, 15 descriptors, {n_pts} points, "
            "0 unreadable scans") in r.stderr
    assert "db_num and db_files: 12, 12" in r.stderr and "Read poses with size: 15" in r.stderr
    # a truncated descriptor table is an error, as is a missing scan
    open(model, "wb").write(feats[:5].tobytes())
    r = subprocess.run([BIN, valset, poses, model], capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "MODEL must hold (db_num + q_num) x 512 float32" in r.stderr
    feats.tofile(model)
    os.remove(files[2])
    r = subprocess.run([BIN, valset, poses, model], capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "1 unreadable scans" in r.stderr


def test_driver_levels_scans_without_gpu(tmp_path):
    """4th argument = ground alignment (global_localization.cpp:431-440, :495-499): in parse-only
    mode the driver runs the host half of it -- every scan through the ground estimator."""
    from test_ground_host import make_scene

    build()
    tmp = str(tmp_path)
    heights = (1.73, 1.80, 1.65, 1.73)
    files = []
    for i, h in enumerate(heights):
        pts, _ = make_scene(40 + i, 0.02 * (i - 1), -0.015 * i, h, n_ground=15000, n_wall=6000)
        f = os.path.join(tmp, f"{i:06d}.bin")
        pts.tofile(f)
        files.append(f)
    with open(os.path.join(tmp, "valset.txt"), "w") as f:
        f.write("3 1\n" + "".join(p + "\n" for p in files) + "0:0 1\n")
    with open(os.path.join(tmp, "poses.txt"), "w") as f:
        f.write("0 0 0 1 0 0 0\n" * 4)
    np.zeros((4, 512), np.float32).tofile(os.path.join(tmp, "descriptors.bin"))
    env = dict(os.environ, GLOC_DRIVER_PARSE_ONLY="1")
    args = [BIN, os.path.join(tmp, "valset.txt"), os.path.join(tmp, "poses.txt"), os.path.join(tmp, "descriptors.bin")]
    r = subprocess.run(args + ["align"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"ground alignment: (\d+) of (\d+) scans levelled, mean sensor height ([0-9.]+) m", r.stderr)
    assert m and (int(m.group(1)), int(m.group(2))) == (4, 4), r.stderr
    assert abs(float(m.group(3)) - np.mean(heights)) < 0.05
    r = subprocess.run(args, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "ground alignment" not in r.stderr


def test_driver_accepts_the_network_weights_as_model(tmp_path):
    """MODEL = the weight container tools/export_weights.py writes (gloc3d_b200/weights.py): the
    driver reads and validates it on the host (parse-only mode; the forward needs the GPU)."""
    from gloc3d_b200 import synth, weights

    build()
    tmp = str(tmp_path)
    valset, poses, _, files, feats, db_pose, q_pose = make_drive(tmp, n_db=4, n_q=1)
    ws, bs = synth.hashed_vgg_weights(1)
    cw, cent, hid = synth.hashed_vlad_weights(64, 512, 512, 2)
    model = os.path.join(tmp, "model.glocw")
    weights.save_weights(model, ws, bs, cw, cent, hid)
    env = dict(os.environ, GLOC_DRIVER_PARSE_ONLY="1")
    r = subprocess.run([BIN, valset, poses, model], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "MODEL: 13 convolutions + NetVLAD_fc (64 clusters, 512-d)" in r.stderr
    assert "positives, network, " in r.stderr
    # a container with a wrong layer shape is refused
    ws[3] = ws[3][:, :64]
    weights.save_weights(model, ws, bs, cw, cent, hid)
    r = subprocess.run([BIN, valset, poses, model], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 1 and "layer 3 missing or of the wrong shape" in r.stderr


@pytest.mark.gpu
def test_driver_matches_python_pipeline(tmp_path):
    import gloc3d_b200 as g

    build()
    tmp = str(tmp_path)
    valset, poses, model, files, feats, db_pose, q_pose = make_drive(tmp)
    n_db, n_q = len(db_pose), len(q_pose)
    r = subprocess.run([BIN, valset, poses, model], capture_output=True, text=True, cwd=tmp, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    log = r.stderr
    rec = {int(k): float(v) for k, v in re.findall(r"Recall @ (\d+): ([0-9.eE+-]+)", log)}
    succ, total = map(int, re.search(r"\] (\d+), (\d+)\n", log).groups())
    for line in ("db_num and db_files: 60, 60", "q_num and q_files: 6, 6", "Read poses with size: 66",
                 "Success rate:", "Rot error:", "Pos error:", "Average 2D match costs", "Each query cost:",
                 "time cost for feature extraction:"):
        assert line in log, line
    assert os.path.exists(os.path.join(tmp, "failed_detect_indices.txt"))
    assert os.path.exists(os.path.join(tmp, "failed_registration_indices.txt"))

    # the same pipeline through the Python mirrors of the C ABI
    bev, st, ix = g.BevProjector(0), g.CsmStore(0), g.KnnIndex(512, 0)
    gids = []
    for f in files[:n_db]:
        bev.project(np.fromfile(f, np.float32).reshape(-1, 4))
        gids.append(bev.add_to_store_aligned(st))
    ix.set_db(feats[:n_db])
    ok, hits = 0, {1: 0, 5: 0, 10: 0, 20: 0}
    valid = 0
    for qi in range(n_q):
        bev.project(np.fromfile(files[n_db + qi], np.float32).reshape(-1, 4))
        pts = bev.occupied_points()
        idx, _ = ix.query(feats[n_db + qi:n_db + qi + 1], 20)
        cand = [int(v) for v in idx[0]]
        pos = {i for i, (dx, dy, _) in enumerate(db_pose)
               if (dx - q_pose[qi][0]) ** 2 + (dy - q_pose[qi][1]) ** 2 < 16.0}
        if pos:
            valid += 1
            for k in hits:
                hits[k] += any(c in pos for c in cand[:k])
        res = st.match_batch([pts], [gids[c] for c in cand], [0] * 20, [(0.0, 0.0, 0.0)] * 20, 100, 180,
                             2 * np.pi / 360, 5, 0.35)
        first = next((i for i, rr in enumerate(res) if rr.found), None)
        if first is None:
            continue
        db = cand[first]
        # ground truth q -> db (yaw-only poses) against the located pose
        dx, dy = q_pose[qi][0] - db_pose[db][0], q_pose[qi][1] - db_pose[db][1]
        c, s = np.cos(-db_pose[db][2]), np.sin(-db_pose[db][2])
        gt = (c * dx - s * dy, s * dx + c * dy, q_pose[qi][2] - db_pose[db][2])
        err_pos = np.hypot(gt[0] - res[first].pose_x, gt[1] - res[first].pose_y)
        err_rot = abs((gt[2] - res[first].pose_yaw + np.pi) % (2 * np.pi) - np.pi) * 180 / np.pi
        ok += bool(err_pos < 1.0 and err_rot < 5.0)
    assert total == n_q and succ == ok
    for k in hits:
        assert abs(rec[k] - hits[k] / valid) < 1e-6
    assert rec[1] == 1.0 and succ >= n_q - 1     # the planted drive is recoverable
    for h in (bev, st, ix):
        h.close()
