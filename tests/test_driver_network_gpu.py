"""The driver with MODEL = network weights (tools/global_localization.cpp, gloc_desc_extract)
against the Python mirror of the same path.  Opt-in until the encoder has run on a GPU once."""
import os
import re
import subprocess

import numpy as np
import pytest

import gloc3d_b200 as g
from gloc3d_b200 import synth, weights
from test_driver import BIN, build, make_drive



@pytest.mark.gpu
def test_driver_computes_its_descriptors(tmp_path):
    build()
    tmp = str(tmp_path)
    valset, poses, _, files, _, db_pose, q_pose = make_drive(tmp, n_db=56, n_q=4)
    ws, bs = synth.hashed_vgg_weights(1)
    cw, cent, hid = synth.hashed_vlad_weights(64, 512, 512, 2)
    model = os.path.join(tmp, "model.glocw")
    weights.save_weights(model, ws, bs, cw, cent, hid)
    r = subprocess.run([BIN, valset, poses, model], capture_output=True, text=True, cwd=tmp, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "MODEL: 13 convolutions + NetVLAD_fc" in r.stderr and "Recall @ 1:" in r.stderr
    rec1 = float(re.search(r"Recall @ 1: ([0-9.eE+-]+)", r.stderr).group(1))

    # the same descriptors through the Python mirror: BEV plane -> DescriptorExtractor -> retrieval
    n_db = len(db_pose)
    bev = g.BevProjector(0)
    planes = []
    for f in files:
        bev.project(np.fromfile(f, np.float32).reshape(-1, 4))
        planes.append(bev.cnn_input(768, 768))
    ex = g.DescriptorExtractor(ws, bs, cw, cent, hid)
    desc = np.concatenate([ex.describe(np.stack(planes[i:i + 16])) for i in range(0, len(planes), 16)])
    ix = g.KnnIndex(512, 0)
    ix.set_db(desc[:n_db])
    idx, _ = ix.query(desc[n_db:], 20)
    hits = valid = 0
    for qi, (x, y, _) in enumerate(q_pose):
        pos = {i for i, (dx, dy, _) in enumerate(db_pose) if (dx - x) ** 2 + (dy - y) ** 2 < 16.0}
        if pos:
            valid += 1
            hits += int(idx[qi, 0]) in pos
    assert abs(rec1 - hits / max(valid, 1)) < 1e-6
    for h in (bev, ex, ix):
        h.close()
