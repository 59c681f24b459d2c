"""Grid store file (include/gloc3d.h "grid store file"): the C ABI's host-only reader/writer
against an independent numpy implementation of the documented layout.  No device needed; the
store-level save/load (device round trip) is at the bottom, GPU-marked."""
import os
import struct

import numpy as np
import pytest

import gloc3d_b200 as g
from gloc3d_b200 import _lib


def np_write(path, grids):
    """The documented layout, written with struct + numpy.packbits."""
    with open(path, "wb") as f:
        f.write(b"GLOCGRD1" + struct.pack("<IIQ", 1, 0, len(grids)))
        for cells, res, mx, my in grids:
            ny, nx = cells.shape
            flat = np.ascontiguousarray(cells, np.uint8).ravel()
            binary = bool(np.all((flat == 0) | (flat == 255)))
            payload = np.packbits(flat != 0, bitorder="little").tobytes() if binary else flat.tobytes()
            f.write(struct.pack("<iidddIIQ", nx, ny, res, mx, my, 1 if binary else 0, 0, len(payload)))
            f.write(payload)


def np_read(path):
    out = []
    with open(path, "rb") as f:
        assert f.read(8) == b"GLOCGRD1"
        version, _, n = struct.unpack("<IIQ", f.read(16))
        assert version == 1
        for _ in range(n):
            nx, ny, res, mx, my, enc, _, nbytes = struct.unpack("<iidddIIQ", f.read(48))
            raw = np.frombuffer(f.read(nbytes), np.uint8)
            if enc == 1:
                cells = np.unpackbits(raw, bitorder="little")[:nx * ny].astype(np.uint8) * 255
            else:
                cells = raw.copy()
            out.append((cells.reshape(ny, nx), res, mx, my))
        assert f.read() == b""
    return out


def make_grids(seed=0):
    rng = np.random.default_rng(seed)
    grids = []
    for nx, ny in ((781, 504), (13, 7), (1, 1), (64, 64), (9, 1)):      # ragged sizes, bits not a multiple of 8
        cells = ((rng.random((ny, nx)) < 0.02) * 255).astype(np.uint8)
        grids.append((cells, 0.2, float(rng.uniform(-50, 50)), float(rng.uniform(-50, 50))))
    graded = rng.integers(0, 256, (31, 17)).astype(np.uint8)            # non-binary: stored raw
    grids.append((graded, 0.05, 1.25, -3.5))
    return grids


def same(a, b):
    return (len(a) == len(b) and
            all(np.array_equal(x[0], y[0]) and tuple(x[1:]) == tuple(y[1:]) for x, y in zip(a, b)))


def test_c_writer_produces_the_documented_layout(tmp_path):
    grids = make_grids()
    p = str(tmp_path / "c.grd")
    g.write_grid_file(p, grids)
    assert same(np_read(p), grids)
    q = str(tmp_path / "np.grd")
    np_write(q, grids)
    assert open(p, "rb").read() == open(q, "rb").read()                 # byte for byte
    # bit-packed: a KITTI-sized grid takes nx*ny/8 bytes + 48
    assert os.path.getsize(p) < sum(c.size for c, *_ in grids[:5]) // 8 + 31 * 17 + 24 + 48 * 6 + 8


def test_c_reader_reads_the_documented_layout(tmp_path):
    grids = make_grids(3)
    p = str(tmp_path / "np.grd")
    np_write(p, grids)
    assert same(g.read_grid_file(p), grids)
    np_write(p, [])
    assert g.read_grid_file(p) == []


def test_sizing_call_does_not_consume_a_record(tmp_path):
    import ctypes as C

    grids = make_grids(5)[:2]
    p = str(tmp_path / "a.grd")
    g.write_grid_file(p, grids)
    L = _lib.lib()
    h, n = C.c_void_p(), C.c_size_t()
    assert L.gloc_grid_file_open(p.encode(), C.byref(h), C.byref(n)) == 0 and n.value == 2
    info = _lib.GridInfo()
    for _ in range(3):                                                   # as often as one likes
        assert L.gloc_grid_file_next(h, C.byref(info), None, 0) == 0
        assert (info.nx, info.ny) == grids[0][0].shape[::-1]
    small = np.empty(10, np.uint8)                                        # too small: still not consumed
    assert L.gloc_grid_file_next(h, C.byref(info), small.ctypes.data, small.size) == 0
    for cells, *_ in grids:
        buf = np.empty(cells.shape, np.uint8)
        assert L.gloc_grid_file_next(h, C.byref(info), buf.ctypes.data, buf.size) == 0
        assert np.array_equal(buf, cells)
    assert L.gloc_grid_file_next(h, C.byref(info), None, 0) == _lib.GLOC_ERR_RANGE
    L.gloc_grid_file_close(h)
    L.gloc_grid_file_close(None)


def test_malformed_files_are_rejected(tmp_path):
    grids = make_grids(7)[:2]
    p = str(tmp_path / "a.grd")
    g.write_grid_file(p, grids)
    blob = open(p, "rb").read()
    with pytest.raises(g.GlocError, match="cannot open"):
        g.read_grid_file(str(tmp_path / "missing.grd"))
    bad = str(tmp_path / "bad.grd")
    open(bad, "wb").write(b"NOTAGRID" + blob[8:])
    with pytest.raises(g.GlocError, match="not a grid store file"):
        g.read_grid_file(bad)
    open(bad, "wb").write(blob[:8] + struct.pack("<I", 2) + blob[12:])
    with pytest.raises(g.GlocError, match="unsupported version"):
        g.read_grid_file(bad)
    open(bad, "wb").write(blob[:len(blob) - 5])
    with pytest.raises(g.GlocError, match="truncated"):
        g.read_grid_file(bad)
    open(bad, "wb").write(blob[:24] + struct.pack("<ii", -4, 3) + blob[32:])
    with pytest.raises(g.GlocError, match="corrupt record 0"):
        g.read_grid_file(bad)
    with pytest.raises(g.GlocError):
        g.write_grid_file(str(tmp_path / "no_such_dir" / "x.grd"), grids)
    with pytest.raises(g.GlocError, match="is empty"):
        g.write_grid_file(bad, [(np.zeros((0, 4), np.uint8), 0.2, 0.0, 0.0)])


def test_store_level_entry_points_validate_arguments_without_gpu():
    L = _lib.lib()
    assert L.gloc_csm_save_grids(None, b"x") == _lib.GLOC_ERR_INVALID
    assert L.gloc_csm_load_grids(None, b"x", None, None) == _lib.GLOC_ERR_INVALID
    assert L.gloc_csm_get_grid_info(None, 0, None) == _lib.GLOC_ERR_INVALID


@pytest.mark.gpu
def test_store_round_trip_through_the_device(tmp_path):
    grids = make_grids(11)
    st = g.CsmStore(0)
    for cells, res, mx, my in grids:
        st.add_grid_u8(cells, res, mx, my)
    p = str(tmp_path / "store.grd")
    st.save_grids(p)
    assert same(np_read(p), grids)
    st2 = g.CsmStore(0)
    ids = st2.load_grids(p)
    assert ids == list(range(len(grids))) and len(st2) == len(grids)
    for gid, (cells, res, mx, my) in zip(ids, grids):
        lim = st2.grid_info(gid)
        assert (lim.num_x_cells, lim.num_y_cells) == cells.shape[::-1]
        assert (lim.resolution, lim.max_x, lim.max_y) == (res, mx, my)
        assert np.array_equal(st2.precomputation_grid(gid, 1), cells)
    st.close()
    st2.close()
