"""Executable model (numpy) of the column arithmetic of the paired bit-plane layout in
gloc3d_b200/csrc/csm.cu (csm_coarse_bits_kernel<NP, true>, pmb_write_planes) and of the rule in
csm_api.cu::csm_make_plan that admits it.  A plane row holds plane column c of the coarsest level for
c in [0, 64); only columns [c_lo, c_hi] can hold data.  The layout stores two 32-column halves starting at
b0 = c_lo and b1 = max(c_lo, c_hi - 31); a point whose window of ncx candidate columns starts at plane
column ax reads ONE half: half 0 when the window ends inside it, else half 1, shifted by 32 + a with
a = ax - b_h through `(row << 32) >> (32 + a)` -- right for a >= 0, LEFT (zeros in) for a < 0 -- and a
start beyond half 1 reads the all-zero rows.  The model checks, for every grid width the rule admits and
every window position, that the 16-bit field the kernel feeds to its counters equals the true columns."""
import numpy as np


def plan(nx, n_lin, depth):
    """csm_make_plan: (eligible, b0, b1, max_side)."""
    top = depth - 1
    w = 1 << top
    max_side = (2 * n_lin) // w + 1
    wide_nx = nx + w - 1
    if max_side > 16 or ((wide_nx + n_lin - 1) >> top) >= 64:
        return False, 0, -1, max_side
    c_lo, c_hi = n_lin // w, (wide_nx - 1 + n_lin) // w
    b1 = max(c_lo, c_hi - 31)
    ok = max_side <= 14 and b1 - c_lo <= 33 - max_side
    return ok, c_lo, b1, max_side


def kernel_field(row64, ax, ncx, b0, b1):
    """point_addr4 + point_words for one row: the low 16 bits the counters see (None: zero rows)."""
    a0 = ax - b0
    second = a0 > 32 - ncx
    a = a0 - (b1 - b0) if second else a0
    if a > 31:
        return 0                                   # row index replaced by the all-zero rows
    sh = max(a, -31) + 32
    base = b1 if second else b0
    half = (row64 >> base) & 0xFFFFFFFF            # pmb_write_planes: (uint32_t)(bits >> b_h)
    return ((half << 32) >> sh) & 0xFFFF


def test_every_admitted_width_and_window():
    rng = np.random.default_rng(0)
    checked = 0
    for depth, n_lin in ((5, 100), (5, 104), (4, 50), (5, 60), (3, 20), (6, 200), (5, 7), (2, 13)):
        w = 1 << (depth - 1)
        for nx in list(range(1, 64)) + list(range(64, 1100, 7)) + [781, 800, 813, 814, 815, 816]:
            ok, b0, b1, max_side = plan(nx, n_lin, depth)
            if not ok:
                continue
            wide_nx = nx + w - 1
            c_lo, c_hi = n_lin // w, (wide_nx - 1 + n_lin) // w
            # data only in [c_lo, c_hi]: bit c of a plane row = level bit w c + rx - n_lin, 0 <= . < wide_nx
            for rx in (0, w // 2, w - 1):
                for _ in range(2):
                    row = 0
                    for c in range(64):
                        lx = w * c + rx - n_lin
                        if 0 <= lx < wide_nx and rng.random() < 0.5:
                            row |= 1 << c
                    assert row >> (c_hi + 1) == 0 and row & ((1 << c_lo) - 1) == 0
                    for ncx in {1, max_side // 2 + 1, max_side}:
                        for ax in range(-40, 110):
                            got = kernel_field(row, ax, ncx, b0, b1)
                            want = sum(((row >> (ax + i)) & 1) << i for i in range(ncx) if 0 <= ax + i < 64)
                            assert got & ((1 << ncx) - 1) == want, (depth, n_lin, nx, rx, ncx, ax)
                            checked += 1
    assert checked > 100000


def test_the_widest_kitti_grid_is_admitted_and_one_more_column_is_not():
    assert plan(800, 100, 5) == (True, 6, 26, 13)
    assert plan(813, 100, 5)[0] and not plan(814, 100, 5)[0]
    assert not plan(800, 120, 5)[0]            # 16 candidates per axis: the 64-bit-row layout
    assert plan(150, 24, 4)[0]


def test_lane_row_split_covers_the_lattice():
    """Lane 0 takes candidate rows from 0, lane 1 from 2 NP - 1; the last counter row of either lane is unused:
    4 NP - 2 usable rows >= max_side, and for either parity of the first row the NP stored pairs a lane loads
    contain all of its usable rows."""
    for max_side in range(1, 15):
        NP = (max_side + 5) // 4
        assert 4 * NP - 2 >= max_side and NP <= 4
        for ay in range(0, 6):
            for half in (0, 1):
                row0 = half * (2 * NP - 1)
                par0 = ay & 1
                first_word = (ay >> 1) + half * (NP - 1 + par0)          # point_words: addr + lane_words + (par0 & half)
                par = par0 ^ half
                loaded = {2 * (first_word + j) + k for j in range(NP) for k in (0, 1)}      # plane rows
                usable = [ay + row0 + ly for ly in range(2 * NP - 1)]
                assert set(usable) <= loaded, (max_side, ay, half)
                # re-pairing: counter word j = plane rows (2 first_word + 2 j + par, + 1)
                for j in range(NP):
                    lo_row = 2 * first_word + 2 * j + par
                    assert lo_row == ay + row0 + 2 * j


def test_row_major_storage_transposes_to_plane_major_staging():
    """pmb_write_planes stores word (rp, plane, h) at (rp w^2 + plane) 2 + h (64-bit rows: (r, plane) at
    r w^2 + plane); the scorer stages global word i at (i & (np2 - 1)) * per_plane + (i >> lp2) and addresses
    ((plane << 1) | h) * rpc + rp (64-bit rows: plane * rows + r)."""
    for log2w in (1, 2, 4):
        w2 = 1 << (2 * log2w)
        for rows in (6, 18, 90):
            rpc = rows // 2
            # paired
            np2, lp2, per_plane = 2 * w2, 2 * log2w + 1, rpc
            seen = set()
            for rp in range(rpc):
                for plane in range(w2):
                    for h in (0, 1):
                        i = (rp * w2 + plane) * 2 + h
                        staged = (i & (np2 - 1)) * per_plane + (i >> lp2)
                        assert staged == ((plane << 1) | h) * rpc + rp
                        seen.add(staged)
            assert seen == set(range(w2 * rows))
            # 64-bit rows
            np2, lp2, per_plane = w2, 2 * log2w, rows
            for r in range(rows):
                for plane in range(w2):
                    i = r * w2 + plane
                    assert (i & (np2 - 1)) * per_plane + (i >> lp2) == plane * rows + r
