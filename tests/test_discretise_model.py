"""Executable model (numpy) of the float fast path of the scan discretisation in
gloc3d_b200/csrc/csm.cu (point_addr4 / the expand kernel): the cell of a transformed point is
the reference's double-precision GetCellIndex (map_limits.h:69-76: lround((max - w)/res - 0.5));
the kernels evaluate it in float and accept the float result only when its fractional part is
further than delta = 2^-23 (|max|/res + 3 U) from 0 and 1.  The model checks that an accepted
float cell is never wrong, on random and on adversarial (boundary-hugging) inputs, and that the
fast path is actually taken most of the time."""
import re
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def lround(x):
    return (np.sign(x) * np.floor(np.abs(x) + 0.5)).astype(np.int64)


def exact_cell(w, res, mx):
    """cell_exact(): double arithmetic on the float world coordinate."""
    return lround((mx - w.astype(np.float64)) / res - 0.5)


def fast_cell(w, res, mx, U):
    """The kernels' float path: returns (cell, accepted)."""
    mx_f, ir = f32(mx), f32(1.0 / res)
    delta = f32(1.1920929e-07) * (np.abs(mx_f) * ir + f32(3.0) * f32(U))
    hi1 = f32(1.0) - delta
    u = ((mx_f - w).astype(f32) * ir).astype(f32)
    fl = np.floor(u)
    d = (u - fl).astype(f32)
    ok = (d > delta) & (d < hi1) & (np.abs(u) < f32(U))
    return fl.astype(np.int64), ok, float(delta)


def test_constant_in_the_kernel_is_two_to_the_minus_23():
    src = open(os.path.join(ROOT, "gloc3d_b200", "csrc", "csm.cu")).read()
    consts = set(re.findall(r"const float delta = ([0-9.e+-]+)f \*", src))
    assert consts == {"1.1920929e-07"} and abs(float(consts.pop()) - 2.0 ** -23) < 1e-14


def test_accepted_float_cells_are_exact():
    rng = np.random.default_rng(0)
    taken = []
    for res, mx, n_cells in ((0.2, 80.0, 800), (0.2, -3000.0, 1000), (0.05, 5000.0, 4000), (0.3, 0.1, 300),
                             (0.1, 12345.678, 2000), (0.2, 1.0e5, 800)):
        U = n_cells + 2 * 100 + 32
        # random points over (and a little around) the grid
        w = (mx - rng.uniform(-50, n_cells + 50, 400000) * res).astype(f32)
        # adversarial: world coordinates whose exact cell coordinate hugs an integer
        k = rng.integers(-20, n_cells + 20, 400000)
        eps = rng.choice([0.0, 1e-9, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3], 400000) * rng.choice([-1.0, 1.0], 400000)
        w_adv = (mx - (k + eps) * res).astype(f32)
        for pts in (w, w_adv, np.nextafter(w_adv, f32(np.inf)), np.nextafter(w_adv, f32(-np.inf))):
            cell, ok, delta = fast_cell(pts, res, mx, U)
            ref = exact_cell(pts, res, mx)
            assert np.array_equal(cell[ok], ref[ok]), (res, mx, int((cell[ok] != ref[ok]).sum()))
        cell, ok, delta = fast_cell(w, res, mx, U)
        taken.append((res, mx, delta, ok.mean()))
    # the usual map (KITTI BEV around the origin): the double fallback is rare
    assert taken[0][3] > 0.997, taken   # 2 delta = 0.09 % of the points take the double path
    # far-from-origin maps widen delta but the float path still carries most points
    assert all(t[3] > 0.5 for t in taken if t[2] < 0.2), taken


def test_observed_float_error_stays_under_half_of_delta():
    """delta carries a 2x safety factor over the proven 2^-24 (|max|/res + 3 |u|) bound."""
    rng = np.random.default_rng(1)
    for res, mx, n_cells in ((0.2, 80.0, 800), (0.05, 5000.0, 4000), (0.2, 1.0e5, 800)):
        U = n_cells + 2 * 100 + 32
        w = (mx - rng.uniform(-50, n_cells + 50, 1000000) * res).astype(f32)
        mx_f, ir = f32(mx), f32(1.0 / res)
        u = ((mx_f - w).astype(f32) * ir).astype(np.float64)
        u_exact = (mx - w.astype(np.float64)) / res
        _, _, delta = fast_cell(w, res, mx, U)
        assert np.abs(u - u_exact).max() <= 0.5 * delta * 1.0001, (res, mx, np.abs(u - u_exact).max(), delta)
