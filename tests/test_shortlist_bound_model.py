"""Executable model (numpy) of the shortlist error bound of DESIGN.md §2.4 /
gloc3d_b200/csrc/knn_shortlist.cu: FP16 operands after a power-of-two scaling, FP32 truncating
accumulation in 32 steps of 16 exact products, score = ||x||^2 - 2 dot, compared with the
reference's evalMetric distance.  Checks |D_apx - D_ref| <= eps(q) on friendly and hostile
data, with the constants read from the kernel source."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "gloc3d_b200", "csrc", "knn_shortlist.cu")).read()


def const(name):
    m = re.search(r"constexpr float %s = ([^;]+);" % name, SRC)
    return float(eval(m.group(1).replace("f", "")))


kU, kCAcc, kC2, kInfl = const("kU"), const("kCAcc"), const("kC2"), const("kInfl")
f32 = np.float32


def pow2_scale_for(max_abs):
    if not (max_abs > 0 and max_abs < 3.0e38):
        return f32(1)
    _, e = np.frexp(f32(max_abs))
    return f32(np.ldexp(1.0, 14 - int(e)))


def eval_metric(q, x):
    """nanoflann L2_Simple_Adaptor::evalMetric in float32: groups of four, left-associated."""
    d = (q[None, :] - x).astype(f32)
    sq = (d * d).astype(f32)
    g = ((sq[:, 0::4] + sq[:, 1::4]).astype(f32) + sq[:, 2::4]).astype(f32)
    g = (g + sq[:, 3::4]).astype(f32)
    acc = np.zeros(len(x), f32)
    for j in range(g.shape[1]):
        acc = (acc + g[:, j]).astype(f32)
    return acc


def trunc_f32(v):
    """float64 -> float32 rounding toward zero (the pessimistic accumulator)."""
    r = v.astype(f32)
    over = np.abs(r.astype(np.float64)) > np.abs(v)
    r[over] = np.nextafter(r[over], f32(0))
    return r


def model_scores(q, db):
    sx = pow2_scale_for(np.abs(db).max())
    sq = pow2_scale_for(np.abs(q).max())
    xh = (db * sx).astype(np.float16)
    qh = (q * sq).astype(np.float16)
    dx2 = (((db - xh.astype(f32) / sx).astype(np.float64)) ** 2).sum(axis=1)
    dq2 = (((q - qh.astype(f32) / sq).astype(np.float64)) ** 2).sum()
    acc = np.zeros(len(db), f32)
    for kb in range(0, db.shape[1], 16):     # 16 exact products per step, truncating add
        part = (xh[:, kb:kb + 16].astype(np.float64) * qh[kb:kb + 16].astype(np.float64)).sum(axis=1)
        acc = trunc_f32(acc.astype(np.float64) + part)
    xn = (db.astype(f32) ** 2).sum(axis=1, dtype=f32)
    qn = f32((q.astype(f32) ** 2).sum(dtype=f32))
    cm = f32(-2) / sx / sq
    s = (xn + cm * acc).astype(f32)           # the epilogue's FFMA, modelled with two roundings
    d_apx = (qn + s).astype(f32)
    xmax = f32(np.sqrt(xn.max())) * f32(kInfl)
    dxmax = f32(np.sqrt(dx2.max())) * f32(kInfl)
    qnorm = f32(np.sqrt(qn)) * f32(kInfl)
    dq = f32(np.sqrt(dq2)) * f32(kInfl)
    eps = (2 * (dq * xmax + (1 + kU) * qnorm * dxmax) + kCAcc * qnorm * xmax
           + kC2 * (qnorm + xmax) ** 2 + 1e-30)
    return d_apx, float(eps)


def cases():
    rng = np.random.default_rng(7)
    n, d = 2000, 512
    unit = rng.standard_normal((n, d)).astype(f32)
    unit /= np.linalg.norm(unit, axis=1, keepdims=True)
    yield "unit-norm", unit, unit[3] + 0.01 * rng.standard_normal(d).astype(f32)
    yield "near-duplicates", np.repeat(unit[:250], 8, axis=0) + 1e-4 * rng.standard_normal((n, d)).astype(f32), unit[5]
    spiky = unit.copy()
    spiky[:, 0] *= 300.0                      # one dominant dimension pushes the rest towards FP16 underflow
    yield "dominant-dim", spiky, spiky[11] * f32(1.001)
    wide = (rng.standard_normal((n, d)) * np.exp(rng.uniform(-12, 0, (n, d)))).astype(f32)
    yield "12-decades", wide, wide[9] + f32(1e-6)
    yield "tiny", unit * f32(1e-18), unit[0] * f32(1e-18)
    yield "huge", unit * f32(1e15), unit[1] * f32(1e15)
    yield "mismatched-scales", unit * f32(1e-3), unit[2] * f32(50.0)
    nonneg = np.abs(unit)                     # same-sign products: the truncation bias accumulates
    yield "non-negative", nonneg, nonneg[4]


def test_bound_holds_under_the_pessimistic_accumulator():
    for name, db, q in cases():
        db = np.ascontiguousarray(db, f32)
        q = np.ascontiguousarray(q, f32)
        d_ref = eval_metric(q, db).astype(np.float64)
        d_apx, eps = model_scores(q, db)
        err = np.abs(d_apx.astype(np.float64) - d_ref).max()
        assert np.isfinite(eps) and err <= eps, (name, err, eps)


def test_bound_is_not_vacuous_on_descriptor_like_data():
    # at config 1 the shortlist keeps ~30 rows per query because 2 eps is ~2e-3 of the distances
    name, db, q = next(cases())
    d_ref = eval_metric(q, db)
    _, eps = model_scores(q, db)
    assert 2 * eps < 0.01 * float(np.median(d_ref)), (eps, float(np.median(d_ref)))
