"""Seeded synthetic inputs of KITTI shape: netvlad_fc-like
descriptors and BEV occupancy grids with planted query scans.  numpy only; the
same arrays feed the CPU oracle and the GPU path byte for byte.

Conventions follow the reference: descriptors are float32 [n, 512] row-major and
NOT L2-normalised (model/netvlad_fc.py:99-109); BEV grids are the matcher's
uint8 width-1 precomputation grid (0 = free ... 255 = occupied,
registration/2d/fast_correlative_scan_matcher_2d.cpp:184-190) stored
[num_y_cells][num_x_cells] so that flat index = num_x_cells*y + x
(registration/2d/grid_2d.cpp:168-171).
"""
from __future__ import annotations

import numpy as np

DIM = 512


def make_descriptors(n: int, dim: int = DIM, seed: int = 1234, dup_run: int = 0,
                     dup_sigma: float = 0.001) -> np.ndarray:
    """iid N(0, 1/dim) rows; with dup_run > 1 consecutive rows form random-walk
    runs of near-duplicates (KITTI-like consecutive frames)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    chunk = 65536
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        out[s:e] = (rng.standard_normal((e - s, dim)) * (1.0 / np.sqrt(dim))).astype(np.float32)
    if dup_run > 1:
        for s in range(0, n, dup_run):
            e = min(n, s + dup_run)
            steps = (rng.standard_normal((e - s - 1, dim)) * dup_sigma).astype(np.float32)
            out[s + 1:e] = out[s] + np.cumsum(steps, axis=0, dtype=np.float32)
    return out


MT_BLOCK = 16384     # rows per independent stream of make_descriptors_mt


def make_descriptors_mt(n: int, dim: int = DIM, seed: int = 1234, dup_run: int = 8,
                        dup_sigma: float = 0.001, threads: int = 0, rows=None) -> np.ndarray:
    """Same distribution as make_descriptors (iid N(0, 1/dim) run heads, random-walk runs of
    near-duplicates), generated block by block from independent streams `default_rng([seed, block])`
    on a thread pool -- million-row databases in seconds; the result does not depend on the
    thread count.  Not byte-identical to make_descriptors (different streams).  rows = (lo, hi):
    only rows [lo, hi) of the n-row database (a shard; identical to the slice of the whole)."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    run = max(1, dup_run)
    block = MT_BLOCK // run * run
    lo, hi = (0, n) if rows is None else rows
    out = np.empty((hi - lo, dim), np.float32)
    scale = np.float32(1.0 / np.sqrt(dim))

    def fill(b):
        s, e = b * block, min(n, (b + 1) * block)
        rng = np.random.default_rng([seed, b])
        rows = e - s            # noqa: F841 (shadows the parameter on purpose: rows of this block)
        heads = -(-rows // run)
        base = rng.standard_normal((heads, dim), dtype=np.float32) * scale
        if run == 1:
            full = base
        else:
            steps = rng.standard_normal((heads, run - 1, dim), dtype=np.float32) * np.float32(dup_sigma)
            full = np.empty((heads, run, dim), np.float32)
            full[:, 0] = base
            full[:, 1:] = base[:, None, :] + np.cumsum(steps, axis=1, dtype=np.float32)
            full = full.reshape(heads * run, dim)[:rows]
        a, z = max(s, lo), min(e, hi)            # the part of this block inside [lo, hi)
        out[a - lo:z - lo] = full[a - s:z - s]

    blocks = range(lo // block, -(-hi // block))
    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(fill, blocks))
    return out


def make_queries(db: np.ndarray, nq: int, seed: int = 5678, sigma: float | None = None) -> np.ndarray:
    """sigma None: independent draws (set A).  sigma > 0: perturbed copies of
    random DB rows (set B), the near-tie stress case."""
    rng = np.random.default_rng(seed)
    dim = db.shape[1]
    if sigma is None:
        return (rng.standard_normal((nq, dim)) * (1.0 / np.sqrt(dim))).astype(np.float32)
    rows = rng.integers(0, db.shape[0], nq)
    noise = (np.random.default_rng(91011).standard_normal((nq, dim)) * sigma).astype(np.float32)
    return (db[rows] + noise).astype(np.float32)


def make_bev_grid(nx: int = 800, ny: int = 800, seed: int = 2222, n_segments: int = 60,
                  n_blobs: int = 40, graded: bool = False) -> np.ndarray:
    """uint8 [ny, nx] occupancy grid: oblique wall segments + small blobs,
    about 1.2 % occupied at the default size (about 4.7k cells at 800x800 is reached
    with the defaults).  graded=True fills occupied cells with values 1..255 instead
    of 255 (general Cartographer semantics)."""
    rng = np.random.default_rng(seed)
    g = np.zeros((ny, nx), np.uint8)
    scale = np.sqrt(nx * ny) / 800.0
    for _ in range(n_segments):
        x0, y0 = rng.uniform(0.08 * nx, 0.92 * nx), rng.uniform(0.08 * ny, 0.92 * ny)
        ang = rng.uniform(0, np.pi)
        length = rng.uniform(20, 110) * scale
        n = int(length * 2) + 2
        t = np.linspace(0, length, n)
        xs = np.clip(np.rint(x0 + t * np.cos(ang)).astype(int), 0, nx - 1)
        ys = np.clip(np.rint(y0 + t * np.sin(ang)).astype(int), 0, ny - 1)
        g[ys, xs] = 255
    for _ in range(n_blobs):
        x0, y0 = rng.integers(10, nx - 10), rng.integers(10, ny - 10)
        r = rng.integers(1, 4)
        yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
        m = (xx * xx + yy * yy) <= r * r
        ys = np.clip(y0 + yy[m], 0, ny - 1)
        xs = np.clip(x0 + xx[m], 0, nx - 1)
        g[ys, xs] = 255
    if graded:
        vals = rng.integers(1, 256, size=g.shape).astype(np.uint8)
        g = np.where(g > 0, vals, 0).astype(np.uint8)
    return g


def centered_limits(nx: int, ny: int, resolution: float = 0.2):
    """MapLimits::max() that puts the world origin at the grid centre.  Cell x
    runs along world -y and cell y along world -x (registration/2d/map_limits.h:69-76),
    hence max_y spans num_x_cells and max_x spans num_y_cells."""
    return 0.5 * ny * resolution, 0.5 * nx * resolution  # (max_x, max_y)


def cells_to_world(cx: np.ndarray, cy: np.ndarray, resolution: float, max_x: float, max_y: float):
    """Centre of cell (cx, cy) in world coordinates: inverse of GetCellIndex."""
    y = max_y - (cx.astype(np.float64) + 0.5) * resolution
    x = max_x - (cy.astype(np.float64) + 0.5) * resolution
    return x, y


def grid_points_world(level1: np.ndarray, resolution: float, max_x: float, max_y: float,
                      threshold: int = 128) -> np.ndarray:
    """World-frame (x, y, 0) float32 points at the centres of occupied cells."""
    cy, cx = np.nonzero(level1 >= threshold)
    x, y = cells_to_world(cx, cy, resolution, max_x, max_y)
    return np.stack([x, y, np.zeros_like(x)], axis=1).astype(np.float32)


def planted_scan(level1: np.ndarray, resolution: float, max_x: float, max_y: float,
                 yaw: float, dx: float, dy: float, dropout: float = 0.2,
                 jitter_cells: float = 0.0, seed: int = 3333) -> np.ndarray:
    """Sensor-frame scan whose true pose in the map is (dx, dy, yaw): the map's
    occupied-cell centres m are mapped to s = R(-yaw) (m - t), thinned by
    `dropout`, optionally jittered by up to +-jitter_cells cells."""
    rng = np.random.default_rng(seed)
    m = grid_points_world(level1, resolution, max_x, max_y).astype(np.float64)
    keep = rng.random(m.shape[0]) >= dropout
    m = m[keep]
    if jitter_cells > 0:
        m[:, :2] += rng.uniform(-jitter_cells, jitter_cells, (m.shape[0], 2)) * resolution
    c, s = np.cos(-yaw), np.sin(-yaw)
    px, py = m[:, 0] - dx, m[:, 1] - dy
    out = np.stack([c * px - s * py, s * px + c * py, np.zeros_like(px)], axis=1)
    return out.astype(np.float32)


def level1_to_cells(level1: np.ndarray) -> np.ndarray:
    """A uint16 Grid2D cell array whose width-1 precomputation grid is `level1`
    for binary grids (255 -> cost value 1 = kMinCorrespondenceCost, 0 -> unknown);
    graded values are mapped through the inverse of ComputeCellValue approximately
    (used only to exercise the uint16 entry point)."""
    lv = level1.astype(np.float64) / 255.0
    v = np.rint((1.0 - lv) * 32766.0).astype(np.int64) + 1
    v = np.clip(v, 1, 32767).astype(np.uint16)
    return np.where(level1 == 0, np.uint16(0), v).astype(np.uint16)


def make_lidar_scan(n_walls: int = 40, seed: int = 4444, max_range: float = 90.0,
                    ground: bool = True) -> np.ndarray:
    """A KITTI-like [n, 4] float32 scan (x, y, z, intensity): vertical wall segments several
    voxels tall (they become occupied BEV pixels: >= 2 hit voxels per column), a flat ground
    disc (one voxel per column: stays free), a few points beyond 100 m (misses) and some
    exactly on voxel boundaries (rounding half away from zero)."""
    rng = np.random.default_rng(seed)
    parts = []
    for _ in range(n_walls):
        x0, y0 = rng.uniform(-max_range * 0.7, max_range * 0.7, 2)
        ang = rng.uniform(0, np.pi)
        length = rng.uniform(3, 25)
        n = int(length * 12)
        t = rng.uniform(0, length, n)
        z = rng.uniform(-1.5, rng.uniform(-1.0, 2.5), n)
        parts.append(np.stack([x0 + t * np.cos(ang), y0 + t * np.sin(ang), z], axis=1))
    if ground:
        r = np.sqrt(rng.uniform(4, (max_range * 0.8) ** 2, 20000))
        a = rng.uniform(0, 2 * np.pi, 20000)
        parts.append(np.stack([r * np.cos(a), r * np.sin(a), np.full(20000, -1.73)], axis=1))
    far = rng.uniform(-1, 1, (200, 3))
    far = far / np.linalg.norm(far, axis=1, keepdims=True) * rng.uniform(99.5, 130, (200, 1))
    parts.append(far)
    # exact half-voxel coordinates (0.1, 0.3, -0.1 ... in float32) and the origin
    k = np.arange(-20, 21)
    parts.append(np.stack([(k + 0.5) * 0.2, (k - 0.5) * 0.2, (k % 3) * 0.2], axis=1))
    parts.append(np.zeros((2, 3)))
    xyz = np.concatenate(parts).astype(np.float32)
    inten = rng.uniform(0, 1, (xyz.shape[0], 1)).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([xyz, inten], axis=1))


# ------------------------------------------------------------------ descriptor network
VGG16_COUT = (64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512)


def hash_uniform(shape, seed):
    """Deterministic pseudo-random float32 values in [-1, 1) from integer arithmetic only (a
    splitmix64 finaliser of the element index): identical on every platform and numpy version,
    so fixtures need not store large weight tensors."""
    n = int(np.prod(shape))
    off = np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) + off
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u24 = (z >> np.uint64(40)).astype(np.float32)          # 24 random bits: exact in float32
    return (u24 / np.float32(1 << 23) - np.float32(1.0)).reshape(shape)


def hashed_vlad_weights(K, C, D, seed):
    """conv_w [K, C] ~ U(-1, 1)/sqrt(C) (the Conv2d default init range, netvlad_fc.py:34),
    centroids [K, C] ~ U(0, 1) (:35), hidden_w [K*C, D] ~ U(-1, 1) sqrt(3/C) (same variance as
    the reference's randn/sqrt(dim), :37-38)."""
    conv_w = hash_uniform((K, C), seed) / np.float32(np.sqrt(C))
    centroids = (hash_uniform((K, C), seed + 1) + np.float32(1.0)) * np.float32(0.5)
    hidden_w = hash_uniform((K * C, D), seed + 2) * np.float32(np.sqrt(3.0 / C))
    return conv_w.astype(np.float32), centroids.astype(np.float32), hidden_w.astype(np.float32)


def hashed_features(B, C, S, seed):
    """A feature map [B, C, S] with conv5_3-like statistics (no ReLU: both signs)."""
    return (hash_uniform((B, C, S), seed) * np.float32(3.0)).astype(np.float32)


def hashed_vgg_weights(seed, gain=1.0):
    """13 (weight [Cout, Cin, 3, 3], bias [Cout]) pairs from the integer-hash generator of
    above (platform independent), He-scaled so that activations keep their magnitude."""
    ws, bs, cin = [], [], 3
    for l, cout in enumerate(VGG16_COUT):
        std = gain * np.sqrt(2.0 / (9 * cin))
        ws.append((hash_uniform((cout, cin, 3, 3), seed + 2 * l) * np.float32(std * np.sqrt(3.0))).astype(np.float32))
        bs.append((hash_uniform((cout,), seed + 2 * l + 1) * np.float32(0.05)).astype(np.float32))
        cin = cout
    return ws, bs
