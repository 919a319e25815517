"""Seeded synthetic inputs of SURVEY.md 8(d), shared by tests and bench.py.  Pure numpy; the device generator
(hbsm_generate_decay) produces bit-identical values from the same hash and the same exp(-lambda d) table."""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = np.asarray(x, np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def hash_u01(seed, i, j):
    """u(seed,i,j) in [0,1): same arithmetic as csrc/matrix.cu hash_u01."""
    i = np.asarray(i, np.uint64); j = np.asarray(j, np.uint64)
    with np.errstate(over="ignore"):
        h = splitmix64(np.uint64(seed) ^ splitmix64(i * np.uint64(0x100000001B3) + splitmix64(j)))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def decay_width(lam, eps=1e-12):
    return int(np.floor(np.log(1.0 / eps) / lam))


def decay_coo(n, lam, W, seed, symmetric=False, dtype=np.float64, rows=None):
    """Banded decay a_ij = (0.5 + 0.5 u) exp(-lam |i-j|), |i-j| <= W as COO triplets (optionally only `rows`)."""
    table = np.exp(-float(lam) * np.arange(W + 1, dtype=np.float64))
    rr = np.arange(n, dtype=np.int64) if rows is None else np.asarray(rows, np.int64)
    off = np.arange(-W, W + 1, dtype=np.int64)
    r = np.repeat(rr, len(off))
    c = r + np.tile(off, len(rr))
    keep = (c >= 0) & (c < n)
    r = r[keep]; c = c[keep]
    d = np.abs(r - c)
    if symmetric:
        u = hash_u01(seed, np.minimum(r, c), np.maximum(r, c))
    else:
        u = hash_u01(seed, r, c)
    v = (0.5 + 0.5 * u) * table[d]
    return r.astype(np.int32), c.astype(np.int32), v.astype(dtype)


def decay_coo_block(n, lam, W, seed, r0, r1, c0, c1, symmetric=False, dtype=np.float64):
    """The entries of the same banded decay matrix inside rows [r0,r1) x columns [c0,c1), as COO triplets with GLOBAL indices."""
    table = np.exp(-float(lam) * np.arange(W + 1, dtype=np.float64))
    r0, r1, c0, c1 = max(0, r0), min(n, r1), max(0, c0), min(n, c1)
    if r0 >= r1 or c0 >= c1:
        z = np.zeros(0, np.int32)
        return z, z.copy(), np.zeros(0, dtype)
    r, c = np.meshgrid(np.arange(r0, r1, dtype=np.int64), np.arange(c0, c1, dtype=np.int64), indexing="ij")
    r = r.ravel(); c = c.ravel()
    d = np.abs(r - c)
    keep = d <= W
    r = r[keep]; c = c[keep]; d = d[keep]
    u = hash_u01(seed, np.minimum(r, c), np.maximum(r, c)) if symmetric else hash_u01(seed, r, c)
    v = (0.5 + 0.5 * u) * table[d]
    return r.astype(np.int32), c.astype(np.int32), v.astype(dtype)


def task_hash(ci, cj, k):
    """Per-product hash of the engine's hbsm_task_checksum: splitmix64(ci << 42 | cj << 21 | k); the checksum of a set of
    products is the sum of these modulo 2^64."""
    ci = np.asarray(ci, np.uint64); cj = np.asarray(cj, np.uint64); k = np.asarray(k, np.uint64)
    return splitmix64((ci << np.uint64(42)) | (cj << np.uint64(21)) | k)


def task_checksum(ci, cj, k):
    with np.errstate(over="ignore"):
        return int(np.sum(task_hash(ci, cj, k), dtype=np.uint64))


def decay_tiles(n, b, lam, W, seed, symmetric=False, dtype=np.float64, tile_rows=None):
    """Same matrix as decay_coo but as whole column-major tiles: returns (bi, bj, tiles[n_tiles, b*b])."""
    g = -(-n // b)
    wb = min(g, -(-W // b))
    table = np.exp(-float(lam) * np.arange(W + 1, dtype=np.float64))
    tr = range(g) if tile_rows is None else tile_rows
    bis, bjs, tiles = [], [], []
    li = np.arange(b, dtype=np.int64)
    for bi in tr:
        lo, hi = max(0, bi - wb), min(g - 1, bi + wb)
        for bj in range(lo, hi + 1):
            r = (bi * b + li)[:, None] + np.zeros((1, b), np.int64)
            c = (bj * b + li)[None, :] + np.zeros((b, 1), np.int64)
            d = np.abs(r - c)
            inside = (r < n) & (c < n) & (d <= W)
            if symmetric:
                u = hash_u01(seed, np.minimum(r, c), np.maximum(r, c))
            else:
                u = hash_u01(seed, r, c)
            v = np.where(inside, (0.5 + 0.5 * u) * table[np.minimum(d, W)], 0.0)
            bis.append(bi); bjs.append(bj)
            tiles.append(np.asarray(v, dtype).T.reshape(-1))   # column-major
    return (np.asarray(bis, np.int32), np.asarray(bjs, np.int32),
            np.stack(tiles) if tiles else np.zeros((0, b * b), dtype))


def random_block_sparse_coo(n, b, fill, seed, dtype=np.float64):
    """cfg 1: block (bi,bj) present iff u(seed,bi,bj) < fill; entries uniform [-1,1)."""
    g = -(-n // b)
    bi, bj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    present = hash_u01(seed, bi.ravel(), bj.ravel()) < fill
    bi = bi.ravel()[present]; bj = bj.ravel()[present]
    li = np.arange(b)
    rr, cc = np.meshgrid(li, li, indexing="ij")
    r = (bi[:, None] * b + rr.ravel()[None, :]).ravel()
    c = (bj[:, None] * b + cc.ravel()[None, :]).ravel()
    keep = (r < n) & (c < n)
    r = r[keep]; c = c[keep]
    v = 2.0 * hash_u01(seed + 7919, r, c) - 1.0
    return r.astype(np.int32), c.astype(np.int32), v.astype(dtype)
