"""Shared helpers for the parity tests: build the same matrix in the CUDA engine (through the C ABI) and in the CPU
oracle from one set of COO triplets, and compare results."""
import numpy as np

import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import generators as G
from oracle import pyoracle as po

HBSM = hb.HierarchicalBlockSparseMatrix


def gpu_from_coo(b, m, n, r, c, v, dtype=np.float64, update=True):
    A = HBSM(dtype, b)
    A.resize(m, n)
    A.assign_from_vectors(r, c, v)
    if update:
        A.update_internal_info()
    return A


def both_from_coo(b, m, n, r, c, v, dtype=np.float64, cls=None):
    cls = cls or po.OrcMatrix
    return gpu_from_coo(b, m, n, r, c, v, dtype), po.from_coo(cls, b, m, n, r, c, v, dtype)


def both_from_dense(b, D, dtype=np.float64, cls=None):
    D = np.asarray(D, dtype)
    m, n = D.shape
    r, c = np.meshgrid(np.arange(m), np.arange(n), indexing="ij")
    return both_from_coo(b, m, n, r.ravel(), c.ravel(), D.ravel(), dtype, cls)


def sort_tasks(t):
    t = np.asarray(t, np.int64).reshape(-1, 3)
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]


def rel_frob(x, y):
    d = np.linalg.norm(np.asarray(x, np.float64) - np.asarray(y, np.float64))
    s = np.linalg.norm(np.asarray(y, np.float64))
    return d / s if s > 0 else d


def leaves_equal_structure(g, o):
    gbi, gbj, _, _ = g.export_leaves(tiles=False)
    obi, obj, _, _ = o.leaves(tiles=False)
    return np.array_equal(gbi, obi) and np.array_equal(gbj, obj)


def decay_pair(n, lam, dtype=np.float64, seeds=(1, 2), eps=1e-12):
    W = min(G.decay_width(lam, eps), n - 1)
    return G.decay_coo(n, lam, W, seeds[0], dtype=dtype), G.decay_coo(n, lam, W, seeds[1], dtype=dtype)
