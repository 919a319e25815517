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


# ---------------------------------------------------------------------------------------------------
# backend adapters: one scenario (tests/known_answers.py, the golden fixtures) runs through the CPU oracle port,
# the unmodified reference (oracle/_ref) or the CUDA engine with the same calls
# ---------------------------------------------------------------------------------------------------
class OracleBackend:
    """OrcMatrix (plain-C port) or RefMatrix (unmodified reference compiled in place)."""

    def __init__(self, cls, dtype=np.float64):
        self.cls = cls
        self.dtype = np.dtype(dtype)
        self.name = cls.kind

    def sized(self, b, m, n):
        A = self.cls(b, self.dtype); A.resize(m, n); return A

    def coo(self, b, m, n, r, c, v, update=True):
        return po.from_coo(self.cls, b, m, n, r, c, np.asarray(v, self.dtype), self.dtype, update)

    def dense(self, b, D, update=True):
        return po.from_dense(self.cls, b, D, self.dtype, True, update)

    def to_dense(self, A): return A.to_dense()
    def depth(self, A): return A.depth()
    def shape(self, A): return tuple(A.shape())
    def consistent(self, A): return A.consistent()
    def n_blocks(self, A): return A.n_blocks()
    def n_mults(self, A): return A.n_mults()
    def nnz(self, A): return A.nnz()
    def frob_sq(self, A): return float(A.frob_sq())
    def get(self, A, r, c): return A.get(r, c)
    def get_all(self, A): return A.get_all()
    def leaves(self, A, tiles=True): return A.leaves(tiles)
    def update(self, A): A.update()

    def product(self, A, tA, B, tB, spamm=False, tau=0.0, want_tasks=False):
        return self.cls.product(A, tA, B, tB, spamm=spamm, tau=tau, want_tasks=want_tasks)

    def worth_to_multiply(self, A, tA, B, tB): return self.cls.worth(A, tA, B, tB)
    def worth_to_spamm(self, A, tA, B, tB, tau): return self.cls.worth(A, tA, B, tB, True, tau)
    def add(self, A, B): return self.cls.add(A, B)
    def transpose(self, A): return self.cls.transpose(A)
    def upper(self, A): return self.cls.upper(A)
    def rescale(self, A, alpha): return self.cls.rescale(A, alpha)
    def copy(self, A): return self.cls.copy(A)
    def trunc(self, A, t): return self.cls.trunc(A, t)
    def can_estimate(self): return hasattr(self.cls, "count_skips")
    def count_skips(self, A, tA, B, tB, taus, tr, sp): return self.cls.count_skips(A, tA, B, tB, taus, int(tr), int(sp))
    def spamm_errors(self, A, tA, B, tB, taus): return self.cls.spamm_errors(A, tA, B, tB, taus)
    def can_serialize(self): return hasattr(self.cls, "write_to_buffer")
    def serialize(self, A): return A.write_to_buffer()
    def deserialize(self, b, data): A = self.cls(b, self.dtype); A.assign_from_buffer(data); return A
    def symm_multiply(self, A, sA, B, sB): return self.cls.symm_multiply(A, int(sA), B, int(sB))
    def symm_square(self, A): return self.cls.symm_square(A)
    def symm_rk(self, A, transposed): return self.cls.symm_rk(A, int(transposed))


class GpuBackend:
    """The CUDA engine through the C ABI (Python mirror of the reference class)."""
    name = "cuda"

    def __init__(self, dtype=np.float64):
        self.dtype = np.dtype(dtype)

    def sized(self, b, m, n):
        A = HBSM(self.dtype, b); A.resize(m, n); return A

    def coo(self, b, m, n, r, c, v, update=True):
        return gpu_from_coo(b, m, n, r, c, np.asarray(v, self.dtype), self.dtype, update)

    def dense(self, b, D, update=True):
        D = np.asarray(D, self.dtype)
        m, n = D.shape
        r, c = np.meshgrid(np.arange(m), np.arange(n), indexing="ij")
        return self.coo(b, m, n, r.ravel(), c.ravel(), D.ravel(), update)

    def to_dense(self, A): return A.to_dense()
    def depth(self, A): return A.get_depth()
    def shape(self, A): return (A.get_n_rows(), A.get_n_cols())
    def consistent(self, A): return A.check_if_matrix_is_consistent()
    def n_blocks(self, A): return A.get_n_blocks()
    def n_mults(self, A): return A.get_n_block_multiplications()
    def nnz(self, A): return A.get_nnz()
    def frob_sq(self, A): return float(A.get_frob_squared())
    def get(self, A, r, c): return A.get_values(r, c)
    def get_all(self, A): return A.get_all_values()
    def update(self, A): A.update_internal_info()

    def leaves(self, A, tiles=True):
        return A.export_leaves(tiles=tiles)

    def product(self, A, tA, B, tB, spamm=False, tau=0.0, want_tasks=False):
        Cm = HBSM(self.dtype)
        if spamm:
            nm, nb = HBSM.spamm(A, tA, B, tB, Cm, tau, True)
        else:
            nm, nb = HBSM.multiply(A, tA, B, tB, Cm)
        return Cm, nm, nb, (Cm.export_tasks() if want_tasks else None)

    def worth_to_multiply(self, A, tA, B, tB): return HBSM.worth_to_multiply(A, tA, B, tB)
    def worth_to_spamm(self, A, tA, B, tB, tau): return HBSM.worth_to_spamm(A, tA, B, tB, tau)

    def _out(self, fn, *args):
        Cm = HBSM(self.dtype); fn(*args, Cm); return Cm

    def add(self, A, B): return self._out(HBSM.add, A, B)
    def transpose(self, A): return self._out(HBSM.transpose, A)
    def upper(self, A): Cm = HBSM(self.dtype); A.get_upper_triangle(Cm); return Cm
    def rescale(self, A, alpha): Cm = HBSM(self.dtype); Cm.rescale(A, alpha); return Cm
    def copy(self, A): Cm = HBSM(self.dtype); Cm.copy(A); return Cm
    def trunc(self, A, t): Cm = HBSM(self.dtype); r = A.frob_block_trunc(Cm, t); return Cm, r
    def can_estimate(self): return True
    def count_skips(self, A, tA, B, tB, taus, tr, sp): return HBSM.count_skips(A, tA, B, tB, taus, tr, sp)
    def spamm_errors(self, A, tA, B, tB, taus): return HBSM.get_spamm_errors(A, tA, B, tB, taus)
    def can_serialize(self): return True
    def serialize(self, A): return A.write_to_buffer()
    def deserialize(self, b, data): A = HBSM(self.dtype); A.assign_from_buffer(data); return A
    def symm_multiply(self, A, sA, B, sB): return self._out(HBSM.symm_multiply, A, sA, B, sB)
    def symm_square(self, A): return self._out(HBSM.symm_square, A)
    def symm_rk(self, A, transposed): return self._out(HBSM.symm_rk, A, transposed)
