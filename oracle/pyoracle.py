"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the two CPU checkers.

* ``OrcMatrix``  -> oracle/_build/libhbsm_oracle.so (plain-C restatement, hbsm_oracle.c)
* ``RefMatrix``  -> oracle/_ref/libhbsm_ref.so     (the unmodified reference compiled in place)

Both expose the same small Python surface so tests can run one scenario through either.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product
package (hierarchical_block_sparse_lib_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORC_SO = os.path.join(HERE, "_build", "libhbsm_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libhbsm_ref.so")

_libs = {}


def build(ref=True):
    """Compile the checkers (gcc/g++ only).  oracle/_ref is rebuilt only when /root/reference exists."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-f", os.path.join(HERE, "Makefile")] + targets, check=True,
                   env={**os.environ, "CC": "/usr/bin/gcc", "CXX": "/usr/bin/g++"})


def _load(path):
    if path not in _libs:
        if not os.path.exists(path):
            build(ref=(path == REF_SO))
        _libs[path] = C.CDLL(path)
    return _libs[path]


def have_ref():
    return os.path.exists(REF_SO) or os.path.exists("/root/reference/source/HierarchicalBlockSparseMatrix.h")


_SUF = {np.dtype(np.float64): "d", np.dtype(np.float32): "s"}
_CT = {np.dtype(np.float64): C.c_double, np.dtype(np.float32): C.c_float}
_P = C.c_void_p
_L = C.c_long


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


class _Base:
    """Shared Python surface; subclasses bind names to a library."""

    def __init__(self, b, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.b = int(b)
        self.h = self._create(self.b)

    # -- to be provided: _fn(name) -> ctypes function with restype/argtypes loose
    def __del__(self):
        try:
            if getattr(self, "h", None):
                self._destroy()
                self.h = None
        except Exception:
            pass

    def leaves(self, tiles=True):
        n = self.n_blocks()
        bi = np.zeros(n, np.int64); bj = np.zeros(n, np.int64)
        nrm = np.zeros(n, self.dtype)
        t = np.zeros((n, self.b * self.b), self.dtype) if tiles else None
        self._export(bi, bj, nrm, t)
        return bi, bj, nrm, t

    def to_dense(self):
        m, n = self.shape()
        bi, bj, _, t = self.leaves()
        b = self.b
        g = max(1, -(-max(m, n) // b))
        out = np.zeros((g * b + b, g * b + b), self.dtype)
        for i in range(len(bi)):
            out[bi[i] * b:(bi[i] + 1) * b, bj[i] * b:(bj[i] + 1) * b] = t[i].reshape(b, b).T
        return out[:m, :n]


class OrcMatrix(_Base):
    kind = "port"

    def _f(self, name, restype=C.c_int):
        f = getattr(_load(ORC_SO), "orc_%s_%s" % (name, _SUF[self.dtype]))
        f.restype = restype
        return f

    def _create(self, b):
        return C.c_void_p(self._f("create", _P)(C.c_int(b)))

    def _destroy(self):
        self._f("destroy", None)(self.h)

    def resize(self, m, n): self._f("resize", None)(self.h, C.c_int(m), C.c_int(n))
    def clear(self): self._f("clear", None)(self.h)
    def empty(self): return bool(self._f("empty")(self.h))
    def shape(self): return self._f("rows")(self.h), self._f("cols")(self.h)
    def depth(self): return self._f("depth")(self.h)
    def n_blocks(self): return self._f("n_blocks", _L)(self.h)
    def n_mults(self): return self._f("n_mults", _L)(self.h)
    def update(self): self._f("update", None)(self.h)
    def consistent(self): return bool(self._f("consistent")(self.h))
    def nnz(self): return self._f("nnz", _L)(self.h)
    def frob_sq(self): return self.dtype.type(self._f("frob_sq", _CT[self.dtype])(self.h))
    def frob_sq_cached(self): return self.dtype.type(self._f("frob_sq_cached", _CT[self.dtype])(self.h))

    def assign(self, rows, cols, vals, use_max=False):
        r = np.ascontiguousarray(rows, np.int32); c = np.ascontiguousarray(cols, np.int32)
        v = np.ascontiguousarray(vals, self.dtype)
        rc = self._f("assign")(self.h, _L(len(v)), _ptr(r), _ptr(c), _ptr(v), C.c_int(use_max))
        if rc:
            raise RuntimeError("orc_assign rc=%d" % rc)

    def get(self, rows, cols):
        f = self._f("get", _CT[self.dtype])
        return np.array([f(self.h, C.c_int(int(r)), C.c_int(int(c))) for r, c in zip(rows, cols)], self.dtype)

    def get_all(self):
        f = self._f("get_all", _L)
        n = f(self.h, _L(0), None, None, None)
        r = np.zeros(n, np.int32); c = np.zeros(n, np.int32); v = np.zeros(n, self.dtype)
        f(self.h, _L(n), _ptr(r), _ptr(c), _ptr(v))
        return r, c, v

    def _export(self, bi, bj, nrm, t):
        self._f("export_leaves", _L)(self.h, _ptr(bi), _ptr(bj), _ptr(nrm), _ptr(t))

    # ---- operations (static style: result written into a fresh matrix) ----
    @classmethod
    def product(cls, A, tA, B, tB, spamm=False, tau=0.0, want_tasks=False):
        Cm = cls(A.b, A.dtype)
        nm = _L(0); nb = _L(0)
        f = A._f("product")
        cap = 0; ci = cj = kk = None
        if want_tasks:
            rc = f(A.h, C.c_int(tA), B.h, C.c_int(tB), Cm.h, C.c_int(spamm), _CT[A.dtype](tau),
                   C.byref(nm), C.byref(nb), _L(0), None, None, None)
            if rc: raise RuntimeError("orc_product rc=%d" % rc)
            cap = nm.value
            ci = np.zeros(cap, np.int64); cj = np.zeros(cap, np.int64); kk = np.zeros(cap, np.int64)
            Cm = cls(A.b, A.dtype)
        rc = f(A.h, C.c_int(tA), B.h, C.c_int(tB), Cm.h, C.c_int(spamm), _CT[A.dtype](tau),
               C.byref(nm), C.byref(nb), _L(cap), _ptr(ci), _ptr(cj), _ptr(kk))
        if rc: raise RuntimeError("orc_product rc=%d" % rc)
        tasks = np.stack([ci, cj, kk], 1) if want_tasks else None
        return Cm, nm.value, nb.value, tasks

    @classmethod
    def worth(cls, A, tA, B, tB, spamm=False, tau=0.0):
        return bool(A._f("worth")(A.h, C.c_int(tA), B.h, C.c_int(tB), C.c_int(spamm), _CT[A.dtype](tau)))

    @classmethod
    def copy(cls, A):
        Cm = cls(A.b, A.dtype)
        rc = A._f("copy")(Cm.h, A.h)
        if rc: raise RuntimeError("orc_copy rc=%d" % rc)
        return Cm

    @classmethod
    def trunc(cls, A, trunc_value):
        Cm = cls(A.b, A.dtype)
        r = C.c_int(0)
        rc = A._f("trunc")(A.h, Cm.h, _CT[A.dtype](trunc_value), C.byref(r))
        if rc: raise RuntimeError("orc_trunc rc=%d" % rc)
        return Cm, bool(r.value)

    @classmethod
    def _unary(cls, name, A, *extra):
        Cm = cls(A.b, A.dtype)
        rc = A._f(name)(A.h, *extra, Cm.h) if name not in ("rescale",) else None
        if rc: raise RuntimeError("orc_%s rc=%d" % (name, rc))
        return Cm

    @classmethod
    def add(cls, A, B):
        Cm = cls(A.b, A.dtype)
        rc = A._f("add")(A.h, B.h, Cm.h)
        if rc: raise RuntimeError("orc_add rc=%d" % rc)
        return Cm

    @classmethod
    def transpose(cls, A): return cls._unary("transpose", A)
    @classmethod
    def upper(cls, A): return cls._unary("upper", A)
    @classmethod
    def symm_square(cls, A): return cls._unary("symm_square", A)
    @classmethod
    def symm_rk(cls, A, transposed): return cls._unary("symm_rk", A, C.c_int(transposed))

    @classmethod
    def rescale(cls, A, alpha):
        Cm = cls(A.b, A.dtype)
        rc = A._f("rescale")(Cm.h, A.h, _CT[A.dtype](alpha))
        if rc: raise RuntimeError("orc_rescale rc=%d" % rc)
        return Cm

    @classmethod
    def symm_multiply(cls, A, sA, B, sB):
        Cm = cls(A.b, A.dtype)
        rc = A._f("symm_multiply")(A.h, C.c_int(sA), B.h, C.c_int(sB), Cm.h)
        if rc: raise RuntimeError("orc_symm_multiply rc=%d" % rc)
        return Cm


class RefMatrix(_Base):
    kind = "reference"

    def _f(self, name, restype=C.c_int):
        f = getattr(_load(REF_SO), "ref_%s_%s" % (name, _SUF[self.dtype]))
        f.restype = restype
        return f

    @staticmethod
    def last_error():
        f = _load(REF_SO).ref_last_error
        f.restype = C.c_char_p
        return f().decode()

    @staticmethod
    def blas_kind():
        f = _load(REF_SO).ref_blas_kind
        f.restype = C.c_char_p
        return f().decode()

    def _ck(self, rc, what):
        if rc:
            raise RuntimeError("%s: %s" % (what, self.last_error()))

    def _create(self, b):
        return C.c_void_p(self._f("create", _P)(C.c_int(b)))

    def _destroy(self):
        self._f("destroy", None)(self.h)

    def resize(self, m, n): self._ck(self._f("resize")(self.h, C.c_int(m), C.c_int(n)), "resize")
    def clear(self): self._ck(self._f("clear")(self.h), "clear")
    def empty(self): return bool(self._f("empty")(self.h))
    def shape(self): return self._f("n_rows")(self.h), self._f("n_cols")(self.h)
    def depth(self): return self._f("depth")(self.h)
    def consistent(self): return bool(self._f("consistent")(self.h))
    def n_blocks(self): return self._f("n_blocks", _L)(self.h)
    def n_mults(self): return self._f("n_mults", _L)(self.h)
    def update(self): self._ck(self._f("update")(self.h), "update")
    def size_bytes(self): return self._f("size_bytes", _L)(self.h)

    def write_to_buffer(self):
        n = self.size_bytes()
        buf = np.zeros(n, np.uint8)
        self._ck(self._f("write_to_buffer")(self.h, _ptr(buf), _L(n)), "write_to_buffer")
        return buf.tobytes()

    def assign_from_buffer(self, data):
        buf = np.frombuffer(bytes(data), np.uint8)
        self._ck(self._f("assign_from_buffer")(self.h, _ptr(buf), _L(len(buf))), "assign_from_buffer")

    def frob_sq(self):
        out = _CT[self.dtype](0)
        self._ck(self._f("frob_sq")(self.h, C.byref(out)), "frob_sq")
        return self.dtype.type(out.value)

    def frob_sq_cached(self):
        out = _CT[self.dtype](0)
        self._ck(self._f("frob_sq_cached")(self.h, C.byref(out)), "frob_sq_cached")
        return self.dtype.type(out.value)

    def nnz(self):
        out = _L(0)
        self._ck(self._f("nnz")(self.h, C.byref(out)), "nnz")
        return out.value

    def assign(self, rows, cols, vals, use_max=False):
        r = np.ascontiguousarray(rows, np.int32); c = np.ascontiguousarray(cols, np.int32)
        v = np.ascontiguousarray(vals, self.dtype)
        self._ck(self._f("assign")(self.h, _L(len(v)), _ptr(r), _ptr(c), _ptr(v), C.c_int(use_max)), "assign")

    def get(self, rows, cols):
        r = np.ascontiguousarray(rows, np.int32); c = np.ascontiguousarray(cols, np.int32)
        out = np.zeros(len(r), self.dtype)
        self._ck(self._f("get_values")(self.h, _L(len(r)), _ptr(r), _ptr(c), _ptr(out)), "get_values")
        return out

    def get_all(self):
        f = self._f("get_all_values", _L)
        n = f(self.h, _L(0), None, None, None)
        if n < 0: raise RuntimeError(self.last_error())
        r = np.zeros(n, np.int32); c = np.zeros(n, np.int32); v = np.zeros(n, self.dtype)
        f(self.h, _L(n), _ptr(r), _ptr(c), _ptr(v))
        return r, c, v

    def n_leaves(self): return self._f("n_leaves", _L)(self.h)

    def leaves(self, tiles=True):
        n = self.n_leaves()
        bi = np.zeros(n, np.int64); bj = np.zeros(n, np.int64)
        nrm = np.zeros(n, self.dtype)
        t = np.zeros((n, self.b * self.b), self.dtype) if tiles else None
        self._f("export_leaves", _L)(self.h, _ptr(bi), _ptr(bj), _ptr(nrm), _ptr(t))
        return bi, bj, nrm, t

    @classmethod
    def count_skips(cls, A, tA, B, tB, taus, apply_truncation, apply_spamm):
        t = np.ascontiguousarray(taus, A.dtype); out = np.zeros(len(t), np.uint64)
        A._ck(A._f("count_skips")(A.h, C.c_int(tA), B.h, C.c_int(tB), _L(len(t)), _ptr(t), C.c_int(apply_truncation),
                                  C.c_int(apply_spamm), _ptr(out)), "count_skips")
        return out

    @classmethod
    def spamm_errors(cls, A, tA, B, tB, taus):
        t = np.ascontiguousarray(taus, A.dtype); out = np.zeros(len(t), A.dtype)
        n = A._f("spamm_errors", _L)(A.h, C.c_int(tA), B.h, C.c_int(tB), _L(len(t)), _ptr(t), _ptr(out))
        if n < 0: raise RuntimeError(cls.last_error())
        return out[:n]

    @classmethod
    def task_set(cls, A, tA, B, tB, spamm=False, tau=0.0):
        f = A._f("task_set", _L)
        args = (A.h, C.c_int(tA), B.h, C.c_int(tB), C.c_int(spamm), _CT[A.dtype](tau))
        n = f(*args, _L(0), None, None, None)
        if n < 0: raise RuntimeError(cls.last_error())
        ci = np.zeros(n, np.int64); cj = np.zeros(n, np.int64); kk = np.zeros(n, np.int64)
        f(*args, _L(n), _ptr(ci), _ptr(cj), _ptr(kk))
        return np.stack([ci, cj, kk], 1)

    @classmethod
    def product(cls, A, tA, B, tB, spamm=False, tau=0.0, want_tasks=False, timed=False):
        Cm = cls(A.b, A.dtype)
        nm = _L(0); nb = _L(0)
        t3 = (C.c_double * 3)()
        if timed:
            rc = A._f("product_timed")(A.h, C.c_int(tA), B.h, C.c_int(tB), Cm.h, C.c_int(spamm),
                                       _CT[A.dtype](tau), C.byref(nm), C.byref(nb), t3)
        elif spamm:
            rc = A._f("spamm")(A.h, C.c_int(tA), B.h, C.c_int(tB), Cm.h, _CT[A.dtype](tau), C.byref(nm), C.byref(nb))
        else:
            rc = A._f("multiply")(A.h, C.c_int(tA), B.h, C.c_int(tB), Cm.h, C.byref(nm), C.byref(nb))
        A._ck(rc, "product")
        tasks = cls.task_set(A, tA, B, tB, spamm, tau) if want_tasks else None
        if timed:
            return Cm, nm.value, nb.value, tasks, tuple(t3)
        return Cm, nm.value, nb.value, tasks

    @classmethod
    def worth(cls, A, tA, B, tB, spamm=False, tau=0.0):
        if spamm:
            return bool(A._f("worth_to_spamm")(A.h, C.c_int(tA), B.h, C.c_int(tB), _CT[A.dtype](tau)))
        return bool(A._f("worth_to_multiply")(A.h, C.c_int(tA), B.h, C.c_int(tB)))

    @classmethod
    def copy(cls, A):
        Cm = cls(A.b, A.dtype); A._ck(A._f("copy")(Cm.h, A.h), "copy"); return Cm

    @classmethod
    def trunc(cls, A, trunc_value):
        Cm = cls(A.b, A.dtype)
        r = C.c_int(0)
        A._ck(A._f("trunc")(A.h, Cm.h, _CT[A.dtype](trunc_value), C.byref(r)), "frob_block_trunc")
        return Cm, bool(r.value)

    @classmethod
    def add(cls, A, B):
        Cm = cls(A.b, A.dtype); A._ck(A._f("add")(A.h, B.h, Cm.h), "add"); return Cm
    @classmethod
    def transpose(cls, A):
        Cm = cls(A.b, A.dtype); A._ck(A._f("transpose")(A.h, Cm.h), "transpose"); return Cm
    @classmethod
    def upper(cls, A):
        Cm = cls(A.b, A.dtype); A._ck(A._f("upper")(A.h, Cm.h), "upper"); return Cm
    @classmethod
    def rescale(cls, A, alpha):
        Cm = cls(A.b, A.dtype); A._ck(A._f("rescale")(Cm.h, A.h, _CT[A.dtype](alpha)), "rescale"); return Cm
    @classmethod
    def symm_multiply(cls, A, sA, B, sB):
        Cm = cls(A.b, A.dtype); A._ck(A._f("symm_multiply")(A.h, C.c_int(sA), B.h, C.c_int(sB), Cm.h), "symm_multiply"); return Cm
    @classmethod
    def symm_square(cls, A):
        Cm = cls(A.b, A.dtype); A._ck(A._f("symm_square")(A.h, Cm.h), "symm_square"); return Cm
    @classmethod
    def symm_rk(cls, A, transposed):
        Cm = cls(A.b, A.dtype); A._ck(A._f("symm_rk")(A.h, C.c_int(transposed), Cm.h), "symm_rk"); return Cm


def from_coo(cls, b, m, n, rows, cols, vals, dtype=np.float64, update=True):
    A = cls(b, dtype)
    A.resize(m, n)
    A.assign(rows, cols, vals)
    if update:
        A.update()
    return A


def from_dense(cls, b, D, dtype=np.float64, keep_zeros=True, update=True):
    """Assign every entry of a small dense matrix (like the reference tests' set_row, which stores zeros too)."""
    D = np.asarray(D, dtype)
    m, n = D.shape
    r, c = np.meshgrid(np.arange(m), np.arange(n), indexing="ij")
    r = r.ravel(); c = c.ravel(); v = D.ravel()
    if not keep_zeros:
        k = v != 0
        r, c, v = r[k], c[k], v[k]
    return from_coo(cls, b, m, n, r, c, v, dtype, update)
