#!/usr/bin/env python
"""BASELINE.json configs 1-5 on one B200: time per call (CUDA events inside the engine), product counts, leaf TFLOP/s
and size-independent property checks.  One JSON line per case -> stdout (copy into profiles/).
Usage: python tools/run_configs.py [cfg1 cfg2 cfg3 cfg4 cfg5]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import generators as G
H = hb.HierarchicalBlockSparseMatrix
hb.init(0)


def out(**kw):
    print(json.dumps(kw), flush=True)


def best_of(fn, reps=5, warm=2):
    """best of `reps` after `warm` untimed calls whose results are dropped at once (the stream-ordered memory pool grows on
    the first calls, and the engine's GEMM stage timer includes the allocation of C's tiles)"""
    for _ in range(warm):
        r = fn()
        del r
    best = None
    for _ in range(reps):
        r = fn()
        st = hb.stage_times()
        if best is None or st["total_ms"] < best[1]["total_ms"]:
            best = (r, st)
        del r
    return best


def rel(x, y):
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))


def tflops(b, nm, ms):
    return 2.0 * b ** 3 * nm / ms / 1e9 if ms > 0 else 0.0


def cfg1():
    """exact multiply C=A*B, random block-sparse fp64 N=1024 leaf=32 (30% block fill)"""
    n, b = 1024, 32
    ra, ca, va = G.random_block_sparse_coo(n, b, 0.3, 1); rb, cb, vb = G.random_block_sparse_coo(n, b, 0.3, 2)
    A = H(np.float64, b); A.resize(n, n); A.assign_from_vectors(ra, ca, va); A.update_internal_info()
    B = H(np.float64, b); B.resize(n, n); B.assign_from_vectors(rb, cb, vb); B.update_internal_info()
    def run():
        C = H(np.float64); return (C,) + H.multiply(A, 0, B, 0, C)
    (C, nm, nr), st = best_of(run)
    Ad = A.to_dense(); Bd = B.to_dense()
    res = dict(cfg=1, op="multiply NN", dtype="f64", n=n, b=b, products=nm, c_tiles=nr, total_ms=st["total_ms"], gemm_ms=st["gemm_ms"],
               tasklist_ms=st["tasklist_ms"], gemm_tflops=tflops(b, nm, st["gemm_ms"]), rel_err_vs_dense_fp64=rel(C.to_dense(), Ad @ Bd))
    try:
        from oracle import pyoracle as po
        if os.path.exists(po.REF_SO):
            Ar = po.from_coo(po.RefMatrix, b, n, n, ra, ca, va); Br = po.from_coo(po.RefMatrix, b, n, n, rb, cb, vb)
            t0 = time.perf_counter(); Cr, rnm, rnb, rt = po.RefMatrix.product(Ar, 0, Br, 0, want_tasks=True); dt = time.perf_counter() - t0
            t0 = time.perf_counter(); po.RefMatrix.product(Ar, 0, Br, 0); dt = min(dt, time.perf_counter() - t0)
            gt = C.export_tasks()
            key = lambda t: t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]
            res.update(reference_ms=1e3 * dt, reference_products=rnm, task_set_equal=bool(np.array_equal(key(gt), key(rt))),
                       rel_err_vs_reference=rel(C.to_dense(), Cr.to_dense()), reference_threads=int(os.environ.get("OMP_NUM_THREADS", os.cpu_count())))
    except Exception as ex:
        res["reference"] = repr(ex)
    out(**res)


def flat_rule_tasks(A, tA, B, tB, tau, b, n):
    abi, abj, an, _ = A.export_leaves(tiles=False); bbi, bbj, bn, _ = B.export_leaves(tiles=False)
    if tA: abi, abj = abj, abi
    if tB: bbi, bbj = bbj, bbi
    g = -(-n // b)
    tau2 = an.dtype.type(tau) * an.dtype.type(tau)
    order = np.argsort(bbi, kind="stable"); bbi = bbi[order]; bbj = bbj[order]; bn = bn[order]
    start = np.searchsorted(bbi, np.arange(g + 1))
    cnt = 0
    for i, k, na in zip(abi, abj, an):
        s, e = start[k], start[k + 1]
        cnt += int(np.count_nonzero(na * bn[s:e] > tau2))
    return cnt


def cfg2():
    """exponential-decay fp64 N=16384 leaf=64 SpAMM tau=1e-6"""
    n, b, tau = 16384, 64, 1e-6
    for lam in (0.05, 0.01):
        W = G.decay_width(lam)
        A = H(np.float64, b); A.generate_decay(n, lam, W, 1); A.update_internal_info()
        B = H(np.float64, b); B.generate_decay(n, lam, W, 2); B.update_internal_info()
        def run():
            C = H(np.float64); return (C,) + H.spamm(A, 0, B, 0, C, tau, True)
        (C, nm, nr), st = best_of(run)
        out(cfg=2, op="spamm NN", dtype="f64", n=n, b=b, lam=lam, tau=tau, products=nm, candidates=st["n_candidates"], c_tiles=nr,
            total_ms=st["total_ms"], gemm_ms=st["gemm_ms"], tasklist_ms=st["tasklist_ms"], gemm_tflops=tflops(b, nm, st["gemm_ms"]),
            total_tflops=tflops(b, nm, st["total_ms"]), flat_rule_product_count=flat_rule_tasks(A, 0, B, 0, tau, b, n))


def cfg3():
    """symm_square of banded symmetric decay fp64 N=65536 leaf=64, tau sweep 1e-4..1e-10 (SpAMM-pruned symmetric square =
    triu(spamm(sym(A), sym(A), tau)), SURVEY 8 note on cfg 3) plus the reference's exact symm_square"""
    n, b, lam = 65536, 64, 0.05
    W = G.decay_width(lam)
    F = H(np.float64, b); F.generate_decay(n, lam, W, 3, symmetric=True); F.update_internal_info()
    U = H(np.float64); F.get_upper_triangle(U); U.update_internal_info()
    def run_exact():
        C = H(np.float64); H.symm_square(U, C); return C
    C, st = best_of(run_exact)
    Cf = H(np.float64); nmf, nrf = H.multiply(F, 0, F, 0, Cf)
    Cu = H(np.float64); Cf.get_upper_triangle(Cu)
    _, _, _, t1 = C.export_leaves(norms=False); _, _, _, t2 = Cu.export_leaves(norms=False)
    out(cfg=3, op="symm_square (exact)", dtype="f64", n=n, b=b, lam=lam, products=st["n_products"], c_tiles=st["n_ctiles"],
        total_ms=st["total_ms"], gemm_ms=st["gemm_ms"], gemm_tflops=tflops(b, st["n_products"], st["gemm_ms"]),
        full_multiply_products=nmf, rel_err_vs_triu_of_full_multiply=rel(t1, t2) if t1.shape == t2.shape else None)
    del C, Cf, Cu, t1, t2
    for tau in (1e-4, 1e-6, 1e-8, 1e-10):
        def run():
            C = H(np.float64); return (C,) + H.symm_square_spamm(U, C, tau)
        (C, nm, nr), st = best_of(run)
        Cf = H(np.float64); nmf, nrf = H.spamm(F, 0, F, 0, Cf, tau, True)
        Cu = H(np.float64); Cf.get_upper_triangle(Cu)
        _, _, _, t1 = C.export_leaves(norms=False); _, _, _, t2 = Cu.export_leaves(norms=False)
        out(cfg=3, op="symm_square_spamm", dtype="f64", n=n, b=b, lam=lam, tau=tau, products=nm, c_tiles=nr, total_ms=st["total_ms"],
            gemm_ms=st["gemm_ms"], tasklist_ms=st["tasklist_ms"], gemm_tflops=tflops(b, nm, st["gemm_ms"]), full_spamm_products=nmf,
            structure_equals_triu_of_full_spamm=bool(nr == Cu.get_n_blocks()), rel_err_vs_triu_of_full_spamm=rel(t1, t2) if t1.shape == t2.shape else None)
        del C, Cf, Cu, t1, t2


def cfg4():
    """decay-matrix SpAMM fp64 N=262144 leaf=128, tau=1e-6 (single GPU here; sharded: bench.py --gpus N --n 262144 --leaf 128)"""
    n, b, lam, tau = 262144, 128, 0.01, 1e-6
    W = G.decay_width(lam)
    A = H(np.float64, b); A.generate_decay(n, lam, W, 1); A.update_internal_info()
    B = H(np.float64, b); B.generate_decay(n, lam, W, 2); B.update_internal_info()
    def run():
        C = H(np.float64); r = H.spamm(A, 0, B, 0, C, tau, True); del C; return r
    (nm, nr), st = best_of(run)
    out(cfg=4, op="spamm NN", dtype="f64", n=n, b=b, lam=lam, tau=tau, a_tiles=A.get_n_blocks(), products=nm, candidates=st["n_candidates"],
        c_tiles=nr, total_ms=st["total_ms"], gemm_ms=st["gemm_ms"], tasklist_ms=st["tasklist_ms"], gemm_tflops=tflops(b, nm, st["gemm_ms"]),
        total_tflops=tflops(b, nm, st["total_ms"]))


def cfg5():
    """transposed variants A^T*B and A*B^T plus add, fp32 N=65536, leaf sweep 32/64/128/256"""
    n, lam, tau = 65536, 0.02, 1e-6
    W = G.decay_width(lam)
    for b in (32, 64, 128, 256):
        A = H(np.float32, b); A.generate_decay(n, lam, W, 1); A.update_internal_info()
        B = H(np.float32, b); B.generate_decay(n, lam, W, 2); B.update_internal_info()
        for name, tA, tB in (("A^T*B", 1, 0), ("A*B^T", 0, 1)):
            def run():
                C = H(np.float32); return (C,) + H.spamm(A, tA, B, tB, C, tau, True)
            (C, nm, nr), st = best_of(run)
            # property: op(A) op(B) computed with an explicitly transposed operand gives the same tiles
            X = H(np.float32); H.transpose(A if tA else B, X); X.update_internal_info()
            C2 = H(np.float32)
            nm2, _ = H.spamm(X, 0, B, 0, C2, tau, True) if tA else H.spamm(A, 0, X, 0, C2, tau, True)
            _, _, _, t1 = C.export_leaves(norms=False); _, _, _, t2 = C2.export_leaves(norms=False)
            out(cfg=5, op="spamm " + name, dtype="f32", n=n, b=b, lam=lam, tau=tau, products=nm, c_tiles=nr, total_ms=st["total_ms"],
                gemm_ms=st["gemm_ms"], tasklist_ms=st["tasklist_ms"], gemm_kernel=st["gemm_kernel"], gemm_tflops=tflops(b, nm, st["gemm_ms"]),
                same_products_as_explicit_transpose=bool(nm == nm2), rel_err_vs_explicit_transpose=rel(t1, t2) if t1.shape == t2.shape else None)
            del C, C2, X, t1, t2
        import torch
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        stream = torch.cuda.ExternalStream(hb._capi_stream()) if hasattr(hb, "_capi_stream") else None
        times = []
        for _ in range(3):
            C = H(np.float32); t0 = time.perf_counter(); H.add(A, B, C); times.append(time.perf_counter() - t0)
            nt = C.get_n_blocks(); del C
        bytes_moved = 3.0 * nt * b * b * 4
        out(cfg=5, op="add", dtype="f32", n=n, b=b, tiles=nt, wall_ms=1e3 * min(times), algorithmic_GBps=bytes_moved / min(times) / 1e9)
        del A, B


if __name__ == "__main__":
    want = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    info = hb.device_info()
    out(device=info["name"], sms=info["sm_count"])
    for w in want:
        globals()[w]()
