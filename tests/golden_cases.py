"""Golden-fixture scenarios: seeded inputs -> a dict of named result arrays.  tests/golden/make_golden.py runs them
through the UNMODIFIED reference (oracle/_ref) and commits the results as tests/golden/<id>.npz; the CPU tests check
the oracle port against the fixtures, the GPU tests check the CUDA engine against the same fixtures.
Keys starting with 'x_' are exact (bit-for-bit / integer); keys starting with 't_' are floating-point sums whose
order is implementation-defined (BLAS, bucket order: SURVEY 3.1) and are compared by relative Frobenius norm."""
import os

import numpy as np

from hierarchical_block_sparse_lib_b200 import generators as G

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}

CASES = [
    dict(id="mult_randblock_f64", kind="product", dtype=np.float64, gen="randblock", n=128, b=8, fill=0.3, taus=[None]),
    dict(id="spamm_decay_f64", kind="product", dtype=np.float64, gen="decay", n=128, b=16, lam=0.3, taus=[1e-6, 1e-3, 1e-1]),
    dict(id="spamm_decay_f32", kind="product", dtype=np.float32, gen="decay", n=128, b=16, lam=0.3, taus=[None, 1e-4, 1e-2]),
    dict(id="mult_nonpow2_f64", kind="product", dtype=np.float64, gen="decay", n=100, b=8, lam=0.2, taus=[None, 1e-3]),
    dict(id="mult_rect_f64", kind="rect", dtype=np.float64, b=8, m=70, k=50, n=30),
    dict(id="structure_f64", kind="structure", dtype=np.float64, n=100, b=8, lam=0.2),
    dict(id="structure_f32", kind="structure", dtype=np.float32, n=100, b=8, lam=0.2),
    dict(id="symmetric_f64", kind="symmetric", dtype=np.float64, n=72, b=8, lam=0.15),
    dict(id="assembly_f64", kind="assembly", dtype=np.float64, m=29, n=37, b=4),
    dict(id="estimators_f64", kind="estimators", dtype=np.float64, needs_serialize=True),
    dict(id="estimators_f32", kind="estimators", dtype=np.float32, needs_serialize=True),
    dict(id="wire_format_f64", kind="wire", dtype=np.float64, needs_serialize=True),
    dict(id="wire_format_f32", kind="wire", dtype=np.float32, needs_serialize=True),
]


def _inputs(case, seeds=(1, 2)):
    dt = case["dtype"]
    if case["gen"] == "randblock":
        return [G.random_block_sparse_coo(case["n"], case["b"], case["fill"], s, dt) for s in seeds]
    W = min(G.decay_width(case["lam"]), case["n"] - 1)
    return [G.decay_coo(case["n"], case["lam"], W, s, dtype=dt) for s in seeds]


def _tasks(t):
    t = np.asarray(t, np.int64).reshape(-1, 3)
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]


def run_case(K, case):
    out = {}
    kind = case["kind"]
    if kind == "product":
        n, b = case["n"], case["b"]
        (ra, ca, va), (rb, cb, vb) = _inputs(case)
        A = K.coo(b, n, n, ra, ca, va); B = K.coo(b, n, n, rb, cb, vb)
        bi, bj, nrm, _ = K.leaves(A, False)
        out["x_A_bi"], out["x_A_bj"], out["x_A_norms"] = bi, bj, nrm
        out["x_B_norms"] = K.leaves(B, False)[2]
        for ti, tau in enumerate(case["taus"]):
            for tA, tB in ((0, 0), (0, 1), (1, 0), (1, 1)):
                Cm, nm, nb, t = K.product(A, tA, B, tB, spamm=tau is not None, tau=tau or 0.0, want_tasks=True)
                tag = "tau%d_%d%d" % (ti, tA, tB)
                out["x_tasks_" + tag] = _tasks(t)
                out["x_counts_" + tag] = np.array([nm, nb, K.depth(Cm)], np.int64)
                out["t_C_" + tag] = np.asarray(K.to_dense(Cm))
    elif kind == "rect":
        # integers: products are exact in any summation order.  The reference only supports operands of different
        # depth for the plain NN product (its transposed depth-mismatch branches throw "bad sizes", H:5486-5509), so
        # the four (tA,tB) variants run on an equal-depth rectangular triple and NN additionally on a mismatched one.
        b = case["b"]
        rng = np.random.default_rng(7)

        def ints(m, n):
            D = np.round(rng.standard_normal((m, n)) * 4)
            D[np.abs(D) < 3] = 0
            r, c = np.nonzero(D)
            return K.coo(b, m, n, r, c, D[r, c])

        m, k, n = case["m"], case["k"], case["n"]
        Cm, nm, nb, t = K.product(ints(m, k), 0, ints(k, n), 0, want_tasks=True)
        out["x_C_mismatch_nn"] = np.asarray(K.to_dense(Cm))
        out["x_tasks_mismatch_nn"] = _tasks(t)
        out["x_counts_mismatch_nn"] = np.array([nm, nb, K.depth(Cm)], np.int64)
        A = ints(40, 60); B = ints(60, 50)
        At = K.transpose(A); Bt = K.transpose(B)
        for tag, (X, tX, Y, tY) in {"nn": (A, 0, B, 0), "tn": (At, 1, B, 0), "nt": (A, 0, Bt, 1), "tt": (At, 1, Bt, 1)}.items():
            Cm, nm, nb, t = K.product(X, tX, Y, tY, want_tasks=True)
            out["x_C_" + tag] = np.asarray(K.to_dense(Cm))
            out["x_tasks_" + tag] = _tasks(t)
            out["x_counts_" + tag] = np.array([nm, nb, K.depth(Cm)], np.int64)
    elif kind == "structure":
        n, b, dt = case["n"], case["b"], case["dtype"]
        W = min(G.decay_width(case["lam"]), n - 1)
        ra, ca, va = G.decay_coo(n, case["lam"], W, 1, dtype=dt)
        rb, cb, vb = G.decay_coo(n, case["lam"], W, 2, dtype=dt)
        keep = (ra // b + ca // b) % 3 != 0
        A = K.coo(b, n, n, ra[keep], ca[keep], va[keep]); B = K.coo(b, n, n, rb, cb, vb)
        S = K.add(A, B)
        out["x_add"] = np.asarray(K.to_dense(S)); out["x_add_bi"], out["x_add_bj"] = K.leaves(S, False)[:2]
        T = K.transpose(A)
        out["x_transpose"] = np.asarray(K.to_dense(T)); out["x_tr_bi"], out["x_tr_bj"] = K.leaves(T, False)[:2]
        U = K.upper(B)
        out["x_upper"] = np.asarray(K.to_dense(U)); out["x_up_bi"], out["x_up_bj"] = K.leaves(U, False)[:2]
        out["x_rescale"] = np.asarray(K.to_dense(K.rescale(A, -0.37)))
        out["x_copy"] = np.asarray(K.to_dense(K.copy(A)))
        for ti, tv in enumerate((1e-3, 0.3, 1.0)):          # frob_block_trunc, H:4935 (leaf norms here span 1e-6 .. 3)
            Tm, removed = K.trunc(B, tv)
            out["x_trunc%d" % ti] = np.asarray(K.to_dense(Tm))
            out["x_trunc%d_bi" % ti], out["x_trunc%d_bj" % ti] = K.leaves(Tm, False)[:2]
            out["x_trunc%d_removed" % ti] = np.array([int(removed), K.n_blocks(Tm)], np.int64)
        out["x_frob"] = np.array([K.frob_sq(A), K.frob_sq(B)], dt)
        out["x_counts"] = np.array([K.n_blocks(A), K.n_blocks(B), K.nnz(A), K.nnz(B)], np.int64)
    elif kind == "symmetric":
        n, b, dt = case["n"], case["b"], case["dtype"]
        W = min(G.decay_width(case["lam"]), n - 1)
        r, c, v = G.decay_coo(n, case["lam"], W, 3, symmetric=True, dtype=dt)
        up = r <= c
        U = K.coo(b, n, n, r[up], c[up], v[up])
        rb, cb, vb = G.decay_coo(n, case["lam"], W, 4, dtype=dt)
        B = K.coo(b, n, n, rb, cb, vb)
        Q = K.symm_square(U)
        out["t_symm_square"] = np.asarray(K.to_dense(Q)); out["x_sq_bi"], out["x_sq_bj"] = K.leaves(Q, False)[:2]
        out["t_symm_mul_AB"] = np.asarray(K.to_dense(K.symm_multiply(U, True, B, False)))
        out["t_symm_mul_BA"] = np.asarray(K.to_dense(K.symm_multiply(B, False, U, True)))
        out["t_symm_rk_n"] = np.asarray(K.to_dense(K.symm_rk(B, False)))
        out["t_symm_rk_t"] = np.asarray(K.to_dense(K.symm_rk(B, True)))
    elif kind == "estimators":
        # count_skips H:4945 / get_spamm_errors H:5236 on the reference's own example (TO:634-702: skips 0 1 2 2 2 3 4) and on
        # a decay pair with three levels, all four transpositions, the three skip rules
        dt = case["dtype"]
        from known_answers import SP_A, SP_B
        As = K.dense(2, SP_A); Bs = K.dense(2, SP_B)
        taus = np.array([0.0125, 0.025, 0.05, 0.1, 0.2, 0.4, 0.8], dt)
        out["x_to_skips_spamm"] = np.asarray(K.count_skips(As, 0, Bs, 0, taus, False, True), np.int64)
        out["t_to_spamm_errors"] = np.asarray(K.spamm_errors(As, 0, Bs, 0, taus), np.float64)
        n, b = 96, 8
        W = G.decay_width(0.25)
        A = K.coo(b, n, n, *G.decay_coo(n, 0.25, min(W, n - 1), 1, dtype=dt))
        B = K.coo(b, n, n, *G.decay_coo(n, 0.25, min(W, n - 1), 2, dtype=dt))
        taus2 = np.array([1e-6, 1e-4, 1e-3, 1e-2, 0.1, 1.0, 10.0], dt)
        for tA, tB in ((0, 0), (0, 1), (1, 0), (1, 1)):
            for name, tr, sp in (("spamm", False, True), ("trunc", True, False), ("hybrid", True, True), ("none", False, False)):
                out["x_skips_%s_%d%d" % (name, tA, tB)] = np.asarray(K.count_skips(A, tA, B, tB, taus2, tr, sp), np.int64)
            out["t_spamm_errors_%d%d" % (tA, tB)] = np.asarray(K.spamm_errors(A, tA, B, tB, taus2), np.float64)
    elif kind == "wire":
        # the reference's serialisation (H:1124-1487), byte for byte: TC:131-133 size 528 (fp64), a product result with
        # its multiply counter and stale norms, a refreshed matrix, a single-leaf matrix, sized-but-childless, empty
        dt = case["dtype"]
        M1 = K.coo(4, 14, 14, [0, 6], [0, 7], np.array([7.7, 1.1], dt), update=False)
        out["x_bytes_two_leaves_stale"] = np.frombuffer(K.serialize(M1), np.uint8).copy()
        K.update(M1)
        out["x_bytes_two_leaves_updated"] = np.frombuffer(K.serialize(M1), np.uint8).copy()
        rng = np.random.default_rng(5)
        D = np.round(rng.standard_normal((11, 9)) * 3).astype(dt); D[np.abs(D) < 2] = 0
        r, c = np.nonzero(D)
        A = K.coo(3, 11, 9, r, c, D[r, c])
        P, nm, nb, _ = K.product(A, 1, A, 0)                 # 9x9, integer values: exact in any order
        out["x_bytes_product"] = np.frombuffer(K.serialize(P), np.uint8).copy()
        out["x_product_counts"] = np.array([nm, nb], np.int64)
        out["x_bytes_single_leaf"] = np.frombuffer(K.serialize(K.coo(4, 3, 2, [0, 2], [1, 0], np.array([1.5, -2.0], dt))), np.uint8).copy()
        out["x_bytes_childless"] = np.frombuffer(K.serialize(K.sized(2, 8, 8)), np.uint8).copy()
        # round trip through the reader
        R = K.deserialize(3, out["x_bytes_product"].tobytes())
        out["x_roundtrip_dense"] = np.asarray(K.to_dense(R))
        out["x_roundtrip_meta"] = np.array([K.n_mults(R), K.n_blocks(R), K.depth(R), K.shape(R)[0], K.shape(R)[1]], np.int64)
        R1 = K.deserialize(4, out["x_bytes_two_leaves_updated"].tobytes())
        out["x_roundtrip_norm"] = np.array([R1.frob_sq_cached() if hasattr(R1, "frob_sq_cached") else R1.get_frob_norm_squared_internal()], dt)
    elif kind == "assembly":
        m, n, b = case["m"], case["n"], case["b"]
        rng = np.random.default_rng(11)
        cnt = 400
        r = rng.integers(0, m, cnt); c = rng.integers(0, n, cnt)
        v = rng.standard_normal(cnt)
        v[::9] = 0.0
        A = K.coo(b, m, n, r, c, v)                         # duplicates sum in input order (H:718)
        rr, cc, vv = K.get_all(A)
        out["x_all_r"], out["x_all_c"], out["x_all_v"] = np.asarray(rr, np.int64), np.asarray(cc, np.int64), np.asarray(vv)
        bi, bj, nrm, _ = K.leaves(A, False)
        out["x_bi"], out["x_bj"], out["x_norms"] = bi, bj, nrm
        out["x_counts"] = np.array([K.n_blocks(A), K.nnz(A), K.depth(A)], np.int64)
        out["x_frob"] = np.array([K.frob_sq(A)])
        qr = rng.integers(0, m, 64); qc = rng.integers(0, n, 64)
        out["x_get"] = np.asarray(K.get(A, qr, qc))
    return out


def compare(got, want, case):
    tol = TOL[np.dtype(case["dtype"])]
    assert set(got) == set(want), set(got) ^ set(want)
    for k in sorted(want):
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.shape == w.shape, (k, g.shape, w.shape)
        if k.startswith("x_"):
            assert np.array_equal(g, w), "%s differs (exact key)" % k
        else:
            d = np.linalg.norm(g.astype(np.float64) - w.astype(np.float64))
            s = np.linalg.norm(w.astype(np.float64))
            assert d <= tol * max(s, 1e-300), "%s: rel err %.3e > %.1e" % (k, d / max(s, 1e-300), tol)
