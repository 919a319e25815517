"""Small-size sweep of every hot kernel (compute-sanitizer target; no torch): leaf GEMMs fp64/fp32 at every tensor-core
leaf size and transposition, generic kernel, norms, add, transpose, symmetric square, COO assembly, the host-to-host
pipeline.  Values are checked against dense numpy so that a sanitizer-clean run is also a correct one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import generators as G
H = hb.HierarchicalBlockSparseMatrix
hb.init(0)
worst = 0.0
for dtype, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
    for b in (32, 64, 128, 256, 5):
        n = b * 6 if b >= 32 else 37
        lam = 4.0 / n
        W = n // 3
        A = H(dtype, b); A.generate_decay(n, lam, W, 1) if b >= 32 else None
        if b < 32:
            r, c, v = G.decay_coo(n, lam, W, 1, dtype=dtype)
            A = H(dtype, b); A.resize(n, n); A.assign_from_vectors(r, c, v)
            r, c, v = G.decay_coo(n, lam, W, 2, dtype=dtype)
            B = H(dtype, b); B.resize(n, n); B.assign_from_vectors(r, c, v)
        else:
            B = H(dtype, b); B.generate_decay(n, lam, W, 2)
        A.update_internal_info(); B.update_internal_info()
        Ad = A.to_dense().astype(np.float64); Bd = B.to_dense().astype(np.float64)
        for tA in (0, 1):
            for tB in (0, 1):
                C = H(dtype); H.multiply(A, tA, B, tB, C)
                ref = (Ad.T if tA else Ad) @ (Bd.T if tB else Bd)
                err = np.linalg.norm(C.to_dense() - ref) / np.linalg.norm(ref)
                worst = max(worst, err / tol)
                assert err <= tol, (dtype, b, tA, tB, err)
                Cs = H(dtype); H.spamm(A, tA, B, tB, Cs, 1e-3, True)
        S = H(dtype); H.add(A, B, S); assert np.array_equal(S.to_dense(), (A.to_dense() + B.to_dense()))
        T = H(dtype); H.transpose(A, T); assert np.array_equal(T.to_dense(), A.to_dense().T)
        F = H(dtype, b)
        if b >= 32:
            F.generate_decay(n, lam, W, 3, symmetric=True)
            U = H(dtype); F.get_upper_triangle(U); U.update_internal_info()
            Q = H(dtype); H.symm_square(U, Q)
            Fd = F.to_dense().astype(np.float64)
            assert np.linalg.norm(Q.to_dense() - np.triu(Fd @ Fd)) / np.linalg.norm(Fd @ Fd) <= tol
            # host-to-host pipeline, 2 slabs
            abi, abj, _, at = A.export_leaves(norms=False); bbi, bbj, _, bt = B.export_leaves(norms=False)
            A2 = H(dtype, b); A2.resize(n, n); B2 = H(dtype, b); B2.resize(n, n); C2 = H(dtype)
            out = np.zeros((64, b * b), dtype)
            nm, nr, cbi, cbj = H.product_from_host(A2, abi, abj, at, 0, B2, bbi, bbj, bt, 1, C2, True, 1e-3, out, 2)
            C3 = H(dtype); H.spamm(A, 0, B, 1, C3, 1e-3, True)
            assert np.array_equal(C2.to_dense(), C3.to_dense())
print("sanity_small ok; worst err/tol %.3f" % worst)
