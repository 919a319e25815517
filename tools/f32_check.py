#!/usr/bin/env python
"""fp32 leaf GEMM: tcgen05 3xTF32 kernel (variant 0) against the generic FMA kernel (variant 1) and fp64 numpy."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import generators as G
H = hb.HierarchicalBlockSparseMatrix
hb.init(0)
bs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 32, 128]
big = len(sys.argv) > 2
worst = 0.0
for b in bs:
    n = b * 16
    lam = 0.02 * 64 / b
    W = min(G.decay_width(lam), n - 1)
    ra, ca, va = G.decay_coo(n, lam, W, 1, dtype=np.float32)
    rb, cb, vb = G.decay_coo(n, lam, W, 2, dtype=np.float32)
    A = H(np.float32, b); A.resize(n, n); A.assign_from_vectors(ra, ca, va); A.update_internal_info()
    B = H(np.float32, b); B.resize(n, n); B.assign_from_vectors(rb, cb, vb); B.update_internal_info()
    Ad = A.to_dense().astype(np.float64); Bd = B.to_dense().astype(np.float64)
    for tA in (0, 1):
        for tB in (0, 1):
            res = {}
            for variant in (0, 1):
                hb.set_gemm_variant(variant)
                C = H(np.float32)
                nm, nr = H.multiply(A, tA, B, tB, C)
                res[variant] = C.to_dense().astype(np.float64)
                kern = hb.stage_times()["gemm_kernel"]
                if variant == 0:
                    assert kern == 3, "tcgen05 kernel was not used (gemm_kernel=%d)" % kern
            hb.set_gemm_variant(0)
            ref = (Ad.T if tA else Ad) @ (Bd.T if tB else Bd)
            e_tc = np.linalg.norm(res[0] - ref) / np.linalg.norm(ref)
            e_fma = np.linalg.norm(res[1] - ref) / np.linalg.norm(ref)
            worst = max(worst, e_tc)
            print("b=%3d tA=%d tB=%d products=%6d  rel err vs fp64: tcgen05 %.2e  fma %.2e" % (b, tA, tB, nm, e_tc, e_fma), flush=True)
if big:
    for b, n, lam in ((64, 16384, 0.01), (128, 16384, 0.01), (32, 8192, 0.02)):
        W = G.decay_width(lam)
        A = H(np.float32, b); A.generate_decay(n, lam, W, 1); A.update_internal_info()
        B = H(np.float32, b); B.generate_decay(n, lam, W, 2); B.update_internal_info()
        for it in range(3):
            C = H(np.float32); nm, nr = H.spamm(A, 0, B, 0, C, 1e-6, True); st = hb.stage_times(); del C
        print(json.dumps({"b": b, "n": n, "products": nm, "gemm_ms": st["gemm_ms"], "tasklist_ms": st["tasklist_ms"],
                          "fp32_equiv_tflops": 2.0 * b ** 3 * nm / st["gemm_ms"] / 1e9, "kernel": st["gemm_kernel"]}), flush=True)
print("worst tcgen05 rel err %.2e" % worst)
sys.exit(0 if worst <= 1e-5 else 1)
