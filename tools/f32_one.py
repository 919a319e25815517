"""One fp32 product at a given leaf size (profiling target): python tools/f32_one.py <b> <n> <lam> [tA tB]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import generators as G
H = hb.HierarchicalBlockSparseMatrix
hb.init(0)
b, n, lam = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
tA = int(sys.argv[4]) if len(sys.argv) > 4 else 0
tB = int(sys.argv[5]) if len(sys.argv) > 5 else 0
W = G.decay_width(lam)
A = H(np.float32, b); A.generate_decay(n, lam, W, 1); A.update_internal_info()
B = H(np.float32, b); B.generate_decay(n, lam, W, 2); B.update_internal_info()
for it in range(3):
    C = H(np.float32); nm, nr = H.spamm(A, tA, B, tB, C, 1e-6, True); st = hb.stage_times(); del C
print("b", b, "tA", tA, "tB", tB, "products", nm, "gemm_ms %.3f" % st["gemm_ms"], "TF %.1f" % (2.0 * b ** 3 * nm / st["gemm_ms"] / 1e9), flush=True)
