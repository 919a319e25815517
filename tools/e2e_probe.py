"""Development probe: wall time of hbsm_product_from_host (pipelined host-to-host SpAMM) over repeated calls."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import generators as G
H = hb.HierarchicalBlockSparseMatrix
hb.init(0)
n, b, lam, tau = int(os.environ.get("N", 65536)), 64, 0.01, 1e-6
W = G.decay_width(lam)
A = H(np.float64, b); A.generate_decay(n, lam, W, 1)
B = H(np.float64, b); B.generate_decay(n, lam, W, 2)
def pinned(Mx):
    bi, bj, _, t = Mx.export_leaves(norms=False)
    pt = torch.empty(t.shape, dtype=torch.float64, pin_memory=True); pt.numpy()[...] = t
    return bi.astype(np.int32), bj.astype(np.int32), pt
abi, abj, at = pinned(A); bbi, bbj, bt = pinned(B)
del A, B
out = torch.empty((80000 * n // 65536, b * b), dtype=torch.float64, pin_memory=True)
for slabs in [int(x) for x in (sys.argv[1:] or ["16", "16", "16", "32", "8", "16"])]:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    A2 = H(np.float64, b); A2.resize(n, n); B2 = H(np.float64, b); B2.resize(n, n); C = H(np.float64)
    t1 = time.perf_counter()
    nm, nr, cbi, cbj = H.product_from_host(A2, abi, abj, at.numpy(), 0, B2, bbi, bbj, bt.numpy(), 0, C, True, tau, out.numpy(), slabs)
    t2 = time.perf_counter()
    st = hb.stage_times()
    del A2, B2, C
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(json.dumps({"slabs": slabs, "setup_ms": 1e3 * (t1 - t0), "call_ms": 1e3 * (t2 - t1), "free_ms": 1e3 * (t3 - t2),
                      "device_total_ms": st["total_ms"], "gemm_ms": st["gemm_ms"], "products": nm, "c_tiles": nr}), flush=True)
