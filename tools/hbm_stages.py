#!/usr/bin/env python
"""The HBM-bound stages (leaf norms, add, transpose, task list) on the bench workload, timed with CUDA events on the
engine stream; algorithmic bytes per SURVEY 8(d).  Run plain for the timings, under ncu for the per-kernel DRAM bytes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import generators as G, _capi
H = hb.HierarchicalBlockSparseMatrix
hb.init(0)
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
stream = torch.cuda.ExternalStream(_capi.lib().hbsm_stream())


def timed(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream); r = fn(); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1)); del r
    return best


for dtype, s in ((np.float64, 8), (np.float32, 4)):
    n, b, lam, tau = 65536, 64, 0.01, 1e-6
    W = G.decay_width(lam)
    A = H(dtype, b); A.generate_decay(n, lam, W, 1); A.update_internal_info()
    B = H(dtype, b); B.generate_decay(n, lam, W, 2); B.update_internal_info()
    L = A.get_n_blocks(); tb = b * b * s
    ms = timed(lambda: A.update_internal_info())
    print(json.dumps({"stage": "update_internal_info (k_leaf_norms + level reduce)", "dtype": np.dtype(dtype).name, "tiles": L, "ms": ms,
                      "algorithmic_bytes": L * tb + L * s, "GBps": (L * tb + L * s) / ms / 1e6, "frac_of_measured_hbm": (L * tb + L * s) / ms / 1e6 / PEAK}), flush=True)
    def add():
        C = H(dtype); H.add(A, B, C); return C
    ms = timed(add)
    print(json.dumps({"stage": "add (union + k_add_tiles)", "dtype": np.dtype(dtype).name, "tiles": L, "ms": ms, "algorithmic_bytes": 3 * L * tb,
                      "GBps": 3 * L * tb / ms / 1e6, "frac_of_measured_hbm": 3 * L * tb / ms / 1e6 / PEAK}), flush=True)
    def tr():
        C = H(dtype); H.transpose(A, C); return C
    ms = timed(tr)
    print(json.dumps({"stage": "transpose (key sort + k_transpose_tiles)", "dtype": np.dtype(dtype).name, "tiles": L, "ms": ms, "algorithmic_bytes": 2 * L * tb,
                      "GBps": 2 * L * tb / ms / 1e6, "frac_of_measured_hbm": 2 * L * tb / ms / 1e6 / PEAK}), flush=True)
    C = H(dtype); nm, nr = H.spamm(A, 0, B, 0, C, tau, True); st = hb.stage_times()
    Q, P = st["n_candidates"], st["n_products"]
    tl_bytes = Q * (2 * s + 8) + P * 16 + 2 * P * 12 * 5
    print(json.dumps({"stage": "task list (k_join count+fill, radix sort, finish)", "dtype": np.dtype(dtype).name, "candidates": Q, "products": P,
                      "ms": st["tasklist_ms"], "algorithmic_bytes": tl_bytes, "GBps": tl_bytes / st["tasklist_ms"] / 1e6,
                      "frac_of_measured_hbm": tl_bytes / st["tasklist_ms"] / 1e6 / PEAK, "note": "latency/launch-bound: 30+ small launches"}), flush=True)
    del A, B, C
