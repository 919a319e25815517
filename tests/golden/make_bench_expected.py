#!/usr/bin/env python
"""Writes tests/golden/bench_expected.json: executed-product checksum, product / C-tile counts and ||C||_F^2 of the
benchmarked SpAMM workloads computed on ONE B200 (bench.py compares every world size against these; the single-GPU
values themselves are checked against the unmodified reference by bench.py's sampled check and by
tests/test_gpu_parity_r2.py).  Run on a GPU box:  python tests/golden/make_bench_expected.py > gpurun_out/bench_expected.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import generators as G

H = hb.HierarchicalBlockSparseMatrix
hb.init(0)
out = {}
for cfg in ("headline", "2", "4"):
    for c in bench.config_cases(cfg):
        ops = bench.build_operands(H, G, c)
        Cm, nm, nr = bench.make_step(H, c, ops)()
        out[bench.expected_key(c)] = {"task_checksum": Cm.task_checksum(), "c_frob_sq": float(Cm.get_frob_squared()),
                                      "products": int(nm), "c_tiles": int(nr)}
        del Cm, ops
print(json.dumps(out, indent=1))
