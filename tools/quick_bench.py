"""Quick device-resident timing of one product (not the judged bench; a development probe)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import generators as G
H = hb.HierarchicalBlockSparseMatrix

def run(n, b, lam, tau, reps=5, dtype=np.float64, tA=0, tB=0, exact=False):
    W = G.decay_width(lam)
    t0 = time.time()
    A = H(dtype, b); A.generate_decay(n, lam, W, 1)
    B = H(dtype, b); B.generate_decay(n, lam, W, 2)
    t1 = time.time()
    A.update_internal_info(); B.update_internal_info()
    t2 = time.time()
    best = None
    for i in range(reps):
        C = H(dtype)
        if exact: nm, nr = H.multiply(A, tA, B, tB, C)
        else: nm, nr = H.spamm(A, tA, B, tB, C, tau, True)
        st = hb.stage_times()
        if best is None or st["total_ms"] < best["total_ms"]: best = st
        del C
    fl = 2.0 * b ** 3 * nm
    out = dict(n=n, b=b, lam=lam, tau=tau, dtype=np.dtype(dtype).name, tA=tA, tB=tB, exact=exact, leaves=A.get_n_blocks(),
               gen_s=round(t1 - t0, 3), norms_s=round((t2 - t1) / 2, 4), products=nm, ctiles=nr, **{k: round(v, 4) if isinstance(v, float) else v for k, v in best.items()},
               gemm_tflops=round(fl / best["gemm_ms"] / 1e9, 3) if best["gemm_ms"] > 0 else 0,
               total_tflops=round(fl / best["total_ms"] / 1e9, 3))
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    hb.init(0)
    cfgs = sys.argv[1:] or ["16384,64,0.05,1e-6", "16384,64,0.01,1e-6", "65536,64,0.05,1e-6", "65536,64,0.01,1e-6"]
    for c in cfgs:
        p = c.split(",")
        run(int(p[0]), int(p[1]), float(p[2]), float(p[3]), exact=(len(p) > 4 and p[4] == "exact"),
            dtype=np.float32 if (len(p) > 5 and p[5] == "f32") else np.float64)
