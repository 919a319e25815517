"""TEST INFRASTRUCTURE ONLY -- sampled parity of a full-size product against the UNMODIFIED reference.

At the benchmarked sizes (N = 65536, leaf 64) the reference cannot run the whole multiply in seconds: it reserves and
sweeps (N/b+1)^3 hash buckets (reference source/HierarchicalBlockSparseMatrix.h:3968-3970, :7249-7250).  A single C
tile, however, is a complete reference problem of its own:

    C(ci, cj) = A(ci, :) * B(:, cj)        A(ci, :)  = block row ci of A   (b x N, same virtual depth as A)
                                           B(:, cj)  = block column cj of B (N x b)

so this module assembles those two slices with the reference's own assign_from_vectors (H:838), refreshes their norms
(update_internal_info, H:3905) and runs the reference's own spamm()/multiply() (H:3931 / H:2142: its hierarchical
prune H:6649-6651 and its leaf gemm H:7273) on them.  The sub-problem has exactly the leaves (and therefore the leaf
norms) the full problem has for that C tile, so its executed k-list and its tile values are the reference's answers
for C(ci, cj) of the full product ("sample C tiles and recompute them with reference gemm on reference
tiles").  The engine's k-list (hbsm_export_tile_tasks), tile values (hbsm_export_tile) and cached leaf norms are
compared with them: k-lists and norms bit-exact, values within the stated relative Frobenius tolerance.

Only tests/ and bench.py's check leg import this file.
"""
import numpy as np

from hierarchical_block_sparse_lib_b200 import generators as G
from . import pyoracle as po


def checker_class():
    import os
    return po.RefMatrix if os.path.exists(po.REF_SO) else po.OrcMatrix


def reference_tile(cls, n, b, lam, W, seeds, ci, cj, spamm, tau, dtype=np.float64, symmetric=False):
    """The reference's answer for C tile (ci, cj) of C = A*B (decay law of SURVEY 8d, seeds = (seed_A, seed_B)).
    Returns (k_list ascending, tile[b,b] or None if nothing executed, {k: nsq(A_ci,k)}, {k: nsq(B_k,cj)})."""
    dtype = np.dtype(dtype)
    r0, r1 = ci * b, min(n, (ci + 1) * b)
    c0, c1 = cj * b, min(n, (cj + 1) * b)
    ra, ca, va = G.decay_coo_block(n, lam, W, seeds[0], r0, r1, max(0, r0 - W), min(n, r1 + W), symmetric, dtype)
    rb, cb, vb = G.decay_coo_block(n, lam, W, seeds[1], max(0, c0 - W), min(n, c1 + W), c0, c1, symmetric, dtype)
    A = po.from_coo(cls, b, b, n, ra - r0, ca, va, dtype)      # b x n: the tiles of block row ci sit at (0, k)
    B = po.from_coo(cls, b, n, b, rb, cb - c0, vb, dtype)      # n x b: the tiles of block column cj sit at (k, 0)
    Cm, nm, nb, tasks = cls.product(A, 0, B, 0, spamm=spamm, tau=tau, want_tasks=True)
    abi, abj, an, _ = A.leaves(tiles=False)
    bbi, bbj, bn, _ = B.leaves(tiles=False)
    a_norms = {int(k): an[i] for i, k in enumerate(abj)}
    b_norms = {int(k): bn[i] for i, k in enumerate(bbi)}
    ks = np.sort(np.asarray(tasks, np.int64).reshape(-1, 3)[:, 2])
    tile = None
    if nm > 0:
        tile = np.asarray(Cm.to_dense(), dtype)[:b, :b]
    return ks, tile, a_norms, b_norms


def sampled_check(Cg, Ag, Bg, n, b, lam, W, seeds, spamm, tau, dtype=np.float64, n_samples=16, seed=12345,
                  symmetric=False, absent_probes=2):
    """Cg = engine result of op(A)*op(B) with both operands untransposed (Ag, Bg engine operands with refreshed norms).
    Samples n_samples existing C tiles (first, last and seeded random ones) plus a few coordinates just outside C's
    structure, and compares each with the reference's answer for that tile."""
    cls = checker_class()
    dtype = np.dtype(dtype)
    cbi, cbj, _, _ = Cg.export_leaves(tiles=False, norms=False)
    out = {"checker": cls.kind, "sampled_c_tiles": 0, "task_set_equal": True, "rel_err_max": 0.0,
           "leaf_norms_bit_equal": True, "leaf_norms_compared": 0, "absent_tiles_confirmed": 0}
    if len(cbi) == 0:
        return out
    rng = np.random.default_rng(seed)
    picks = {0, len(cbi) - 1}
    while len(picks) < min(n_samples, len(cbi)):
        picks.add(int(rng.integers(0, len(cbi))))
    abi, abj, an, _ = Ag.export_leaves(tiles=False)
    bbi, bbj, bn, _ = Bg.export_leaves(tiles=False)
    g_an = {(int(i), int(j)): an[t] for t, (i, j) in enumerate(zip(abi, abj))} if len(abi) < 4_000_000 else None
    g_bn = {(int(i), int(j)): bn[t] for t, (i, j) in enumerate(zip(bbi, bbj))} if len(bbi) < 4_000_000 else None
    for t in sorted(picks):
        ci, cj = int(cbi[t]), int(cbj[t])
        ks, tile, a_n, b_n = reference_tile(cls, n, b, lam, W, seeds, ci, cj, spamm, tau, dtype, symmetric)
        gk = Cg.tile_tasks(ci, cj)
        if gk is None or not np.array_equal(np.sort(gk), ks):
            out["task_set_equal"] = False
        gt = Cg.get_tile(ci, cj)
        if tile is None or gt is None:
            out["task_set_equal"] = False
        else:
            den = np.linalg.norm(tile.astype(np.float64))
            err = np.linalg.norm(gt.astype(np.float64) - tile.astype(np.float64)) / (den if den > 0 else 1.0)
            out["rel_err_max"] = max(out["rel_err_max"], float(err))
        if g_an is not None:
            for k, v in a_n.items():
                if (ci, k) in g_an:       # this rank may hold only a slab of A
                    out["leaf_norms_compared"] += 1
                    if g_an[(ci, k)] != v:
                        out["leaf_norms_bit_equal"] = False
        if g_bn is not None:
            for k, v in b_n.items():
                if (k, cj) in g_bn:
                    out["leaf_norms_compared"] += 1
                    if g_bn[(k, cj)] != v:
                        out["leaf_norms_bit_equal"] = False
        out["sampled_c_tiles"] += 1
    # coordinates next to C's structure that C does NOT hold: the reference must execute nothing there either
    have = set(zip(cbi.tolist(), cbj.tolist()))
    g = -(-n // b)
    probes = 0
    for t in sorted(picks):
        if probes >= absent_probes:
            break
        ci = int(cbi[t])
        row = [j for (i, j) in have if i == ci] if len(have) < 2_000_000 else []
        if not row:
            continue
        for cj in (max(row) + 1, min(row) - 1):
            if 0 <= cj < g and (ci, cj) not in have and probes < absent_probes:
                ks, tile, _, _ = reference_tile(cls, n, b, lam, W, seeds, ci, cj, spamm, tau, dtype, symmetric)
                if len(ks) != 0 or Cg.tile_tasks(ci, cj) is not None:
                    out["task_set_equal"] = False
                else:
                    out["absent_tiles_confirmed"] += 1
                probes += 1
    return out


def flat_rule_checksum(abi, abj, an, bbi, bbj, bn, spamm, tau, dtype=np.float64, row_lo=0, row_hi=None):
    """Checksum (hbsm_task_checksum convention) and size of the executed set predicted by the flat leaf-pair rule
    fl(nsq(A_ik) * nsq(B_kj)) > fl(tau*tau) from leaf norms, for the C rows [row_lo, row_hi)."""
    dtype = np.dtype(dtype)
    tau_t = dtype.type(tau)
    tau2 = tau_t * tau_t
    abi = np.asarray(abi, np.int64); abj = np.asarray(abj, np.int64)
    bbi = np.asarray(bbi, np.int64); bbj = np.asarray(bbj, np.int64)
    an = np.asarray(an, dtype); bn = np.asarray(bn, dtype)
    if row_hi is not None:
        m = (abi >= row_lo) & (abi < row_hi)
        abi, abj, an = abi[m], abj[m], an[m]
    oa = np.argsort(abj, kind="stable"); ob = np.argsort(bbi, kind="stable")
    ak = abj[oa]; bk = bbi[ob]
    ks = np.intersect1d(ak, bk)
    a_lo = np.searchsorted(ak, ks, "left"); a_hi = np.searchsorted(ak, ks, "right")
    b_lo = np.searchsorted(bk, ks, "left"); b_hi = np.searchsorted(bk, ks, "right")
    total = np.uint64(0)
    count = 0
    with np.errstate(over="ignore"):
        for k, al, ah, bl, bh in zip(ks, a_lo, a_hi, b_lo, b_hi):
            ia = oa[al:ah]; ib = ob[bl:bh]
            if spamm:
                keep = (an[ia][:, None] * bn[ib][None, :]) > tau2
            else:
                keep = np.ones((len(ia), len(ib)), bool)
            x, y = np.nonzero(keep)
            if len(x):
                total = total + np.sum(G.task_hash(abi[ia][x], bbj[ib][y], np.full(len(x), k)), dtype=np.uint64)
                count += len(x)
    return int(total), count
