"""-m gpu parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for structure, task sets, norms, add/transpose/copy; tolerance for GEMM sums (1e-12 fp64, 1e-5 fp32)."""
import numpy as np
import pytest

import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import generators as G
from oracle import pyoracle as po
from helpers import (HBSM, both_from_coo, both_from_dense, gpu_from_coo, sort_tasks, rel_frob,
                     leaves_equal_structure, decay_pair)

pytestmark = pytest.mark.gpu
TOL = {np.float64: 1e-12, np.float32: 1e-5}


@pytest.fixture(scope="module", autouse=True)
def _built(oracle_built):
    hb.init(0)
    return oracle_built


def test_device_is_b200():
    info = hb.device_info()
    assert info["cc"][0] == 10 and info["sm_count"] > 100


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("b,m,n", [(3, 14, 17), (4, 14, 14), (1, 1, 1), (2, 5, 7), (32, 100, 70), (64, 300, 300)])
def test_assembly_readback(dtype, b, m, n):
    rng = np.random.default_rng(b * 1000 + m)
    nn = max(1, (m * n) // 3)
    r = rng.integers(0, m, nn); c = rng.integers(0, n, nn)
    v = rng.standard_normal(nn).astype(dtype)
    v[::7] = 0.0   # explicit zeros still create tiles but are dropped by get_all_values
    g, o = both_from_coo(b, m, n, r, c, v, dtype)
    assert g.get_n_rows() == m and g.get_n_cols() == n
    assert g.get_depth() == o.depth()
    assert g.get_n_blocks() == o.n_blocks()
    assert leaves_equal_structure(g, o)
    gr, gc, gv = g.get_all_values()
    orr, oc, ov = o.get_all()
    assert np.array_equal(gr, orr) and np.array_equal(gc, oc) and np.array_equal(gv, ov)   # order and bits
    assert g.get_nnz() == len(ov)
    qr = rng.integers(0, m, 50); qc = rng.integers(0, n, 50)
    assert np.array_equal(g.get_values(qr, qc), o.get(qr, qc))
    assert g.get_frob_squared() == o.frob_sq()
    assert g.get_frob_norm_squared_internal() == o.frob_sq_cached()


def test_assign_max_and_errors():
    A = HBSM(np.float64, 4)
    A.resize(10, 10)
    A.assign_from_vectors_max([1, 1, 5], [2, 2, 9], [3.0, 7.0, -2.0])
    assert np.array_equal(A.get_values([1, 5], [2, 9]), [7.0, 0.0])   # max against the initial 0.0 (H:715)
    with pytest.raises(hb.HbsmError, match="index outside matrix boundaries"):
        B = HBSM(np.float64, 4); B.resize(10, 10); B.assign_from_vectors([10], [0], [1.0])
    with pytest.raises(hb.HbsmError, match="non-null child"):
        A.assign_from_vectors([1], [1], [1.0])
    with pytest.raises(hb.HbsmError, match="Matrix must be empty"):
        A.set_params(hb.Params(8))
    E = HBSM(np.float64, 4)
    with pytest.raises(hb.HbsmError, match="empty matrix occured"):
        E.get_frob_squared()
    assert E.empty() and E.get_all_values()[0].size == 0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("b", [32, 64, 128])
def test_leaf_norms_bit_exact(dtype, b):
    n = b * 8
    (r, c, v), _ = decay_pair(n, 0.03, dtype)
    g, o = both_from_coo(b, n, n, r, c, v, dtype)
    _, _, gn, _ = g.export_leaves(tiles=False)
    _, _, on, _ = o.leaves(tiles=False)
    assert np.array_equal(gn, on)
    assert g.get_frob_norm_squared_internal() == o.frob_sq_cached()


SMALL = np.array([[1, 2, 3], [4, 5, 6]], float)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_multiply_small_all_transposes(dtype, tA, tB):
    A = SMALL if not tA else SMALL.T            # op(A) is 2x3
    B = (SMALL.T if not tB else SMALL) + 1.0    # op(B) is 3x2
    g_a, o_a = both_from_dense(2, A, dtype)
    g_b, o_b = both_from_dense(2, B, dtype)
    Cg = HBSM(dtype)
    nm, nr = HBSM.multiply(g_a, tA, g_b, tB, Cg)
    Co, onm, onb, ot = po.OrcMatrix.product(o_a, tA, o_b, tB, want_tasks=True)
    assert (nm, nr) == (onm, onb)
    assert np.array_equal(Cg.export_tasks(), sort_tasks(ot)) or np.array_equal(sort_tasks(Cg.export_tasks()), sort_tasks(ot))
    assert np.array_equal(Cg.to_dense(), Co.to_dense())   # small integers: exact
    assert Cg.get_depth() == Co.depth()
    with pytest.raises(hb.HbsmError, match="non-empty matrix to write result"):
        HBSM.multiply(g_a, tA, g_b, tB, Cg)


def test_multiply_bad_sizes():
    g_a, _ = both_from_dense(2, SMALL)
    C = HBSM(np.float64)
    with pytest.raises(hb.HbsmError, match="matrices have bad sizes"):
        HBSM.multiply(g_a, 0, g_a, 0, C)


@pytest.mark.parametrize("dtype,b,n,lam", [(np.float64, 64, 1024, 0.02), (np.float64, 32, 512, 0.05),
                                           (np.float64, 128, 1024, 0.02), (np.float32, 32, 512, 0.05),
                                           (np.float64, 16, 200, 0.1)])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("tau", [None, 1e-6, 1e-2])
def test_product_task_set_and_values(dtype, b, n, lam, tA, tB, tau):
    (ra, ca, va), (rb, cb, vb) = decay_pair(n, lam, dtype)
    g_a, o_a = both_from_coo(b, n, n, ra, ca, va, dtype)
    g_b, o_b = both_from_coo(b, n, n, rb, cb, vb, dtype)
    Cg = HBSM(dtype)
    if tau is None:
        nm, nr = HBSM.multiply(g_a, tA, g_b, tB, Cg)
        Co, onm, onb, ot = po.OrcMatrix.product(o_a, tA, o_b, tB, want_tasks=True)
    else:
        nm, nr = HBSM.spamm(g_a, tA, g_b, tB, Cg, tau, True)
        Co, onm, onb, ot = po.OrcMatrix.product(o_a, tA, o_b, tB, spamm=True, tau=tau, want_tasks=True)
    assert (nm, nr) == (onm, onb)
    gt = Cg.export_tasks()
    assert np.array_equal(sort_tasks(gt), sort_tasks(ot))          # bit-exact executed-product set
    assert leaves_equal_structure(Cg, Co)
    assert rel_frob(Cg.to_dense(), Co.to_dense()) <= TOL[dtype]
    assert Cg.get_n_block_multiplications() == onm
    st = hb.stage_times()
    assert st["n_products"] == onm and st["gpu_launches"] > 0


@pytest.mark.parametrize("b", [32, 64, 128, 256])
def test_dmma_kernel_matches_generic_kernel(b):
    n = b * (16 if b < 256 else 6)
    (ra, ca, va), (rb, cb, vb) = decay_pair(n, 0.02)
    A = gpu_from_coo(b, n, n, ra, ca, va); B = gpu_from_coo(b, n, n, rb, cb, vb)
    res = []
    for variant in (0, 1):
        hb.set_gemm_variant(variant)
        C = HBSM(np.float64)
        HBSM.multiply(A, 0, B, 1, C)
        res.append(C.to_dense())
    hb.set_gemm_variant(0)
    assert rel_frob(res[0], res[1]) <= 1e-14


@pytest.mark.parametrize("b", [32, 64, 128, 256])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_fp32_tcgen05_kernel(b, tA, tB):
    """fp32 leaf GEMM on the tensor cores (3xTF32, gemm_kernel == 3) against the plain FMA kernel and fp64 numpy:
    relative Frobenius error <= 1e-5 (the fp32 parity bar of BASELINE.json), for every operand orientation."""
    n = b * (8 if b < 256 else 5)
    lam = 0.02 * 64 / b
    (ra, ca, va), (rb, cb, vb) = decay_pair(n, lam, np.float32)
    A = gpu_from_coo(b, n, n, ra, ca, va, np.float32); B = gpu_from_coo(b, n, n, rb, cb, vb, np.float32)
    res = []
    for variant in (0, 1):
        hb.set_gemm_variant(variant)
        C = HBSM(np.float32)
        try:
            HBSM.spamm(A, tA, B, tB, C, 1e-7, True)
        finally:
            hb.set_gemm_variant(0)
        res.append(C.to_dense().astype(np.float64))
        if variant == 0:
            assert hb.stage_times()["gemm_kernel"] == 3, "the tcgen05 kernel did not run"
    Ad = A.to_dense().astype(np.float64); Bd = B.to_dense().astype(np.float64)
    assert rel_frob(res[0], res[1]) <= 1e-5
    # tau = 1e-7 prunes nothing visible at this size: compare with the dense fp64 product as well
    assert rel_frob(res[0], (Ad.T if tA else Ad) @ (Bd.T if tB else Bd)) <= 1e-5


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_structure_ops_exact(dtype):
    n, b = 300, 32
    (ra, ca, va), (rb, cb, vb) = decay_pair(n, 0.05, dtype)
    keep = (ra // b + ca // b) % 3 != 0          # make the two tile sets differ
    g_a, o_a = both_from_coo(b, n, n, ra[keep], ca[keep], va[keep], dtype)
    g_b, o_b = both_from_coo(b, n, n, rb, cb, vb, dtype)
    C = HBSM(dtype); HBSM.add(g_a, g_b, C)
    Co = po.OrcMatrix.add(o_a, o_b)
    assert leaves_equal_structure(C, Co) and np.array_equal(C.to_dense(), Co.to_dense())
    T = HBSM(dtype); HBSM.transpose(g_a, T)
    To = po.OrcMatrix.transpose(o_a)
    assert leaves_equal_structure(T, To) and np.array_equal(T.to_dense(), To.to_dense())
    U = HBSM(dtype); g_b.get_upper_triangle(U)
    Uo = po.OrcMatrix.upper(o_b)
    assert leaves_equal_structure(U, Uo) and np.array_equal(U.to_dense(), Uo.to_dense())
    R = HBSM(dtype); R.rescale(g_a, -0.37)
    Ro = po.OrcMatrix.rescale(o_a, -0.37)
    assert leaves_equal_structure(R, Ro) and np.array_equal(R.to_dense(), Ro.to_dense())
    K = HBSM(dtype); K.copy(g_a)
    assert np.array_equal(K.to_dense(), g_a.to_dense()) and K.get_n_blocks() == g_a.get_n_blocks()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("b,n", [(2, 7), (32, 200), (64, 512)])
def test_symmetric_family(dtype, b, n):
    W = min(G.decay_width(0.05), n - 1)
    r, c, v = G.decay_coo(n, 0.05, W, 3, symmetric=True, dtype=dtype)
    up = r <= c
    g_u, o_u = both_from_coo(b, n, n, r[up], c[up], v[up], dtype)       # upper storage
    (rb, cb, vb), _ = decay_pair(n, 0.05, dtype, seeds=(5, 6))
    g_b, o_b = both_from_coo(b, n, n, rb, cb, vb, dtype)
    full = np.zeros((n, n)); full[r, c] = v
    tol = TOL[dtype] * 10
    C = HBSM(dtype); HBSM.symm_square(g_u, C)
    Co = po.OrcMatrix.symm_square(o_u)
    assert leaves_equal_structure(C, Co)
    assert rel_frob(C.to_dense(), Co.to_dense()) <= tol
    assert rel_frob(C.to_dense(), np.triu(full @ full)) <= tol * 10
    C1 = HBSM(dtype); HBSM.symm_multiply(g_u, True, g_b, False, C1)
    assert rel_frob(C1.to_dense(), po.OrcMatrix.symm_multiply(o_u, 1, o_b, 0).to_dense()) <= tol
    C2 = HBSM(dtype); HBSM.symm_multiply(g_b, False, g_u, True, C2)
    assert rel_frob(C2.to_dense(), po.OrcMatrix.symm_multiply(o_b, 0, o_u, 1).to_dense()) <= tol
    for tr in (False, True):
        C3 = HBSM(dtype); HBSM.symm_rk(g_b, tr, C3)
        assert rel_frob(C3.to_dense(), po.OrcMatrix.symm_rk(o_b, int(tr)).to_dense()) <= tol
    with pytest.raises(hb.HbsmError, match="one and only one"):
        HBSM.symm_multiply(g_u, True, g_b, True, HBSM(dtype))


def test_device_generator_matches_numpy():
    n, b, lam = 1000, 64, 0.05
    W = G.decay_width(lam)
    for sym in (False, True):
        A = HBSM(np.float64, b)
        A.generate_decay(n, lam, W, 11, symmetric=sym)
        r, c, v = G.decay_coo(n, lam, W, 11, symmetric=sym)
        D = np.zeros((n, n)); D[r, c] = v
        assert np.array_equal(A.to_dense(), D)


def test_full_size_properties_cfg2():
    """BASELINE config 2 at full size (N=16384, b=64, tau=1e-6): size-independent checks -- linearity in alpha,
    transposition identity (A B)^T = B^T A^T on the executed set, and SpAMM(tau=0) == multiply on non-zero tiles."""
    n, b, lam, tau = 16384, 64, 0.05, 1e-6
    W = G.decay_width(lam)
    A = HBSM(np.float64, b); A.generate_decay(n, lam, W, 1); A.update_internal_info()
    B = HBSM(np.float64, b); B.generate_decay(n, lam, W, 2); B.update_internal_info()
    C = HBSM(np.float64); nm, nr = HBSM.spamm(A, 0, B, 0, C, tau, True)
    assert nm > 0 and nr == C.get_n_blocks()
    t = C.export_tasks()
    # flat rule recomputed on the host from the exported leaf norms (the light oracle of SURVEY 8c)
    abi, abj, an, _ = A.export_leaves(tiles=False)
    bbi, bbj, bn, _ = B.export_leaves(tiles=False)
    g = n // b
    NA = np.zeros((g, g)); NA[abi, abj] = an
    NB = np.zeros((g, g)); NB[bbi, bbj] = bn
    EA = np.zeros((g, g), bool); EA[abi, abj] = True
    EB = np.zeros((g, g), bool); EB[bbi, bbj] = True
    tau2 = np.float64(tau) * np.float64(tau)
    exp = []
    for k in range(g):
        ii = np.nonzero(EA[:, k])[0]; jj = np.nonzero(EB[k, :])[0]
        keep = (NA[ii, k][:, None] * NB[k, jj][None, :]) > tau2
        a, bcol = np.nonzero(keep)
        exp.append(np.stack([ii[a], jj[bcol], np.full(len(a), k)], 1))
    exp = np.concatenate(exp)
    assert np.array_equal(sort_tasks(t), sort_tasks(exp))
    # (A B)^T == B^T A^T : same products, transposed tiles
    Ct = HBSM(np.float64); HBSM.spamm(B, 1, A, 1, Ct, tau, True)
    T = HBSM(np.float64); HBSM.transpose(C, T)
    _, _, _, t1 = Ct.export_leaves(norms=False)
    _, _, _, t2 = T.export_leaves(norms=False)
    assert t1.shape == t2.shape and rel_frob(t1, t2) <= 1e-13
    # linearity: spamm(2A, B, 2 tau) has the same executed set and twice the values
    A2 = HBSM(np.float64); A2.rescale(A, 2.0); A2.update_internal_info()
    C2 = HBSM(np.float64); nm2, _ = HBSM.spamm(A2, 0, B, 0, C2, 2 * tau, True)
    assert nm2 == nm
    _, _, _, c1 = C.export_leaves(norms=False)
    _, _, _, c2 = C2.export_leaves(norms=False)
    assert np.array_equal(2.0 * c1, c2)


def test_product_to_host_streams_the_same_tiles():
    """hbsm_product_to_host: chunked leaf GEMMs + D2H on a second stream deliver exactly the tiles of C."""
    import ctypes as C
    import torch
    from hierarchical_block_sparse_lib_b200 import _capi
    n, b, lam, tau = 8192, 64, 0.02, 1e-6
    W = G.decay_width(lam)
    A = HBSM(np.float64, b); A.generate_decay(n, lam, W, 1); A.update_internal_info()
    B = HBSM(np.float64, b); B.generate_decay(n, lam, W, 2); B.update_internal_info()
    Cref = HBSM(np.float64); nm0, nr0 = HBSM.spamm(A, 0, B, 1, Cref, tau, True)
    _, _, _, want = Cref.export_leaves(norms=False)
    host = torch.zeros((nr0 + 5, b * b), dtype=torch.float64, pin_memory=True)
    Cs = HBSM(np.float64)
    nm = C.c_size_t(0); nr = C.c_size_t(0)
    _capi.check(_capi.lib().hbsm_product_to_host(A._h, 0, B._h, 1, Cs._h, 1, tau, 1, C.c_void_p(host.data_ptr()), host.shape[0],
                                                  C.byref(nm), C.byref(nr)))
    assert (nm.value, nr.value) == (nm0, nr0)
    assert np.array_equal(host.numpy()[:nr0], want)
    assert np.array_equal(Cs.export_leaves(norms=False)[3], want)
    with pytest.raises(hb.HbsmError):       # too small a buffer: C is still completed, the call reports it
        C2 = HBSM(np.float64)
        _capi.check(_capi.lib().hbsm_product_to_host(A._h, 0, B._h, 1, C2._h, 1, tau, 1, C.c_void_p(host.data_ptr()), 3,
                                                      C.byref(nm), C.byref(nr)))


@pytest.mark.parametrize("dtype,b,n,lam", [(np.float64, 64, 8192, 0.02), (np.float32, 32, 2048, 0.05), (np.float64, 32, 1000, 0.05)])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("n_slabs,shuffled", [(8, False), (4, True), (1, False), (0, False)])
def test_product_from_host_pipeline(dtype, b, n, lam, tA, tB, n_slabs, shuffled):
    """hbsm_product_from_host (slab-pipelined upload / norms / task list / GEMM / download) leaves A, B, C exactly as
    assign_tiles + update_norms + spamm do, and delivers C's tiles with their coordinates to the host buffer."""
    tau = 1e-6
    W = G.decay_width(lam)
    A = HBSM(dtype, b); A.generate_decay(n, lam, W, 1); A.update_internal_info()
    B = HBSM(dtype, b); B.generate_decay(n, lam, W, 2); B.update_internal_info()
    Cref = HBSM(dtype); nm0, nr0 = HBSM.spamm(A, tA, B, tB, Cref, tau, True)
    ci0, cj0, _, want = Cref.export_leaves(norms=False)
    abi, abj, an, at = A.export_leaves()
    bbi, bbj, bn, bt = B.export_leaves()
    if shuffled:   # unordered host tiles: more (shorter) upload runs, or the bulk fallback
        rng = np.random.default_rng(5)
        pa = rng.permutation(len(abi)); pb = rng.permutation(len(bbi))
        abi, abj, at = abi[pa], abj[pa], at[pa]
        bbi, bbj, bt = bbi[pb], bbj[pb], bt[pb]
    A2 = HBSM(dtype, b); A2.resize(n, n)
    B2 = HBSM(dtype, b); B2.resize(n, n)
    C2 = HBSM(dtype)
    out = np.zeros((nr0 + 3, b * b), dtype)
    nm, nr, cbi, cbj = HBSM.product_from_host(A2, abi, abj, at, tA, B2, bbi, bbj, bt, tB, C2, True, tau, out, n_slabs)
    assert (nm, nr) == (nm0, nr0)
    # operands: same table, same (bit-exact) leaf norms and root norm as the two-call path
    for X, Y in ((A, A2), (B, B2)):
        xi, xj, xn, xt = X.export_leaves(); yi, yj, yn, yt = Y.export_leaves()
        assert np.array_equal(xi, yi) and np.array_equal(xj, yj) and np.array_equal(xn, yn) and np.array_equal(xt, yt)
        assert X.get_frob_norm_squared_internal() == Y.get_frob_norm_squared_internal()
    # C on the device: identical table and bitwise identical tiles (same products in the same k order per C tile)
    ci2, cj2, _, got = C2.export_leaves(norms=False)
    assert np.array_equal(ci0, ci2) and np.array_equal(cj0, cj2) and np.array_equal(want, got)
    # C on the host: the same tiles, slab-major, labelled by (c_bi, c_bj)
    order = np.lexsort((cbj, cbi))
    ref_order = np.lexsort((cj0, ci0))
    assert np.array_equal(cbi[order], ci0[ref_order]) and np.array_equal(cbj[order], cj0[ref_order])
    assert np.array_equal(out[:nr][order], want[ref_order])
    assert C2.get_n_block_multiplications() == nm0


def test_product_from_host_errors():
    b, n = 32, 256
    A = HBSM(np.float64, b); A.generate_decay(n, 0.1, 100, 1)
    bi, bj, _, t = A.export_leaves(norms=False)
    A2 = HBSM(np.float64, b); A2.resize(n, n); B2 = HBSM(np.float64, b); B2.resize(n, n); C2 = HBSM(np.float64)
    out = np.zeros((2, b * b))
    with pytest.raises(hb.HbsmError):      # host buffer too small: reported after C is complete
        HBSM.product_from_host(A2, bi, bj, t, 0, B2, bi, bj, t, 0, C2, False, 0.0, out, 2)
    assert C2.get_n_blocks() > 2 and A2.get_n_blocks() == len(bi)
    A3 = HBSM(np.float64, b); A3.resize(n, n); B3 = HBSM(np.float64, b); B3.resize(n, n)
    with pytest.raises(hb.HbsmError):      # C must be empty (H:5681)
        HBSM.product_from_host(A3, bi, bj, t, 0, B3, bi, bj, t, 0, C2, False, 0.0, None, 2)
    bad = bi.copy(); bad[0] = n // b + 5
    with pytest.raises(hb.HbsmError):      # coordinates outside the matrix
        HBSM.product_from_host(A3, bad, bj, t, 0, B3, bi, bj, t, 0, HBSM(np.float64), False, 0.0, None, 2)
    with pytest.raises(hb.HbsmError):      # operands already populated
        HBSM.product_from_host(A2, bi, bj, t, 0, B2, bi, bj, t, 0, HBSM(np.float64), False, 0.0, None, 2)
