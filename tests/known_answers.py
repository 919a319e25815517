"""Known-answer scenarios of the reference's own tests, restated as data so that ONE scenario runs through any
backend: the CPU oracle port (OrcMatrix), the unmodified reference (RefMatrix, oracle/_ref) and the CUDA engine
(through the C ABI).  Citations: TO: = test_source/test_matrix_operations.cc, TC: = test_source/test_matrix_creation.cc
of the reference.  Inputs are assigned the way the reference's SparseMatrix::set_row/assign helper does
(test_utils.h:27-43): EVERY entry of every row, zeros included, through assign_from_vectors.
All values are small integers / short decimals, so results are exact (`==`) regardless of BLAS or summation order,
which is how the reference compares them (verify_that_matrices_are_equal, test_utils.h:51-74).
"""
import numpy as np

A_UP = [[2, 2, 3, 5, 2],      # TO:25-31 (upper-triangular storage of a symmetric matrix)
        [0, 1, 2, 4, 3],
        [0, 0, 3, 1, 2],
        [0, 0, 0, 4, 5],
        [0, 0, 0, 0, 1]]
B57 = [[5, 3, 1, 5, 0, 3, 3],  # TO:37-43
       [1, 5, 5, 4, 1, 5, 1],
       [2, 1, 3, 2, 1, 1, 4],
       [2, 2, 3, 3, 1, 4, 2],
       [5, 1, 1, 2, 1, 2, 3]]
C25 = [[5, 3, 0, 3, 3],        # TO:49-53
       [1, 4, 1, 5, 1]]
AxB57 = [[38, 31, 38, 43, 12, 43, 36],   # TO:68-74
         [38, 24, 28, 36, 10, 35, 32],
         [35, 26, 27, 36, 8, 30, 31],
         [64, 49, 45, 65, 14, 62, 46],
         [32, 34, 39, 43, 11, 45, 30]]
CxA25 = [[37, 34, 30, 64, 37],           # TO:90-93
         [40, 31, 21, 47, 42]]
SQ_UP = [[46, 38, 28, 51, 43],           # TO:122-128
         [0, 34, 24, 47, 34],
         [0, 0, 27, 40, 25],
         [0, 0, 0, 83, 49],
         [0, 0, 0, 0, 43]]
A23 = [[2, 3, 5], [0, 1, 2]]             # TO:143-146 and TO:240-244
B23 = [[1, 3, 2], [6, 2, 4]]             # TO:266-269
D32 = [[2, 1], [7, 3], [3, 5]]           # TO:292-296
A55 = [[2, 2, 3, 5, 2],                  # TO:437-443 (full, 2-level at b=3)
       [2, 1, 2, 4, 3],
       [3, 2, 3, 1, 2],
       [5, 4, 1, 4, 5],
       [2, 3, 2, 5, 1]]
SP_A = [[1, 2, 0.1, 0.1],                # TO:641-646
        [2, 1, 0.1, 0.1],
        [3, 1, 0, 0],
        [5, 1, 0, 0.1]]
SP_B = [[1, 6, 5, 0],                    # TO:655-660
        [0, 1, 2, 1],
        [2, 1, 0.1, 0.1],
        [0, 1, 0.1, 0.1]]
SP_C = [[1.2, 8.2, 9, 2],                # TO:675-680 (tau = 0.2)
        [2.2, 13.2, 12, 1],
        [3, 19, 17, 1],
        [5, 31.1, 27, 1]]


def _eq(got, want, what):
    got = np.asarray(got, float); want = np.asarray(want, float)
    assert got.shape == want.shape, "%s: shape %s != %s" % (what, got.shape, want.shape)
    assert np.array_equal(got, want), "%s:\n%s\n!=\n%s" % (what, got, want)


def run_all(K):
    """K is a backend adapter (tests/helpers.py: OracleBackend / GpuBackend).  Returns the number of checks made."""
    n = 0
    # ---- symm_multiply, TO:20-95 (b = 2, TO:194) ----
    A = K.dense(2, A_UP); B = K.dense(2, B57); C = K.dense(2, C25)
    _eq(K.to_dense(K.symm_multiply(A, True, B, False)), AxB57, "symm_multiply sym(A)*B"); n += 1
    _eq(K.to_dense(K.symm_multiply(C, False, A, True)), CxA25, "symm_multiply C*sym(A)"); n += 1
    # ---- symm_square, TO:98-134 ----
    _eq(K.to_dense(K.symm_square(A)), SQ_UP, "symm_square"); n += 1
    # ---- symm_rk, TO:137-185 ----
    R = K.dense(2, A23)
    _eq(K.to_dense(K.symm_rk(R, False)), [[38, 13], [0, 5]], "symm_rk A*A'"); n += 1
    _eq(K.to_dense(K.symm_rk(R, True)), [[4, 6, 10], [0, 10, 17], [0, 0, 29]], "symm_rk A'*A"); n += 1
    # ---- transpose / add / multiply NN NT TN TT, TO:236-384 (b = 2) ----
    A2 = K.dense(2, A23); B2 = K.dense(2, B23); D = K.dense(2, D32)
    _eq(K.to_dense(K.transpose(A2)), np.array(A23).T, "transpose"); n += 1
    _eq(K.to_dense(K.add(A2, B2)), [[3, 6, 7], [6, 3, 6]], "add"); n += 1
    _eq(K.to_dense(K.product(A2, 0, D, 0)[0]), [[40, 36], [13, 13]], "multiply NN A*D"); n += 1
    _eq(K.to_dense(K.product(D, 0, A2, 0)[0]), [[4, 7, 12], [14, 24, 41], [6, 14, 25]], "multiply NN D*A"); n += 1
    _eq(K.to_dense(K.product(A2, 0, A2, 1)[0]), [[38, 13], [13, 5]], "multiply NT"); n += 1
    _eq(K.to_dense(K.product(A2, 1, A2, 0)[0]), [[4, 6, 10], [6, 10, 17], [10, 17, 29]], "multiply TN"); n += 1
    _eq(K.to_dense(K.product(A2, 1, D, 1)[0]), [[4, 14, 6], [7, 24, 14], [12, 41, 25]], "multiply TT"); n += 1
    # ---- depth mismatch 1x1 * 1x2, b = 1, TO:389-426 (the reference only prints; value is 2*[1 2]) ----
    a11 = K.dense(1, [[2]]); b12 = K.dense(1, [[1, 2]])
    assert K.depth(a11) == 0 and K.depth(b12) == 1
    _eq(K.to_dense(K.product(a11, 0, b12, 0)[0]), [[2, 4]], "1x1 * 1x2"); n += 1
    # ---- 2-level x 3-level at b = 3, TO:430-509 ----
    A5 = K.dense(3, A55); B5 = K.dense(3, B57)
    assert K.depth(A5) == 1 and K.depth(B5) == 2
    assert K.worth_to_multiply(A5, 0, B5, 0) and K.worth_to_multiply(B5, 1, A5, 0)   # TO:459, TO:501
    P, nm, nb, _ = K.product(A5, 0, B5, 0)
    _eq(K.to_dense(P), AxB57, "A_2level * B_3level"); n += 1
    assert K.depth(P) == K.depth(K.dense(2, AxB57)) == 2                             # TO:480 (ref built at b = 2)
    P, nm, nb, _ = K.product(B5, 1, A5, 0)
    _eq(K.to_dense(P), np.array(AxB57).T, "B_3level' * A_2level"); n += 1
    assert nm == 12 and K.n_mults(P) == 12                                           # TO:507
    # ---- rescale, TO:513-526 ----
    _eq(K.to_dense(K.rescale(A2, -1.0)), -np.array(A23, float), "rescale"); n += 1
    # ---- SpAMM tau = 0.2, b = 2, TO:634-685, and the counters the reference prints (SURVEY 4 [probe]) ----
    As = K.dense(2, SP_A); Bs = K.dense(2, SP_B)
    P, nm, nb, _ = K.product(As, 0, Bs, 0, spamm=True, tau=0.2)
    _eq(K.to_dense(P), SP_C, "spamm tau=0.2"); n += 1
    assert (nm, nb) == (6, 4), (nm, nb)
    for (tA, tB), want in {(0, 1): 6, (1, 0): 7, (1, 1): 7}.items():                 # TO:731-742
        _, nm, nb, _ = K.product(As, tA, Bs, tB, spamm=True, tau=0.2)
        assert (nm, nb) == (want, 4), ((tA, tB), nm, nb)
        n += 1
    # error of SpAMM vs tau = 0 is monotone in tau and zero at tau -> 0 (TO:714-723 pattern: approx + (-exact))
    exact = K.product(As, 0, Bs, 0, spamm=True, tau=0.0)[0]
    prev = -1.0
    for tau in (0.0125, 0.025, 0.05, 0.1, 0.2, 0.4, 0.8):                            # TO:691-697
        approx = K.product(As, 0, Bs, 0, spamm=True, tau=tau)[0]
        err = K.frob_sq(K.add(approx, K.rescale(exact, -1.0)))
        assert err >= prev
        prev = err
        n += 1
    # ---- frob_block_trunc(0.21), TO:746-778: the two blocks of 0.1-entries go, and only they ----
    T, removed = K.trunc(K.dense(2, SP_A, update=False), 0.21)
    _eq(K.to_dense(T), [[1, 2, 0, 0], [2, 1, 0, 0], [3, 1, 0, 0], [5, 1, 0, 0]], "frob_block_trunc(0.21)"); n += 1
    assert removed and K.n_blocks(T) == 2
    T, removed = K.trunc(K.dense(2, SP_A, update=False), 0.0)
    assert not removed and K.n_blocks(T) == 4
    n += 1
    # ---- dummy-level squeeze: 1x4 * 4x1 -> 1x1 at b = 2, TO:782-862 ----
    r1 = K.coo(2, 1, 4, [0, 0, 0, 0], [0, 1, 2, 3], [1, 2, 3, 4])
    c1 = K.coo(2, 4, 1, [0, 1, 2, 3], [0, 0, 0, 0], [5, 6, 7, 8])
    P = K.product(r1, 0, c1, 0)[0]
    _eq(K.to_dense(P), [[70]], "1x4 * 4x1"); n += 1
    assert K.depth(P) == K.depth(K.coo(2, 1, 1, [0], [0], [70])) == 0               # TO:846
    assert not K.consistent(K.sized(2, 4, 4)) and K.consistent(K.sized(2, 2, 2))    # TO:856-861
    # ---- 2x6 * 6x2 at b = 1 (zeros assigned explicitly), TO:865-912 ----
    a26 = K.dense(1, [[1, 2, 3, 4, 5, 6], [7, 8, 9, 10, 11, 12]])
    b62 = K.dense(1, [[0, 0], [0, 0], [0, 0], [0, 0], [1, 1], [1, 1]])
    P = K.product(a26, 0, b62, 0)[0]
    _eq(K.to_dense(P), [[11, 11], [23, 23]], "2x6 * 6x2"); n += 1
    assert K.depth(P) == K.depth(K.dense(1, [[11, 11], [23, 23]])) == 1             # TO:909
    # ---- creation: TC:60-128 (b = 4, 14x14, two entries) ----
    M = K.coo(4, 14, 14, [0, 6], [0, 7], [7.7, 1.1])
    assert abs(K.frob_sq(M) - (7.7 * 7.7 + 1.1 * 1.1)) <= 1e-7                      # TC:86-89
    assert K.nnz(M) == 2 and K.n_blocks(M) == 2
    _eq(K.get(M, [0, 6, 3], [0, 7, 3]), [7.7, 1.1, 0.0], "get_values"); n += 1      # TC:98-118
    r, c, v = K.get_all(M)
    assert list(r) == [0, 6] and list(c) == [0, 7] and list(v) == [7.7, 1.1]        # TC:120-127 (order!)
    n += 1
    # ---- creation: non-square 14x17 at b = 3, TC:208-231 pattern ----
    M = K.coo(3, 14, 17, [0, 13, 5], [0, 16, 9], [1.5, -2.5, 4.0])
    _eq(K.get(M, [0, 13, 5, 1], [0, 16, 9, 1]), [1.5, -2.5, 4.0, 0.0], "14x17 get_values"); n += 1
    assert K.depth(M) == 3 and K.shape(M) == (14, 17)
    # ---- 1x1 at b = 1 round trip, TC:280-304 ----
    M = K.coo(1, 1, 1, [0], [0], [3.25])
    assert K.depth(M) == 0 and K.nnz(M) == 1 and K.get(M, [0], [0])[0] == 3.25
    n += 1
    return n
