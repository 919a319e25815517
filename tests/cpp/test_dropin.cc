// test_dropin.cc -- the reference's known-answer tests (test_source/test_matrix_operations.cc = TO:,
// test_source/test_matrix_creation.cc = TC:) written against the drop-in C++ class
// include/hbsm/HierarchicalBlockSparseMatrix.h, i.e. through the C ABI onto the B200.
// Usage is deliberately that of a reference caller: same includes (via the umbrella header), same type name,
// same method calls, exact `==` comparison of dense expansions.  Exit code 0 = all passed.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

#include "hbsm/hierarchical_block_sparse_lib.h"

template <class T>
struct Coo {
    std::vector<int> r, c;
    std::vector<T> v;
    void row(int i, std::initializer_list<double> vals) {   // like the reference's set_row: zeros are assigned too
        int j = 0;
        for (double x : vals) { r.push_back(i); c.push_back(j++); v.push_back((T)x); }
    }
};

template <class M>
void fill(M& A, int b, int m, int n, std::initializer_list<std::initializer_list<double> > rows) {
    typename M::Params p; p.blocksize = b;
    A.set_params(p);
    A.resize(m, n);
    Coo<typename M::real> s;
    int i = 0;
    for (auto& r : rows) s.row(i++, r);
    A.assign_from_vectors(s.r, s.c, s.v);
}

template <class M>
std::vector<double> dense(const M& A) {
    std::vector<int> r, c;
    std::vector<typename M::real> v;
    A.get_all_values(r, c, v);
    std::vector<double> d((size_t)A.get_n_rows() * A.get_n_cols(), 0.0);
    for (size_t i = 0; i < r.size(); ++i) d[(size_t)r[i] * A.get_n_cols() + c[i]] = v[i];
    return d;
}

static int g_checks = 0;
#define REQUIRE(cond) do { ++g_checks; if (!(cond)) throw std::runtime_error(std::string("FAILED: ") + #cond + " at line " + std::to_string(__LINE__)); } while (0)

template <class M>
void expect(const M& A, int m, int n, std::initializer_list<std::initializer_list<double> > rows) {
    REQUIRE(A.get_n_rows() == m && A.get_n_cols() == n);
    std::vector<double> d = dense(A);
    size_t k = 0;
    for (auto& r : rows)
        for (double x : r) {
            // exact for doubles (integers / short decimals computed in a fixed order); one rounding for float inputs
            REQUIRE(d[k] == (double)(typename M::real)x || std::fabs(d[k] - x) <= 1e-6 * std::fabs(x));
            ++k;
        }
    REQUIRE(k == d.size());
}

template <class T>
void run() {
    typedef hbsm::HierarchicalBlockSparseMatrix<T> M;
    // ---- creation, TC:27-128 ----
    {
        M A;
        REQUIRE(A.empty());
        typename M::Params p; p.blocksize = 4;
        A.set_params(p);
        REQUIRE(A.get_params().blocksize == 4);
        A.resize(14, 14);
        REQUIRE(!A.empty() && A.get_n_rows() == 14 && A.get_n_cols() == 14 && !A.children_exist());
        std::vector<int> r = {0, 6}, c = {0, 7};
        std::vector<T> v = {(T)7.7, (T)1.1};
        A.assign_from_vectors(r, c, v);
        REQUIRE(A.children_exist());
        REQUIRE(std::fabs((double)A.get_frob_squared() - (7.7 * 7.7 + 1.1 * 1.1)) < 1e-5);
        REQUIRE(A.get_nnz() == 2);
        std::vector<int> qr = {0, 6, 3}, qc = {0, 7, 3};
        std::vector<T> out;
        A.get_values(qr, qc, out);
        REQUIRE(out[0] == (T)7.7 && out[1] == (T)1.1 && out[2] == (T)0);
        std::vector<int> ar, ac; std::vector<T> av;
        A.get_all_values(ar, ac, av);
        REQUIRE(ar.size() == 2 && ar[0] == 0 && ac[0] == 0 && ar[1] == 6 && ac[1] == 7);
        A.clear();
        REQUIRE(A.empty());
        bool thrown = false;
        try { A.get_frob_squared(); } catch (const std::runtime_error& e) { thrown = std::string(e.what()).find("empty matrix occured") != std::string::npos; }
        REQUIRE(thrown);
    }
    // ---- transpose / add / multiply NN NT TN TT, TO:236-384 ----
    M A, B, D;
    fill(A, 2, 2, 3, {{2, 3, 5}, {0, 1, 2}});
    fill(B, 2, 2, 3, {{1, 3, 2}, {6, 2, 4}});
    fill(D, 2, 3, 2, {{2, 1}, {7, 3}, {3, 5}});
    { M AT; M::transpose(A, AT); expect(AT, 3, 2, {{2, 0}, {3, 1}, {5, 2}}); }
    { M C; M::add(A, B, C); expect(C, 2, 3, {{3, 6, 7}, {6, 3, 6}}); }
    size_t nm = 0, nr = 0;
    { M C; M::multiply(A, false, D, false, C, &nm, &nr); expect(C, 2, 2, {{40, 36}, {13, 13}}); REQUIRE(nm == 2 && nr == 1); }
    { M C; M::multiply(D, false, A, false, C); expect(C, 3, 3, {{4, 7, 12}, {14, 24, 41}, {6, 14, 25}}); }
    { M C; M::multiply(A, false, A, true, C); expect(C, 2, 2, {{38, 13}, {13, 5}}); }
    { M C; M::multiply(A, true, A, false, C); expect(C, 3, 3, {{4, 6, 10}, {6, 10, 17}, {10, 17, 29}}); }
    { M C; M::multiply(A, true, D, true, C); expect(C, 3, 3, {{4, 14, 6}, {7, 24, 14}, {12, 41, 25}}); }
    {   // C must be empty on entry (H:5681)
        M C; M::multiply(A, false, D, false, C);
        bool thrown = false;
        try { M::multiply(A, false, D, false, C); } catch (const std::runtime_error& e) { thrown = std::string(e.what()).find("non-empty matrix to write result") != std::string::npos; }
        REQUIRE(thrown);
        thrown = false;
        M C2;
        try { M::multiply(A, false, A, false, C2); } catch (const std::runtime_error& e) { thrown = std::string(e.what()).find("bad sizes") != std::string::npos; }
        REQUIRE(thrown);
    }
    // ---- 2-level x 3-level at b = 3, TO:430-509 ----
    {
        M A5, B5;
        fill(A5, 3, 5, 5, {{2, 2, 3, 5, 2}, {2, 1, 2, 4, 3}, {3, 2, 3, 1, 2}, {5, 4, 1, 4, 5}, {2, 3, 2, 5, 1}});
        fill(B5, 3, 5, 7, {{5, 3, 1, 5, 0, 3, 3}, {1, 5, 5, 4, 1, 5, 1}, {2, 1, 3, 2, 1, 1, 4}, {2, 2, 3, 3, 1, 4, 2}, {5, 1, 1, 2, 1, 2, 3}});
        REQUIRE(A5.get_depth() == 1 && B5.get_depth() == 2);
        REQUIRE(M::worth_to_multiply(A5, false, B5, false));
        M P; M::multiply(A5, false, B5, false, P);
        expect(P, 5, 7, {{38, 31, 38, 43, 12, 43, 36}, {38, 24, 28, 36, 10, 35, 32}, {35, 26, 27, 36, 8, 30, 31}, {64, 49, 45, 65, 14, 62, 46}, {32, 34, 39, 43, 11, 45, 30}});
        REQUIRE(P.get_depth() == 2);
        M Q; M::multiply(B5, true, A5, false, Q);
        expect(Q, 7, 5, {{38, 38, 35, 64, 32}, {31, 24, 26, 49, 34}, {38, 28, 27, 45, 39}, {43, 36, 36, 65, 43}, {12, 10, 8, 14, 11}, {43, 35, 30, 62, 45}, {36, 32, 31, 46, 30}});
        REQUIRE(Q.get_n_block_multiplications() == 12);   // TO:507
    }
    // ---- rescale, TO:513-526 ----
    { M R; R.rescale(A, (T)-1.0); expect(R, 2, 3, {{-2, -3, -5}, {0, -1, -2}}); }
    // ---- symmetric family, TO:20-185 ----
    {
        M U, B7, C2;
        fill(U, 2, 5, 5, {{2, 2, 3, 5, 2}, {0, 1, 2, 4, 3}, {0, 0, 3, 1, 2}, {0, 0, 0, 4, 5}, {0, 0, 0, 0, 1}});
        fill(B7, 2, 5, 7, {{5, 3, 1, 5, 0, 3, 3}, {1, 5, 5, 4, 1, 5, 1}, {2, 1, 3, 2, 1, 1, 4}, {2, 2, 3, 3, 1, 4, 2}, {5, 1, 1, 2, 1, 2, 3}});
        fill(C2, 2, 2, 5, {{5, 3, 0, 3, 3}, {1, 4, 1, 5, 1}});
        { M P; M::symm_multiply(U, true, B7, false, P);
          expect(P, 5, 7, {{38, 31, 38, 43, 12, 43, 36}, {38, 24, 28, 36, 10, 35, 32}, {35, 26, 27, 36, 8, 30, 31}, {64, 49, 45, 65, 14, 62, 46}, {32, 34, 39, 43, 11, 45, 30}}); }
        { M P; M::symm_multiply(C2, false, U, true, P); expect(P, 2, 5, {{37, 34, 30, 64, 37}, {40, 31, 21, 47, 42}}); }
        { M P; M::symm_square(U, P);
          expect(P, 5, 5, {{46, 38, 28, 51, 43}, {0, 34, 24, 47, 34}, {0, 0, 27, 40, 25}, {0, 0, 0, 83, 49}, {0, 0, 0, 0, 43}}); }
        { M P; M::symm_rk(A, false, P); expect(P, 2, 2, {{38, 13}, {0, 5}}); }
        { M P; M::symm_rk(A, true, P); expect(P, 3, 3, {{4, 6, 10}, {0, 10, 17}, {0, 0, 29}}); }
        bool thrown = false;
        try { M P; M::symm_multiply(U, true, B7, true, P); } catch (const std::runtime_error&) { thrown = true; }
        REQUIRE(thrown);
    }
    // ---- SpAMM, TO:634-742 ----
    {
        M As, Bs;
        fill(As, 2, 4, 4, {{1, 2, 0.1, 0.1}, {2, 1, 0.1, 0.1}, {3, 1, 0, 0}, {5, 1, 0, 0.1}});
        fill(Bs, 2, 4, 4, {{1, 6, 5, 0}, {0, 1, 2, 1}, {2, 1, 0.1, 0.1}, {0, 1, 0.1, 0.1}});
        As.update_internal_info();
        Bs.update_internal_info();
        { M P; M::spamm(As, false, Bs, false, P, (T)0.2, true, &nm, &nr);
          expect(P, 4, 4, {{1.2, 8.2, 9, 2}, {2.2, 13.2, 12, 1}, {3, 19, 17, 1}, {5, 31.1, 27, 1}});
          REQUIRE(nm == 6 && nr == 4); }
        { M P; M::spamm(As, false, Bs, true, P, (T)0.2, true, &nm, &nr); REQUIRE(nm == 6 && nr == 4); }
        { M P; M::spamm(As, true, Bs, false, P, (T)0.2, true, &nm, &nr); REQUIRE(nm == 7 && nr == 4); }
        { M P; M::spamm(As, true, Bs, true, P, (T)0.2, true, &nm, &nr); REQUIRE(nm == 7 && nr == 4); }
        {   // a-priori skip counts and error bounds, TO:687-723 (reference prints S = 0 1 2 2 2 3 4)
            std::vector<T> taus = {(T)0.0125, (T)0.025, (T)0.05, (T)0.1, (T)0.2, (T)0.4, (T)0.8};
            std::vector<unsigned long int> skips = M::count_skips(As, false, Bs, false, taus, false, true);
            const unsigned long want[7] = {0, 1, 2, 2, 2, 3, 4};
            for (int i = 0; i < 7; ++i) REQUIRE(skips[i] == want[i]);
            std::vector<T> bound = M::get_errors_of_approx_multiplication(As, false, Bs, false, taus, false, true);
            std::vector<T> est = M::get_spamm_errors(As, false, Bs, false, taus);
            REQUIRE(bound.size() == 7 && est.size() == 7);
            for (int i = 0; i < 7; ++i) {
                M approx, exact, minus, err;
                M::spamm(As, false, Bs, false, approx, taus[i], true);
                M::spamm(As, false, Bs, false, exact, (T)0.0, true);
                minus.rescale(exact, (T)-1.0);
                M::add(approx, minus, err);
                if (bound[i] > 0) REQUIRE(std::sqrt((double)err.get_frob_squared()) < (double)bound[i]);   // TO:722
            }
        }
        REQUIRE(M::worth_to_spamm(As, false, Bs, false, (T)0.2));
        REQUIRE(!M::worth_to_spamm(As, false, Bs, false, (T)1e6));
        // error vs tau = 0 via rescale + add + get_frob_squared (TO:714-723)
        double prev = -1;
        for (double tau : {0.0125, 0.025, 0.05, 0.1, 0.2, 0.4, 0.8}) {
            M approx, exact, minus, err;
            M::spamm(As, false, Bs, false, approx, (T)tau, true);
            M::spamm(As, false, Bs, false, exact, (T)0.0, true);
            minus.rescale(exact, (T)-1.0);
            M::add(approx, minus, err);
            double e = (double)err.get_frob_squared();
            REQUIRE(e >= prev);
            prev = e;
        }
    }
    // ---- dummy-level squeeze and consistency, TO:782-912 ----
    {
        typename M::Params p; p.blocksize = 2;
        M As, Bs;
        As.set_params(p); As.resize(1, 4);
        Bs.set_params(p); Bs.resize(4, 1);
        std::vector<int> z4 = {0, 0, 0, 0}, i4 = {0, 1, 2, 3};
        std::vector<T> v1 = {1, 2, 3, 4}, v2 = {5, 6, 7, 8};
        As.assign_from_vectors(z4, i4, v1);
        Bs.assign_from_vectors(i4, z4, v2);
        M P; M::multiply(As, false, Bs, false, P);
        expect(P, 1, 1, {{70}});
        REQUIRE(P.get_depth() == 0);
        M X; X.set_params(p); X.resize(4, 4);
        REQUIRE(!X.check_if_matrix_is_consistent());
        M Y; Y.set_params(p); Y.resize(2, 2);
        REQUIRE(Y.check_if_matrix_is_consistent());
        M A26, B62;
        fill(A26, 1, 2, 6, {{1, 2, 3, 4, 5, 6}, {7, 8, 9, 10, 11, 12}});
        fill(B62, 1, 6, 2, {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {1, 1}, {1, 1}});
        M Q; M::multiply(A26, false, B62, false, Q);
        expect(Q, 2, 2, {{11, 11}, {23, 23}});
        REQUIRE(Q.get_depth() == 1);
    }
    // ---- frob_block_trunc(0.21), TO:746-778 ----
    {
        M As, Bs;
        fill(As, 2, 4, 4, {{1, 2, 0.1, 0.1}, {2, 1, 0.1, 0.1}, {3, 1, 0, 0}, {5, 1, 0, 0.1}});
        REQUIRE(As.frob_block_trunc(Bs, (T)0.21));
        expect(Bs, 4, 4, {{1, 2, 0, 0}, {2, 1, 0, 0}, {3, 1, 0, 0}, {5, 1, 0, 0}});
        REQUIRE(Bs.get_n_blocks() == 2);
        M Cs;
        REQUIRE(!As.frob_block_trunc(Cs, (T)0.0));     // nothing to remove
        REQUIRE(Cs.get_n_blocks() == 4);
        std::vector<T> errs, taus = {(T)0.05, (T)0.21, (T)100};
        As.get_frob_squared_of_error_matrix(errs, taus);
        REQUIRE(errs.size() == 3 && errs[0] == (T)0);
        REQUIRE(std::fabs((double)errs[1] - 0.05) < 1e-6);   // the two 0.1-blocks: 4*0.01 + 1*0.01
        REQUIRE(std::fabs((double)errs[2] - (double)As.get_frob_squared()) < 1e-4);
    }
    // ---- get_trace / set_to_identity / get_nnz_diag_lowest_level / random_blocks, TO:568-630 ----
    {
        M A3; fill(A3, 2, 3, 3, {{2, 3, 5}, {0, 1, 2}, {5, 8, 9}});
        REQUIRE(A3.get_trace() == (T)12);
        typename M::Params p; p.blocksize = 2;
        M I17; I17.set_params(p);
        M::set_to_identity(I17, 17);
        REQUIRE(I17.get_trace() == (T)17 && I17.get_nnz() == 17 && I17.get_n_rows() == 17);
        M U3; fill(U3, 2, 3, 3, {{1, 2, 3}, {0, 6, 4}, {0, 0, 5}});
        REQUIRE(U3.get_nnz_diag_lowest_level() == 8);
        REQUIRE(U3.get_max_abs_value() == (T)6);
        p.blocksize = 3;
        M X; X.set_params(p); X.resize(6, 6);
        X.random_blocks(3);
        REQUIRE(X.get_nnz() == 27 && X.get_n_blocks() == 3);
    }
    // ---- add_scaled_identity, TC:233-279 ----
    {
        typename M::Params p; p.blocksize = 4;
        M E, D;
        E.set_params(p); E.resize(33, 33);
        std::vector<int> r = {5, 2}, c = {5, 24};
        std::vector<T> v = {(T)2.2, (T)-2.2};
        E.assign_from_vectors(r, c, v);
        D.add_scaled_identity(E, (T)0.5);
        std::vector<int> qr, qc; std::vector<T> out;
        for (int i = 0; i < 33; ++i) { qr.push_back(i); qc.push_back(i); }
        qr.push_back(2); qc.push_back(24); qr.push_back(3); qc.push_back(7);
        D.get_values(qr, qc, out);
        for (int i = 0; i < 33; ++i) REQUIRE(out[i] == (i == 5 ? (T)2.2 + (T)0.5 : (T)0.5));
        REQUIRE(out[33] == (T)-2.2 && out[34] == (T)0 && D.get_nnz() == 34);
    }
    // ---- inv_chol, TO:188-223 (5x5) and TO:527-562 (4x4), both at b = 2, to 1e-10 (fp64) ----
    {
        const double tol = sizeof(T) == 8 ? 1e-10 : 2e-6;
        M A5, Z5;
        fill(A5, 2, 5, 5, {{5, 0, 1, 0, 0}, {0, 6, 2, 0, 0}, {1, 2, 5, 3, 0}, {0, 0, 3, 8, 0}, {0, 0, 0, 0, 7}});
        M::inv_chol(A5, Z5);
        const double z5[5][5] = {{0.447213595499958, 0, -0.098373875367593, 0.060157954894827, 0},
                                 {0, 0.408248290463863, -0.163956458945988, 0.100263258158045, 0},
                                 {0, 0, 0.491869376837965, -0.300789774474136, 0},
                                 {0, 0, 0, 0.414421467053253, 0},
                                 {0, 0, 0, 0, 0.377964473009227}};
        std::vector<double> d5 = dense(Z5);
        REQUIRE(Z5.get_n_rows() == 5 && Z5.get_n_cols() == 5);
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) REQUIRE(std::fabs(d5[i * 5 + j] - z5[i][j]) <= tol);
        M A4, Z4;
        fill(A4, 2, 4, 4, {{5.2, 0.1, 0.4, 0.7}, {0.1, 5.8, 0.25, 0.55}, {0.4, 0.25, 6.7, 0.6}, {0.7, 0.55, 0.6, 5.5}});
        M::inv_chol(A4, Z4);
        const double z4[4][4] = {{0.438529009653515, -0.007986466419713, -0.029497652767486, -0.055022297341530},
                                 {0, 0.415296253825059, -0.016194789754698, -0.038713453134371},
                                 {0, 0, 0.387518183415998, -0.034114893784155},
                                 {0, 0, 0, 0.433761784290076}};
        std::vector<double> d4 = dense(Z4);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) REQUIRE(std::fabs(d4[i * 4 + j] - z4[i][j]) <= tol);
        // a larger SPD band through three levels with a ragged edge (37 = 4*8 + 5): Z^T A Z = I
        const int n = 37;
        typename M::Params p; p.blocksize = 8;
        M A, Z; A.set_params(p); A.resize(n, n);
        std::vector<int> r, c; std::vector<T> v;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
            const int d = i > j ? i - j : j - i;
            if (d <= 6) { r.push_back(i); c.push_back(j); v.push_back((T)(d == 0 ? 4.0 + 0.01 * i : 0.5 / (1 + d) + 0.001 * ((i * j) % 7))); }
        }
        A.assign_from_vectors(r, c, v);
        M::inv_chol(A, Z);
        M AZ, ZtAZ;
        M::multiply(A, false, Z, false, AZ);
        M::multiply(Z, true, AZ, false, ZtAZ);
        std::vector<double> id = dense(ZtAZ);
        double worst = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) worst = std::max(worst, std::fabs(id[i * n + j] - (i == j ? 1.0 : 0.0)));
        REQUIRE(worst <= (sizeof(T) == 8 ? 1e-12 : 5e-5));
        std::vector<double> zd = dense(Z);
        for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) REQUIRE(zd[i * n + j] == 0.0);   // upper triangular
    }
    // ---- host-to-host product in one pipelined call (engine extra): same result as assign + multiply ----
    {
        const int n = 24, b = 4, g = n / b;
        typename M::Params p; p.blocksize = b;
        std::vector<int> bi, bj;
        std::vector<T> tiles;
        for (int i = 0; i < g; ++i) for (int j = 0; j < g; ++j) {
            if (std::abs(i - j) > 1) continue;
            bi.push_back(i); bj.push_back(j);
            for (int e = 0; e < b * b; ++e) tiles.push_back((T)(1 + ((i * 7 + j * 3 + e) % 5)));   // small integers: exact in any order
        }
        M A2, B2, C2;
        A2.set_params(p); A2.resize(n, n); B2.set_params(p); B2.resize(n, n);
        std::vector<T> ct((size_t)g * g * b * b);
        std::vector<int> cbi(g * g), cbj(g * g);
        size_t nm = 0, nr = 0;
        M::product_from_host_tiles(A2, bi.size(), bi.data(), bj.data(), tiles.data(), false, B2, bi.size(), bi.data(), bj.data(),
                                   tiles.data(), true, C2, false, (T)0, ct.data(), (size_t)g * g, cbi.data(), cbj.data(), &nm, &nr);
        M C3;
        size_t nm3 = 0, nr3 = 0;
        M::multiply(A2, false, B2, true, C3, &nm3, &nr3);
        REQUIRE(nm == nm3 && nr == nr3 && nr == C2.get_n_blocks());
        std::vector<double> d2 = dense(C2), d3 = dense(C3);
        REQUIRE(d2 == d3);
        for (size_t t = 0; t < nr; ++t)   // host copy: column-major tiles labelled by block coordinates
            for (int c = 0; c < b; ++c) for (int r = 0; r < b; ++r)
                REQUIRE((double)ct[t * b * b + c * b + r] == d3[(size_t)(cbi[t] * b + r) * n + cbj[t] * b + c]);
    }
    // ---- value semantics of the drop-in: deep copy ----
    {
        M C1(A);
        M C2; C2 = A;
        expect(C1, 2, 3, {{2, 3, 5}, {0, 1, 2}});
        expect(C2, 2, 3, {{2, 3, 5}, {0, 1, 2}});
        bool thrown = false;
        try { M U; A.get_upper_triangle(U); } catch (const std::runtime_error& e) {   // non-square: throws (H:3517)
            thrown = std::string(e.what()).find("non-square") != std::string::npos;
        }
        REQUIRE(thrown);
    }
}

int main() {
    try {
        if (hbsm_init(0) != HBSM_OK) { std::fprintf(stderr, "hbsm_init: %s\n", hbsm_last_error()); return 2; }
        run<double>();
        int nd = g_checks;
        run<float>();
        std::printf("test_dropin: %d checks passed (double %d, float %d)\n", g_checks, nd, g_checks - nd);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "test_dropin: %s\n", e.what());
        return 1;
    }
}
