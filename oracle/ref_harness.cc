// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// oracle/_ref harness: compiles the UNMODIFIED reference header where it lies
// (/root/reference/source/HierarchicalBlockSparseMatrix.h, passed with -I by oracle/Makefile)
// into oracle/_ref/libhbsm_ref.so and exposes a flat C API for tests/ and bench.py's
// cpu_baseline / --impl reference legs.  No reference source is copied into this repo.
//
// Usage rules inherited from SURVEY.md 8(c):
//   * spamm is always called with updated=true after update_internal_info()
//     (updated=false is a use-after-free in the batched build, H:6294-6307);
//   * built with -DNDEBUG so symm_* run on sparse inputs (H:5689 assert);
//   * leaves are read directly from the tree (get_all_values reserves nRows*nCols ints, H:1043).
//
// The reference's private members are reached with the classic "#define private public"
// after all standard headers are already included (their include guards keep them untouched).
#include <vector>
#include <list>
#include <iterator>
#include <stdexcept>
#include <cstring>
#include <iostream>
#include <cstdio>
#include <cstdlib>
#include <cassert>
#include <cmath>
#include <random>
#include <algorithm>
#include <functional>
#include <memory>
#include <unordered_map>
#include <map>
#include <string>
#include <chrono>
#include <cstdint>

#define private public
#include "hierarchical_block_sparse_lib.h"
#undef private

namespace {

thread_local std::string g_err;

template <class T> using Mat = hbsm::HierarchicalBlockSparseMatrix<T>;

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// tile coordinates (block row, block col) of a node from its position code, H:906 digit = 2*colbit+rowbit
inline void code_to_rc(const std::string& code, long& r, long& c) {
    r = 0; c = 0;
    for (char ch : code) {
        int d = ch - '0';
        r = (r << 1) | (d & 1);
        c = (c << 1) | ((d >> 1) & 1);
    }
}

template <class T>
void walk_leaves(const Mat<T>& m, long r, long c, std::vector<long>& bi, std::vector<long>& bj,
                 std::vector<const Mat<T>*>& leaves) {
    if (m.lowest_level()) {
        bi.push_back(r); bj.push_back(c); leaves.push_back(&m);
        return;
    }
    for (int ch = 0; ch < 4; ++ch)
        if (m.children[ch] != NULL)
            walk_leaves<T>(*m.children[ch], 2 * r + (ch & 1), 2 * c + ((ch >> 1) & 1), bi, bj, leaves);
}

template <class T> struct Api {
    static void* create(int blocksize) {
        Mat<T>* m = new Mat<T>();
        typename Mat<T>::Params p; p.blocksize = blocksize;
        m->set_params(p);
        return m;
    }
    static void destroy(void* h) { delete static_cast<Mat<T>*>(h); }
    static Mat<T>& M(void* h) { return *static_cast<Mat<T>*>(h); }

    static long n_leaves(void* h) {
        if (M(h).empty()) return 0;
        std::vector<long> bi, bj; std::vector<const Mat<T>*> lv;
        walk_leaves<T>(M(h), 0, 0, bi, bj, lv);
        return (long)lv.size();
    }
    // leaves in reference child order (= ascending Morton key); any output pointer may be NULL
    static long export_leaves(void* h, long* obi, long* obj, T* onorm_cached, T* otiles) {
        if (M(h).empty()) return 0;
        std::vector<long> bi, bj; std::vector<const Mat<T>*> lv;
        walk_leaves<T>(M(h), 0, 0, bi, bj, lv);
        const size_t bb = (size_t)M(h).blocksize * M(h).blocksize;
        for (size_t i = 0; i < lv.size(); ++i) {
            if (obi) obi[i] = bi[i];
            if (obj) obj[i] = bj[i];
            if (onorm_cached) onorm_cached[i] = lv[i]->frob_norm_squared_internal;
            if (otiles) std::memcpy(otiles + i * bb, &lv[i]->submatrix[0], bb * sizeof(T));
        }
        return (long)lv.size();
    }
    // executed-product set of multiply (is_spamm=0) or spamm: triples (ci, cj, k), unsorted.
    // Runs only the reference's symbolic phase (H:5478 / H:6291) on a scratch C.
    static long task_set(void* a, int tA, void* b, int tB, int is_spamm, T tau, long cap,
                         long* ci, long* cj, long* kk) {
        typename Mat<T>::BatchMapMultiply batches;
        Mat<T> C;
        if (is_spamm) Mat<T>::get_batches_spamm(M(a), tA, M(b), tB, C, tau, true, batches, 0);
        else Mat<T>::get_batches_multiply(M(a), tA, M(b), tB, C, batches, 0);
        // leaf coordinates come from a walk from each operand's root, not from get_position_code(): the latter follows
        // `parent` pointers, which transpose()/remove_dummy_levels leave unset or dangling
        std::unordered_map<const Mat<T>*, std::pair<long, long> > pos_a, pos_b;
        {
            std::vector<long> bi, bj; std::vector<const Mat<T>*> lv;
            walk_leaves<T>(M(a), 0, 0, bi, bj, lv);
            for (size_t i = 0; i < lv.size(); ++i) pos_a[lv[i]] = std::make_pair(bi[i], bj[i]);
            bi.clear(); bj.clear(); lv.clear();
            walk_leaves<T>(M(b), 0, 0, bi, bj, lv);
            for (size_t i = 0; i < lv.size(); ++i) pos_b[lv[i]] = std::make_pair(bi[i], bj[i]);
        }
        long n = 0;
        for (auto it = batches.begin(); it != batches.end(); ++it) {
            auto fa = pos_a.find(it->second.a); auto fb = pos_b.find(it->second.b);
            if (fa == pos_a.end() || fb == pos_b.end()) throw std::runtime_error("ref_harness: triplet operand is not a leaf of A/B");
            long ar = fa->second.first, ac = fa->second.second, br = fb->second.first, bc = fb->second.second;
            long i = tA ? ac : ar, k = tA ? ar : ac, k2 = tB ? bc : br, j = tB ? br : bc;
            if (k != k2) throw std::runtime_error("ref_harness: inconsistent k in triplet");
            if (n < cap) { ci[n] = i; cj[n] = j; kk[n] = k; }
            ++n;
        }
        return n;
    }
};

}  // namespace

#define GUARD(...)                                               \
    try { __VA_ARGS__; return 0; }                               \
    catch (const std::exception& e) { g_err = e.what(); return 1; } \
    catch (...) { g_err = "unknown exception"; return 2; }

#define DEFINE_API(SUF, T)                                                                                   \
    extern "C" void* ref_create_##SUF(int bs) { return Api<T>::create(bs); }                                 \
    extern "C" void ref_destroy_##SUF(void* h) { Api<T>::destroy(h); }                                       \
    extern "C" int ref_resize_##SUF(void* h, int m, int n) { GUARD(Api<T>::M(h).resize(m, n)) }              \
    extern "C" int ref_clear_##SUF(void* h) { GUARD(Api<T>::M(h).clear()) }                                  \
    extern "C" int ref_empty_##SUF(void* h) { return Api<T>::M(h).empty() ? 1 : 0; }                         \
    extern "C" int ref_n_rows_##SUF(void* h) { return Api<T>::M(h).get_n_rows(); }                           \
    extern "C" int ref_n_cols_##SUF(void* h) { return Api<T>::M(h).get_n_cols(); }                           \
    extern "C" int ref_depth_##SUF(void* h) { return Api<T>::M(h).get_depth(); }                             \
    extern "C" int ref_consistent_##SUF(void* h) { return Api<T>::M(h).check_if_matrix_is_consistent(); }    \
    extern "C" long ref_n_blocks_##SUF(void* h) { return (long)Api<T>::M(h).get_n_blocks(); }                \
    extern "C" long ref_n_mults_##SUF(void* h) { return (long)Api<T>::M(h).get_n_block_multiplications(); }  \
    extern "C" int ref_assign_##SUF(void* h, long n, const int* r, const int* c, const T* v, int use_max) {  \
        GUARD(std::vector<int> rr(r, r + n), cc(c, c + n); std::vector<T> vv(v, v + n);                      \
              if (use_max) Api<T>::M(h).assign_from_vectors_max(rr, cc, vv);                                 \
              else Api<T>::M(h).assign_from_vectors(rr, cc, vv)) }                                           \
    extern "C" int ref_update_##SUF(void* h) { GUARD(Api<T>::M(h).update_internal_info()) }                  \
    extern "C" int ref_frob_sq_##SUF(void* h, T* out) { GUARD(*out = Api<T>::M(h).get_frob_squared()) }      \
    extern "C" int ref_frob_sq_cached_##SUF(void* h, T* out) {                                               \
        GUARD(*out = Api<T>::M(h).get_frob_norm_squared_internal()) }                                        \
    extern "C" int ref_nnz_##SUF(void* h, long* out) { GUARD(*out = (long)Api<T>::M(h).get_nnz()) }          \
    extern "C" int ref_get_values_##SUF(void* h, long n, const int* r, const int* c, T* out) {               \
        GUARD(std::vector<int> rr(r, r + n), cc(c, c + n); std::vector<T> vv;                                \
              Api<T>::M(h).get_values(rr, cc, vv);                                                           \
              for (long i = 0; i < (long)vv.size(); ++i) out[i] = vv[i]) }                                   \
    /* two-call protocol: cap=0 returns the count */                                                         \
    extern "C" long ref_get_all_values_##SUF(void* h, long cap, int* r, int* c, T* v) {                      \
        try { std::vector<int> rr, cc; std::vector<T> vv; Api<T>::M(h).get_all_values(rr, cc, vv);           \
              long n = (long)vv.size();                                                                      \
              if (cap >= n) for (long i = 0; i < n; ++i) { r[i] = rr[i]; c[i] = cc[i]; v[i] = vv[i]; }       \
              return n; } catch (const std::exception& e) { g_err = e.what(); return -1; } }                 \
    extern "C" long ref_n_leaves_##SUF(void* h) { return Api<T>::n_leaves(h); }                              \
    extern "C" long ref_export_leaves_##SUF(void* h, long* bi, long* bj, T* nrm, T* tiles) {                 \
        return Api<T>::export_leaves(h, bi, bj, nrm, tiles); }                                               \
    extern "C" long ref_task_set_##SUF(void* a, int tA, void* b, int tB, int is_spamm, T tau, long cap,      \
                                       long* ci, long* cj, long* kk) {                                       \
        try { return Api<T>::task_set(a, tA, b, tB, is_spamm, tau, cap, ci, cj, kk); }                       \
        catch (const std::exception& e) { g_err = e.what(); return -1; } }                                   \
    extern "C" int ref_multiply_##SUF(void* a, int tA, void* b, int tB, void* c, long* nm, long* nb) {       \
        GUARD(size_t m = 0, r = 0;                                                                           \
              Mat<T>::multiply(Api<T>::M(a), tA, Api<T>::M(b), tB, Api<T>::M(c), &m, &r);                    \
              if (nm) *nm = (long)m; if (nb) *nb = (long)r) }                                                \
    extern "C" int ref_spamm_##SUF(void* a, int tA, void* b, int tB, void* c, T tau, long* nm, long* nb) {   \
        GUARD(size_t m = 0, r = 0;                                                                           \
              Mat<T>::spamm(Api<T>::M(a), tA, Api<T>::M(b), tB, Api<T>::M(c), tau, true, &m, &r);            \
              if (nm) *nm = (long)m; if (nb) *nb = (long)r) }                                                \
    /* phase-split timing of the reference's own two phases (H:3931-3989 / H:2142-2199) */                   \
    extern "C" int ref_product_timed_##SUF(void* a, int tA, void* b, int tB, void* c, int is_spamm, T tau,   \
                                           long* nm, long* nb, double* t3) {                                 \
        GUARD(Mat<T>& A = Api<T>::M(a); Mat<T>& B = Api<T>::M(b); Mat<T>& C = Api<T>::M(c);                  \
              typename Mat<T>::BatchMapMultiply batches;                                                     \
              int bs = A.blocksize;                                                                          \
              int AM = tA ? A.nCols_orig : A.nRows_orig, AN = tA ? A.nRows_orig : A.nCols_orig;              \
              int BN = tB ? B.nRows_orig : B.nCols_orig;                                                     \
              double t0 = now_s();                                                                           \
              batches.reserve(((size_t)(AM / bs) + 1) * ((size_t)(BN / bs) + 1) * ((size_t)(AN / bs) + 1));  \
              double t1 = now_s();                                                                           \
              if (is_spamm) Mat<T>::get_batches_spamm(A, tA, B, tB, C, tau, true, batches, 0);               \
              else Mat<T>::get_batches_multiply(A, tA, B, tB, C, batches, 0);                                \
              double t2 = now_s();                                                                           \
              Mat<T>::multiply_batches(A, B, C, batches);                                                    \
              double t3e = now_s();                                                                          \
              if (nm) *nm = (long)batches.size(); if (nb) *nb = (long)C.get_n_blocks();                      \
              C.n_block_multiplies = batches.size();                                                         \
              if (t3) { t3[0] = t1 - t0; t3[1] = t2 - t1; t3[2] = t3e - t2; }) }                             \
    extern "C" int ref_add_##SUF(void* a, void* b, void* c) {                                                \
        GUARD(Mat<T>::add(Api<T>::M(a), Api<T>::M(b), Api<T>::M(c))) }                                       \
    extern "C" int ref_transpose_##SUF(void* a, void* c) {                                                   \
        GUARD(Mat<T>::transpose(Api<T>::M(a), Api<T>::M(c))) }                                               \
    extern "C" int ref_upper_##SUF(void* a, void* c) { GUARD(Api<T>::M(a).get_upper_triangle(Api<T>::M(c))) } \
    extern "C" int ref_rescale_##SUF(void* c, void* a, T alpha) { GUARD(Api<T>::M(c).rescale(Api<T>::M(a), alpha)) } \
    extern "C" int ref_copy_##SUF(void* c, void* a) { GUARD(Api<T>::M(c).copy(Api<T>::M(a))) }               \
    extern "C" int ref_symm_multiply_##SUF(void* a, int sA, void* b, int sB, void* c) {                      \
        GUARD(Mat<T>::symm_multiply(Api<T>::M(a), sA, Api<T>::M(b), sB, Api<T>::M(c))) }                     \
    extern "C" int ref_symm_square_##SUF(void* a, void* c) {                                                 \
        GUARD(Mat<T>::symm_square(Api<T>::M(a), Api<T>::M(c))) }                                             \
    extern "C" int ref_symm_rk_##SUF(void* a, int tr, void* c) {                                             \
        GUARD(Mat<T>::symm_rk(Api<T>::M(a), tr, Api<T>::M(c))) }                                             \
    extern "C" int ref_worth_to_multiply_##SUF(void* a, int tA, void* b, int tB) {                           \
        return Mat<T>::worth_to_multiply(Api<T>::M(a), tA, Api<T>::M(b), tB); }                              \
    extern "C" int ref_worth_to_spamm_##SUF(void* a, int tA, void* b, int tB, T tau) {                       \
        return Mat<T>::worth_to_spamm(Api<T>::M(a), tA, Api<T>::M(b), tB, tau); }                            \
    extern "C" int ref_trunc_##SUF(void* a, void* c, T tau, int* removed) {                                   \
        GUARD(*removed = Api<T>::M(a).frob_block_trunc(Api<T>::M(c), tau) ? 1 : 0) }                         \
    extern "C" int ref_write_to_buffer_##SUF(void* h, char* buf, long cap) {                                  \
        GUARD(Api<T>::M(h).write_to_buffer(buf, (size_t)cap)) }                                              \
    extern "C" int ref_assign_from_buffer_##SUF(void* h, const char* buf, long n) {                           \
        GUARD(Api<T>::M(h).assign_from_buffer(buf, (size_t)n)) }                                             \
    extern "C" int ref_count_skips_##SUF(void* a, int tA, void* b, int tB, long n, const T* taus, int tr, int sp,   \
                                         unsigned long* out) {                                                \
        GUARD(std::vector<T> t(taus, taus + n);                                                              \
              std::vector<unsigned long> r = Mat<T>::count_skips(Api<T>::M(a), tA, Api<T>::M(b), tB, t, tr, sp); \
              for (long i = 0; i < n; ++i) out[i] = r[i]) }                                                  \
    extern "C" long ref_spamm_errors_##SUF(void* a, int tA, void* b, int tB, long n, const T* taus, T* out) { \
        try { std::vector<T> t(taus, taus + n);                                                              \
              std::vector<T> r = Mat<T>::get_spamm_errors(Api<T>::M(a), tA, Api<T>::M(b), tB, t);            \
              for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];                                           \
              return (long)r.size(); }                                                                       \
        catch (const std::exception& e) { g_err = e.what(); return -1; } }                                   \
    extern "C" long ref_size_bytes_##SUF(void* h) { return (long)Api<T>::M(h).get_size(); }

DEFINE_API(d, double)
DEFINE_API(s, float)

extern "C" const char* ref_last_error() { return g_err.c_str(); }
