// estimators.cu -- the reference's a-priori analysis tools for approximate products:
//   count_skips H:4945, get_errors_of_approx_multiplication H:5193, get_spamm_errors H:5236 (+ sum_skips H:5158,
//   sum_spamm_errors H:5455).
// They walk the same node-pair lattice as the multiply (C quadrant (r,c), contraction half k -> op(A) child (r,k),
// op(B) child (k,c)) but only look at cached Frobenius norms, so they run on the host over a per-level view of the two
// block tables: level-l nodes = distinct (key >> 2l), norm^2 = sum of the children's cached norm^2 in child order, in
// Treal -- exactly what update_internal_info leaves in the reference's nodes (H:3918-3923).
#include "matrix.cuh"
#include <cmath>

namespace hbsm_b200 {

namespace {

struct Levels {
    int depth = 0;                                 // levels above the leaves
    std::vector<std::vector<uint64_t>> key;        // [level][node], ascending
    std::vector<std::vector<double>> nsq;          // cached norm^2, rounded to Treal at every level
};

Levels build_levels(const Matrix& A, int depth) {
    Levels t;
    t.depth = depth;
    t.key.resize(depth + 1); t.nsq.resize(depth + 1);
    if (A.L == 0) return t;
    std::vector<uint64_t> keys(A.L);
    std::vector<char> raw(A.L * A.esize());
    HB_CUDA(cudaMemcpyAsync(keys.data(), A.keys.p, A.L * sizeof(uint64_t), cudaMemcpyDeviceToHost, engine().stream));
    HB_CUDA(cudaMemcpyAsync(raw.data(), A.norms.p, A.L * A.esize(), cudaMemcpyDeviceToHost, engine().stream));
    sync_stream();
    const bool f32 = A.dtype == HBSM_F32;
    t.key[0] = keys;
    t.nsq[0].resize(A.L);
    for (size_t i = 0; i < A.L; ++i)
        t.nsq[0][i] = f32 ? (double)reinterpret_cast<const float*>(raw.data())[i] : reinterpret_cast<const double*>(raw.data())[i];
    for (int l = 1; l <= depth; ++l) {
        const auto& ck = t.key[l - 1];
        const auto& cn = t.nsq[l - 1];
        for (size_t i = 0; i < ck.size();) {
            const uint64_t pk = ck[i] >> 2;
            double sd = 0.0; float sf = 0.0f;
            for (; i < ck.size() && (ck[i] >> 2) == pk; ++i) { sd += cn[i]; sf += (float)cn[i]; }
            t.key[l].push_back(pk);
            t.nsq[l].push_back(f32 ? (double)sf : sd);
        }
    }
    return t;
}

long find_node(const Levels& t, int level, uint64_t key) {
    const auto& k = t.key[level];
    auto it = std::lower_bound(k.begin(), k.end(), key);
    return (it != k.end() && *it == key) ? (long)(it - k.begin()) : -1;
}

struct Ctx {
    const Levels* A; const Levels* B;
    bool tA, tB, f32;
    std::vector<double> taus;
    bool trunc, spamm;
};

// child of op(X) at (row bit rb, col bit cb): digit = 2*colbit + rowbit, transposed operands swap the roles
uint64_t op_child(uint64_t key, bool t, int rb, int cb) { return (key << 2) | (uint64_t)(t ? 2 * rb + cb : 2 * cb + rb); }

double rnd(const Ctx& c, double v) { return c.f32 ? (double)(float)v : v; }

void node_norms(const Ctx& c, int level, long ia, long ib, double& na, double& nb, double& prod) {
    na = rnd(c, std::sqrt(c.A->nsq[level][ia]));
    nb = rnd(c, std::sqrt(c.B->nsq[level][ib]));
    if (c.f32) { na = (double)sqrtf((float)c.A->nsq[level][ia]); nb = (double)sqrtf((float)c.B->nsq[level][ib]); }
    prod = rnd(c, na * nb);
}

std::vector<unsigned long> skips_rec(const Ctx& c, int level, uint64_t ka, uint64_t kb) {
    const long ia = find_node(*c.A, level, ka), ib = find_node(*c.B, level, kb);
    const size_t n = c.taus.size();
    std::vector<unsigned long> cur(n, 0);
    double na, nb, prod;
    node_norms(c, level, ia, ib, na, nb, prod);
    for (size_t i = 0; i < n; ++i) {   // H:4960-4982
        const double tau = c.taus[i];
        bool skip = false;
        if (c.trunc && !c.spamm) skip = na < tau || nb < tau;
        else if (!c.trunc && c.spamm) skip = prod < tau;
        else if (c.trunc && c.spamm) skip = na < tau || nb < tau || prod < tau;
        cur[i] = skip ? 1 : 0;
    }
    if (level == 0) return cur;
    std::vector<unsigned long> sub(n, 0);
    for (int q = 0; q < 4; ++q)
        for (int kb2 = 0; kb2 < 2; ++kb2) {
            const int rb = q & 1, cb = (q >> 1) & 1;
            const uint64_t ca = op_child(ka, c.tA, rb, kb2), cbk = op_child(kb, c.tB, kb2, cb);
            if (find_node(*c.A, level - 1, ca) < 0 || find_node(*c.B, level - 1, cbk) < 0) continue;
            std::vector<unsigned long> s = skips_rec(c, level - 1, ca, cbk);
            for (size_t i = 0; i < n; ++i) sub[i] += s[i];
        }
    for (size_t i = 0; i < n; ++i) cur[i] = cur[i] == 1 ? 1 : sub[i];   // sum_skips H:5158
    return cur;
}

// worth_to_multiply H:1873 on a node pair: does a chain of existing children reach a leaf pair?
bool worth_rec(const Ctx& c, int level, uint64_t ka, uint64_t kb) {
    if (find_node(*c.A, level, ka) < 0 || find_node(*c.B, level, kb) < 0) return false;
    if (level == 0) return true;
    for (int q = 0; q < 4; ++q)
        for (int k2 = 0; k2 < 2; ++k2)
            if (worth_rec(c, level - 1, op_child(ka, c.tA, q & 1, k2), op_child(kb, c.tB, k2, (q >> 1) & 1))) return true;
    return false;
}

// get_spamm_errors H:5236: empty vector = "this pair contributes no product"
std::vector<double> errors_rec(const Ctx& c, int level, uint64_t ka, uint64_t kb) {
    std::vector<double> cur;
    if (!worth_rec(c, level, ka, kb)) return cur;
    const size_t n = c.taus.size();
    cur.assign(n, 0.0);
    const long ia = find_node(*c.A, level, ka), ib = find_node(*c.B, level, kb);
    double na, nb, prod;
    node_norms(c, level, ia, ib, na, nb, prod);
    for (size_t i = 0; i < n; ++i)
        if (prod < c.taus[i]) cur[i] = prod;
    if (level == 0) return cur;
    // per C quadrant the two k-halves add (operator+ H:435), then the quadrants combine in quadrature (sum_spamm_errors H:5455)
    std::vector<std::vector<double>> quad;
    bool any = false;
    for (int q = 0; q < 4; ++q) {
        const int rb = q & 1, cb = (q >> 1) & 1;
        std::vector<double> s[2];
        for (int k2 = 0; k2 < 2; ++k2) {
            const uint64_t ca = op_child(ka, c.tA, rb, k2), cbk = op_child(kb, c.tB, k2, cb);
            if (find_node(*c.A, level - 1, ca) >= 0 && find_node(*c.B, level - 1, cbk) >= 0) s[k2] = errors_rec(c, level - 1, ca, cbk);
        }
        std::vector<double> both;
        if (!s[0].empty() && !s[1].empty()) { both.resize(n); for (size_t i = 0; i < n; ++i) both[i] = rnd(c, s[0][i] + s[1][i]); }
        else if (!s[0].empty()) both = s[0];
        else if (!s[1].empty()) both = s[1];
        any = any || !both.empty();
        quad.push_back(both);
    }
    if (!any) return cur;
    std::vector<double> tot(n, 0.0);
    for (size_t j = 0; j < n; ++j) {
        double acc = 0.0;
        for (const auto& v : quad)
            if (!v.empty()) acc = rnd(c, acc + rnd(c, v[j] * v[j]));
        tot[j] = rnd(c, c.f32 ? (double)sqrtf((float)acc) : std::sqrt(acc));
    }
    return tot;
}

Ctx make_ctx(const Matrix& A, bool tA, const Matrix& B, bool tB, size_t n, const double* taus, Levels& la, Levels& lb) {
    if (A.empty() || B.empty()) throw Error(HBSM_E_ARG, "hbsm_b200: estimator on an empty matrix");
    if (A.dtype != B.dtype || A.b != B.b) throw Error(HBSM_E_ARG, "hbsm_b200: operands differ in dtype or blocksize");
    ensure_engine();
    // operands of different depth: the shallower tree is block (0,0) of the deeper grid -- keys unchanged, more levels
    const int depth = std::max(A.vdepth(), B.vdepth());
    la = build_levels(A, depth);
    lb = build_levels(B, depth);
    Ctx c;
    c.A = &la; c.B = &lb; c.tA = tA; c.tB = tB; c.f32 = A.dtype == HBSM_F32;
    c.taus.assign(taus, taus + n);
    c.trunc = false; c.spamm = true;
    return c;
}

}  // namespace

void count_skips(const Matrix& A, bool tA, const Matrix& B, bool tB, size_t n, const double* taus, bool apply_truncation, bool apply_spamm,
                 unsigned long* out) {
    Levels la, lb;
    Ctx c = make_ctx(A, tA, B, tB, n, taus, la, lb);
    c.trunc = apply_truncation; c.spamm = apply_spamm;
    for (size_t i = 0; i < n; ++i) out[i] = 0;
    if (A.L == 0 || B.L == 0) return;
    std::vector<unsigned long> r = skips_rec(c, c.A->depth, 0, 0);
    for (size_t i = 0; i < n; ++i) out[i] = r[i];
}

size_t spamm_errors(const Matrix& A, bool tA, const Matrix& B, bool tB, size_t n, const double* taus, double* out) {
    Levels la, lb;
    Ctx c = make_ctx(A, tA, B, tB, n, taus, la, lb);
    if (A.L == 0 || B.L == 0) return 0;
    std::vector<double> r = errors_rec(c, c.A->depth, 0, 0);
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return r.size();
}

}  // namespace hbsm_b200
