// test_sharded.cc -- the multi-GPU product through the C++ drop-in class: `world` processes are FORKED here (before any CUDA
// call), one per GPU, rank 0 creates the NCCL id and pipes its bytes to the others (the host program's own transport), every
// rank builds its slab of two banded decay matrices with the library's generator, publishes B and calls the sharded statics.
// Each rank also computes the FULL product on its own GPU and compares its slab of C with it: identical structure, identical
// executed-product count, bit-identical tile values (same kernels, same k order).
// Usage: test_sharded <world>        exit 0 = ok, 77 = skipped (fewer GPUs than ranks)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

#include "hbsm/HierarchicalBlockSparseMatrix.h"

typedef hbsm::HierarchicalBlockSparseMatrix<double> Matrix;

static void make_decay(Matrix& m, int n, int b, double lam, int seed, bool symmetric, int lo, int hi) {
    const int W = (int)std::floor(std::log(1e12) / lam);
    std::vector<double> table(W + 1);
    for (int d = 0; d <= W; ++d) table[d] = std::exp(-lam * d);
    Matrix::Params p; p.blocksize = b;
    m.set_params(p);
    hbsm::detail::check(hbsm_generate_decay(m.handle(), n, table.data(), W, (uint64_t)seed, symmetric ? 1 : 0, lo, hi));
    m.update_internal_info();
}

struct Leaves { std::vector<int64_t> bi, bj; std::vector<double> tiles; };
static Leaves leaves_of(const Matrix& m, int b) {
    Leaves L;
    size_t n = 0;
    hbsm::detail::check(hbsm_export_leaves(m.handle(), 0, NULL, NULL, NULL, NULL, &n));
    L.bi.resize(n); L.bj.resize(n); L.tiles.resize(n * (size_t)b * b);
    if (n) hbsm::detail::check(hbsm_export_leaves(m.handle(), n, L.bi.data(), L.bj.data(), NULL, L.tiles.data(), &n));
    return L;
}

static int check_slab(const Matrix& Cl, const Matrix& Cf, int b, int lo, int hi, const char* what) {
    Leaves l = leaves_of(Cl, b), f = leaves_of(Cf, b);
    size_t at = 0;
    const size_t te = (size_t)b * b;
    for (size_t i = 0; i < f.bi.size(); ++i) {
        if (f.bi[i] < lo || f.bi[i] >= hi) continue;
        if (at >= l.bi.size() || l.bi[at] != f.bi[i] || l.bj[at] != f.bj[i]) { fprintf(stderr, "%s: structure differs at tile %zu\n", what, at); return 1; }
        if (memcmp(&l.tiles[at * te], &f.tiles[i * te], te * sizeof(double)) != 0) { fprintf(stderr, "%s: tile (%ld,%ld) differs\n", what, (long)f.bi[i], (long)f.bj[i]); return 1; }
        ++at;
    }
    if (at != l.bi.size()) { fprintf(stderr, "%s: %zu extra tiles in the slab\n", what, l.bi.size() - at); return 1; }
    return 0;
}

static int rank_main(int rank, int world, const std::vector<unsigned char>& id) {
    const int n = 4096, b = 64, g = n / b;
    const double lam = 0.02, tau = 1e-6;
    hbsm::comm::init(id, rank, world);
    int lo = 0, hi = 0;
    hbsm::comm::rows_of(g, world, rank, lo, hi);
    int bad = 0;
    {
        Matrix A, B, Af, Bf;
        make_decay(A, n, b, lam, 1, false, lo, hi);
        make_decay(B, n, b, lam, 2, false, lo, hi);
        make_decay(Af, n, b, lam, 1, false, 0, -1);
        make_decay(Bf, n, b, lam, 2, false, 0, -1);
        B.publish();
        for (int mode = 0; mode < 3 && !bad; ++mode) {       // spamm NN, exact multiply NN, spamm N T
            const bool tB = mode == 2;
            Matrix Bt, Btf;
            Matrix* Bl = &B; Matrix* Bfull = &Bf;
            if (tB) {   // op(B) = B^T is sharded by ITS contraction index, the column of the stored matrix: use B^T's transpose slab
                Matrix::transpose(Bf, Btf); Btf.update_internal_info();       // stored matrix whose transpose is the operand
                // slab of Btf by COLUMN k in [lo,hi) = transpose of rows [lo,hi) of Bf
                Matrix::transpose(B, Bt); Bt.update_internal_info(); Bt.publish();
                Bl = &Bt; Bfull = &Btf;
            }
            Matrix C, Cf;
            size_t nm = 0, nr = 0, nmf = 0, nrf = 0;
            if (mode == 1) { Matrix::sharded_multiply(A, false, *Bl, tB, C, &nm, &nr); Matrix::multiply(Af, false, *Bfull, tB, Cf, &nmf, &nrf); }
            else { Matrix::sharded_spamm(A, false, *Bl, tB, C, tau, &nm, &nr); Matrix::spamm(Af, false, *Bfull, tB, Cf, tau, true, &nmf, &nrf); }
            const double tot = hbsm::comm::sum((double)nm), totb = hbsm::comm::sum((double)nr);
            if (tot != (double)nmf || totb != (double)nrf) { fprintf(stderr, "rank %d mode %d: %g products / %g tiles over the ranks, single GPU %zu / %zu\n", rank, mode, tot, totb, nmf, nrf); bad = 1; }
            if (!bad) bad = check_slab(C, Cf, b, lo, hi, mode == 0 ? "spamm NN" : mode == 1 ? "multiply NN" : "spamm NT");
        }
    }
    if (!bad) {   // symmetric square of a full-storage symmetric matrix, sharded, against symm_square of the upper triangle
        Matrix F, Ff, U, C, Cf;
        make_decay(F, n, b, lam, 3, true, lo, hi);
        make_decay(Ff, n, b, lam, 3, true, 0, -1);
        F.publish();
        Ff.get_upper_triangle(U); U.update_internal_info();
        size_t nm = 0, nr = 0, nmf = 0, nrf = 0;
        Matrix::sharded_symm_square(F, C, true, tau, &nm, &nr);
        Matrix::symm_square_spamm(U, Cf, tau, &nmf, &nrf);
        if (hbsm::comm::sum((double)nm) != (double)nmf) { fprintf(stderr, "rank %d: symm_square product count differs\n", rank); bad = 1; }
        if (!bad) bad = check_slab(C, Cf, b, lo, hi, "symm_square_spamm");
    }
    const double any_bad = hbsm::comm::max((double)bad);
    hbsm::comm::finalize();
    if (rank == 0 && any_bad == 0) printf("sharded c++ ok world=%d\n", world);
    return any_bad != 0 ? 1 : 0;
}

int main(int argc, char** argv) {
    const int world = argc > 1 ? atoi(argv[1]) : 2;
    if (world < 1 || world > 8) return 2;
    std::vector<int> rd(world, -1), wr(world, -1);
    for (int r = 1; r < world; ++r) {
        int fd[2];
        if (pipe(fd) != 0) return 2;
        rd[r] = fd[0]; wr[r] = fd[1];
    }
    std::vector<pid_t> kids;
    for (int r = 0; r < world; ++r) {
        pid_t pid = fork();          // before the first CUDA call of the process
        if (pid < 0) return 2;
        if (pid == 0) {
            int rc = 1;
            try {
                if (hbsm_init(r) != HBSM_OK) { fprintf(stderr, "rank %d: %s\n", r, hbsm_last_error()); _exit(77); }
                std::vector<unsigned char> id(HBSM_COMM_ID_BYTES);
                if (r == 0) {
                    id = hbsm::comm::unique_id();
                    for (int q = 1; q < world; ++q)
                        if (write(wr[q], id.data(), id.size()) != (ssize_t)id.size()) _exit(2);
                } else if (read(rd[r], id.data(), id.size()) != (ssize_t)id.size()) _exit(2);
                rc = rank_main(r, world, id);
            } catch (const std::exception& ex) {
                fprintf(stderr, "rank %d: exception: %s\n", r, ex.what());
                rc = 1;
            }
            fflush(stdout); fflush(stderr);
            _exit(rc);
        }
        kids.push_back(pid);
    }
    int worst = 0;
    for (size_t i = 0; i < kids.size(); ++i) {
        int st = 0;
        waitpid(kids[i], &st, 0);
        const int rc = WIFEXITED(st) ? WEXITSTATUS(st) : 1;
        if (rc == 77 && worst == 0) worst = 77;
        else if (rc != 0 && rc != 77) worst = 1;
    }
    return worst;
}
