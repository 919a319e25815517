/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Fortran-BLAS entry points the reference's hot path reaches through source/gblas.h
 * (gemm G:85-98, symm G:99-112, axpy G:130-143): dgemm_/sgemm_/dsymm_/ssymm_/daxpy_/saxpy_.
 * The reference links "-lblas -llapack" (Makefile.include:3), an un-vendored, un-pinned system
 * library that does not exist in this image.  This shim provides those symbols to oracle/_ref:
 * at first use it dlopen()s an LP64 OpenBLAS if one is found (HBSM_REF_BLAS=<path>, else the
 * copy bundled with opencv in the venv) and forwards; otherwise it falls back to the textbook
 * loops below (column-major, Fortran argument conventions).  ref_blas_kind() reports which.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <glob.h>

typedef void (*dgemm_fn)(const char*, const char*, const int*, const int*, const int*, const double*,
                         const double*, const int*, const double*, const int*, const double*, double*, const int*);
typedef void (*sgemm_fn)(const char*, const char*, const int*, const int*, const int*, const float*,
                         const float*, const int*, const float*, const int*, const float*, float*, const int*);

static dgemm_fn p_dgemm = NULL;
static sgemm_fn p_sgemm = NULL;
static int g_state = 0; /* 0 = unresolved, 1 = openblas, 2 = builtin */
static char g_path[1024] = "builtin";

/* wheels bundle libgfortran/libquadmath beside OpenBLAS without an rpath: preload them from the same directory */
static void preload_siblings(const char* path) {
    char dir[1024];
    strncpy(dir, path, sizeof(dir) - 1); dir[sizeof(dir) - 1] = 0;
    char* slash = strrchr(dir, '/');
    if (!slash) return;
    *slash = 0;
    const char* names[] = {"libquadmath*.so*", "libgfortran*.so*", NULL};
    for (int i = 0; names[i]; ++i) {
        char pat[1200];
        snprintf(pat, sizeof(pat), "%s/%s", dir, names[i]);
        glob_t g;
        if (glob(pat, 0, NULL, &g) == 0) {
            for (size_t j = 0; j < g.gl_pathc; ++j) dlopen(g.gl_pathv[j], RTLD_NOW | RTLD_GLOBAL);
            globfree(&g);
        }
    }
}

static void try_open(const char* path) {
    preload_siblings(path);
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
    dgemm_fn d = (dgemm_fn)dlsym(h, "dgemm_");
    sgemm_fn s = (sgemm_fn)dlsym(h, "sgemm_");
    if (d && s) {
        p_dgemm = d; p_sgemm = s; g_state = 1;
        strncpy(g_path, path, sizeof(g_path) - 1);
    }
}

static void resolve(void) {
    if (g_state) return;
    if (!getenv("OPENBLAS_NUM_THREADS")) setenv("OPENBLAS_NUM_THREADS", "1", 1);
    const char* env = getenv("HBSM_REF_BLAS");
    if (env && strcmp(env, "builtin") == 0) { g_state = 2; return; }
    if (env) try_open(env);
    if (!g_state) {
        glob_t g;
        const char* pats[] = {
            "/opt/prime-rl/.venv/lib/python3*/site-packages/opencv_python_headless.libs/libopenblas*.so*",
            "/usr/lib/x86_64-linux-gnu/libopenblas.so*", "/usr/lib/x86_64-linux-gnu/libblas.so*", NULL};
        for (int i = 0; pats[i] && !g_state; ++i) {
            if (glob(pats[i], 0, NULL, &g) == 0) {
                for (size_t j = 0; j < g.gl_pathc && !g_state; ++j) try_open(g.gl_pathv[j]);
                globfree(&g);
            }
        }
    }
    if (!g_state) g_state = 2;
}

const char* ref_blas_kind(void) { resolve(); return g_state == 1 ? g_path : "builtin"; }

#define GEMM_BODY(T)                                                                              \
    const int M = *m, N = *n, K = *k, LDA = *lda, LDB = *ldb, LDC = *ldc;                         \
    const int ta = (*transa == 'T' || *transa == 't'), tb = (*transb == 'T' || *transb == 't');    \
    for (int j = 0; j < N; ++j) {                                                                 \
        for (int i = 0; i < M; ++i) C[i + (size_t)j * LDC] *= *beta;                              \
        for (int l = 0; l < K; ++l) {                                                             \
            const T bv = *alpha * (tb ? B[j + (size_t)l * LDB] : B[l + (size_t)j * LDB]);         \
            if (!ta) for (int i = 0; i < M; ++i) C[i + (size_t)j * LDC] += A[i + (size_t)l * LDA] * bv; \
            else     for (int i = 0; i < M; ++i) C[i + (size_t)j * LDC] += A[l + (size_t)i * LDA] * bv; \
        }                                                                                         \
    }

void dgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k,
            const double* alpha, const double* A, const int* lda, const double* B, const int* ldb,
            const double* beta, double* C, const int* ldc) {
    resolve();
    if (p_dgemm) { p_dgemm(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc); return; }
    GEMM_BODY(double)
}

void sgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k,
            const float* alpha, const float* A, const int* lda, const float* B, const int* ldb,
            const float* beta, float* C, const int* ldc) {
    resolve();
    if (p_sgemm) { p_sgemm(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc); return; }
    GEMM_BODY(float)
}

/* C = alpha*sym(A)*B + beta*C (side L) or alpha*B*sym(A) + beta*C (side R); only `uplo` triangle of A read */
#define SYMM_BODY(T)                                                                              \
    const int M = *m, N = *n, LDA = *lda, LDB = *ldb, LDC = *ldc;                                 \
    const int left = (*side == 'L' || *side == 'l'), up = (*uplo == 'U' || *uplo == 'u');         \
    for (int j = 0; j < N; ++j)                                                                   \
        for (int i = 0; i < M; ++i) {                                                             \
            T acc = 0;                                                                            \
            if (left) {                                                                           \
                for (int l = 0; l < M; ++l) {                                                     \
                    int r = i, c = l;                                                             \
                    if ((up && r > c) || (!up && r < c)) { int t = r; r = c; c = t; }             \
                    acc += A[r + (size_t)c * LDA] * B[l + (size_t)j * LDB];                       \
                }                                                                                 \
            } else {                                                                              \
                for (int l = 0; l < N; ++l) {                                                     \
                    int r = l, c = j;                                                             \
                    if ((up && r > c) || (!up && r < c)) { int t = r; r = c; c = t; }             \
                    acc += B[i + (size_t)l * LDB] * A[r + (size_t)c * LDA];                       \
                }                                                                                 \
            }                                                                                     \
            C[i + (size_t)j * LDC] = *alpha * acc + *beta * C[i + (size_t)j * LDC];               \
        }

void dsymm_(const char* side, const char* uplo, const int* m, const int* n, const double* alpha,
            const double* A, const int* lda, const double* B, const int* ldb, const double* beta,
            double* C, const int* ldc) { SYMM_BODY(double) }
void ssymm_(const char* side, const char* uplo, const int* m, const int* n, const float* alpha,
            const float* A, const int* lda, const float* B, const int* ldb, const float* beta,
            float* C, const int* ldc) { SYMM_BODY(float) }

void daxpy_(const int* n, const double* da, const double* dx, const int* incx, double* dy, const int* incy) {
    for (int i = 0; i < *n; ++i) dy[(size_t)i * *incy] += *da * dx[(size_t)i * *incx];
}
void saxpy_(const int* n, const float* da, const float* dx, const int* incx, float* dy, const int* incy) {
    for (int i = 0; i < *n; ++i) dy[(size_t)i * *incy] += *da * dx[(size_t)i * *incx];
}
