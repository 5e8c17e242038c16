/* cpu_lapack.c -- host-CPU baseline of the batched path (SURVEY.md 8d "Baseline 2").
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (like everything under oracle/): bench.py's cpu_baseline leg and tests/ may load it;
 * the product never does.
 *
 * GPUtils has no CPU path of its own, so the host baseline SURVEY.md prescribes is a loop over the matrices of the batch
 * calling single-threaded LAPACK / BLAS -- the OpenBLAS that ships with scipy (symbols prefixed scipy_, LP64), found with
 * dlopen at run time -- under `omp parallel for`, one matrix per iteration. Layout is the reference's: column-major
 * matrices, mats axis slowest (tensor.cuh:1278-1284). Every routine works in place like the cuSOLVER call it mirrors and
 * returns the wall-clock seconds of the loop (negative on error).
 *
 * build: gcc -O2 -fopenmp -shared -fPIC cpu_lapack.c -o _build/libcpu_lapack.so -ldl     (oracle/Makefile: cpulapack)
 */
#include <dlfcn.h>
#include <omp.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef void (*potrf_fn)(const char *, const int *, void *, const int *, int *);
typedef void (*potrs_fn)(const char *, const int *, const int *, const void *, const int *, void *, const int *, int *);
typedef void (*gemm_fn)(const char *, const char *, const int *, const int *, const int *, const void *, const void *, const int *,
                        const void *, const int *, const void *, void *, const int *);
typedef void (*gels_fn)(const char *, const int *, const int *, const int *, void *, const int *, void *, const int *, void *,
                        const int *, int *);
typedef void (*geqrf_fn)(const int *, const int *, void *, const int *, void *, void *, const int *, int *);
typedef void (*gesvd_fn)(const char *, const char *, const int *, const int *, void *, const int *, void *, void *, const int *,
                         void *, const int *, void *, const int *, int *);

static void *g_lib = NULL;
static potrf_fn dpotrf, spotrf;
static potrs_fn dpotrs, spotrs;
static gemm_fn dgemm, sgemm;
static gels_fn dgels, sgels;
static geqrf_fn dgeqrf, sgeqrf;
static gesvd_fn dgesvd, sgesvd;

/* path = the libscipy_openblas*.so found by the caller; 0 on success */
int cpu_lapack_open(const char *path) {
    if (g_lib) return 0;
    g_lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!g_lib) return -1;
    void (*set_threads)(int) = (void (*)(int)) dlsym(g_lib, "scipy_openblas_set_num_threads");
    if (set_threads) set_threads(1); /* one LAPACK thread per matrix; the parallelism is over the batch */
#define SYM(var, name) do { *(void **) (&var) = dlsym(g_lib, name); if (!var) return -2; } while (0)
    SYM(dpotrf, "scipy_dpotrf_"); SYM(spotrf, "scipy_spotrf_");
    SYM(dpotrs, "scipy_dpotrs_"); SYM(spotrs, "scipy_spotrs_");
    SYM(dgemm, "scipy_dgemm_");   SYM(sgemm, "scipy_sgemm_");
    SYM(dgels, "scipy_dgels_");   SYM(sgels, "scipy_sgels_");
    SYM(dgeqrf, "scipy_dgeqrf_"); SYM(sgeqrf, "scipy_sgeqrf_");
    SYM(dgesvd, "scipy_dgesvd_"); SYM(sgesvd, "scipy_sgesvd_");
#undef SYM
    return 0;
}

int cpu_lapack_threads(void) { return omp_get_max_threads(); }
void cpu_lapack_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* CholeskyBatchFactoriser::factorise + solve (tensor.cuh:2135-2197): potrf('L') then potrs, one rhs */
double cpu_chol_batch(int is_f64, int n, size_t batch, void *A, void *b, int *info) {
    if (!g_lib) return -1.0;
    const size_t es = is_f64 ? 8 : 4;
    const int one = 1;
    const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < batch; i++) {
        char *Ai = (char *) A + i * (size_t) n * n * es;
        int inf = 0;
        if (is_f64) dpotrf("L", &n, Ai, &n, &inf); else spotrf("L", &n, Ai, &n, &inf);
        if (info) info[i] = inf;
        if (b && inf == 0) {
            char *bi = (char *) b + i * (size_t) n * es;
            int inf2 = 0;
            if (is_f64) dpotrs("L", &n, &one, Ai, &n, bi, &n, &inf2); else spotrs("L", &n, &one, Ai, &n, bi, &n, &inf2);
        }
    }
    return omp_get_wtime() - t0;
}

/* DTensor::addAB (tensor.cuh:1286-1338): C_i = alpha A_i B_i + beta C_i, NN */
double cpu_gemm_batch(int is_f64, int m, int n, int k, size_t batch, const void *A, const void *B, void *C, double alpha, double beta) {
    if (!g_lib) return -1.0;
    const size_t es = is_f64 ? 8 : 4;
    const float af = (float) alpha, bf = (float) beta;
    const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < batch; i++) {
        const char *Ai = (const char *) A + i * (size_t) m * k * es, *Bi = (const char *) B + i * (size_t) k * n * es;
        char *Ci = (char *) C + i * (size_t) m * n * es;
        if (is_f64) dgemm("N", "N", &m, &n, &k, &alpha, Ai, &m, Bi, &k, &beta, Ci, &m);
        else sgemm("N", "N", &m, &n, &k, &af, Ai, &m, Bi, &k, &bf, Ci, &m);
    }
    return omp_get_wtime() - t0;
}

/* DTensor::leastSquaresBatched (tensor.cuh:1340-1394): gels('N'), one rhs, in place */
double cpu_gels_batch(int is_f64, int m, int n, size_t batch, void *A, void *b) {
    if (!g_lib) return -1.0;
    const size_t es = is_f64 ? 8 : 4;
    const int one = 1, lwork = 64 * (m + n) + 64;
    int failed = 0;
    const double t0 = omp_get_wtime();
#pragma omp parallel
    {
        void *work = malloc((size_t) lwork * es);
#pragma omp for schedule(static)
        for (size_t i = 0; i < batch; i++) {
            char *Ai = (char *) A + i * (size_t) m * n * es, *bi = (char *) b + i * (size_t) m * es;
            int inf = 0;
            if (is_f64) dgels("N", &m, &n, &one, Ai, &m, bi, &m, work, &lwork, &inf);
            else sgels("N", &m, &n, &one, Ai, &m, bi, &m, work, &lwork, &inf);
            if (inf) failed = 1;
        }
        free(work);
    }
    return failed ? -2.0 : omp_get_wtime() - t0;
}

/* QRFactoriser::factorise (tensor.cuh:1866-1889): geqrf; tau is n per matrix */
double cpu_geqrf_batch(int is_f64, int m, int n, size_t batch, void *A, void *tau) {
    if (!g_lib) return -1.0;
    const size_t es = is_f64 ? 8 : 4;
    const int lwork = 64 * n + 64;
    const double t0 = omp_get_wtime();
#pragma omp parallel
    {
        void *work = malloc((size_t) lwork * es);
#pragma omp for schedule(static)
        for (size_t i = 0; i < batch; i++) {
            int inf = 0;
            if (is_f64) dgeqrf(&m, &n, (char *) A + i * (size_t) m * n * es, &m, (char *) tau + i * (size_t) n * es, work, &lwork, &inf);
            else sgeqrf(&m, &n, (char *) A + i * (size_t) m * n * es, &m, (char *) tau + i * (size_t) n * es, work, &lwork, &inf);
        }
        free(work);
    }
    return omp_get_wtime() - t0;
}

/* Svd::factorise (tensor.cuh:1624-1676): gesvd, jobu = 'A' or 'N', jobvt = 'A'; U may be NULL */
double cpu_gesvd_batch(int is_f64, int m, int n, size_t batch, void *A, void *S, void *U, void *Vt) {
    if (!g_lib) return -1.0;
    const size_t es = is_f64 ? 8 : 4;
    const int mn = m < n ? m : n, mx = m > n ? m : n;
    const int lwork = 8 * (3 * mn + mx) + 64 * (m + n);
    const char *jobu = U ? "A" : "N";
    const double t0 = omp_get_wtime();
#pragma omp parallel
    {
        void *work = malloc((size_t) lwork * es);
#pragma omp for schedule(static)
        for (size_t i = 0; i < batch; i++) {
            int inf = 0;
            char *Ui = U ? (char *) U + i * (size_t) m * m * es : NULL;
            if (is_f64) dgesvd(jobu, "A", &m, &n, (char *) A + i * (size_t) m * n * es, &m, (char *) S + i * (size_t) mn * es, Ui, &m,
                               (char *) Vt + i * (size_t) n * n * es, &n, work, &lwork, &inf);
            else sgesvd(jobu, "A", &m, &n, (char *) A + i * (size_t) m * n * es, &m, (char *) S + i * (size_t) mn * es, Ui, &m,
                        (char *) Vt + i * (size_t) n * n * es, &n, work, &lwork, &inf);
        }
        free(work);
    }
    return omp_get_wtime() - t0;
}
