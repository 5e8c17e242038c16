/*
 * oracle.c -- CPU restatement of the batched linear-algebra hot path of GPUtils.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked, imported or executed by the product
 * (libgputils_b200 / include/tensor.cuh). It is used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / reference legs as the checker and the CPU baseline, never as the thing measured as ours.
 *
 * Where the algorithm lives. The reference (include/tensor.cuh) contains no numerical kernel of its own:
 * the arithmetic of this path is in closed-source third-party libraries that are not under
 * /root/reference -- cuBLAS 12.9.1.4 (gemmBatched, gelsBatched, L1 routines) and cuSOLVER 11.7.5.82
 * (potrf/potrs[Batched], geqrf, ormqr, gesvd). The reference pins no version (bare `cublas cusolver` link
 * names, CMakeLists.txt:56-57). Their published contract is the BLAS / LAPACK one, so each function below
 * restates the corresponding unblocked LAPACK algorithm and cites the reference call site it stands for.
 *
 * Parity pinning: PINNED. tests/test_oracle_golden.py checks these routines against every golden vector the
 * reference's own tests hold for the path (tests/golden/reference_vectors.json, transcribed with file:line
 * from test/testTensor.cu), and tests/test_gpu_vs_reference.py compares the CUDA path with the reference
 * itself (oracle/_ref/libgputils_ref.so: the untouched reference header built against cuBLAS/cuSOLVER) on the
 * same inputs on the GPU box.
 *
 * Layout: column-major, leading dimension = rows, batch stride = rows*cols (ref: tensor.cuh:672-688, 1278-1284).
 * Every routine comes as _f64 and _f32 (fp32 arithmetic is plain float, no widening, like SGEMM etc.).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX(i, j, ld) ((size_t) (i) + (size_t) (j) * (size_t) (ld))

#define ORACLE_DEFINE(T, SUF, SQRT, FABS)                                                                          \
                                                                                                                   \
    /* C_i <- beta C_i + alpha A_i B_i   (ref: tensor.cuh:1286-1338, cublas?gemmBatched NN) */                     \
    void oracle_gemm_batched_##SUF(size_t m, size_t n, size_t k, T alpha, const T *A, const T *B, T beta, T *C,    \
                                   size_t batch) {                                                                 \
        _Pragma("omp parallel for schedule(static)") for (long long b = 0; b < (long long) batch; b++) {           \
            const T *a = A + (size_t) b * m * k, *bb = B + (size_t) b * k * n;                                     \
            T *c = C + (size_t) b * m * n;                                                                         \
            for (size_t j = 0; j < n; j++)                                                                         \
                for (size_t i = 0; i < m; i++) {                                                                   \
                    T acc = 0;                                                                                     \
                    for (size_t l = 0; l < k; l++) acc += a[IDX(i, l, m)] * bb[IDX(l, j, k)];                      \
                    c[IDX(i, j, m)] = (beta == (T) 0) ? alpha * acc : alpha * acc + beta * c[IDX(i, j, m)];        \
                }                                                                                                  \
        }                                                                                                          \
    }                                                                                                              \
                                                                                                                   \
    /* A_i = L_i L_i^T, lower, in place; strict upper triangle untouched; info = first bad pivot (1-based).      \
     * Unblocked right-looking potf2.  (ref: tensor.cuh:2135-2159 potrfBatched, 1742-1761 potrf) */                \
    void oracle_potrf_batched_##SUF(size_t n, T *A, int *info, size_t batch) {                                     \
        _Pragma("omp parallel for schedule(static)") for (long long b = 0; b < (long long) batch; b++) {           \
            T *a = A + (size_t) b * n * n;                                                                         \
            int bad = 0;                                                                                           \
            for (size_t j = 0; j < n && !bad; j++) {                                                               \
                T d = a[IDX(j, j, n)];                                                                             \
                if (!(d > (T) 0)) {                                                                                \
                    bad = (int) j + 1;                                                                             \
                    break;                                                                                         \
                }                                                                                                  \
                d = SQRT(d);                                                                                       \
                a[IDX(j, j, n)] = d;                                                                               \
                for (size_t i = j + 1; i < n; i++) a[IDX(i, j, n)] /= d;                                           \
                for (size_t c = j + 1; c < n; c++) {                                                               \
                    T lcj = a[IDX(c, j, n)];                                                                       \
                    for (size_t i = c; i < n; i++) a[IDX(i, c, n)] -= a[IDX(i, j, n)] * lcj;                       \
                }                                                                                                  \
            }                                                                                                      \
            info[b] = bad;                                                                                         \
        }                                                                                                          \
    }                                                                                                              \
                                                                                                                   \
    /* L_i L_i^T x = b_i in place, one rhs  (ref: tensor.cuh:2161-2197 potrsBatched, 1763-1783 potrs) */           \
    void oracle_potrs_batched_##SUF(size_t n, const T *L, T *B, size_t batch) {                                    \
        _Pragma("omp parallel for schedule(static)") for (long long b = 0; b < (long long) batch; b++) {           \
            const T *l = L + (size_t) b * n * n;                                                                   \
            T *x = B + (size_t) b * n;                                                                             \
            for (size_t j = 0; j < n; j++) {                                                                       \
                x[j] /= l[IDX(j, j, n)];                                                                           \
                for (size_t i = j + 1; i < n; i++) x[i] -= l[IDX(i, j, n)] * x[j];                                 \
            }                                                                                                      \
            for (size_t jj = n; jj-- > 0;) {                                                                       \
                T s = x[jj];                                                                                       \
                for (size_t i = jj + 1; i < n; i++) s -= l[IDX(i, jj, n)] * x[i];                                  \
                x[jj] = s / l[IDX(jj, jj, n)];                                                                     \
            }                                                                                                      \
        }                                                                                                          \
    }                                                                                                              \
                                                                                                                   \
    /* Householder reflector (larfg): returns tau, alpha <- beta, x <- v */                                        \
    static T oracle_larfg_##SUF(size_t n, T *alpha, T *x) {                                                        \
        if (n <= 1) return (T) 0;                                                                                  \
        T ss = 0;                                                                                                  \
        for (size_t i = 0; i + 1 < n; i++) ss += x[i] * x[i];                                                      \
        if (ss == (T) 0) return (T) 0;                                                                             \
        T nrm = SQRT((*alpha) * (*alpha) + ss);                                                                    \
        T beta = (*alpha >= (T) 0) ? -nrm : nrm;                                                                   \
        T tau = (beta - *alpha) / beta;                                                                            \
        T s = (T) 1 / (*alpha - beta);                                                                             \
        for (size_t i = 0; i + 1 < n; i++) x[i] *= s;                                                              \
        *alpha = beta;                                                                                             \
        return tau;                                                                                                \
    }                                                                                                              \
                                                                                                                   \
    /* C(rows x cols, ldc) <- (I - tau v v^T) C with v = (1, vtail) */                                             \
    static void oracle_apply_left_##SUF(size_t rows, size_t cols, const T *vtail, T tau, T *C, size_t ldc) {       \
        if (tau == (T) 0) return;                                                                                  \
        for (size_t c = 0; c < cols; c++) {                                                                        \
            T *cc = C + c * ldc;                                                                                   \
            T w = cc[0];                                                                                           \
            for (size_t r = 1; r < rows; r++) w += vtail[r - 1] * cc[r];                                           \
            w *= tau;                                                                                              \
            cc[0] -= w;                                                                                            \
            for (size_t r = 1; r < rows; r++) cc[r] -= w * vtail[r - 1];                                           \
        }                                                                                                          \
    }                                                                                                              \
                                                                                                                   \
    /* geqr2: A_i <- (R above, reflectors below), tau_i[n]  (ref: tensor.cuh:1866-1889 geqrf) */                   \
    void oracle_geqrf_batched_##SUF(size_t m, size_t n, T *A, T *tau, size_t batch) {                              \
        _Pragma("omp parallel for schedule(static)") for (long long b = 0; b < (long long) batch; b++) {           \
            T *a = A + (size_t) b * m * n, *t = tau + (size_t) b * n;                                              \
            for (size_t j = 0; j < n; j++) {                                                                       \
                t[j] = oracle_larfg_##SUF(m - j, &a[IDX(j, j, m)], &a[IDX(j + 1 < m ? j + 1 : j, j, m)]);          \
                if (j + 1 < n) oracle_apply_left_##SUF(m - j, n - j - 1, &a[IDX(j + 1 < m ? j + 1 : j, j, m)], t[j], &a[IDX(j, j + 1, m)], m); \
            }                                                                                                      \
        }                                                                                                          \
    }                                                                                                              \
                                                                                                                   \
    /* orm2r, side = left: C_i <- Q_i^T C_i (trans) or Q_i C_i  (ref: tensor.cuh:1896-1902, 1946-1952 ormqr) */    \
    void oracle_ormqr_batched_##SUF(int trans, size_t m, size_t ncols, size_t k, const T *A, const T *tau, T *C,   \
                                    size_t batch) {                                                                \
        _Pragma("omp parallel for schedule(static)") for (long long b = 0; b < (long long) batch; b++) {           \
            const T *a = A + (size_t) b * m * k, *t = tau + (size_t) b * k;                                        \
            T *c = C + (size_t) b * m * ncols;                                                                     \
            for (size_t jj = 0; jj < k; jj++) {                                                                    \
                size_t j = trans ? jj : k - 1 - jj;                                                                \
                oracle_apply_left_##SUF(m - j, ncols, &a[IDX(j + 1 < m ? j + 1 : j, j, m)], t[j], &c[j], m);       \
            }                                                                                                      \
        }                                                                                                          \
    }                                                                                                              \
                                                                                                                   \
    /* R_i x = b_i, upper, non-unit, in place (ref: tensor.cuh:1903-1907 trsm LEFT UPPER N NONUNIT, nrhs 1) */     \
    void oracle_trsv_upper_batched_##SUF(size_t n, const T *R, size_t ldr, size_t strideR, T *B, size_t strideB,   \
                                         size_t batch) {                                                           \
        _Pragma("omp parallel for schedule(static)") for (long long b = 0; b < (long long) batch; b++) {           \
            const T *r = R + (size_t) b * strideR;                                                                 \
            T *x = B + (size_t) b * strideB;                                                                       \
            for (size_t jj = n; jj-- > 0;) {                                                                       \
                x[jj] /= r[IDX(jj, jj, ldr)];                                                                      \
                for (size_t i = 0; i < jj; i++) x[i] -= r[IDX(i, jj, ldr)] * x[jj];                                \
            }                                                                                                      \
        }                                                                                                          \
    }                                                                                                              \
                                                                                                                   \
    /* gels, m >= n, one rhs: A_i <- QR factors, b_i[0:n] <- x, b_i[n:m] <- tail of Q^T b                         \
     * (ref: tensor.cuh:1340-1394 gelsBatched) */                                                                  \
    void oracle_gels_batched_##SUF(size_t m, size_t n, T *A, T *B, int *info, size_t batch) {                      \
        _Pragma("omp parallel for schedule(static)") for (long long b = 0; b < (long long) batch; b++) {           \
            T *a = A + (size_t) b * m * n, *x = B + (size_t) b * m;                                                \
            int bad = 0;                                                                                           \
            for (size_t j = 0; j < n; j++) {                                                                       \
                T *vt = &a[IDX(j + 1 < m ? j + 1 : j, j, m)];                                                      \
                T tau = oracle_larfg_##SUF(m - j, &a[IDX(j, j, m)], vt);                                           \
                if (j + 1 < n) oracle_apply_left_##SUF(m - j, n - j - 1, vt, tau, &a[IDX(j, j + 1, m)], m);        \
                oracle_apply_left_##SUF(m - j, 1, vt, tau, &x[j], m);                                              \
            }                                                                                                      \
            for (size_t jj = n; jj-- > 0;) {                                                                       \
                if (a[IDX(jj, jj, m)] == (T) 0 && !bad) bad = (int) jj + 1;                                        \
                x[jj] /= a[IDX(jj, jj, m)];                                                                        \
                for (size_t i = 0; i < jj; i++) x[i] -= a[IDX(i, jj, m)] * x[jj];                                  \
            }                                                                                                      \
            if (info) info[b] = bad;                                                                               \
        }                                                                                                          \
    }                                                                                                              \
                                                                                                                   \
    /* flat reductions (ref: tensor.cuh:968-1072 dot, nrm2, asum, iamax, iamin) */                                 \
    double oracle_dot_##SUF(size_t n, const T *x, const T *y) {                                                    \
        double s = 0;                                                                                              \
        for (size_t i = 0; i < n; i++) s += (double) x[i] * (double) y[i];                                         \
        return s;                                                                                                  \
    }                                                                                                              \
    double oracle_nrm2_##SUF(size_t n, const T *x) {                                                               \
        double s = 0;                                                                                              \
        for (size_t i = 0; i < n; i++) s += (double) x[i] * (double) x[i];                                         \
        return sqrt(s);                                                                                            \
    }                                                                                                              \
    double oracle_asum_##SUF(size_t n, const T *x) {                                                               \
        double s = 0;                                                                                              \
        for (size_t i = 0; i < n; i++) s += fabs((double) x[i]);                                                   \
        return s;                                                                                                  \
    }                                                                                                              \
    /* out-of-place batched transpose (ref: tensor.cuh:1167-1197 geam T) */                                        \
    void oracle_transpose_batched_##SUF(size_t m, size_t n, const T *A, T *At, size_t batch) {                     \
        for (size_t b = 0; b < batch; b++)                                                                         \
            for (size_t j = 0; j < n; j++)                                                                         \
                for (size_t i = 0; i < m; i++) At[b * m * n + IDX(j, i, n)] = A[b * m * n + IDX(i, j, m)];         \
    }

ORACLE_DEFINE(double, f64, sqrt, fabs)
ORACLE_DEFINE(float, f32, sqrtf, fabsf)

int oracle_num_threads(void) {
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
#endif
    return n;
}
