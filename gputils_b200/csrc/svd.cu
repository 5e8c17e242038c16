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

int gpub_internal_gemm_plus_e_f64(gpub_ctx_t ctx, int sidx, size_t m, size_t n, size_t k, double alpha, const double *A, size_t lda, size_t sA,
                                  const double *B, size_t ldb, size_t sB, const double *Ublk, size_t ne, size_t sUblk, double *C, size_t ldc,
                                  size_t sC, size_t batch);   // gemm.cu

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

// ------------------------------------------------------------------------------------------
// U assembly with ONE block reflector (fp64, m and n multiples of 64): Q = H_0 ... H_{n-1} = I - V T V' with the explicit
// unit-lower-trapezoidal V (m x n) and the n x n compact-WY factor T of ALL n reflectors, so
//     U = Q E = E - V (T (V' E)),   E = blockdiag(Ur, I_{m-n}),   V' E = [ V(0:n, :)' Ur | V(n:m, :)' ]
// is four plain batched GEMMs on the FP64 tensor pipe (V'V for T, V'(0:n) Ur, T X, V Y) instead of n / 16 panel sweeps of
// ormqr over an m x m matrix: 0.34 GFLOP per 1024 x 128 matrix instead of 0.54, all of it GEMM-shaped.
// k_make_v: geqrf output -> explicit V in place (the SVD owns A and R has been consumed by then).
// k_tfactor: T from G = V'V and tau by LAPACK's dlarft recurrence, T(0:j, j) = -tau_j T(0:j, 0:j) G(0:j, j), T(j, j) = tau_j
// (tau_j = 0 gives a zero column, as in dlarft); one CTA per matrix, T built in shared memory, zeros below the diagonal.
// ------------------------------------------------------------------------------------------
template<typename T>
__global__ void k_make_v(int n, T *A, size_t lda, size_t sA, size_t batch) {
    const size_t nn = (size_t) n * n;
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < nn * batch; e += (size_t) gridDim.x * blockDim.x) {
        const size_t b = e / nn, r = e - b * nn;
        const int i = (int) (r % n), j = (int) (r / n);
        if (i <= j) A[b * sA + i + (size_t) j * lda] = i == j ? T(1) : T(0);
    }
}

// Blocked: the 16 x 16 diagonal blocks of T come from the recurrence (one warp per block, lane = row: no cross-lane dependency), then
// neighbouring blocks are merged level by level, T = [T11 T12; 0 T22] with T12 = -T11 (V1'V2) T22 = -T11 G12 T22 (the block form of the
// same recurrence): log2(n / 16) levels of small products in shared memory instead of n sequential steps with a CTA barrier each
// (n = 128: 0.49 -> 0.0x ms for 256 matrices). Everything happens in ONE n x n shared-memory matrix that starts as G and ends as T:
// a level only overwrites its own off-diagonal blocks, which still hold G until then. n must be 16 * 2^l.
template<typename T>
__global__ void __launch_bounds__(256) k_tfactor(int n, T *G, size_t sG, const T *__restrict__ tau, size_t sTau, size_t batch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *M = reinterpret_cast<T *>(smem_raw);              // [n][n + 1]: entry (i, j) at j * (n + 1) + i
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, ld = n + 1;
    T *W = M + (size_t) n * ld;                           // [n / 2][n / 2 + 1] scratch of a merge (its largest block is n / 2 x n / 2)
    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        T *g = G + mat * sG;
        const T *tg = tau + mat * sTau;
        for (int e = tid; e < n * n; e += 256) {
            const int i = e % n, j = e / n;
            if (i <= j) M[(size_t) j * ld + i] = g[e];
        }
        __syncthreads();
        // diagonal blocks: T(0:j, j) = -tau_j T(0:j, 0:j) G(0:j, j), T(j, j) = tau_j inside each block of 16
        for (int blk = warp; blk < n / 16; blk += 8) {
            const int o = 16 * blk;
            T trow[16];
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const T tj = tg[o + j];
                    T acc = 0;
#pragma unroll
                    for (int k = 0; k < j; k++)
                        if (k >= lane) acc = fma(trow[k], M[(size_t) (o + j) * ld + o + k], acc);
                    trow[j] = j == lane ? tj : (j > lane ? -tj * acc : T(0));
                }
            }
            __syncwarp();                                   // every lane has read the G entries the rows are about to replace
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (j >= lane) M[(size_t) (o + j) * ld + o + lane] = trow[j];
            }
        }
        __syncthreads();
        for (int b = 16; b < n; b *= 2) {
            const int nm = n / (2 * b), ldw = b + 1;
            // W = G12 T22 for every merge of the level: W(i, j) = sum_{k <= j} G12(i, k) T22(k, j)
            for (int e = tid; e < nm * b * b; e += 256) {
                const int mg = e / (b * b), r = e - mg * b * b, i = r % b, j = r / b;
                const int o = 2 * b * mg;
                const T *g12 = M + (size_t) (o + b) * ld + o + i;        // row i of G12: entry k at + k * ld
                const T *t22 = M + (size_t) (o + b + j) * ld + o + b;    // column j of T22
                T a0 = 0, a1 = 0;
                int k = 0;
                for (; k + 1 <= j; k += 2) {
                    a0 = fma(g12[(size_t) k * ld], t22[k], a0);
                    a1 = fma(g12[(size_t) (k + 1) * ld], t22[k + 1], a1);
                }
                if (k <= j) a0 = fma(g12[(size_t) k * ld], t22[k], a0);
                W[(size_t) mg * b * ldw + (size_t) j * ldw + i] = a0 + a1;
            }
            __syncthreads();
            // T12 = -T11 W: T12(i, j) = -sum_{k >= i} T11(i, k) W(k, j)
            for (int e = tid; e < nm * b * b; e += 256) {
                const int mg = e / (b * b), r = e - mg * b * b, i = r % b, j = r / b;
                const int o = 2 * b * mg;
                const T *t11 = M + (size_t) o * ld + o + i;              // row i of T11: entry k at + k * ld
                const T *w = W + (size_t) mg * b * ldw + (size_t) j * ldw;
                T a0 = 0, a1 = 0;
                int k = i;
                for (; k + 1 < b; k += 2) {
                    a0 = fma(t11[(size_t) k * ld], w[k], a0);
                    a1 = fma(t11[(size_t) (k + 1) * ld], w[k + 1], a1);
                }
                if (k < b) a0 = fma(t11[(size_t) k * ld], w[k], a0);
                M[(size_t) (o + b + j) * ld + o + i] = -(a0 + a1);
            }
            __syncthreads();
        }
        for (int e = tid; e < n * n; e += 256) {
            const int i = e % n, j = e / n;
            g[e] = i <= j ? M[(size_t) j * ld + i] : T(0);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// One-sided Jacobi SVD of R (n x n upper triangular, the geqrf output), one CTA per matrix.
// Works on X = R^T (lower triangular: the better-conditioned choice after a QR step) held in shared memory,
// column p of X = row p of R. Right rotations orthogonalise the columns: X J = W D, so
//     R = J D W^T  =>  singular values D, Vt rows = normalised columns of X J, Ur = J (accumulated only when
// U is wanted; J lives in global/L2 memory). Pairs follow the round-robin (circle) ordering: n/2 disjoint
// pairs per round, one warp per pair, a CTA barrier per round, sweeps until no pair rotates.
// Columns with a negligible norm (rank deficiency) are replaced by an orthonormal completion so Vt is always
// an orthogonal matrix, which Nullspace relies on.
// ------------------------------------------------------------------------------------------
#ifndef GPUB_JACOBI_THREADS
#define GPUB_JACOBI_THREADS 512
#endif
#ifndef GPUB_JACOBI_PAIRS
#define GPUB_JACOBI_PAIRS 4
#endif
constexpr int JT = GPUB_JACOBI_THREADS; // threads per CTA
constexpr int JP = GPUB_JACOBI_PAIRS;   // pairs a warp rotates at once (4 or 2)
constexpr int JRP = JP == 4 ? 16 : 8;   // values in the transpose-reduce (3 per pair, padded to a power of two)

template<typename T> struct JacobiEps;
// floor: a pair whose product of squared norms (of the matrix scaled to order one) is below it is not rotated: the product -- and with it
// the relative test c^2 > tol^2 a b -- underflows there, and columns that small only span the null space (replaced by the orthonormal completion)
template<> struct JacobiEps<double> { static constexpr double v = 1.1102230246251565e-16; static constexpr double big = 1e100; static constexpr double huge = 1.7e308; static constexpr double floor = 1e-280; };
template<> struct JacobiEps<float> { static constexpr float v = 5.9604645e-08f; static constexpr float big = 1e15f; static constexpr float huge = 3.4e38f; static constexpr float floor = 1e-30f; };

// reciprocal square root / reciprocal from the MUFU seed plus Newton steps (~1 ulp, no slow-path call)
template<typename T> __device__ __forceinline__ T jac_rsqrt(T x);
template<> __device__ __forceinline__ double jac_rsqrt<double>(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    y = fma(y, fma(-h, y * y, 0.5), y);
    y = fma(y, fma(-h, y * y, 0.5), y);
    return y;
}
template<> __device__ __forceinline__ float jac_rsqrt<float>(float x) {
    const float y = rsqrtf(x);
    return fmaf(0.5f * y, fmaf(-x * y, y, 1.0f), y);
}
template<typename T> __device__ __forceinline__ T jac_rcp(T x);
template<> __device__ __forceinline__ double jac_rcp<double>(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(fma(-x, y, 1.0), y, y);
    y = fma(fma(-x, y, 1.0), y, y);
    return fma(fma(-x, y, 1.0), y, y);
}
template<> __device__ __forceinline__ float jac_rcp<float>(float x) { return __frcp_rn(x); }
// two Newton steps: relative error ~1e-12 (fp64), for quantities that do not need the last bits
template<typename T> __device__ __forceinline__ T jac_rcp2(T x);
template<> __device__ __forceinline__ double jac_rcp2<double>(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(fma(-x, y, 1.0), y, y);
    return fma(fma(-x, y, 1.0), y, y);
}
template<> __device__ __forceinline__ float jac_rcp2<float>(float x) { return __frcp_rn(x); }

template<typename T>
__device__ __forceinline__ T jwarp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// X = (R / 2^e)' into shared memory, 2^e the power of two that brings the largest entry of R into [1/2, 1): the rotations work with
// squared column norms and their products (fourth powers of the entries), which leave the exponent range long before the entries do
// (LAPACK's gesvd scales for the same reason). The scaling is exact; the singular values are multiplied back by 2^e at the end.
// Returns 2^e. s_red: NT / 32 values of scratch. Ends with a CTA barrier.
template<typename T, int NT>
__device__ __forceinline__ T jacobi_load_scaled(int n, int ldx, const T *__restrict__ a_g, size_t lda, T *X, T *s_red) {
    const int tid = threadIdx.x;
    T amax = T(0);
    for (int e = tid; e < n * n; e += NT) {
        const int p = e / n, r = e % n;                  // X(r, p) = R(p, r)
        const T v = (p <= r) ? a_g[p + (size_t) r * lda] : T(0);
        X[(size_t) p * ldx + r] = v;
        const T av = fabs(v);
        amax = (av > amax && av <= JacobiEps<T>::huge) ? av : amax;   // (NaN / inf entries do not take part)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T other = __shfl_xor_sync(0xffffffffu, amax, o);
        amax = other > amax ? other : amax;
    }
    if ((tid & 31) == 0) s_red[tid >> 5] = amax;
    __syncthreads();
    amax = s_red[0];
    for (int w = 1; w < NT / 32; w++) amax = s_red[w] > amax ? s_red[w] : amax;
    int ex = 0;
    if (amax > T(0)) (void) frexp(amax, &ex);
    const T down = ldexp(T(1), -ex), up = ldexp(T(1), ex);
    if (ex != 0) {
        for (int e = tid; e < n * n; e += NT) {
            const int p = e / n, r = e % n;
            if (p <= r) X[(size_t) p * ldx + r] *= down;
        }
    }
    __syncthreads();
    return up;
}

// Tail of the Jacobi kernels: X (shared memory, column p at X + p * ldx) holds the orthogonalised columns W D. Singular values = column
// norms, sorted descending; Vt rows = normalised columns; columns without a singular value (rank deficiency) are replaced by an
// orthonormal completion so that Vt is always orthogonal (Nullspace relies on it). NT = threads of the CTA.
template<typename T, int NT>
__device__ __forceinline__ void jacobi_finish(int n, int ldx, T *X, T *s_sig, T *s_coef, int *s_perm, T *s_val, int *s_idx, T *s_g, T *vt_g,
                                              size_t ldvt, T *J, int want_u, T unscale) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = NT / 32;
    constexpr int JT = NT;
    // column norms
    for (int p = warp; p < n; p += NW) {
        T aa = 0;
        for (int r = lane; r < n; r += 32) aa = fma(X[(size_t) p * ldx + r], X[(size_t) p * ldx + r], aa);
        aa = jwarp_sum(aa);
        if (lane == 0) s_sig[p] = sqrt(aa);
    }
    __syncthreads();
    // rank sort, descending (ties broken by column index)
    // (a NaN norm -- NaN in the input -- sorts first, like +inf: the ranks must stay a permutation, s_perm indexes shared memory)
    for (int p = tid; p < n; p += JT) {
        const T sp = s_sig[p] == s_sig[p] ? s_sig[p] : (T) JacobiEps<T>::huge;
        int rank = 0;
        for (int q = 0; q < n; q++) {
            const T sq = s_sig[q] == s_sig[q] ? s_sig[q] : (T) JacobiEps<T>::huge;
            rank += (sq > sp || (sq == sp && q < p)) ? 1 : 0;
        }
        s_perm[rank] = p;
    }
    __syncthreads();
    for (int i = tid; i < n; i += JT) s_g[i] = s_sig[s_perm[i]] * unscale;
    const T smax = s_sig[s_perm[0]];
    const T thr = smax * (T) n * (T) JacobiEps<T>::v;
    // normalise the columns that carry a singular value; count them
    int nfull = 0;
    for (int i = 0; i < n; i++) nfull += (s_sig[s_perm[i]] > thr) ? 1 : 0;   // sorted: the first nfull positions
    for (int i = warp; i < nfull; i += NW) {
        T *xp = X + (size_t) s_perm[i] * ldx;
        const T inv = T(1) / s_sig[s_perm[i]];
        for (int r = lane; r < n; r += 32) xp[r] *= inv;
    }
    __syncthreads();
    // orthonormal completion of the null columns, one at a time
    for (int i = nfull; i < n; i++) {
        T *xz = X + (size_t) s_perm[i] * ldx;
        // pick the unit vector e_k with the largest component outside span(final columns): 1 - sum_w w[k]^2
        T best = T(-1);
        int bestk = 0;
        for (int k = tid; k < n; k += JT) {
            T acc = T(1);
            for (int w = 0; w < i; w++) {
                const T x = X[(size_t) s_perm[w] * ldx + k];
                acc = fma(-x, x, acc);
            }
            if (acc > best) { best = acc; bestk = k; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const T ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
            if (ob > best || (ob == best && ok < bestk)) { best = ob; bestk = ok; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bestk; }
        __syncthreads();
        best = s_val[0]; bestk = s_idx[0];
        for (int w = 1; w < NW; w++)
            if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bestk)) { best = s_val[w]; bestk = s_idx[w]; }
        __syncthreads();
        for (int r = tid; r < n; r += JT) xz[r] = (r == bestk) ? T(1) : T(0);
        __syncthreads();
        for (int pass = 0; pass < 2; pass++) {     // project out the final columns, twice
            // coefficients g_w = w . z, one final column per warp iteration
            for (int w = warp; w < i; w += NW) {
                const T *xw = X + (size_t) s_perm[w] * ldx;
                T g = 0;
                for (int r = lane; r < n; r += 32) g = fma(xw[r], xz[r], g);
                g = jwarp_sum(g);
                if (lane == 0) s_coef[w] = g;
            }
            __syncthreads();
            for (int r = tid; r < n; r += JT) {
                T acc = xz[r];
                for (int w = 0; w < i; w++) acc = fma(-s_coef[w], X[(size_t) s_perm[w] * ldx + r], acc);
                xz[r] = acc;
            }
            __syncthreads();
        }
        T nn = 0;
        for (int r = tid; r < n; r += JT) nn = fma(xz[r], xz[r], nn);
        nn = jwarp_sum(nn);
        if (lane == 0) s_val[warp] = nn;
        __syncthreads();
        T tot = 0;
        for (int w = 0; w < NW; w++) tot += s_val[w];
        __syncthreads();
        const T inv = T(1) / sqrt(tot);
        for (int r = tid; r < n; r += JT) xz[r] *= inv;
        __syncthreads();
    }
    for (int e = tid; e < n * n; e += JT) {
        const int i = e % n, c = e / n;             // Vt(i, c) = W(c, perm i)
        vt_g[i + (size_t) c * ldvt] = X[(size_t) s_perm[i] * ldx + c];
    }
    if (want_u) {
        // Ur(:, i) = J(:, perm i): permute the columns in place through shared memory (X is free now)
        __syncthreads();
        for (int e = tid; e < n * n; e += JT) {
            const int r = e % n, i = e / n;
            X[(size_t) i * ldx + r] = J[(size_t) s_perm[i] * n + r];
        }
        __syncthreads();
        for (int e = tid; e < n * n; e += JT) {
            const int r = e % n, i = e / n;
            J[(size_t) i * n + r] = X[(size_t) i * ldx + r];
        }
    }
}

// JE: rows per lane (n <= 32 * JE). FULL: n == 32 * JE (even, no bye, every lane owns JE rows of every column): the row / pair guards
// of the round loop are compile-time true (BASELINE config 4: n = 128, JE = 4)
template<typename T, int JE, bool FULL>
__global__ void __launch_bounds__(JT) k_jacobi_rt(int n, const T *__restrict__ A, size_t lda, size_t sA, T *S, size_t sS, T *Vt, size_t ldvt,
                                                  size_t sVt, T *Ur, size_t sUr, int want_u, int *info, size_t batch, int ldx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *X = reinterpret_cast<T *>(smem_raw);            // [n][ldx]
    T *s_sig = X + (size_t) n * ldx;                    // [n] column norms
    T *s_coef = s_sig + n;                              // [n] projection coefficients (completion)
    int *s_perm = reinterpret_cast<int *>(s_coef + n);  // [n] sorted position -> column
    __shared__ int s_rot;
    __shared__ T s_val[JT / 32];
    __shared__ int s_idx[JT / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = JT / 32;
    const int npad = n + (n & 1);                       // circle method needs an even count; index n is a bye
    const T tol = (T) JacobiEps<T>::v * sqrt((T) n);
    const T tol2 = tol * tol;

    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        const T *a_g = A + mat * sA;
        T *J = want_u ? Ur + mat * sUr : nullptr;
        if (want_u) {
            for (int e = tid; e < n * n; e += JT) J[e] = (e / n == e % n) ? T(1) : T(0);   // J(r, p) at J[r + p*n]
        }
        const T unscale = jacobi_load_scaled<T, JT>(n, ldx, a_g, lda, X, s_val);
        int sweep = 0;
        for (; sweep < 40; sweep++) {
            if (tid == 0) s_rot = 0;
            __syncthreads();
            for (int round = 0; round < npad - 1; round++) {
                // a warp works on (up to) JP disjoint pairs of the round at once: their columns are held in registers, the 3 * JP
                // dot products go through one transpose-reduce butterfly, and the JP rotations are independent instruction
                // streams -- the dependent chain (reduce -> rotation -> update) of one pair no longer idles the warp
                for (int base = warp; base < npad / 2; base += JP * NW) {
                    int pp[JP], qq[JP];
                    unsigned vmask = 0;
                    T cu[JP][JE], cv[JP][JE];
                    T red[JRP];
#pragma unroll
                    for (int e = 0; e < JRP; e++) red[e] = T(0);
#pragma unroll
                    for (int u = 0; u < JP; u++) {
                        const int i = base + u * NW;
                        int p = -1, q = -1;
                        if (i < npad / 2) {
                            if (i == 0) { p = npad - 1; q = round; }
                            else {   // (round +- i) mod (npad - 1) without a division: both terms are below npad - 1
                                p = round + i; p -= p >= npad - 1 ? npad - 1 : 0;
                                q = round - i; q += q < 0 ? npad - 1 : 0;
                            }
                            if (!FULL && (p >= n || q >= n)) { p = -1; q = -1; }      // bye
                            else if (p > q) { const int t = p; p = q; q = t; }
                        }
                        pp[u] = p; qq[u] = q;
                        vmask |= (p >= 0 ? 1u : 0u) << u;
                        T aa = 0, bb = 0, cc = 0;
#pragma unroll
                        for (int e = 0; e < JE; e++) {
                            const int r = lane + 32 * e;
                            const bool ok = FULL || (p >= 0 && r < n);
                            cu[u][e] = ok ? X[(size_t) p * ldx + r] : T(0);
                            cv[u][e] = ok ? X[(size_t) q * ldx + r] : T(0);
                            aa = fma(cu[u][e], cu[u][e], aa);
                            bb = fma(cv[u][e], cv[u][e], bb);
                            cc = fma(cu[u][e], cv[u][e], cc);
                        }
                        red[3 * u] = aa; red[3 * u + 1] = bb; red[3 * u + 2] = cc;
                    }
                    TReduce<T, JRP, 16>::run(red, lane);          // lane l now holds the total of value l / (32 / JRP)
                    const T tot = red[0];
                    // lane u (< JP) gathers the three sums of pair u and computes ITS rotation: the parameters of the JP pairs
                    // cost one instruction stream instead of JP (rsqrt / rcp seeds + Newton steps, no sqrt / divide slow paths)
                    constexpr int HS = 32 / JRP;                 // lanes per value
                    const int src = 3 * HS * (lane % JP);
                    const T aa = __shfl_sync(0xffffffffu, tot, src), bb = __shfl_sync(0xffffffffu, tot, src + HS),
                            cc = __shfl_sync(0xffffffffu, tot, src + 2 * HS);
                    T cs = T(1), sn = T(0);
                    bool rot = false;
                    if (((vmask >> (lane % JP)) & 1u) && cc != T(0) && cc * cc > tol2 * (aa * bb) && aa * bb > (T) JacobiEps<T>::floor) {
                        const T zeta = (bb - aa) * jac_rcp<T>(T(2) * cc);
                        const T az = fabs(zeta);
                        T t;
                        if (az > JacobiEps<T>::big) {
                            t = jac_rcp<T>(T(2) * az);
                        } else {
                            const T w = fma(zeta, zeta, T(1));
                            t = jac_rcp<T>(az + w * jac_rsqrt<T>(w));
                        }
                        t = zeta >= T(0) ? t : -t;
                        cs = jac_rsqrt<T>(fma(t, t, T(1)));
                        sn = cs * t;
                        rot = true;
                    }
                    const unsigned rmask = __ballot_sync(0xffffffffu, rot) & ((1u << JP) - 1u);
#pragma unroll
                    for (int u = 0; u < JP; u++) {
                        const T cu_ = __shfl_sync(0xffffffffu, cs, u), su_ = __shfl_sync(0xffffffffu, sn, u);
                        if ((rmask >> u) & 1u) {
                            T *xp = X + (size_t) pp[u] * ldx, *xq = X + (size_t) qq[u] * ldx;
#pragma unroll
                            for (int e = 0; e < JE; e++) {
                                const int r = lane + 32 * e;
                                if (FULL || r < n) {   // explicit fma: this file is compiled with --fmad=false (three instructions per entry otherwise)
                                    xp[r] = fma(cu_, cu[u][e], -(su_ * cv[u][e]));
                                    xq[r] = fma(su_, cu[u][e], cu_ * cv[u][e]);
                                }
                            }
                            if (want_u) {
                                T *jp = J + (size_t) pp[u] * n, *jq = J + (size_t) qq[u] * n;
                                for (int r = lane; r < n; r += 32) {
                                    const T ju = jp[r], jv = jq[r];
                                    jp[r] = fma(cu_, ju, -(su_ * jv));
                                    jq[r] = fma(su_, ju, cu_ * jv);
                                }
                            }
                        }
                    }
                    const bool any = rmask != 0;
                    if (any && lane == 0) s_rot = 1;
                }
                __syncthreads();
            }
            if (s_rot == 0) break;
            __syncthreads();
        }
        jacobi_finish<T, JT>(n, ldx, X, s_sig, s_coef, s_perm, s_val, s_idx, S + mat * sS, Vt + mat * sVt, ldvt, J, want_u, unscale);
        if (tid == 0 && info) info[mat] = sweep >= 40 ? 1 : 0;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// k_jacobi_blk<T, JE>: the same one-sided Jacobi for n = 32 JE (64, 96, 128: BASELINE config 4), columns resident in REGISTERS.
// k_jacobi_rt moves every column shared memory -> registers -> shared memory in every round (2 n^2 s bytes per round: 2048 cycles of
// shared-memory bandwidth at n = 128, fp64), recomputes both column norms of every pair and meets a CTA barrier per round. Here the
// columns are grouped into half-blocks of 4; a warp holds TWO half-blocks (8 columns, JE rows per lane) and the ordering has two
// levels: the round-robin (circle) schedule runs over the 2 NW half-blocks -- 2 NW - 1 block-rounds per sweep, one CTA barrier and one
// trip of the columns through shared memory per BLOCK-round -- and inside a block-round the warp rotates the 16 cross pairs of its
// two half-blocks in 4 rounds of 4 disjoint pairs without leaving its registers (plus, in the first block-round of a sweep, the
// 2 x 6 pairs inside each half-block in 3 rounds): the same n (n - 1) / 2 pairs per sweep in the same number of 4-pair rounds, with a
// quarter of the shared-memory traffic and barriers. Column norms are computed once per block-round and carried through its
// rotations by a' = a - t c, b' = b + t c (exact for the rotation that annihilates c; at most 7 updates before the next
// recomputation), so a round needs ONE dot product per pair instead of three.
// Convergence: a sweep without a rotation, or -- quadratic convergence -- a sweep in which every rotated pair had both cos^2 and
// tan^2 of its angle below tol / (16 n): such rotations change the other cosines by less than n cos tan < tol, the threshold under
// which a pair is not rotated at all (the angle is tested too: between columns of equal norm a tiny cosine still asks for a large
// rotation, which would carry first-order changes to the neighbours).
// ------------------------------------------------------------------------------------------
#ifndef GPUB_JBLK_T32
#define GPUB_JBLK_T32 1
#endif
// Rotation of one pair from its squared norms aa, bb and dot product cc: cs, sn, the norm transfer dl = t cc, tt = t^2.
// t = sign(d) g / (|d| + sqrt(d^2 + g^2)) with d = bb - aa, g = 2 cc -- one rsqrt and one reciprocal instead of the two reciprocals
// and the rsqrt of the textbook form through zeta = d / g; the rotation stays orthogonal to working precision for ANY t because
// cs = (1 + t^2)^(-1/2) and sn = cs t are computed to full accuracy from the t actually used (an inexact t only leaves a residual
// of relative size eps in the pair's dot product). The textbook path serves d^2 + g^2 outside the normal range.
template<typename T>
__device__ __forceinline__ bool jblk_params(T aa, T bb, T cc, T tol2, T &cs, T &sn, T &dl, T &tt, T &c2, T &ab) {
    cs = T(1); sn = T(0); dl = T(0); tt = T(0);
    c2 = cc * cc; ab = aa * bb;
    if (cc == T(0) || c2 <= tol2 * ab || !(ab > (T) JacobiEps<T>::floor)) return false;
    const T d = bb - aa, g = cc + cc;
    T t;
    if constexpr (sizeof(T) == 8 && GPUB_JBLK_T32) {
        // fp64: the tangent in SINGLE precision. t only steers the rotation (see above): a relative error of 1e-7 leaves 1e-7 of the
        // pair's cosine behind, which is below what the other rotations of the sweep feed back (cos^2) until the cosines themselves
        // are under 1e-7 -- one sweep from the end either way. d and g are scaled by a reciprocal seed of |d| + |g| first, so the
        // squares stay in range; 12 short-latency FP32 / conversion instructions instead of a chain of 17 dependent FP64 ones.
        const double m = fabs(d) + fabs(g);
        if (m > 1e-280 && m < 1e280) {
            double inv;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(m));
            const float df = (float) (d * inv), gf = (float) (g * inv);
            if (fabsf(gf) > 1e-30f) {                       // (below that the tangent leaves the single-precision range: FP64 path)
                const float h = fmaf(df, df, gf * gf);
                float rh, rd;                                   // h is in [1/2, 1], den in [1/2, 2]: flush-to-zero forms, no range fix-ups
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rh) : "f"(h));
                const float den = fmaf(h, rh, fabsf(df));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rd) : "f"(den));
                float tf = gf * rd;
                tf = d >= 0.0 ? tf : -tf;
                tt = (double) tf * (double) tf;
                cs = jac_rsqrt<T>(tt + T(1));
                sn = cs * (double) tf;
                dl = (double) tf * cc;
                return true;
            }
        }
    }
    const T h2 = fma(d, d, g * g);
    if (h2 < JacobiEps<T>::big && h2 > T(1) / JacobiEps<T>::big) {
        const T r = jac_rsqrt<T>(h2);
        const T den = fma(h2, r, fabs(d));            // |d| + sqrt(d^2 + g^2)
        t = g * jac_rcp2<T>(den);
        t = d >= T(0) ? t : -t;
    } else {
        const T zeta = d * jac_rcp<T>(g);
        const T az = fabs(zeta);
        if (az > JacobiEps<T>::big) t = jac_rcp<T>(T(2) * az);
        else {
            const T w = fma(zeta, zeta, T(1));
            t = jac_rcp<T>(fma(w, jac_rsqrt<T>(w), az));
        }
        t = zeta >= T(0) ? t : -t;
    }
    tt = t * t;
    cs = jac_rsqrt<T>(tt + T(1));
    sn = cs * t;
    dl = t * cc;
    return true;
}

// squared norms of the 8 columns of a warp: lane l gets the norm of column l / 4
template<typename T, int JE>
__device__ __forceinline__ void jblk_norms(const T (&c)[8][JE], int lane, T &tot) {
    T red[8];
#pragma unroll
    for (int v = 0; v < 8; v++) {
        red[v] = T(0);
#pragma unroll
        for (int e = 0; e < JE; e++) red[v] = fma(c[v][e], c[v][e], red[v]);
    }
    TReduce<T, 8, 16>::run(red, lane);
    tot = red[0];
}

// Cross round S of a block-round: pairs (a_u, b_((u + S) mod 4)), u = 0..3, a_u = c[u], b_v = c[4 + v]. Lane u (mod 4) computes pair u and
// OWNS the two norms it needs: nA = |a_u|^2 stays with it through the four rounds, nB = |b_((u + S) mod 4)|^2 moves on to the lane that
// meets that column next (one shuffle per round), so the carried norms cost no broadcast and no select.
template<typename T, int JE, int S>
__device__ __forceinline__ void jblk_cross(T (&c)[8][JE], T &nA, T &nB, int lane, T tol2, T quad2, unsigned &flags, bool &dirty) {
    T d[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
    for (int e = 0; e < JE; e++) {
#pragma unroll
        for (int u = 0; u < 4; u++) d[u] = fma(c[u][e], c[4 + ((u + S) & 3)][e], d[u]);
    }
    TReduce<T, 4, 16>::run(d, lane);                  // lane l: total of value l / 8
    const T cc = __shfl_sync(0xffffffffu, d[0], 8 * (lane & 3));
    T cs, sn, dl, tt, c2, ab;
    const bool rot = jblk_params<T>(nA, nB, cc, tol2, cs, sn, dl, tt, c2, ab);
    const unsigned rmask = __ballot_sync(0xffffffffu, rot) & 0xfu;
    flags |= rot ? ((!(c2 <= quad2 * ab) || tt > quad2) ? 3u : 1u) : 0u;
    // a norm that lost more than a decimal digit to cancellation is recomputed from the column at the end of the block-round
    // (the transfer t c shrinks the smaller of the two norms: that is the one that can cancel)
    dirty = dirty || !(fabs(dl) < T(0.9375) * (dl > T(0) ? nA : nB));
    nA -= dl;
    nB += dl;
    if (rmask) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const T cu_ = __shfl_sync(0xffffffffu, cs, u), su_ = __shfl_sync(0xffffffffu, sn, u);
#pragma unroll
            for (int e = 0; e < JE; e++) {
                const T x = c[u][e], y = c[4 + ((u + S) & 3)][e];
                c[4 + ((u + S) & 3)][e] = fma(su_, x, cu_ * y);
                c[u][e] = fma(cu_, x, -(su_ * y));
            }
        }
    }
    nB = __shfl_sync(0xffffffffu, nB, (lane & ~3) | ((lane + 1) & 3));
}

template<typename T, int JE, int X0, int Y0, int X1, int Y1, int X2, int Y2, int X3, int Y3>
__device__ __forceinline__ void jblk_round(T (&c)[8][JE], T (&nn)[8], int lane, T tol2, T quad2, unsigned &flags) {
    T d[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
    for (int e = 0; e < JE; e++) {
        d[0] = fma(c[X0][e], c[Y0][e], d[0]);
        d[1] = fma(c[X1][e], c[Y1][e], d[1]);
        d[2] = fma(c[X2][e], c[Y2][e], d[2]);
        d[3] = fma(c[X3][e], c[Y3][e], d[3]);
    }
    TReduce<T, 4, 16>::run(d, lane);                  // lane l: total of value l / 8
    const int u = lane & 3;
    const T cc = __shfl_sync(0xffffffffu, d[0], 8 * u);
    // lane u (mod 4) computes the rotation of pair u: one instruction stream for the four pairs
    const T aa = u == 0 ? nn[X0] : (u == 1 ? nn[X1] : (u == 2 ? nn[X2] : nn[X3]));
    const T bb = u == 0 ? nn[Y0] : (u == 1 ? nn[Y1] : (u == 2 ? nn[Y2] : nn[Y3]));
    T cs = T(1), sn = T(0), dl = T(0), tt = T(0);
    bool rot = false;
    const T c2 = cc * cc, ab = aa * bb;
    if (cc != T(0) && !(c2 <= tol2 * ab) && ab > (T) JacobiEps<T>::floor) {
        const T zeta = (bb - aa) * jac_rcp<T>(T(2) * cc);
        const T az = fabs(zeta);
        T t;
        if (az > JacobiEps<T>::big) {
            t = jac_rcp<T>(T(2) * az);
        } else {
            const T w = fma(zeta, zeta, T(1));
            t = jac_rcp<T>(fma(w, jac_rsqrt<T>(w), az));
        }
        t = zeta >= T(0) ? t : -t;
        tt = t * t;
        cs = jac_rsqrt<T>(tt + T(1));
        sn = cs * t;
        dl = t * cc;
        rot = true;
    }
    const unsigned rmask = __ballot_sync(0xffffffffu, rot) & 0xfu;
    flags |= rot ? ((!(c2 <= quad2 * ab) || tt > quad2) ? 3u : 1u) : 0u;
#define GPUB_JBLK_APPLY(U, XI, YI)                                                                                   \
    if ((rmask >> U) & 1u) {                                                                                         \
        const T cu_ = __shfl_sync(0xffffffffu, cs, U), su_ = __shfl_sync(0xffffffffu, sn, U);                        \
        const T dl_ = __shfl_sync(0xffffffffu, dl, U);                                                               \
        _Pragma("unroll") for (int e = 0; e < JE; e++) {                                                             \
            const T x = c[XI][e], y = c[YI][e];                                                                      \
            c[XI][e] = fma(cu_, x, -(su_ * y));                                                                      \
            c[YI][e] = fma(su_, x, cu_ * y);                                                                         \
        }                                                                                                            \
        nn[XI] -= dl_;                                                                                               \
        nn[YI] += dl_;                                                                                               \
    }
    GPUB_JBLK_APPLY(0, X0, Y0)
    GPUB_JBLK_APPLY(1, X1, Y1)
    GPUB_JBLK_APPLY(2, X2, Y2)
    GPUB_JBLK_APPLY(3, X3, Y3)
#undef GPUB_JBLK_APPLY
}

#ifdef GPUB_JBLK_STATS
__device__ unsigned long long g_jblk_stats[4];   // matrices, sweeps, cycles in the sweep loop, cycles in the tail
#endif

template<typename T, int JE>
__global__ void __launch_bounds__(128 * JE) k_jacobi_blk(const T *__restrict__ A, size_t lda, size_t sA, T *S, size_t sS, T *Vt, size_t ldvt,
                                                         size_t sVt, int *info, size_t batch, int ldx) {
    constexpr int n = 32 * JE, NT = 128 * JE, NW = NT / 32, NH = 2 * NW;   // NW warps, NH half-blocks of 4 columns
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *X = reinterpret_cast<T *>(smem_raw);            // [n][ldx]
    T *s_sig = X + (size_t) n * ldx;
    T *s_coef = s_sig + n;
    int *s_perm = reinterpret_cast<int *>(s_coef + n);
    __shared__ unsigned s_flags;
    __shared__ T s_val[NW];
    __shared__ int s_idx[NW];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T tol = (T) JacobiEps<T>::v * sqrt((T) n);
    const T tol2 = tol * tol;
    const T quad2 = tol / (T) (16 * n);

    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        const T *a_g = A + mat * sA;
        if (tid == 0) s_flags = 0;
        const T unscale = jacobi_load_scaled<T, NT>(n, ldx, a_g, lda, X, s_val);
#ifdef GPUB_JBLK_STATS
        const long long st0 = clock64();
#endif
        int sweep = 0;
        // Which warp rotates which pair of half-blocks is free, so a warp FOLLOWS its first half-block (A) along the top row of the
        // circle schedule and keeps it in registers -- columns and norms -- for as long as it stays there (NW - 1 block-rounds); only
        // the second half-block (B) goes through shared memory every block-round. In block-round br the top row holds the items
        // br + 1 .. br + NW - 1 (mod NH - 1); the item that reaches the head (br + 1) is handed to warp 0, which pairs it with the
        // fixed item NH - 1, and its warp picks up the item that enters at the tail (br + NW). Half the shared-memory traffic.
        constexpr int M = NH - 1;
        int ja = warp;                                      // warp >= 1: the item it follows (block-round 0: item w pairs with item -w)
        bool load_a = true;
        T c[8][JE];
        T nA = T(0), nB;
        for (; sweep < 40; sweep++) {
            unsigned flags = 0;
            for (int br = 0; br < M; br++) {
                int ha, hb;
                if (warp == 0) { ha = M; hb = br; }
                else {
                    ha = ja;
                    hb = 2 * br - ja; hb += hb < 0 ? M : 0; hb -= hb >= M ? M : 0;
                }
                T *xa = X + (size_t) (4 * ha) * ldx + lane, *xb = X + (size_t) (4 * hb) * ldx + lane;
                if (load_a) {
#pragma unroll
                    for (int v = 0; v < 4; v++) {
#pragma unroll
                        for (int e = 0; e < JE; e++) c[v][e] = xa[(size_t) v * ldx + 32 * e];
                    }
                    nA = s_sig[4 * ha + (lane & 3)];        // (ignored in block-round 0, where the norms are recomputed)
                    load_a = false;
                }
#pragma unroll
                for (int v = 0; v < 4; v++) {
#pragma unroll
                    for (int e = 0; e < JE; e++) c[4 + v][e] = xb[(size_t) v * ldx + 32 * e];
                }
                bool dirty = false;
                if (br == 0) {
                    // first block-round of a sweep: exact norms, then the pairs inside each half-block (three rounds, norms held by every lane)
                    T nn[8];
                    jblk_norms<T, JE>(c, lane, nn[0]);
#pragma unroll
                    for (int v = 7; v >= 0; v--) nn[v] = __shfl_sync(0xffffffffu, nn[0], 4 * v);
                    jblk_round<T, JE, 0, 1, 2, 3, 4, 5, 6, 7>(c, nn, lane, tol2, quad2, flags);
                    jblk_round<T, JE, 0, 2, 1, 3, 4, 6, 5, 7>(c, nn, lane, tol2, quad2, flags);
                    jblk_round<T, JE, 0, 3, 1, 2, 4, 7, 5, 6>(c, nn, lane, tol2, quad2, flags);
                    const int u = lane & 3;
                    nA = u == 0 ? nn[0] : (u == 1 ? nn[1] : (u == 2 ? nn[2] : nn[3]));
                    nB = u == 0 ? nn[4] : (u == 1 ? nn[5] : (u == 2 ? nn[6] : nn[7]));
                    dirty = true;
                } else {
                    // the squared norms travel with the columns (s_sig is free until the tail)
                    nB = s_sig[4 * hb + (lane & 3)];
                }
                jblk_cross<T, JE, 0>(c, nA, nB, lane, tol2, quad2, flags, dirty);
                jblk_cross<T, JE, 1>(c, nA, nB, lane, tol2, quad2, flags, dirty);
                jblk_cross<T, JE, 2>(c, nA, nB, lane, tol2, quad2, flags, dirty);
                jblk_cross<T, JE, 3>(c, nA, nB, lane, tol2, quad2, flags, dirty);
                if (__any_sync(0xffffffffu, dirty)) {
                    T tot;
                    jblk_norms<T, JE>(c, lane, tot);
                    nA = __shfl_sync(0xffffffffu, tot, 4 * (lane & 3));
                    nB = __shfl_sync(0xffffffffu, tot, 16 + 4 * (lane & 3));
                }
                // A goes back to shared memory when another warp takes it over next (its item reaches the head of the top row)
                // and at the end of a sweep (the tail reads every column from shared memory)
                const int nbr = br + 1 == M ? 0 : br + 1;
                const bool hand_over = warp != 0 && ja == nbr;
                __syncwarp();                               // every lane has read the norms lanes 0..3 are about to replace
                if (lane < 4) {
                    s_sig[4 * hb + lane] = nB;
                    if (hand_over || br == M - 1) s_sig[4 * ha + lane] = nA;
                }
#pragma unroll
                for (int v = 0; v < 4; v++) {
#pragma unroll
                    for (int e = 0; e < JE; e++) xb[(size_t) v * ldx + 32 * e] = c[4 + v][e];
                }
                if (hand_over || br == M - 1) {
#pragma unroll
                    for (int v = 0; v < 4; v++) {
#pragma unroll
                        for (int e = 0; e < JE; e++) xa[(size_t) v * ldx + 32 * e] = c[v][e];
                    }
                }
                if (hand_over) {
                    ja = br + NW; ja -= ja >= M ? M : 0;    // the item that enters the top row at its tail
                    load_a = true;
                }
                __syncthreads();
            }
            if (flags) atomicOr(&s_flags, flags);
            __syncthreads();
            const unsigned f = s_flags;
            __syncthreads();
            if (tid == 0) s_flags = 0;
            if (!(f & 2u)) break;                                    // no rotation, or only negligible ones: converged
        }
        __syncthreads();
#ifdef GPUB_JBLK_STATS
        const long long st1 = clock64();
#endif
        jacobi_finish<T, NT>(n, ldx, X, s_sig, s_coef, s_perm, s_val, s_idx, S + mat * sS, Vt + mat * sVt, ldvt, (T *) nullptr, 0, unscale);
        if (tid == 0 && info) info[mat] = sweep >= 40 ? 1 : 0;
        __syncthreads();
#ifdef GPUB_JBLK_STATS
        if (tid == 0) {
            atomicAdd(&g_jblk_stats[0], 1ull);
            atomicAdd(&g_jblk_stats[1], (unsigned long long) (sweep + 1));
            atomicAdd(&g_jblk_stats[2], (unsigned long long) (st1 - st0));
            atomicAdd(&g_jblk_stats[3], (unsigned long long) (clock64() - st1));
        }
#endif
    }
}

// ------------------------------------------------------------------------------------------
// Left factor of R without accumulating the rotations: the Jacobi kernel ends with R' J = W D, so J = R W D^-1 -- one small GEMM
// (R W, tensor pipe) plus this kernel, instead of rotating a second n x n matrix in global memory through every sweep (which doubled
// the kernel's time: there is no room for J next to X in shared memory at n = 128). Column i of R W divided by sigma_i is accurate to
// eps * sigma_1 / sigma_i, so that is used while sigma_i >= weak * sigma_1; the columns behind (sorted: a suffix) carry little or no
// information about R and only have to complete the orthonormal basis: each starts from its own (normalised) R w_i, or from the unit
// vector least represented so far when that has no length left, and is orthogonalised twice against everything before it.
// One CTA per matrix, the matrix in shared memory.
// ------------------------------------------------------------------------------------------
template<typename T>
__global__ void __launch_bounds__(256) k_ur_finish(int n, T *Ur, size_t sUr, const T *__restrict__ S, size_t sS, size_t batch, int ldx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *X = reinterpret_cast<T *>(smem_raw);            // [n][ldx] columns
    T *s_coef = X + (size_t) n * ldx;                   // [n]
    __shared__ T s_val[8];
    __shared__ int s_idx[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = 8, NT = 256;
    const T weak = sizeof(T) == 8 ? T(1e-3) : T(1e-2);
    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        T *u = Ur + mat * sUr;
        const T *sg = S + mat * sS;
        const T smax = sg[0];
        int ngood = 0;
        for (int i = 0; i < n; i++) ngood += (sg[i] >= weak * smax && sg[i] > T(0)) ? 1 : 0;   // sorted: a prefix
        for (int e = tid; e < n * n; e += NT) {
            const int r = e % n, c = e / n;
            T v = u[e];
            if (c < ngood) v *= T(1) / sg[c];
            X[(size_t) c * ldx + r] = v;
        }
        __syncthreads();
        auto cta_sum = [&](T v) {
            v = jwarp_sum(v);
            if (lane == 0) s_val[warp] = v;
            __syncthreads();
            T tot = 0;
            for (int w = 0; w < NW; w++) tot += s_val[w];
            __syncthreads();
            return tot;
        };
        auto project_twice = [&](T *xz, int i) {
            for (int pass = 0; pass < 2; pass++) {
                for (int w = warp; w < i; w += NW) {
                    const T *xw = X + (size_t) w * ldx;
                    T g = 0;
                    for (int r = lane; r < n; r += 32) g = fma(xw[r], xz[r], g);
                    g = jwarp_sum(g);
                    if (lane == 0) s_coef[w] = g;
                }
                __syncthreads();
                for (int r = tid; r < n; r += NT) {
                    T acc = xz[r];
                    for (int w = 0; w < i; w++) acc = fma(-s_coef[w], X[(size_t) w * ldx + r], acc);
                    xz[r] = acc;
                }
                __syncthreads();
            }
        };
        for (int i = ngood; i < n; i++) {
            T *xz = X + (size_t) i * ldx;
            T nn = 0;
            for (int r = tid; r < n; r += NT) nn = fma(xz[r], xz[r], nn);
            nn = cta_sum(nn);
            bool own = nn > T(0);
            if (own) {
                const T inv = T(1) / sqrt(nn);
                for (int r = tid; r < n; r += NT) xz[r] *= inv;
                __syncthreads();
                project_twice(xz, i);
                T n2 = 0;
                for (int r = tid; r < n; r += NT) n2 = fma(xz[r], xz[r], n2);
                n2 = cta_sum(n2);
                own = n2 > T(0.25);                     // otherwise R w_i lies (numerically) in the span of the columns before it
                if (own) {
                    const T inv2 = T(1) / sqrt(n2);
                    for (int r = tid; r < n; r += NT) xz[r] *= inv2;
                    __syncthreads();
                }
            }
            if (!own) {
                // the unit vector with the largest component outside the span of the columns so far: 1 - sum_w w[k]^2
                T best = T(-1);
                int bestk = 0;
                for (int k = tid; k < n; k += NT) {
                    T acc = T(1);
                    for (int w = 0; w < i; w++) {
                        const T x = X[(size_t) w * ldx + k];
                        acc = fma(-x, x, acc);
                    }
                    if (acc > best) { best = acc; bestk = k; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const T ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
                    if (ob > best || (ob == best && ok < bestk)) { best = ob; bestk = ok; }
                }
                if (lane == 0) { s_val[warp] = best; s_idx[warp] = bestk; }
                __syncthreads();
                best = s_val[0]; bestk = s_idx[0];
                for (int w = 1; w < NW; w++)
                    if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bestk)) { best = s_val[w]; bestk = s_idx[w]; }
                __syncthreads();
                for (int r = tid; r < n; r += NT) xz[r] = (r == bestk) ? T(1) : T(0);
                __syncthreads();
                project_twice(xz, i);
                T n2 = 0;
                for (int r = tid; r < n; r += NT) n2 = fma(xz[r], xz[r], n2);
                n2 = cta_sum(n2);
                const T inv2 = T(1) / sqrt(n2);
                for (int r = tid; r < n; r += NT) xz[r] *= inv2;
                __syncthreads();
            }
        }
        for (int e = tid; e < n * n; e += NT) u[e] = X[(size_t) (e / n) * ldx + (e % n)];
        __syncthreads();
    }
}

template<typename T> int internal_geqrf(gpub_ctx_t, int, size_t, size_t, T *, size_t, size_t, T *, size_t, size_t);
template<> int internal_geqrf<double>(gpub_ctx_t c, int s, size_t m, size_t n, double *A, size_t lda, size_t sA, double *tau, size_t sT, size_t b) { return gpub_geqrf_batched_f64(c, s, m, n, A, lda, sA, tau, sT, b); }
template<> int internal_geqrf<float>(gpub_ctx_t c, int s, size_t m, size_t n, float *A, size_t lda, size_t sA, float *tau, size_t sT, size_t b) { return gpub_geqrf_batched_f32(c, s, m, n, A, lda, sA, tau, sT, b); }
template<typename T> int internal_ormqr(gpub_ctx_t, int, int, size_t, size_t, size_t, const T *, size_t, size_t, const T *, size_t, T *, size_t, size_t, size_t);
template<> int internal_ormqr<double>(gpub_ctx_t c, int s, int tr, size_t m, size_t nc, size_t k, const double *A, size_t lda, size_t sA, const double *tau, size_t sT, double *C, size_t ldc, size_t sC, size_t b) { return gpub_ormqr_batched_f64(c, s, tr, m, nc, k, A, lda, sA, tau, sT, C, ldc, sC, b); }
template<> int internal_ormqr<float>(gpub_ctx_t c, int s, int tr, size_t m, size_t nc, size_t k, const float *A, size_t lda, size_t sA, const float *tau, size_t sT, float *C, size_t ldc, size_t sC, size_t b) { return gpub_ormqr_batched_f32(c, s, tr, m, nc, k, A, lda, sA, tau, sT, C, ldc, sC, b); }

int internal_gemm(gpub_ctx_t c, int s, size_t m, size_t n, size_t k, double alpha, const double *A, size_t lda, size_t sA, const double *B, size_t ldb,
                  size_t sB, double beta, double *C, size_t ldc, size_t sC, size_t b) {
    return gpub_gemm_batched_f64(c, s, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, b);
}
int internal_gemm(gpub_ctx_t c, int s, size_t m, size_t n, size_t k, float alpha, const float *A, size_t lda, size_t sA, const float *B, size_t ldb,
                  size_t sB, float beta, float *C, size_t ldc, size_t sC, size_t b) {
    return gpub_gemm_batched_f32(c, s, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, b);
}
int internal_transpose(gpub_ctx_t c, int s, size_t m, size_t n, const double *A, size_t sA, double *At, size_t sAt, size_t b) {
    return gpub_transpose_batched_f64(c, s, m, n, A, sA, At, sAt, b);
}
int internal_transpose(gpub_ctx_t c, int s, size_t m, size_t n, const float *A, size_t sA, float *At, size_t sAt, size_t b) {
    return gpub_transpose_batched_f32(c, s, m, n, A, sA, At, sAt, b);
}

inline size_t per_matrix_work_elems(size_t n) { return 3 * n * n + 8 * n + 8; }

inline int wy_final_gemm(gpub_ctx_t ctx, int sidx, size_t m, size_t n, const double *V, size_t lda, size_t sA, const double *Y, const double *Ur, size_t sUr,
                         double *U, size_t ldu, size_t sU, size_t batch) {
    return gpub_internal_gemm_plus_e_f64(ctx, sidx, m, m, n, -1.0, V, lda, sA, Y, n, m * n, Ur, n, sUr, U, ldu, sU, batch);
}
inline int wy_final_gemm(gpub_ctx_t, int, size_t, size_t, const float *, size_t, size_t, const float *, const float *, size_t, float *, size_t, size_t, size_t) {
    return GPUB_ENOTSUP;   // the block-reflector assembly is fp64 only (use_wy_assembly)
}

// U = Q blockdiag(Ur, I) through one block reflector (see k_make_v / k_tfactor): fp64 shapes the tensor-pipe GEMM tiles cover
template<typename T>
inline bool use_wy_assembly(size_t m, size_t n) { return sizeof(T) == 8 && n > 32 && n <= 128 && m % 64 == 0 && n % 64 == 0; }
// extra workspace elements per matrix: V' (n x m), Y (n x m), T (n x n)
inline size_t wy_extra_elems(size_t m, size_t n) { return 2 * m * n + n * n; }

// shared memory of the Jacobi kernel for an n x n factor (sm_100a: 227 KB opt-in per CTA, 2 KB kept for the static part)
template<typename T>
inline size_t jacobi_smem(size_t n) { return (n * (n | 1) + 2 * n) * sizeof(T) + n * sizeof(int) + 64; }

template<typename T>
inline bool shape_supported(size_t m, size_t n) {
    if (m < n) return false;
    if (n <= 32) return true;
    return n <= 256 && jacobi_smem<T>(n) <= (size_t) 227 * 1024 - 2048;
}

#ifndef GPUB_SVD_CHUNKS
#define GPUB_SVD_CHUNKS 4
#endif

// 0 = shape not served by the batched kernels (the header turns that into std::invalid_argument in the Svd constructor)
template<typename T>
size_t worksize(size_t m, size_t n, int jobu, size_t batch) {
    if (!shape_supported<T>(m, n)) return 0;
    const bool wy = (jobu == 'A' || jobu == 'a') && use_wy_assembly<T>(m, n);
    return (per_matrix_work_elems(n) + (wy ? wy_extra_elems(m, n) : 0)) * batch * sizeof(T) + 256;
}

// ------------------------------------------------------------------------------------------
// Range guard (LAPACK's gesvd scales a matrix whose largest entry is below sqrt(safmin) / eps or above its reciprocal, dlascl): the QR
// step and the rotations work with sums of squares, which underflow (or overflow) long before the entries do -- a matrix of
// entries around 1e-150 with small singular values came back with U orthogonal to 1e-4 only, one of 256 x 64 with NaN. One CTA per
// matrix finds the largest entry and, only when it is out of range, multiplies the matrix by the power of two that brings it to order
// one (exact); the factor to multiply the singular values back by is left in the last element of the matrix's workspace slice
// and applied by k_unscale_s at the end. In range -- always, for ordinary data -- this is one read of the batch.
// ------------------------------------------------------------------------------------------
template<typename T> struct VecPair;
template<> struct VecPair<double> { using type = double2; };
template<> struct VecPair<float> { using type = float2; };
template<typename T> struct SvdRange;
template<> struct SvdRange<double> { static constexpr double small = 1.35e-138, big = 7.4e137, huge = 1.7e308; };
template<> struct SvdRange<float> { static constexpr float small = 1.8e-12f, big = 5.5e11f, huge = 3.4e38f; };

template<typename T>
__global__ void __launch_bounds__(512) k_prescale(int m, int n, T *A, size_t lda, size_t sA, T *w, size_t per, size_t batch) {
    __shared__ T s_red[16];
    const int tid = threadIdx.x;
    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        T *a = A + mat * sA;
        T amax = T(0);
        if (lda == (size_t) m && ((size_t) m * n) % (2 * 512 * 8) == 0 && (((uintptr_t) a) & 15u) == 0) {
            // dense matrix: 128-bit loads, eight per thread in flight (this pass is one read of the whole batch)
            using V2 = typename VecPair<T>::type;
            const V2 *a2 = reinterpret_cast<const V2 *>(a);
            const size_t n2 = (size_t) m * n / 2;
            for (size_t i0 = tid; i0 < n2; i0 += 512 * 8) {
                V2 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = a2[i0 + (size_t) u * 512];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const T ax = fabs(v[u].x), ay = fabs(v[u].y);
                    amax = (ax > amax && ax <= SvdRange<T>::huge) ? ax : amax;
                    amax = (ay > amax && ay <= SvdRange<T>::huge) ? ay : amax;
                }
            }
        } else {
            for (int j = 0; j < n; j++)
                for (int i = tid; i < m; i += 512) {
                    const T av = fabs(a[(size_t) j * lda + i]);
                    amax = (av > amax && av <= SvdRange<T>::huge) ? av : amax;
                }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T other = __shfl_xor_sync(0xffffffffu, amax, o);
            amax = other > amax ? other : amax;
        }
        if ((tid & 31) == 0) s_red[tid >> 5] = amax;
        __syncthreads();
        amax = s_red[0];
        for (int k = 1; k < 16; k++) amax = s_red[k] > amax ? s_red[k] : amax;
        T up = T(1);
        if (amax > T(0) && (amax < SvdRange<T>::small || amax > SvdRange<T>::big)) {
            int ex = 0;
            (void) frexp(amax, &ex);
            // two half steps: 2^-ex itself may not be representable when the entries are near the bottom of the range
            const T d1 = ldexp(T(1), -(ex / 2)), d2 = ldexp(T(1), -(ex - ex / 2));
            for (int j = 0; j < n; j++)
                for (int i = tid; i < m; i += 512) a[(size_t) j * lda + i] = (a[(size_t) j * lda + i] * d1) * d2;
            up = T(-1);                                   // marker: the factor is 2^ex, applied in two steps as well
            if (tid == 0) w[mat * per + per - 2] = (T) ex;
        }
        if (tid == 0) w[mat * per + per - 1] = up;
        __syncthreads();
    }
}

template<typename T>
__global__ void k_unscale_s(int n, T *S, size_t sS, const T *__restrict__ w, size_t per, size_t batch) {
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < (size_t) n * batch; e += (size_t) gridDim.x * blockDim.x) {
        const size_t mat = e / n;
        if (w[mat * per + per - 1] < T(0)) {
            const int ex = (int) w[mat * per + per - 2];
            T *sp = S + mat * sS + (e - mat * n);
            *sp = ldexp(ldexp(*sp, ex / 2), ex - ex / 2);
        }
    }
}

// the paths of gesvd_batched on one sub-batch and one stream
template<typename T>
int gesvd_paths(gpub_ctx_t ctx, int sidx, cudaStream_t stream, bool want_u, size_t m, size_t n, T *A, size_t lda, size_t sA, T *S, size_t sS, T *U,
                size_t ldu, size_t sU, T *Vt, size_t ldvt, size_t sVt, T *w, size_t per, int *info, size_t batch) {
    if (m <= 64 && n <= 32) {
        const unsigned grid = (unsigned) gpub_ceil_div(batch, 64);
        k_gesvd_small<T><<<grid, 64, 0, stream>>>((int) m, (int) n, A, lda, sA, S, sS, U, ldu, sU, Vt, ldvt, sVt, w, per, want_u ? 1 : 0,
                                                   info, batch);
        GPUB_LAUNCH_CHECK();
        return GPUB_OK;
    }
    if (n > 32) {
        // jacobi path: per-matrix scratch = [Ur n*n | tau n | R, later V(0:n)' Ur n*n | W n*n]
        const size_t ldx = n | 1;
        const size_t smem = (n * ldx + 2 * n) * sizeof(T) + n * sizeof(int) + 64;
        if (smem > (size_t) ctx->max_smem_optin - 2048) return GPUB_ENOTSUP;
        T *Urj = w, *tauj = w + n * n;
        int e = internal_geqrf<T>(ctx, sidx, m, n, A, lda, sA, tauj, per, batch);
        if (e) return e;
        const size_t cap = (size_t) ctx->sm_count * 2;
        const unsigned jgrid = (unsigned) (batch < cap ? batch : cap);
        // the left factor of R comes from R W D^-1 afterwards (k_ur_finish) unless Vt is strided: then the rotations are accumulated
        const size_t usm = (n * ldx + n) * sizeof(T);
        const bool accumulate = want_u && (ldvt != n || usm > (size_t) ctx->max_smem_optin - 1024);
#define GPUB_JACOBI_LAUNCH(JEV)                                                                                              \
    {                                                                                                                         \
        auto kern = (n == (size_t) 32 * JEV) ? k_jacobi_rt<T, JEV, true> : k_jacobi_rt<T, JEV, false>;                        \
        GPUB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));                       \
        kern<<<jgrid, JT, smem, stream>>>((int) n, A, lda, sA, S, sS, Vt, ldvt, sVt, Urj, per, accumulate ? 1 : 0, info, \
                                                         batch, (int) ldx);                                                   \
    }
        if (!accumulate && (n == 64 || n == 96 || n == 128)) {
            // columns resident in registers, two-level ordering (k_jacobi_blk)
#define GPUB_JBLK_LAUNCH(JEV)                                                                                                \
    {                                                                                                                         \
        GPUB_CUDA(cudaFuncSetAttribute(k_jacobi_blk<T, JEV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));       \
        k_jacobi_blk<T, JEV><<<jgrid, 128 * JEV, smem, stream>>>(A, lda, sA, S, sS, Vt, ldvt, sVt, info, batch, (int) ldx);  \
    }
            if (n == 64) GPUB_JBLK_LAUNCH(2)
            else if (n == 96) GPUB_JBLK_LAUNCH(3)
            else GPUB_JBLK_LAUNCH(4)
#undef GPUB_JBLK_LAUNCH
        } else if (n <= 128) GPUB_JACOBI_LAUNCH(4)
        else if (n <= 192) GPUB_JACOBI_LAUNCH(6)
        else if (n <= 256) GPUB_JACOBI_LAUNCH(8)
        else return GPUB_ENOTSUP;
#undef GPUB_JACOBI_LAUNCH
        GPUB_LAUNCH_CHECK();
        if (want_u && !accumulate) {
            T *Rw = w + n * n + n, *Ww = Rw + n * n;
            const unsigned gr = (unsigned) (gpub_ceil_div(n * n * batch, 256) < 4096 ? gpub_ceil_div(n * n * batch, 256) : 4096);
            k_extract_r<T><<<gr, 256, 0, stream>>>((int) n, A, lda, sA, Rw, per, batch);
            GPUB_LAUNCH_CHECK();
            e = internal_transpose(ctx, sidx, n, n, Vt, sVt, Ww, per, batch);                                           // W = Vt'
            if (e) return e;
            e = internal_gemm(ctx, sidx, n, n, n, T(1), Rw, n, per, Ww, n, per, T(0), Urj, n, per, batch);               // R W = J D
            if (e) return e;
            GPUB_CUDA(cudaFuncSetAttribute(k_ur_finish<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) usm));
            k_ur_finish<T><<<(unsigned) (batch < (size_t) ctx->sm_count * 2 ? batch : (size_t) ctx->sm_count * 2), 256, usm, stream>>>(
                (int) n, Urj, per, S, sS, batch, (int) ldx);
            GPUB_LAUNCH_CHECK();
        }
        if (want_u) {
            auto init_u = [&]() {
                size_t total = m * m * batch;
                unsigned grid = (unsigned) (gpub_ceil_div(total, 256) < 8192 ? gpub_ceil_div(total, 256) : 8192);
                k_init_u<T><<<grid, 256, 0, stream>>>((int) m, (int) n, Urj, per, U, ldu, sU, batch);
            };
            const size_t tsm = (n * (n + 1) + (n / 2) * (n / 2 + 1) * (n > 16 ? 1 : 0) + 16) * sizeof(T);   // T / G in place + the merge scratch
            if (use_wy_assembly<T>(m, n) && lda == m && (sA & 1) == 0 && (ldu & 1) == 0 && (sU & 1) == 0 &&
                ((((uintptr_t) A) | ((uintptr_t) U)) & 15u) == 0 && tsm <= (size_t) ctx->max_smem_optin) {
                T *VtW = w + per * batch, *Yw = VtW + m * n * batch, *Tw = Yw + m * n * batch, *X1 = w + n * n + n;
                const unsigned gv = (unsigned) (gpub_ceil_div(n * n * batch, 256) < 4096 ? gpub_ceil_div(n * n * batch, 256) : 4096);
                k_make_v<T><<<gv, 256, 0, stream>>>((int) n, A, lda, sA, batch);
                GPUB_LAUNCH_CHECK();
                e = internal_transpose(ctx, sidx, m, n, A, sA, VtW, m * n, batch);
                if (e) return e;
                e = internal_gemm(ctx, sidx, n, n, m, T(1), VtW, n, m * n, A, lda, sA, T(0), Tw, n, n * n, batch);      // G = V'V
                if (e) return e;
                GPUB_CUDA(cudaFuncSetAttribute(k_tfactor<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tsm));
                k_tfactor<T><<<(unsigned) (batch < (size_t) ctx->sm_count ? batch : (size_t) ctx->sm_count), 256, tsm, stream>>>((int) n, Tw, n * n, tauj, per, batch);
                GPUB_LAUNCH_CHECK();
                e = internal_gemm(ctx, sidx, n, n, n, T(1), VtW, n, m * n, Urj, n, per, T(0), X1, n, per, batch);           // V(0:n)' Ur
                if (e) return e;
                e = internal_gemm(ctx, sidx, n, n, n, T(1), Tw, n, n * n, X1, n, per, T(0), Yw, n, m * n, batch);            // Y = T X
                if (e) return e;
                if (m > n) {
                    e = internal_gemm(ctx, sidx, n, m - n, n, T(1), Tw, n, n * n, VtW + n * n, n, m * n, T(0), Yw + n * n, n, m * n, batch);
                    if (e) return e;
                }
                // U = E - V Y, E = blockdiag(Ur, I) generated in the GEMM's epilogue: the m x m result is written once and never read
                e = wy_final_gemm(ctx, sidx, m, n, A, lda, sA, Yw, Urj, per, U, ldu, sU, batch);
                if (e) return e;
            } else {
                init_u();
                GPUB_LAUNCH_CHECK();
                e = internal_ormqr<T>(ctx, sidx, 0, m, m, n, A, lda, sA, tauj, per, U, ldu, sU, batch);
                if (e) return e;
            }
        }
        return GPUB_OK;
    }

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

template<typename T>
int gesvd_batched(gpub_ctx_t ctx, int sidx, int jobu, size_t m, size_t n, T *A, size_t lda, size_t sA, T *S, size_t sS, T *U,
                  size_t ldu, size_t sU, T *Vt, size_t ldvt, size_t sVt, void *work, size_t work_bytes, int *info, size_t batch) {
    if (m == 0 || n == 0 || batch == 0) return GPUB_OK;
    const bool want_u = (jobu == 'A' || jobu == 'a');
    if (!want_u && !(jobu == 'N' || jobu == 'n')) return GPUB_EINVAL;
    if (!A || !S || !Vt || (want_u && !U) || !work) return GPUB_EINVAL;
    if (m < n || lda < m || ldvt < n || (want_u && ldu < m)) return GPUB_EINVAL;
    if (!shape_supported<T>(m, n)) return GPUB_ENOTSUP;
    if (work_bytes < worksize<T>(m, n, jobu, batch)) return GPUB_EWORK;
    GPUB_ENTER(ctx, sidx);
    T *w = reinterpret_cast<T *>((((uintptr_t) work) + 15) & ~(uintptr_t) 15);
    const size_t per = per_matrix_work_elems(n);
#if GPUB_SVD_CHUNKS > 1
    if (n > 32 && sidx < GPUB_INTERNAL_SLOT0 && batch > (size_t) ctx->sm_count && batch >= 2 * GPUB_SVD_CHUNKS) {
        // The QR and the Jacobi kernel both run one CTA per matrix and SM: a batch that is not a multiple of the SM count leaves SMs idle
        // in the last wave of each of them (256 matrices on 148 SMs: 40 idle for half of both kernels), and every stage waits for the
        // slowest CTA of the one before. Cut into chunks that run the whole sequence on the library's own side streams, the stages of
        // different chunks overlap: a finished QR CTA makes room for a Jacobi CTA of another chunk, and the GEMMs of the U assembly
        // fill the SMs the last Jacobi wave leaves idle. The chunks are separate sub-batches (own slice of every operand and of the
        // workspace); the caller's stream forks into them and joins them again.
        const bool wy = want_u && use_wy_assembly<T>(m, n);
        const size_t per_all = per + (wy ? wy_extra_elems(m, n) : 0);
        cudaEvent_t ev[GPUB_SVD_CHUNKS + 1];
        int made = 0, e = GPUB_OK;
        cudaError_t c = cudaSuccess;
        for (; made <= GPUB_SVD_CHUNKS && c == cudaSuccess; made++) c = cudaEventCreateWithFlags(&ev[made], cudaEventDisableTiming);
        if (c != cudaSuccess) made--;
        if (c == cudaSuccess) c = cudaEventRecord(ev[GPUB_SVD_CHUNKS], stream);
        for (int ch = 0; ch < GPUB_SVD_CHUNKS && c == cudaSuccess && !e; ch++) {
            const size_t lo = ch * batch / GPUB_SVD_CHUNKS, hi = (ch + 1) * batch / GPUB_SVD_CHUNKS;
            int serr = 0;
            gpub_stream_slot *side = gpub_slot(ctx, GPUB_INTERNAL_SLOT0 + ch, &serr);
            if (!side) { e = serr; break; }
            c = cudaStreamWaitEvent(side->stream, ev[GPUB_SVD_CHUNKS], 0);
            if (c != cudaSuccess) break;
            e = gesvd_batched<T>(ctx, GPUB_INTERNAL_SLOT0 + ch, jobu, m, n, A + lo * sA, lda, sA, S + lo * sS, sS, U ? U + lo * sU : nullptr, ldu, sU,
                                 Vt + lo * sVt, ldvt, sVt, w + per_all * lo, per_all * (hi - lo) * sizeof(T) + 256 /* aligned already: the slack is not touched */,
                                 info ? info + lo : nullptr, hi - lo);
            if (!e) c = cudaEventRecord(ev[ch], side->stream);
        }
        // (the joins come after ALL the launches: on a legacy default stream each of them is a device-wide ordering point)
        for (int ch = 0; ch < GPUB_SVD_CHUNKS && c == cudaSuccess && !e; ch++) c = cudaStreamWaitEvent(stream, ev[ch], 0);
        for (int i = 0; i < made; i++) cudaEventDestroy(ev[i]);
        return e ? e : (int) c;
    }
#endif

    if (m <= 64 && n <= 32)   // LAPACK-faithful thread-per-matrix path: scale-safe building blocks (lartg, lasv2, nrm2 with scaling)
        return gesvd_paths<T>(ctx, sidx, stream, want_u, m, n, A, lda, sA, S, sS, U, ldu, sU, Vt, ldvt, sVt, w, per, info, batch);
    {
        const size_t cap = (size_t) ctx->sm_count * 4;
        k_prescale<T><<<(unsigned) (batch < cap ? batch : cap), 512, 0, stream>>>((int) m, (int) n, A, lda, sA, w, per, batch);
        GPUB_LAUNCH_CHECK();
    }
    const int e = gesvd_paths<T>(ctx, sidx, stream, want_u, m, n, A, lda, sA, S, sS, U, ldu, sU, Vt, ldvt, sVt, w, per, info, batch);
    if (e) return e;
    {
        const size_t tot = gpub_ceil_div(n * batch, 256);
        k_unscale_s<T><<<(unsigned) (tot < 1024 ? tot : 1024), 256, 0, stream>>>((int) n, S, sS, w, per, batch);
        GPUB_LAUNCH_CHECK();
    }
    return GPUB_OK;
}

} // namespace

extern "C" {

#ifdef GPUB_JBLK_STATS
int gpub_debug_jacobi_stats(unsigned long long *out4, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out4, g_jblk_stats, sizeof(unsigned long long) * 4);
    if (reset) { unsigned long long z[4] = {0}; cudaMemcpyToSymbol(g_jblk_stats, z, sizeof(z)); }
    return 0;
}
#endif

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
