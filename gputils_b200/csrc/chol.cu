// Batched Cholesky factorisation (lower) and solve.
// Replaces cusolverDn{D,S}potrfBatched / potrsBatched and the single-matrix potrf / potrs
// (ref: tensor.cuh:2135-2197, 1742-1783).
//
// Roofline: potrf moves 2 n^2 s bytes for n^3/3 flop -> AI = n/(6 s) flop/B, HBM-bound for every n
// of the sweep; potrs moves n^2 s + 2 n s bytes for 2 n^2 flop.
//
//   k_potrf_group<T, NP> : n <= 32. One matrix per group of NP lanes (NP = 4, 8, 16, 32; a warp holds
//        32/NP matrices). Lane i keeps row i in registers (compile-time indexed), loads and stores are
//        column-wise (a warp reads whole 32-byte sectors), the pivot column is broadcast through a
//        per-group shared-memory line read back with 128-bit loads, 1/sqrt(d) comes from one
//        rsqrt + Newton step shared by the diagonal and the column scaling. Right-looking, fully unrolled.
//   k_potrf_cta<T>       : n > 32. One matrix per CTA, staged in shared memory when it fits.
//   k_potrs_group<T, NP> : forward substitution on the row layout, transposition through padded shared
//        memory, backward substitution on the column layout; reciprocals of the diagonal are computed once,
//        in parallel, one per lane.
//   k_potrs_cta<T>       : n > 32.
//
// Only the lower triangle is read and written: the strict upper triangle of A is never touched
// (cuSOLVER semantics, SURVEY.md section 7 hard part 8). info[i] = first non-positive pivot (1-based) or 0.
#include "common.cuh"

namespace {

template<typename T> __device__ __forceinline__ T fast_rsqrt(T d);
template<> __device__ __forceinline__ double fast_rsqrt<double>(double d) {
    // rsqrt() is ~1 ulp; one Newton step on (d, r) makes s = d*r a correctly rounded-quality sqrt
    double r = rsqrt(d);
    double e = fma(-d * r, r, 1.0); // 1 - d r^2
    return fma(0.5 * r, e, r);
}
template<> __device__ __forceinline__ float fast_rsqrt<float>(float d) {
    float r = rsqrtf(d);
    float e = fmaf(-d * r, r, 1.0f);
    return fmaf(0.5f * r, e, r);
}

// ------------------------------------------------------------------------------------------
// potrf, n <= 32: NP lanes per matrix
// ------------------------------------------------------------------------------------------
template<typename T, int NP>
__global__ void __launch_bounds__(256) k_potrf_group(int n, T *A, size_t lda, size_t strideA, int *info, size_t batch) {
    constexpr int GROUPS = 256 / NP;
    __shared__ __align__(16) T s_col[GROUPS][NP];
    const int grp = threadIdx.x / NP;
    const int i = threadIdx.x % NP; // row owned by this lane
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gmask = (NP == 32) ? 0xffffffffu : (((1u << NP) - 1u) << (lane & ~(unsigned) (NP - 1)));
    const size_t ngroups = (size_t) gridDim.x * GROUPS;
    const bool row_ok = i < n;
    T *col = s_col[grp];

    for (size_t mat = (size_t) blockIdx.x * GROUPS + grp; mat < batch; mat += ngroups) {
        T *a_g = A + mat * strideA;
        T a[NP];
        // lower triangle only: lane i needs columns 0..i
#pragma unroll
        for (int c = 0; c < NP; c++) a[c] = (row_ok && c <= i) ? a_g[i + (size_t) c * lda] : T(c == i ? 1 : 0);
        int bad = 0;
#pragma unroll
        for (int j = 0; j < NP; j++) {
            if (j < n) { // warp-uniform
                const T d = __shfl_sync(gmask, a[j], j, NP);
                if (!(d > T(0)) && bad == 0) bad = j + 1;
                const T r = fast_rsqrt<T>(d);
                const T l = a[j] * r; // lane j: sqrt(d); lanes below: L(i,j)
                a[j] = l;
                col[i] = l;
                __syncwarp(gmask);
#pragma unroll
                for (int c = j + 1; c < NP; c++) a[c] = fma(-l, col[c], a[c]);
                __syncwarp(gmask);
            }
        }
        if (row_ok) {
#pragma unroll
            for (int c = 0; c < NP; c++)
                if (c <= i) a_g[i + (size_t) c * lda] = a[c];
        }
        if (i == 0) info[mat] = bad;
    }
}

// ------------------------------------------------------------------------------------------
// potrf, any n: one CTA per matrix (shared memory when it fits, else in place in global memory)
// ------------------------------------------------------------------------------------------
template<typename T>
__global__ void __launch_bounds__(256) k_potrf_cta(int n, T *A, size_t lda, size_t strideA, int *info, size_t batch, int use_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    __shared__ T s_r;
    __shared__ int s_bad;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        T *a_g = A + mat * strideA;
        T *M = use_smem ? sm : a_g;
        const size_t ld = use_smem ? (size_t) n : lda;
        if (use_smem) {
            for (int e = tid; e < n * n; e += nt) {
                int r = e % n, c = e / n;
                if (r >= c) M[r + (size_t) c * ld] = a_g[r + (size_t) c * lda];
            }
        }
        if (tid == 0) s_bad = 0;
        __syncthreads();
        for (int j = 0; j < n; j++) {
            if (tid == 0) {
                T d = M[j + (size_t) j * ld];
                if (!(d > T(0)) && s_bad == 0) s_bad = j + 1;
                s_r = fast_rsqrt<T>(d);
            }
            __syncthreads();
            const T r = s_r;
            for (int rr = j + tid; rr < n; rr += nt) M[rr + (size_t) j * ld] *= r;
            __syncthreads();
            // trailing update of the lower triangle: columns c > j, rows >= c
            const int rem = n - j - 1;
            for (int e = tid; e < rem * rem; e += nt) {
                int rr = j + 1 + e % rem, c = j + 1 + e / rem;
                if (rr >= c) M[rr + (size_t) c * ld] = fma(-M[rr + (size_t) j * ld], M[c + (size_t) j * ld], M[rr + (size_t) c * ld]);
            }
            __syncthreads();
        }
        if (use_smem) {
            for (int e = tid; e < n * n; e += nt) {
                int r = e % n, c = e / n;
                if (r >= c) a_g[r + (size_t) c * lda] = M[r + (size_t) c * ld];
            }
        }
        if (tid == 0) info[mat] = s_bad;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// potrs, n <= 32
// ------------------------------------------------------------------------------------------
template<typename T, int NP>
__global__ void __launch_bounds__(256) k_potrs_group(int n, const T *__restrict__ L, size_t ldl, size_t strideL, T *b, size_t strideB,
                                                      size_t batch) {
    constexpr int GROUPS = 256 / NP;
    constexpr int LDP = NP + 1; // padded: row writes and column reads are both conflict-free
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_t = reinterpret_cast<T *>(smem_raw) + (size_t) (threadIdx.x / NP) * NP * LDP;
    const int grp = threadIdx.x / NP;
    const int i = threadIdx.x % NP;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gmask = (NP == 32) ? 0xffffffffu : (((1u << NP) - 1u) << (lane & ~(unsigned) (NP - 1)));
    const size_t ngroups = (size_t) gridDim.x * GROUPS;
    const bool row_ok = i < n;

    for (size_t mat = (size_t) blockIdx.x * GROUPS + grp; mat < batch; mat += ngroups) {
        const T *l_g = L + mat * strideL;
        T *b_g = b + mat * strideB;
        T l[NP];
#pragma unroll
        for (int c = 0; c < NP; c++) l[c] = (row_ok && c <= i) ? l_g[i + (size_t) c * ldl] : T(c == i ? 1 : 0);
        T x = row_ok ? b_g[i] : T(0);
        // reciprocal of the own diagonal entry, all lanes in parallel
        T dinv = T(1);
#pragma unroll
        for (int c = 0; c < NP; c++)
            if (c == i) dinv = T(1) / l[c];
        // rows -> shared (transposed read below)
#pragma unroll
        for (int c = 0; c < NP; c++) s_t[i * LDP + c] = l[c];
        // forward: L y = b
#pragma unroll
        for (int j = 0; j < NP; j++) {
            if (j < n) {
                const T yj = __shfl_sync(gmask, x * dinv, j, NP);
                if (i == j) x = yj;
                else if (i > j) x = fma(-l[j], yj, x);
            }
        }
        __syncwarp(gmask);
        // column layout: lane j holds L(r, j) for r >= j
#pragma unroll
        for (int r = 0; r < NP; r++) l[r] = s_t[r * LDP + i];
        __syncwarp(gmask);
        // backward: L^T x = y
#pragma unroll
        for (int jj = NP - 1; jj >= 0; jj--) {
            if (jj < n) {
                const T xj = __shfl_sync(gmask, x * dinv, jj, NP);
                if (i == jj) x = xj;
                else if (i < jj) x = fma(-l[jj], xj, x);
            }
        }
        if (row_ok) b_g[i] = x;
    }
}

// ------------------------------------------------------------------------------------------
// potrs, any n: one CTA per matrix, rhs in shared memory, coalesced column sweeps of L
// ------------------------------------------------------------------------------------------
template<typename T>
__global__ void __launch_bounds__(256) k_potrs_cta(int n, const T *__restrict__ L, size_t ldl, size_t strideL, T *b, size_t strideB,
                                                    size_t batch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *x = reinterpret_cast<T *>(smem_raw); // n entries
    __shared__ T s_red[8];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        const T *l_g = L + mat * strideL;
        T *b_g = b + mat * strideB;
        for (int e = tid; e < n; e += nt) x[e] = b_g[e];
        __syncthreads();
        // forward, column oriented
        for (int j = 0; j < n; j++) {
            if (tid == 0) x[j] = x[j] / l_g[j + (size_t) j * ldl];
            __syncthreads();
            const T xj = x[j];
            for (int r = j + 1 + tid; r < n; r += nt) x[r] = fma(-l_g[r + (size_t) j * ldl], xj, x[r]);
            __syncthreads();
        }
        // backward, dot-product oriented (column j of L is row j of L^T, contiguous)
        for (int j = n - 1; j >= 0; j--) {
            T part = 0;
            for (int r = j + 1 + tid; r < n; r += nt) part = fma(l_g[r + (size_t) j * ldl], x[r], part);
            for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
            if ((tid & 31) == 0) s_red[tid >> 5] = part;
            __syncthreads();
            if (tid == 0) {
                T tot = 0;
                for (int w = 0; w < (nt >> 5); w++) tot += s_red[w];
                x[j] = (x[j] - tot) / l_g[j + (size_t) j * ldl];
            }
            __syncthreads();
        }
        for (int e = tid; e < n; e += nt) b_g[e] = x[e];
        __syncthreads();
    }
}

template<typename T>
int potrf_batched(gpub_ctx_t ctx, int sidx, size_t n, T *A, size_t lda, size_t strideA, int *info, size_t batch) {
    if (n == 0 || batch == 0) return GPUB_OK;
    if (!A || !info || lda < n) return GPUB_EINVAL;
    if (n > 8192) return GPUB_ENOTSUP;
    GPUB_ENTER(ctx, sidx);
    if (n <= 32) {
        const int np = n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32;
        const size_t groups = 256 / np;
        const size_t want = gpub_ceil_div(batch, groups);
        const size_t cap = (size_t) ctx->sm_count * 8;
        const unsigned grid = (unsigned) (want < cap ? want : cap);
        switch (np) {
            case 4: k_potrf_group<T, 4><<<grid, 256, 0, stream>>>((int) n, A, lda, strideA, info, batch); break;
            case 8: k_potrf_group<T, 8><<<grid, 256, 0, stream>>>((int) n, A, lda, strideA, info, batch); break;
            case 16: k_potrf_group<T, 16><<<grid, 256, 0, stream>>>((int) n, A, lda, strideA, info, batch); break;
            default: k_potrf_group<T, 32><<<grid, 256, 0, stream>>>((int) n, A, lda, strideA, info, batch); break;
        }
    } else {
        const size_t bytes = n * n * sizeof(T);
        const int use_smem = bytes <= (size_t) ctx->max_smem_optin - 1024 ? 1 : 0;
        const size_t smem = use_smem ? bytes : 0;
        if (smem > 48 * 1024)
            GPUB_CUDA(cudaFuncSetAttribute(k_potrf_cta<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        const size_t cap = (size_t) ctx->sm_count * 4;
        const unsigned grid = (unsigned) (batch < cap ? batch : cap);
        k_potrf_cta<T><<<grid, 256, smem, stream>>>((int) n, A, lda, strideA, info, batch, use_smem);
    }
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

template<typename T>
int potrs_batched(gpub_ctx_t ctx, int sidx, size_t n, const T *L, size_t ldl, size_t strideL, T *b, size_t strideB, size_t batch) {
    if (n == 0 || batch == 0) return GPUB_OK;
    if (!L || !b || ldl < n) return GPUB_EINVAL;
    if (n > 8192) return GPUB_ENOTSUP;
    GPUB_ENTER(ctx, sidx);
    if (n <= 32) {
        const int np = n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32;
        const size_t groups = 256 / np;
        const size_t want = gpub_ceil_div(batch, groups);
        const size_t cap = (size_t) ctx->sm_count * 8;
        const unsigned grid = (unsigned) (want < cap ? want : cap);
        const size_t smem = groups * np * (np + 1) * sizeof(T);
#define GPUB_POTRS_CASE(NPV)                                                                                         \
    {                                                                                                                \
        auto kern = k_potrs_group<T, NPV>;                                                                           \
        if (smem > 48 * 1024) GPUB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
        kern<<<grid, 256, smem, stream>>>((int) n, L, ldl, strideL, b, strideB, batch);                              \
    }
        switch (np) {
            case 4: GPUB_POTRS_CASE(4) break;
            case 8: GPUB_POTRS_CASE(8) break;
            case 16: GPUB_POTRS_CASE(16) break;
            default: GPUB_POTRS_CASE(32) break;
        }
#undef GPUB_POTRS_CASE
    } else {
        const size_t smem = n * sizeof(T);
        if (smem > 48 * 1024)
            GPUB_CUDA(cudaFuncSetAttribute(k_potrs_cta<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        const size_t cap = (size_t) ctx->sm_count * 8;
        const unsigned grid = (unsigned) (batch < cap ? batch : cap);
        k_potrs_cta<T><<<grid, 256, smem, stream>>>((int) n, L, ldl, strideL, b, strideB, batch);
    }
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

} // namespace

extern "C" {

int gpub_potrf_batched_f64(gpub_ctx_t c, int s, size_t n, double *A, size_t lda, size_t sA, int *info, size_t b) { return potrf_batched<double>(c, s, n, A, lda, sA, info, b); }
int gpub_potrf_batched_f32(gpub_ctx_t c, int s, size_t n, float *A, size_t lda, size_t sA, int *info, size_t b) { return potrf_batched<float>(c, s, n, A, lda, sA, info, b); }
int gpub_potrs_batched_f64(gpub_ctx_t c, int s, size_t n, const double *L, size_t ldl, size_t sL, double *b, size_t sB, size_t bt) { return potrs_batched<double>(c, s, n, L, ldl, sL, b, sB, bt); }
int gpub_potrs_batched_f32(gpub_ctx_t c, int s, size_t n, const float *L, size_t ldl, size_t sL, float *b, size_t sB, size_t bt) { return potrs_batched<float>(c, s, n, L, ldl, sL, b, sB, bt); }

} // extern "C"
