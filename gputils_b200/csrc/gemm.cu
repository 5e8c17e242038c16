// Batched GEMM  C_i <- beta*C_i + alpha*A_i*B_i  (column-major, NN).
// Replaces cublas{D,S}gemmBatched / cublas{D,S}gemm (ref: tensor.cuh:1286-1338).
//
// Three kernel families, picked by shape in launch():
//   k_gemm_small  : m, n, k <= 32 and densely packed batches -- HBM-bound (AI = 2n/(3s) flop/B).
//                   A CTA streams a chunk of whole matrices through shared memory with 128-bit loads,
//                   each thread owns RT rows x 1 column of one C_i in registers.
//   k_gemm_dmma   : fp64, tile 64x64 per CTA, FP64 tensor-core tiles (mma.sync m8n8k4 -> DMMA.8x8x4),
//                   cp.async double-buffered K panels. Used once the problem is a dense contraction.
//   k_gemm_tiled  : everything else (any m, n, k, ld, stride), 64x64 tile, 4x4 per thread, FFMA/DFMA.
// FP32 never uses TF32: products and sums are plain FFMA so results track cuBLAS' SGEMM.
// C may alias B (Nullspace::project, tensor.cuh:2084): the small kernel reads every operand of a
// chunk before writing; the tiled kernels go through a stream-ordered scratch copy of B in that case.
#include "common.cuh"
#ifndef GPUB_GRID_WAVES
#define GPUB_GRID_WAVES 2   // persistent grids: resident CTAs per SM x SM count x this
#endif

namespace {

// ------------------------------------------------------------------------------------------
// small matrices: chunk of whole matrices per CTA
// ------------------------------------------------------------------------------------------
inline bool ranges_overlap(const void *p, size_t pbytes, const void *q, size_t qbytes) {
    uintptr_t a = (uintptr_t) p, b = (uintptr_t) q;
    return a < b + qbytes && b < a + pbytes;
}

template<typename T> struct Vec16;
template<> struct Vec16<double> { using type = double2; static constexpr int N = 2; };
template<> struct Vec16<float> { using type = float4; static constexpr int N = 4; };

__device__ __forceinline__ void unpack_v(const double2 &v, double *o) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void unpack_v(const float4 &v, float *o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ double2 pack_v(const double *o) { return make_double2(o[0], o[1]); }
__device__ __forceinline__ float4 pack_v(const float *o) { return make_float4(o[0], o[1], o[2], o[3]); }

// cooperative copy of `count` contiguous elements global -> shared (128-bit when both are aligned)
template<typename T>
__device__ __forceinline__ void stage_in(T *dst, const T *__restrict__ src, size_t count, int tid, int nthreads) {
    using V = typename Vec16<T>::type;
    constexpr int VN = Vec16<T>::N;
    if (((((uintptr_t) src) | ((uintptr_t) dst)) & 15u) == 0) {
        size_t nv = count / VN;
        const V *s = reinterpret_cast<const V *>(src);
        V *d = reinterpret_cast<V *>(dst);
        for (size_t i = tid; i < nv; i += nthreads) d[i] = s[i];
        for (size_t i = nv * VN + tid; i < count; i += nthreads) dst[i] = src[i];
    } else {
        for (size_t i = tid; i < count; i += nthreads) dst[i] = src[i];
    }
}

template<typename T>
__device__ __forceinline__ void stage_out(T *__restrict__ dst, const T *src, size_t count, int tid, int nthreads) {
    using V = typename Vec16<T>::type;
    constexpr int VN = Vec16<T>::N;
    if (((((uintptr_t) src) | ((uintptr_t) dst)) & 15u) == 0) {
        size_t nv = count / VN;
        const V *s = reinterpret_cast<const V *>(src);
        V *d = reinterpret_cast<V *>(dst);
        for (size_t i = tid; i < nv; i += nthreads) d[i] = s[i];
        for (size_t i = nv * VN + tid; i < count; i += nthreads) dst[i] = src[i];
    } else {
        for (size_t i = tid; i < count; i += nthreads) dst[i] = src[i];
    }
}

// RT = rows of C kept in registers per thread (>= m, power of two); one thread = one column of one C_i.
// Shared layout per chunk: A (mpc * m*k), B (mpc * k*n), C (mpc * m*n) exactly as in global memory.
template<typename T, int RT>
__global__ void __launch_bounds__(256) k_gemm_small(int m, int n, int k, T alpha, const T *__restrict__ A,
                                                     const T *__restrict__ B, T beta, T *C, size_t batch, int mpc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    const size_t szA = (size_t) m * k, szB = (size_t) k * n, szC = (size_t) m * n;
    // keep each region 16-byte aligned
    const size_t offB = (szA * mpc * sizeof(T) + 15) / 16 * 16 / sizeof(T);
    const size_t offC = offB + (szB * mpc * sizeof(T) + 15) / 16 * 16 / sizeof(T);
    T *sB = sA + offB;
    T *sC = sA + offC;
    const int tid = threadIdx.x;
    const size_t nchunks = (batch + mpc - 1) / mpc;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const size_t first = ch * mpc;
        const int cnt = (int) ((batch - first) < (size_t) mpc ? (batch - first) : (size_t) mpc);
        stage_in(sA, A + first * szA, szA * cnt, tid, blockDim.x);
        stage_in(sB, B + first * szB, szB * cnt, tid, blockDim.x);
        if (beta != T(0)) stage_in(sC, C + first * szC, szC * cnt, tid, blockDim.x);
        __syncthreads();
        for (int w = tid; w < cnt * n; w += blockDim.x) {
            const int q = w / n, j = w - q * n;
            const T *a = sA + (size_t) q * szA;
            const T *b = sB + (size_t) q * szB + (size_t) j * k;
            T acc[RT];
#pragma unroll
            for (int r = 0; r < RT; r++) acc[r] = T(0);
            for (int kk = 0; kk < k; kk++) {
                const T bv = b[kk];
                const T *ac = a + (size_t) kk * m;
#pragma unroll
                for (int r = 0; r < RT; r++)
                    if (r < m) acc[r] = fma(ac[r], bv, acc[r]);
            }
            T *c = sC + (size_t) q * szC + (size_t) j * m;
            if (beta == T(0)) {
#pragma unroll
                for (int r = 0; r < RT; r++)
                    if (r < m) c[r] = alpha * acc[r];
            } else {
#pragma unroll
                for (int r = 0; r < RT; r++)
                    if (r < m) c[r] = alpha * acc[r] + beta * c[r];
            }
        }
        __syncthreads();
        stage_out(C + first * szC, sC, szC * cnt, tid, blockDim.x);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// square small matrices, N in {4, 8, 16, 32}, dense batch: k_gemm_col<T, N>
// N lanes per matrix, lane = one column of C (32/N matrices per warp, warps fully independent: no CTA
// barrier anywhere). Per warp iteration the 32/N matrices are one contiguous 32*N-element chunk of A, B, C:
//   A : cp.async (16 B per lane, coalesced) into the warp's double-buffered shared slot; the next chunk is
//       in flight while the current one is multiplied. Each matrix gets 16 B of padding so that the 32/N
//       different matrices read in one LDS.128 hit different banks; lanes of the same matrix broadcast.
//   B : the lane's own column, 128-bit loads straight into registers (prefetched one iteration ahead when
//       the register budget allows);  C : 128-bit stores (plus loads when beta != 0).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_ca(void *smem, const void *gmem) {
    unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_grp() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int NPEND> __device__ __forceinline__ void cp_async_wait_grp() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

template<typename T, int N>
struct GemmColCfg {
    static constexpr int VN = 16 / sizeof(T);                 // elements per 128-bit access
    static constexpr int MPW = 32 / N;                        // matrices per warp
    static constexpr int MAT_PAD = N * N + VN;                // padded shared stride of one matrix (elements)
    static constexpr int SLOT = MPW * MAT_PAD;                // one warp slot (elements)
    static constexpr int CHUNKS_PER_LANE = (N * (int) sizeof(T)) / 16; // 16-byte pieces of A per lane
    static constexpr bool PREFETCH_B = (N * sizeof(T) <= 128);
    static constexpr int WARPS = 4;
    static constexpr int MINB = (N * sizeof(T) >= 256) ? 2 : ((N * sizeof(T) >= 128) ? 4 : 6);
};

template<typename T, int N>
__global__ void __launch_bounds__(GemmColCfg<T, N>::WARPS * 32, GemmColCfg<T, N>::MINB)
k_gemm_col(T alpha, const T *__restrict__ A, const T *__restrict__ B, T beta, T *C, size_t batch) {
    using Cfg = GemmColCfg<T, N>;
    using V = typename Vec16<T>::type;
    constexpr int VN = Cfg::VN, MPW = Cfg::MPW, NV = N / VN;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T (*s_a)[2][Cfg::SLOT] = reinterpret_cast<T (*)[2][Cfg::SLOT]>(smem_raw); // [warp][buffer][slot]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane / N;
    const size_t nwarps = (size_t) gridDim.x * Cfg::WARPS;
    const size_t wg = (size_t) blockIdx.x * Cfg::WARPS + warp;
    const size_t nchunks = (batch + MPW - 1) / MPW;
    const size_t iters = (nchunks + nwarps - 1) / nwarps;
    const size_t total_el = batch * (size_t) (N * N);

    auto issue_a = [&](size_t chunk, int buf) {
        // the warp's chunk is 32*N contiguous elements; lane l copies 16-byte pieces l, l+32, ...
        const size_t base = chunk * (size_t) (MPW * N * N);
#pragma unroll
        for (int p = 0; p < Cfg::CHUNKS_PER_LANE; p++) {
            const int piece = lane + 32 * p;                 // 16-byte piece inside the chunk
            const int el = piece * VN;                       // element offset inside the chunk
            const int mq = el / (N * N), within = el - mq * (N * N);
            if (base + el < total_el) cp_async16_ca(&s_a[warp][buf][mq * Cfg::MAT_PAD + within], A + base + el);
        }
        cp_async_commit_grp();
    };
    auto load_col = [&](const T *src, size_t chunk, T *dst) {
        const size_t off = chunk * (size_t) (MPW * N * N) + (size_t) lane * N;
        if (off < total_el) {
            const V *p = reinterpret_cast<const V *>(src + off);
#pragma unroll
            for (int v = 0; v < NV; v++) unpack_v(p[v], dst + v * VN);
        } else {
#pragma unroll
            for (int r = 0; r < N; r++) dst[r] = T(0);
        }
    };

    T b_cur[N], b_nxt[Cfg::PREFETCH_B ? N : 1];
    if (iters > 0) {
        issue_a(wg, 0);
        if (Cfg::PREFETCH_B) load_col(B, wg, b_nxt);
    }
    for (size_t it = 0; it < iters; it++) {
        const size_t chunk = it * nwarps + wg;
        const int buf = (int) (it & 1);
        if (Cfg::PREFETCH_B) {
#pragma unroll
            for (int r = 0; r < N; r++) b_cur[r] = b_nxt[r];
        } else {
            load_col(B, chunk, b_cur);
        }
        if (it + 1 < iters) {
            issue_a(chunk + nwarps, buf ^ 1);
            if (Cfg::PREFETCH_B) load_col(B, chunk + nwarps, b_nxt);
            cp_async_wait_grp<1>();
        } else {
            cp_async_wait_grp<0>();
        }
        __syncwarp();
        const T *a = &s_a[warp][buf][q * Cfg::MAT_PAD];
        T acc[N];
#pragma unroll
        for (int r = 0; r < N; r++) acc[r] = T(0);
#pragma unroll
        for (int k = 0; k < N; k++) {
            const T bk = b_cur[k];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                T av[VN];
                unpack_v(*reinterpret_cast<const V *>(a + k * N + v * VN), av);
#pragma unroll
                for (int e = 0; e < VN; e++) acc[v * VN + e] = fma(av[e], bk, acc[v * VN + e]);
            }
        }
        __syncwarp(); // every lane is done with slot `buf` before the copy two iterations ahead refills it
        const size_t off = chunk * (size_t) (MPW * N * N) + (size_t) lane * N;
        if (off < total_el) {
            V *cp = reinterpret_cast<V *>(C + off);
            if (beta == T(0)) {
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    T o[VN];
#pragma unroll
                    for (int e = 0; e < VN; e++) o[e] = alpha * acc[v * VN + e];
                    cp[v] = pack_v(o);
                }
            } else {
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    T o[VN];
                    unpack_v(cp[v], o);
#pragma unroll
                    for (int e = 0; e < VN; e++) o[e] = alpha * acc[v * VN + e] + beta * o[e];
                    cp[v] = pack_v(o);
                }
            }
        }
    }
}

template<typename T, int N>
int launch_col(gpub_ctx_t ctx, cudaStream_t stream, T alpha, const T *A, const T *B, T beta, T *C, size_t batch) {
    using Cfg = GemmColCfg<T, N>;
    const size_t nchunks = gpub_ceil_div(batch, (size_t) Cfg::MPW);
    const size_t want = gpub_ceil_div(nchunks, (size_t) Cfg::WARPS);
    const size_t cap = (size_t) ctx->sm_count * Cfg::MINB * 4;   // measured: 4 waves beat 2 by 5-10 % for n = 4, 8
    const unsigned grid = (unsigned) (want < cap ? want : cap);
    const size_t smem = sizeof(T) * Cfg::WARPS * 2 * Cfg::SLOT;
    if (smem > 48 * 1024)
        GPUB_CUDA(cudaFuncSetAttribute(k_gemm_col<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    k_gemm_col<T, N><<<grid, Cfg::WARPS * 32, smem, stream>>>(alpha, A, B, beta, C, batch);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

// ------------------------------------------------------------------------------------------
// generic tiled kernel: 64x64 tile of one C_i per CTA, 256 threads, 4x4 outputs per thread
// ------------------------------------------------------------------------------------------
constexpr int TB = 64; // tile edge
constexpr int TK = 16; // K panel

template<typename T>
__global__ void __launch_bounds__(256) k_gemm_tiled(size_t m, size_t n, size_t k, T alpha, const T *__restrict__ A, size_t lda,
                                                     size_t sA, const T *__restrict__ B, size_t ldb, size_t sB, T beta, T *C,
                                                     size_t ldc, size_t sC, size_t tiles_m, size_t tiles_n, size_t batch) {
    __shared__ T As[TK][TB + 4]; // As[kk][row]
    __shared__ T Bs[TK][TB + 4]; // Bs[kk][col]
    const size_t tiles = tiles_m * tiles_n;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4; // tx -> rows, ty -> cols
    for (size_t t = blockIdx.x; t < tiles * batch; t += gridDim.x) {
        const size_t b = t / tiles, r = t - b * tiles;
        const size_t row0 = (r % tiles_m) * TB, col0 = (r / tiles_m) * TB;
        const T *a = A + b * sA;
        const T *bb = B + b * sB;
        T *c = C + b * sC;
        T acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = T(0);
        for (size_t k0 = 0; k0 < k; k0 += TK) {
            // A panel: 64 rows x 16 k ; thread loads 4 elements (consecutive rows => coalesced)
#pragma unroll
            for (int l = 0; l < 4; l++) {
                int e = threadIdx.x + l * 256;
                int rr = e & 63, kk = e >> 6;
                size_t gr = row0 + rr, gk = k0 + kk;
                As[kk][rr] = (gr < m && gk < k) ? a[gr + gk * lda] : T(0);
            }
            // B panel: 16 k x 64 cols ; consecutive threads walk k (contiguous in column-major B)
#pragma unroll
            for (int l = 0; l < 4; l++) {
                int e = threadIdx.x + l * 256;
                int kk = e & 15, cc = e >> 4;
                size_t gk = k0 + kk, gc = col0 + cc;
                Bs[kk][cc] = (gk < k && gc < n) ? bb[gk + gc * ldb] : T(0);
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TK; kk++) {
                T av[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; i++) av[i] = As[kk][tx + 16 * i];
#pragma unroll
                for (int j = 0; j < 4; j++) bv[j] = Bs[kk][ty + 16 * j];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            size_t gc = col0 + ty + 16 * j;
            if (gc >= n) continue;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                size_t gr = row0 + tx + 16 * i;
                if (gr >= m) continue;
                T *p = c + gr + gc * ldc;
                *p = (beta == T(0)) ? alpha * acc[i][j] : alpha * acc[i][j] + beta * (*p);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// fp64 tensor-core kernel: 64x64 tile per CTA, 8 warps (4 x 2), each warp 16 x 32 of C
// = 2 x 4 DMMA tiles (m8n8k4). K panels of 16 staged with cp.async, double buffered.
// Requires m, n multiples of 64 and k a multiple of 16 (the launcher checks).
// ------------------------------------------------------------------------------------------
constexpr int DK = 16;
constexpr int DLD = 64 + 4; // leading dimension == 4 (mod 16) doubles: the 16 lanes of a half-warp hit 16 distinct bank pairs

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// TRB: the second operand is given transposed (B = Bt^T with Bt n x k column-major, ldb its leading dimension): the
// panel then lands as [kk][col] like A's and its fragments are read with A's conflict-free pattern (P = N N^T).
// DKT: depth of a K panel (16 or 32; 32 halves the number of CTA barriers per tile and needs dynamic shared memory)
// WARPS: 8 warps of 16 x 32 warp tiles (6 fragment loads per 8 DMMA) or 4 warps of 32 x 32 (8 per 16)
// SYM (with TRB, A == B): C = A A^T is symmetric -- only the tiles on and below the diagonal are computed, each off-diagonal tile is
// stored twice (itself and mirrored), which halves the flops of the Nullspace projector N N^T
// PROJ (with TRB and SYM): the Nullspace projector P = N N^T from the orthogonal factor it is cut from. N = [ U(:, r:n) | 0 ], so
// N N^T = U2 U2^T = I - U1 U1^T with U1 = U(:, 0:r): per matrix the kernel takes whichever side has FEWER columns (rank r read from
// the device) -- A = N with k = n - r columns, or Ualt = U with k = r columns, alpha = -1 and the identity added in the epilogue.
// For a fat 128 x 1024 matrix (r = 128) that is 128 instead of 896 columns: 7 x fewer flops for the same projector.
// EPI_E (plain NN): C = E + alpha A B with E = blockdiag(Ublk, I) generated in the epilogue (Ublk: ne x ne per matrix, passed through the
// `Ualt` / `sU` / `rank`-less parameters): the U assembly of Svd / Nullspace then neither initialises nor re-reads the m x m result.
template<bool TRB, int DKT, int WARPS = 8, bool SYM = false, bool PROJ = false, bool EPI_E = false>
__global__ void __launch_bounds__(32 * WARPS) k_gemm_dmma(size_t m, size_t n, size_t k, double alpha, const double *__restrict__ A,
                                                    size_t lda, size_t sA, const double *__restrict__ B, size_t ldb, size_t sB,
                                                    double beta, double *C, size_t ldc, size_t sC, size_t tiles_m,
                                                    size_t tiles_n, size_t batch, const unsigned *__restrict__ rank = nullptr,
                                                    const double *__restrict__ Ualt = nullptr, size_t sU = 0, size_t ne = 0) {
    // As[buf][kk][row] (A panel, 16 x 64), Bs[buf][col][kk] (B panel stored k-contiguous per column)
    constexpr int BROWS = TRB ? DKT : 64, BLD = TRB ? DLD : DKT + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double (*As)[DKT][DLD] = reinterpret_cast<double (*)[DKT][DLD]>(smem_raw);
    double (*Bs)[BROWS][BLD] = reinterpret_cast<double (*)[BROWS][BLD]>(smem_raw + sizeof(double) * 2 * DKT * DLD);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NTH = 32 * WARPS, WM = WARPS == 8 ? 16 : 32, MI = WM / 8, WR = 64 / WM;
    const int wr = (warp % WR) * WM, wc = (warp / WR) * 32; // warp origin inside the tile
    const int g = lane >> 2, q = lane & 3;                 // DMMA fragment coordinates
    const size_t tiles = SYM ? tiles_m * (tiles_m + 1) / 2 : tiles_m * tiles_n;
    for (size_t t = blockIdx.x; t < tiles * batch; t += gridDim.x) {
        const size_t b = t / tiles, r = t - b * tiles;
        size_t row0, col0;
        if (SYM) {   // r enumerates the lower-triangular tiles column by column
            size_t rr = r, tj = 0;
            while (rr >= tiles_m - tj) { rr -= tiles_m - tj; tj++; }
            row0 = (tj + rr) * 64;
            col0 = tj * 64;
        } else {
            row0 = (r % tiles_m) * 64;
            col0 = (r / tiles_m) * 64;
        }
        const double *a = A + b * sA + row0;
        const double *bb = B + b * sB + (TRB ? col0 : col0 * ldb);
        size_t kcount = k;
        bool comp = false;
        int klead = 0;                                    // PROJ, U2 side: columns of the first panel that still belong to U1
        if (PROJ) {
            const size_t r = rank[b] < n ? rank[b] : n;
            comp = r <= n - r;
            kcount = comp ? r : n - r;
            // both sides are read from U itself (the packed basis N is not an input: its packing can run beside this kernel).
            // U1 = columns 0 .. r-1, panels from column 0, the tail of the last panel masked; U2 = columns r .. n-1, read as whole
            // panels that END at column n (so nothing is read past the matrix), the head of the first panel masked
            const size_t c0 = comp ? 0 : n - ((kcount + DKT - 1) / DKT) * DKT;
            klead = comp ? 0 : (int) (r - c0);
            a = Ualt + b * sU + row0 + c0 * lda;
            bb = Ualt + b * sU + col0 + c0 * ldb;
            alpha = comp ? -1.0 : 1.0;
        }
        double acc[MI][4][2];
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        auto load_panel = [&](int buf, size_t k0) {
            // A: DKT columns of 64 rows = 32 DKT double2 ; 256 threads x DKT / 8
#pragma unroll
            for (int l = 0; l < 32 * DKT / NTH; l++) {
                int e = threadIdx.x + l * NTH;
                int rr = (e & 31) * 2, kk = e >> 5;
                cp_async16(&As[buf][kk][rr], a + rr + (k0 + kk) * lda);
            }
            // B: 64 columns of DKT k = 32 DKT double2
#pragma unroll
            for (int l = 0; l < 32 * DKT / NTH; l++) {
                int e = threadIdx.x + l * NTH;
                if (TRB) {
                    int cc = (e & 31) * 2, kk = e >> 5;
                    cp_async16(&Bs[buf][kk][cc], bb + cc + (k0 + kk) * ldb);
                } else {
                    int kk = (e % (DKT / 2)) * 2, cc = e / (DKT / 2);
                    cp_async16(&Bs[buf][cc][kk], bb + (k0 + kk) + (size_t) cc * ldb);
                }
            }
            cp_async_commit();
        };

        const size_t npan = PROJ ? (kcount + DKT - 1) / DKT : k / DKT;
        if (npan > 0) load_panel(0, 0);
        for (size_t p = 0; p < npan; p++) {
            const int buf = (int) (p & 1);
            if (p + 1 < npan) {
                load_panel(buf ^ 1, (p + 1) * DKT);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            if (PROJ && comp && p + 1 == npan && kcount % DKT != 0) {
                // the columns of U beyond the rank that came along with the last panel do not belong to U1
                const int kv = (int) (kcount % DKT);
                for (int e = threadIdx.x; e < (DKT - kv) * 64; e += NTH) {
                    As[buf][kv + e / 64][e % 64] = 0.0;
                    Bs[buf][kv + e / 64][e % 64] = 0.0;
                }
                __syncthreads();
            }
            if (PROJ && !comp && p == 0 && klead > 0) {
                // ... and the columns before the rank that came along with the first panel do not belong to U2
                for (int e = threadIdx.x; e < klead * 64; e += NTH) {
                    As[buf][e / 64][e % 64] = 0.0;
                    Bs[buf][e / 64][e % 64] = 0.0;
                }
                __syncthreads();
            }
#pragma unroll
            for (int k4 = 0; k4 < DKT; k4 += 4) {
                double af[MI], bf[4];
#pragma unroll
                for (int i = 0; i < MI; i++) af[i] = As[buf][k4 + q][wr + 8 * i + g]; // A(row g, k q)
#pragma unroll
                for (int j = 0; j < 4; j++) bf[j] = TRB ? Bs[buf][k4 + q][wc + 8 * j + g] : Bs[buf][wc + 8 * j + g][k4 + q]; // B(k q, col g)
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncthreads();
        }
        double *c = C + b * sC;
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                size_t gr = row0 + wr + 8 * i + g;
                size_t gc = col0 + wc + 8 * j + 2 * q; // C fragment: row g, cols 2q, 2q+1
                double *p0 = c + gr + gc * ldc;
                double *p1 = p0 + ldc;
                if (PROJ) {
                    *p0 = alpha * acc[i][j][0] + ((comp && gr == gc) ? 1.0 : 0.0);
                    *p1 = alpha * acc[i][j][1] + ((comp && gr == gc + 1) ? 1.0 : 0.0);
                } else if (EPI_E) {
                    const double *ub = Ualt + b * sU;
                    const double e0 = (gr < ne && gc < ne) ? ub[gr + gc * ne] : (gr == gc ? 1.0 : 0.0);
                    const double e1 = (gr < ne && gc + 1 < ne) ? ub[gr + (gc + 1) * ne] : (gr == gc + 1 ? 1.0 : 0.0);
                    *p0 = fma(alpha, acc[i][j][0], e0);
                    *p1 = fma(alpha, acc[i][j][1], e1);
                } else if (beta == 0.0) {
                    *p0 = alpha * acc[i][j][0];
                    *p1 = alpha * acc[i][j][1];
                } else {
                    *p0 = alpha * acc[i][j][0] + beta * (*p0);
                    *p1 = alpha * acc[i][j][1] + beta * (*p1);
                }
                if (SYM && row0 != col0) {   // the mirrored tile (beta == 0 on this path)
                    c[gc + gr * ldc] = alpha * acc[i][j][0];
                    c[gc + 1 + gr * ldc] = alpha * acc[i][j][1];
                }
            }
    }
}

// ------------------------------------------------------------------------------------------
// fp32 register-tiled kernel (plain FFMA, no TF32): k_sgemm_rt<BM, BN, TM, TN>
// CTA tile BM x BN, thread tile TM x TN, K panels of 16 double-buffered with cp.async. A panels land as
// As[k][row] (row-contiguous, read with 128-bit loads), B panels as Bs[col][k] (k-contiguous, one 128-bit
// load gives 4 consecutive k for one column), so 4 k-steps cost TM/4*4 + TN LDS.128 for 4*TM*TN FFMA.
// Requires m % BM == 0, n % BN == 0, k % 16 == 0 and 16-byte aligned operands (the launcher checks).
// ------------------------------------------------------------------------------------------
template<int BM, int BN, int TM, int TN, int MINB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), MINB) k_sgemm_rt(size_t m, size_t n, size_t k, float alpha, const float *__restrict__ A,
                                                                    size_t lda, size_t sA, const float *__restrict__ B, size_t ldb, size_t sB,
                                                                    float beta, float *C, size_t ldc, size_t sC, size_t tiles_m, size_t tiles_n,
                                                                    size_t batch) {
    constexpr int BK = 16;
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int TX = BM / TM; // threads along rows
    constexpr int LDB_S = BK + 4; // padded k-stride of the B panel (keeps 16-byte alignment, spreads banks)
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BN][LDB_S];
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const size_t tiles = tiles_m * tiles_n;
    for (size_t t = blockIdx.x; t < tiles * batch; t += gridDim.x) {
        const size_t b = t / tiles, r = t - b * tiles;
        const size_t row0 = (r % tiles_m) * BM, col0 = (r / tiles_m) * BN;
        const float *a = A + b * sA + row0;
        const float *bb = B + b * sB + col0 * ldb;
        // accumulators as row pairs: every update is an FFMA2 (fma.rn.f32x2, two IEEE FMAs per issue slot) of a natural pair
        // of A rows with a duplicated B value -- the kernel is issue-bound, FFMA2 frees the slots the LDS need
        float2 acc[TN][TM / 2];
#pragma unroll
        for (int j = 0; j < TN; j++)
#pragma unroll
            for (int i = 0; i < TM / 2; i++) acc[j][i] = make_float2(0.f, 0.f);

        auto load_panel = [&](int buf, size_t k0) {
            // A: BK columns of BM rows -> BK*BM/4 16-byte pieces
            for (int e = tid; e < BK * BM / 4; e += NT) {
                const int rr = (e % (BM / 4)) * 4, kk = e / (BM / 4);
                cp_async16(&As[buf][kk][rr], a + rr + (k0 + kk) * lda);
            }
            // B: BN columns of BK k -> BN*BK/4 pieces
            for (int e = tid; e < BN * BK / 4; e += NT) {
                const int kk = (e % (BK / 4)) * 4, cc = e / (BK / 4);
                cp_async16(&Bs[buf][cc][kk], bb + (k0 + kk) + (size_t) cc * ldb);
            }
            cp_async_commit();
        };

        const size_t npan = k / BK;
        load_panel(0, 0);
        for (size_t p = 0; p < npan; p++) {
            const int buf = (int) (p & 1);
            if (p + 1 < npan) {
                load_panel(buf ^ 1, (p + 1) * BK);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
#pragma unroll
            for (int k4 = 0; k4 < BK; k4 += 4) {
                float bv[TN][4];
#pragma unroll
                for (int j = 0; j < TN; j++) {
                    const float4 v = *reinterpret_cast<const float4 *>(&Bs[buf][ty * TN + j][k4]);
                    bv[j][0] = v.x; bv[j][1] = v.y; bv[j][2] = v.z; bv[j][3] = v.w;
                }
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    float2 av[TM / 2];
#pragma unroll
                    for (int i = 0; i < TM; i += 4) {
                        // rows of a thread are TM/4 groups of 4, the groups BM/(TM/4) apart: consecutive threads read
                        // consecutive 16-byte pieces (conflict-free)
                        const float4 v = *reinterpret_cast<const float4 *>(&As[buf][k4 + kk][(i / 4) * (BM / (TM / 4)) + tx * 4]);
                        av[i / 2] = make_float2(v.x, v.y);
                        av[i / 2 + 1] = make_float2(v.z, v.w);
                    }
#pragma unroll
                    for (int j = 0; j < TN; j++) {
                        const float2 bb = make_float2(bv[j][kk], bv[j][kk]);
#pragma unroll
                        for (int i = 0; i < TM / 2; i++) acc[j][i] = __ffma2_rn(av[i], bb, acc[j][i]);
                    }
                }
            }
            __syncthreads();
        }
        float *c = C + b * sC + row0 + tx * 4 + (col0 + ty * TN) * ldc;
#pragma unroll
        for (int j = 0; j < TN; j++) {
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 *p = reinterpret_cast<float4 *>(c + (i / 4) * (BM / (TM / 4)) + (size_t) j * ldc);
                float4 o;
                const float2 a01 = acc[j][i / 2], a23 = acc[j][i / 2 + 1];
                if (beta == 0.f) {
                    o = make_float4(alpha * a01.x, alpha * a01.y, alpha * a23.x, alpha * a23.y);
                } else {
                    const float4 old = *p;
                    o = make_float4(alpha * a01.x + beta * old.x, alpha * a01.y + beta * old.y, alpha * a23.x + beta * old.z,
                                    alpha * a23.y + beta * old.w);
                }
                *p = o;
            }
        }
    }
}

#ifndef GPUB_DMMA_WARPS
#define GPUB_DMMA_WARPS 4
#endif
#ifndef GPUB_DMMA_DKT32
#define GPUB_DMMA_DKT32 0
#endif
#ifndef GPUB_SGEMM64_TN
#define GPUB_SGEMM64_TN 8
#endif
#ifndef GPUB_SGEMM64_MINB
#define GPUB_SGEMM64_MINB 6
#endif
template<typename T>
bool try_sgemm_rt(gpub_ctx_t, cudaStream_t, size_t, size_t, size_t, T, const T *, size_t, size_t, const T *, size_t, size_t, T, T *,
                  size_t, size_t, size_t) { return false; }

template<>
bool try_sgemm_rt<float>(gpub_ctx_t ctx, cudaStream_t stream, size_t m, size_t n, size_t k, float alpha, const float *A, size_t lda,
                         size_t sA, const float *B, size_t ldb, size_t sB, float beta, float *C, size_t ldc, size_t sC, size_t batch) {
    const bool al = ((((uintptr_t) A) | ((uintptr_t) B) | ((uintptr_t) C)) % 16 == 0) && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 &&
                    sA % 4 == 0 && sB % 4 == 0 && sC % 4 == 0;
    if (!al || k % 16 != 0 || k < 16) return false;
    const size_t cap = (size_t) ctx->sm_count * 8;
    const size_t cap2 = (size_t) ctx->sm_count * GPUB_SGEMM64_MINB * 2;
    (void) cap2;
    if (m % 128 == 0 && n % 128 == 0) {
        const size_t tm = m / 128, tn = n / 128, total = tm * tn * batch;
        k_sgemm_rt<128, 128, 8, 8, 2><<<(unsigned) (total < cap ? total : cap), 256, 0, stream>>>(m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, tm, tn, batch);
        return true;
    }
    if (m % 64 == 0 && n % 64 == 0) {
        const size_t tm = m / 64, tn = n / 64, total = tm * tn * batch;
#if GPUB_SGEMM64_TN == 8
        // 8 x 8 thread tiles on 64 threads: 16 LDS.128 per 128 FFMA2 instead of 12 per 64 with the 8 x 4 tile (0.875 -> 0.78 ms at n = 64)
        k_sgemm_rt<64, 64, 8, 8, GPUB_SGEMM64_MINB><<<(unsigned) (total < cap2 ? total : cap2), 64, 0, stream>>>(m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, tm, tn, batch);
#else
        k_sgemm_rt<64, 64, 8, 4, 1><<<(unsigned) (total < cap ? total : cap), 128, 0, stream>>>(m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, tm, tn, batch);
#endif
        return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------
// fp64, N = 16 or 32, dense batch: k_dgemm_frag<N> -- one warp per matrix, FP64 tensor-core tiles fed
// straight from global memory. The m8n8k4 operand fragments map onto column-major storage so that every
// LDG.64 of a warp reads whole 32-byte sectors (A fragment: 8 consecutive rows x 4 columns, B fragment:
// 4 consecutive k x 8 columns); nothing is staged in shared memory and there is no barrier. The k loop is
// fully unrolled so the loads of later k-steps are in flight while earlier DMMAs issue.
// ------------------------------------------------------------------------------------------
template<int N>
__global__ void __launch_bounds__(128, N == 32 ? 3 : 6) k_dgemm_frag(double alpha, const double *__restrict__ A, const double *__restrict__ B,
                                                                     double beta, double *C, size_t batch) {
    constexpr int NT = N / 8, KS = N / 4;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const size_t nwarps = ((size_t) gridDim.x * blockDim.x) >> 5;
    const size_t wg = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t iters = (batch + nwarps - 1) / nwarps;
    for (size_t it = 0; it < iters; it++) {
        size_t mat = it * nwarps + wg;
        const bool live = mat < batch;
        if (!live) mat = batch - 1;
        const double *a = A + mat * (size_t) (N * N) + g + (size_t) q * N;       // A(8i + g, 4ks + q)
        const double *b = B + mat * (size_t) (N * N) + q + (size_t) g * N;       // B(4ks + q, 8j + g)
        double acc[NT][NT][2];
#pragma unroll
        for (int i = 0; i < NT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            double af[NT], bf[NT];
#pragma unroll
            for (int i = 0; i < NT; i++) af[i] = a[8 * i + (size_t) (4 * ks) * N];
#pragma unroll
            for (int j = 0; j < NT; j++) bf[j] = b[4 * ks + (size_t) (8 * j) * N];
#pragma unroll
            for (int i = 0; i < NT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (live) {
            double *c = C + mat * (size_t) (N * N) + g + (size_t) (2 * q) * N;  // C(8i + g, 8j + 2q + e)
#pragma unroll
            for (int i = 0; i < NT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double *p = c + 8 * i + (size_t) (8 * j + e) * N;
                        *p = (beta == 0.0) ? alpha * acc[i][j][e] : alpha * acc[i][j][e] + beta * (*p);
                    }
        }
    }
}

template<typename T, int N>
bool try_dgemm_frag(gpub_ctx_t, cudaStream_t, T, const T *, const T *, T, T *, size_t) { return false; }
template<>
bool try_dgemm_frag<double, 16>(gpub_ctx_t ctx, cudaStream_t stream, double alpha, const double *A, const double *B, double beta, double *C, size_t batch) {
    const size_t want = gpub_ceil_div(batch, 4), cap = (size_t) ctx->sm_count * 12;
    k_dgemm_frag<16><<<(unsigned) (want < cap ? want : cap), 128, 0, stream>>>(alpha, A, B, beta, C, batch);
    return true;
}
template<>
bool try_dgemm_frag<double, 32>(gpub_ctx_t ctx, cudaStream_t stream, double alpha, const double *A, const double *B, double beta, double *C, size_t batch) {
    const size_t want = gpub_ceil_div(batch, 4), cap = (size_t) ctx->sm_count * 6;
    k_dgemm_frag<32><<<(unsigned) (want < cap ? want : cap), 128, 0, stream>>>(alpha, A, B, beta, C, batch);
    return true;
}

// ------------------------------------------------------------------------------------------
// fp32, N = 32, dense batch: k_sgemm32_rt -- one warp per matrix, 4 x 8 register tile per lane.
// k_gemm_col<float, 32> (lane = column, 8 broadcast LDS.128 of A per 32 FFMA) is bound by the shared-memory
// to register-file path; here lane (rg, cg) owns rows 4rg..4rg+3 and columns cg, cg+4, ..., cg+28, so four
// k-steps cost 4 LDS.128 of A (rows contiguous) + 8 LDS.128 of B (4 consecutive k of one column) for 128 FFMA.
// A and B land in the warp's double-buffered shared slot through cp.async (B columns padded to 36 floats:
// the four column groups of a quarter-warp phase hit different banks); the next matrix is in flight while the
// current one is multiplied; C leaves as 128-bit stores that cover whole 128-byte lines. No CTA barrier.
// ------------------------------------------------------------------------------------------
constexpr int S32_LDB = 36;
constexpr int S32_SLOT = 32 * 32 + 32 * S32_LDB; // floats per buffer: A then padded B
constexpr int S32_WARPS = 4;

__global__ void __launch_bounds__(S32_WARPS * 32, 3) k_sgemm32_rt(float alpha, const float *__restrict__ A, const float *__restrict__ B,
                                                                   float beta, float *C, size_t batch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float (*s_buf)[2][S32_SLOT] = reinterpret_cast<float (*)[2][S32_SLOT]>(smem_raw); // [warp][buffer][slot]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rg = lane & 7, cg = lane >> 3;
    const size_t nwarps = (size_t) gridDim.x * S32_WARPS;
    const size_t wg = (size_t) blockIdx.x * S32_WARPS + warp;

    auto issue = [&](size_t mat, int buf) {
        float *sa = s_buf[warp][buf], *sb = sa + 1024;
        const float *a = A + mat * 1024, *b = B + mat * 1024;
#pragma unroll
        for (int p = 0; p < 8; p++) {
            const int e = (lane + 32 * p) * 4;
            cp_async16_ca(sa + e, a + e);
            cp_async16_ca(sb + (e >> 5) * S32_LDB + (e & 31), b + e);
        }
        cp_async_commit_grp();
    };

    if (wg < batch) issue(wg, 0);
    int buf = 0;
    for (size_t mat = wg; mat < batch; mat += nwarps, buf ^= 1) {
        if (mat + nwarps < batch) {
            issue(mat + nwarps, buf ^ 1);
            cp_async_wait_grp<1>();
        } else {
            cp_async_wait_grp<0>();
        }
        __syncwarp();
        const float *sa = s_buf[warp][buf] + 4 * rg, *sb = s_buf[warp][buf] + 1024 + cg * S32_LDB;
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) acc[j][r] = 0.f;
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 4) {
            float4 bv[8];
#pragma unroll
            for (int j = 0; j < 8; j++) bv[j] = *reinterpret_cast<const float4 *>(sb + 4 * j * S32_LDB + k0);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const float4 av = *reinterpret_cast<const float4 *>(sa + (k0 + kk) * 32);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float bk = kk == 0 ? bv[j].x : (kk == 1 ? bv[j].y : (kk == 2 ? bv[j].z : bv[j].w));
                    acc[j][0] = fmaf(av.x, bk, acc[j][0]);
                    acc[j][1] = fmaf(av.y, bk, acc[j][1]);
                    acc[j][2] = fmaf(av.z, bk, acc[j][2]);
                    acc[j][3] = fmaf(av.w, bk, acc[j][3]);
                }
            }
        }
        __syncwarp(); // every lane is done with slot `buf` before the copy two iterations ahead refills it
        float *c = C + mat * 1024 + cg * 32 + 4 * rg;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float4 *p = reinterpret_cast<float4 *>(c + 4 * j * 32);
            float4 o = make_float4(alpha * acc[j][0], alpha * acc[j][1], alpha * acc[j][2], alpha * acc[j][3]);
            if (beta != 0.f) {
                const float4 old = *p;
                o.x += beta * old.x; o.y += beta * old.y; o.z += beta * old.z; o.w += beta * old.w;
            }
            *p = o;
        }
    }
}

template<typename T>
bool try_sgemm32(gpub_ctx_t, cudaStream_t, T, const T *, const T *, T, T *, size_t, int *) { return false; }
template<>
bool try_sgemm32<float>(gpub_ctx_t ctx, cudaStream_t stream, float alpha, const float *A, const float *B, float beta, float *C, size_t batch, int *err) {
    const size_t smem = sizeof(float) * S32_WARPS * 2 * S32_SLOT;
    cudaError_t e = cudaFuncSetAttribute(k_sgemm32_rt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) { *err = (int) e; return true; }
    const size_t want = gpub_ceil_div(batch, (size_t) S32_WARPS), cap = (size_t) ctx->sm_count * 3;
    k_sgemm32_rt<<<(unsigned) (want < cap ? want : cap), S32_WARPS * 32, smem, stream>>>(alpha, A, B, beta, C, batch);
    *err = GPUB_OK;
    return true;
}

// ------------------------------------------------------------------------------------------
// n == 1 (batched matrix-vector, Nullspace::project and the example's d_A * d_b): k_gemv<T>
// HBM-bound: A is read exactly once. A CTA takes 64 rows of one matrix; its 256 threads are 64 rows x 4 interleaved
// k-slices, so one pass over a column touches 512 contiguous bytes per slice and four columns are in flight per
// CTA (more with the unrolled loop). x is staged in shared memory in chunks, the four slice sums are combined
// through shared memory in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------
constexpr int GEMV_ROWS = 64, GEMV_KS = 4, GEMV_XCHUNK = 2048;

template<typename T>
__global__ void __launch_bounds__(GEMV_ROWS * GEMV_KS) k_gemv(size_t m, size_t k, T alpha, const T *__restrict__ A, size_t lda, size_t sA,
                                                               const T *__restrict__ x, size_t sX, T beta, T *y, size_t sY, size_t row_blocks,
                                                               size_t batch) {
    __shared__ T xs[GEMV_XCHUNK];
    __shared__ T part[GEMV_KS][GEMV_ROWS];
    const int r = threadIdx.x % GEMV_ROWS, ks = threadIdx.x / GEMV_ROWS;
    for (size_t t = blockIdx.x; t < batch * row_blocks; t += gridDim.x) {
        const size_t mat = t / row_blocks, rb = t - mat * row_blocks;
        const size_t row = rb * GEMV_ROWS + r;
        const bool row_ok = row < m;
        const T *a = A + mat * sA + (row_ok ? row : 0);
        const T *xg = x + mat * sX;
        T acc0 = 0, acc1 = 0;
        for (size_t k0 = 0; k0 < k; k0 += GEMV_XCHUNK) {
            const size_t kc = (k - k0) < (size_t) GEMV_XCHUNK ? (k - k0) : (size_t) GEMV_XCHUNK;
            __syncthreads();
            for (size_t j = threadIdx.x; j < kc; j += blockDim.x) xs[j] = xg[k0 + j];
            __syncthreads();
            const T *ac = a + (k0 + ks) * lda;
            size_t j = ks;
#pragma unroll 1
            for (; j + 7 * GEMV_KS < kc; j += 8 * GEMV_KS, ac += 8 * GEMV_KS * lda) {
                T v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = ac[(size_t) u * GEMV_KS * lda];
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    acc0 = fma(v[u], xs[j + u * GEMV_KS], acc0);
                    acc1 = fma(v[u + 1], xs[j + (u + 1) * GEMV_KS], acc1);
                }
            }
            for (; j < kc; j += GEMV_KS, ac += GEMV_KS * lda) acc0 = fma(*ac, xs[j], acc0);
        }
        part[ks][r] = acc0 + acc1;
        __syncthreads();
        if (ks == 0 && row_ok) {
            const T sum = (part[0][r] + part[1][r]) + (part[2][r] + part[3][r]);
            T *yp = y + mat * sY + row;
            *yp = beta == T(0) ? alpha * sum : alpha * sum + beta * (*yp);
        }
    }
}

// Vectorised variant for aligned operands (m a multiple of the CTA's row block, 16-byte aligned A, lda % VN == 0):
// a thread owns VN consecutive rows (one 128-bit load per column) and one of 8 contiguous k-slices, x comes from shared
// memory two / four values at a time. 32 x 8 threads per CTA.
constexpr int GEMV2_KS = 8;

template<typename T>
__global__ void __launch_bounds__(32 * GEMV2_KS) k_gemv_vec(size_t m, size_t k, T alpha, const T *__restrict__ A, size_t lda, size_t sA,
                                                             const T *__restrict__ x, size_t sX, T beta, T *y, size_t sY, size_t row_blocks,
                                                             size_t batch) {
    using V = typename Vec16<T>::type;
    constexpr int VN = Vec16<T>::N;
    constexpr int RB = 32 * VN;
    __shared__ __align__(16) T xs[GEMV_XCHUNK];
    __shared__ T part[GEMV2_KS][RB];
    const int r = threadIdx.x & 31, ks = threadIdx.x >> 5;
    for (size_t t = blockIdx.x; t < batch * row_blocks; t += gridDim.x) {
        const size_t mat = t / row_blocks, rb = t - mat * row_blocks;
        const T *a = A + mat * sA + rb * RB + (size_t) r * VN;
        const T *xg = x + mat * sX;
        T acc[VN];
#pragma unroll
        for (int e = 0; e < VN; e++) acc[e] = T(0);
        for (size_t k0 = 0; k0 < k; k0 += GEMV_XCHUNK) {
            const size_t kc = (k - k0) < (size_t) GEMV_XCHUNK ? (k - k0) : (size_t) GEMV_XCHUNK;
            __syncthreads();
            for (size_t j = threadIdx.x; j < kc; j += blockDim.x) xs[j] = xg[k0 + j];
            __syncthreads();
            // slice [j0, j1) of this chunk, boundaries multiples of 4 columns
            const size_t per = ((kc + GEMV2_KS - 1) / GEMV2_KS + 3) & ~(size_t) 3;
            const size_t j0 = (size_t) ks * per < kc ? (size_t) ks * per : kc;
            const size_t j1 = j0 + per < kc ? j0 + per : kc;
            const T *ac = a + (k0 + j0) * lda;
            size_t j = j0;
#pragma unroll 2
            for (; j + 4 <= j1; j += 4, ac += 4 * lda) {
                V v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = *reinterpret_cast<const V *>(ac + (size_t) u * lda);
                T xv[4];
                if constexpr (VN == 2) {
                    const double2 x01 = *reinterpret_cast<const double2 *>(&xs[j]), x23 = *reinterpret_cast<const double2 *>(&xs[j + 2]);
                    xv[0] = x01.x; xv[1] = x01.y; xv[2] = x23.x; xv[3] = x23.y;
                } else {
                    const float4 x4 = *reinterpret_cast<const float4 *>(&xs[j]);
                    xv[0] = x4.x; xv[1] = x4.y; xv[2] = x4.z; xv[3] = x4.w;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    T av[VN];
                    unpack_v(v[u], av);
#pragma unroll
                    for (int e = 0; e < VN; e++) acc[e] = fma(av[e], xv[u], acc[e]);
                }
            }
            for (; j < j1; j++, ac += lda) {
                T av[VN];
                unpack_v(*reinterpret_cast<const V *>(ac), av);
#pragma unroll
                for (int e = 0; e < VN; e++) acc[e] = fma(av[e], xs[j], acc[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < VN; e++) part[ks][r * VN + e] = acc[e];
        __syncthreads();
        for (int rr = threadIdx.x; rr < RB; rr += blockDim.x) {
            T sum = T(0);
#pragma unroll
            for (int q = 0; q < GEMV2_KS; q++) sum += part[q][rr];
            T *yp = y + mat * sY + rb * RB + rr;
            *yp = beta == T(0) ? alpha * sum : alpha * sum + beta * (*yp);
        }
    }
}

template<typename T> struct UseDmma { static constexpr bool value = false; };
template<> struct UseDmma<double> { static constexpr bool value = true; };

template<typename T>
int launch_small(gpub_ctx_t ctx, cudaStream_t stream, int m, int n, int k, T alpha, const T *A, const T *B, T beta, T *C,
                 size_t batch) {
    const size_t per_mat = ((size_t) m * k + (size_t) k * n + (size_t) m * n) * sizeof(T);
    // matrices per chunk: enough columns of work for 256 threads, bounded by ~64 KB of shared memory per CTA
    int mpc = (int) ((256 + n - 1) / n);
    const size_t budget = 64 * 1024;
    if ((size_t) mpc * per_mat > budget) mpc = (int) (budget / per_mat);
    if (mpc < 1) mpc = 1;
    if ((size_t) mpc > batch) mpc = (int) batch;
    const size_t smem = (size_t) mpc * per_mat + 64;
    const size_t nchunks = gpub_ceil_div(batch, (size_t) mpc);
    const size_t cap = (size_t) ctx->sm_count * 3;
    const unsigned grid = (unsigned) (nchunks < cap ? nchunks : cap);
#define GPUB_SMALL_CASE(RT)                                                                                        \
    {                                                                                                              \
        auto kern = k_gemm_small<T, RT>;                                                                           \
        if (smem > 48 * 1024) GPUB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
        kern<<<grid, 256, smem, stream>>>(m, n, k, alpha, A, B, beta, C, batch, mpc);                              \
    }
    if (m <= 4) GPUB_SMALL_CASE(4)
    else if (m <= 8) GPUB_SMALL_CASE(8)
    else if (m <= 16) GPUB_SMALL_CASE(16)
    else GPUB_SMALL_CASE(32)
#undef GPUB_SMALL_CASE
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

template<typename T>
int gemm_batched(gpub_ctx_t ctx, int sidx, size_t m, size_t n, size_t k, T alpha, const T *A, size_t lda, size_t sA,
                 const T *B, size_t ldb, size_t sB, T beta, T *C, size_t ldc, size_t sC, size_t batch) {
    if (m == 0 || n == 0 || batch == 0) return GPUB_OK;
    if (!A || !B || !C) return GPUB_EINVAL;
    if (lda < m || ldb < k || ldc < m) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);

    const bool dense = lda == m && ldb == k && ldc == m && (batch == 1 || (sA == m * k && sB == k * n && sC == m * n));
    const bool aligned16 = (((uintptr_t) A | (uintptr_t) B | (uintptr_t) C) & 15u) == 0;
    const bool no_alias = !ranges_overlap(C, m * n * batch * sizeof(T), A, m * k * batch * sizeof(T)) &&
                          !ranges_overlap(C, m * n * batch * sizeof(T), B, k * n * batch * sizeof(T));
    if (dense && aligned16 && no_alias && m == n && n == k && batch >= 2) {
        switch (m) {
            case 4: return launch_col<T, 4>(ctx, stream, alpha, A, B, beta, C, batch);
            case 8: return launch_col<T, 8>(ctx, stream, alpha, A, B, beta, C, batch);
            case 16:
                if (try_dgemm_frag<T, 16>(ctx, stream, alpha, A, B, beta, C, batch)) { GPUB_LAUNCH_CHECK(); return GPUB_OK; }
                return launch_col<T, 16>(ctx, stream, alpha, A, B, beta, C, batch);
            case 32:
                if (try_dgemm_frag<T, 32>(ctx, stream, alpha, A, B, beta, C, batch)) { GPUB_LAUNCH_CHECK(); return GPUB_OK; }
                {
                    int e32 = GPUB_OK;
                    if (try_sgemm32<T>(ctx, stream, alpha, A, B, beta, C, batch, &e32)) {
                        if (e32 != GPUB_OK) return e32;
                        GPUB_LAUNCH_CHECK();
                        return GPUB_OK;
                    }
                }
                return launch_col<T, 32>(ctx, stream, alpha, A, B, beta, C, batch);
            default: break;
        }
    }
    if (dense && m <= 32 && n <= 32 && k <= 32 && k > 0 && batch >= 2) {
        // every operand of a chunk is staged before C is written, so C may alias A or B chunk-wise
        return launch_small<T>(ctx, stream, (int) m, (int) n, (int) k, alpha, A, B, beta, C, batch);
    }

    // tiled paths read A/B panels while other CTAs already write C: break aliasing with a scratch copy
    const size_t spanB = ((batch - 1) * sB + (n - 1) * ldb + k) * sizeof(T);
    const size_t spanA = ((batch - 1) * sA + (k ? (k - 1) : 0) * lda + m) * sizeof(T);
    const size_t spanC = ((batch - 1) * sC + (n - 1) * ldc + m) * sizeof(T);
    // scratch: the context's grow-only buffer when both copies fit (no allocation on the hot path), else stream-ordered
    T *tmpA = nullptr, *tmpB = nullptr;
    const bool aliasB = k > 0 && ranges_overlap(C, spanC, B, spanB), aliasA = k > 0 && ranges_overlap(C, spanC, A, spanA);
    bool pooled = false;
    if (aliasA || aliasB) {
        const size_t needB = aliasB ? (spanB + 255) & ~(size_t) 255 : 0, needA = aliasA ? spanA : 0;
        char *big = (char *) gpub_slot_big(slot, needA + needB);
        if (big) {
            pooled = true;
            if (aliasB) tmpB = (T *) big;
            if (aliasA) tmpA = (T *) (big + needB);
        } else {
            if (aliasB) GPUB_CUDA(cudaMallocAsync((void **) &tmpB, spanB, stream));
            if (aliasA) GPUB_CUDA(cudaMallocAsync((void **) &tmpA, spanA, stream));
        }
        if (aliasB) {
            GPUB_CUDA(cudaMemcpyAsync(tmpB, B, spanB, cudaMemcpyDeviceToDevice, stream));
            B = tmpB;
        }
        if (aliasA) {
            GPUB_CUDA(cudaMemcpyAsync(tmpA, A, spanA, cudaMemcpyDeviceToDevice, stream));
            A = tmpA;
        }
    }

    const size_t tm = gpub_ceil_div(m, 64), tn = gpub_ceil_div(n, 64);
    const size_t total = tm * tn * batch;
    const size_t cap = (size_t) ctx->sm_count * 8;
    const unsigned grid = (unsigned) (total < cap ? total : cap);
    bool done = false;
    if (n == 1 && k > 0) {
        constexpr size_t RBV = 32 * Vec16<T>::N;
        const size_t vcap = (size_t) ctx->sm_count * 16;
        if (m % RBV == 0 && lda % Vec16<T>::N == 0 && sA % Vec16<T>::N == 0 && (((uintptr_t) A) & 15u) == 0 && k >= 32) {
            const size_t row_blocks = m / RBV, items = row_blocks * batch;
            k_gemv_vec<T><<<(unsigned) (items < vcap ? items : vcap), 32 * GEMV2_KS, 0, stream>>>(m, k, alpha, A, lda, sA, B, sB, beta, C, sC,
                                                                                                    row_blocks, batch);
        } else {
            const size_t row_blocks = gpub_ceil_div(m, (size_t) GEMV_ROWS), items = row_blocks * batch;
            k_gemv<T><<<(unsigned) (items < vcap ? items : vcap), GEMV_ROWS * GEMV_KS, 0, stream>>>(m, k, alpha, A, lda, sA, B, sB, beta, C, sC,
                                                                                                      row_blocks, batch);
        }
        done = true;
    }
    if (!done && UseDmma<T>::value) {
        const bool ok = (m % 64 == 0) && (n % 64 == 0) && (k % DK == 0) && k >= DK && (lda % 2 == 0) && (ldb % 2 == 0) &&
                        ((((uintptr_t) A) | ((uintptr_t) B)) % 16 == 0) && (sA % 2 == 0) && (sB % 2 == 0);
        if (ok) {
            // measured on 128^3: 32-deep panels are 2 % slower than 16-deep ones (profiles/r1e), and a 3-stage cp.async ring with one
            // barrier per panel instead of two is 1 % slower than the 2-stage one (29.4 against 29.7 TFLOP/s, round 2)
            if (GPUB_DMMA_DKT32 && k % 32 == 0) {
                constexpr size_t smem32 = sizeof(double) * (2 * 32 * DLD + 2 * 64 * (32 + 4));
                GPUB_CUDA(cudaFuncSetAttribute(k_gemm_dmma<false, 32, GPUB_DMMA_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem32));
                k_gemm_dmma<false, 32, GPUB_DMMA_WARPS><<<grid, 32 * GPUB_DMMA_WARPS, smem32, stream>>>(m, n, k, (double) alpha, (const double *) A, lda, sA, (const double *) B,
                                                                      ldb, sB, (double) beta, (double *) C, ldc, sC, tm, tn, batch);
            } else {
                constexpr size_t smem16 = sizeof(double) * (2 * 16 * DLD + 2 * 64 * (16 + 4));
#if GPUB_DMMA_WARPS == 4
                k_gemm_dmma<false, 16, 4><<<grid, 128, smem16, stream>>>(m, n, k, (double) alpha, (const double *) A, lda, sA, (const double *) B,
                                                                         ldb, sB, (double) beta, (double *) C, ldc, sC, tm, tn, batch);
#else
                k_gemm_dmma<false, 16><<<grid, 256, smem16, stream>>>(m, n, k, (double) alpha, (const double *) A, lda, sA, (const double *) B,
                                                                      ldb, sB, (double) beta, (double *) C, ldc, sC, tm, tn, batch);
#endif
            }
            done = true;
        }
    }
    if (!done) done = try_sgemm_rt<T>(ctx, stream, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
    if (!done)
        k_gemm_tiled<T><<<grid, 256, 0, stream>>>(m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, tm, tn, batch);
    GPUB_LAUNCH_CHECK();
    if (!pooled && tmpA) GPUB_CUDA(cudaFreeAsync(tmpA, stream));
    if (!pooled && tmpB) GPUB_CUDA(cudaFreeAsync(tmpB, stream));
    return GPUB_OK;
}

// P_i = N_i N_i^T : one thread per output element for small n, tiled GEMM-like loop otherwise
template<typename T>
__global__ void k_aat(size_t n, const T *__restrict__ N, size_t sN, T *__restrict__ P, size_t sP, size_t batch) {
    const size_t nn = n * n;
    for (size_t b = blockIdx.y; b < batch; b += gridDim.y) {
        const T *nb = N + b * sN;
        for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t) gridDim.x * blockDim.x) {
            size_t i = e % n, j = e / n;
            T acc = 0;
            for (size_t c = 0; c < n; c++) acc = fma(nb[i + c * n], nb[j + c * n], acc);
            P[b * sP + e] = acc;
        }
    }
}

template<typename T> bool try_aat_dmma(gpub_ctx_t, cudaStream_t, size_t, const T *, size_t, T *, size_t, size_t) { return false; }
template<>
bool try_aat_dmma<double>(gpub_ctx_t ctx, cudaStream_t stream, size_t n, const double *N, size_t sN, double *P, size_t sP, size_t batch) {
    if (n % 64 != 0 || (sN & 1) || (((uintptr_t) N) & 15u)) return false;
    const size_t tm = n / 64, total = tm * (tm + 1) / 2 * batch, cap = (size_t) ctx->sm_count * 8;
    constexpr size_t smemT = sizeof(double) * (2 * 16 * DLD + 2 * 16 * DLD);
    k_gemm_dmma<true, 16, 8, true><<<(unsigned) (total < cap ? total : cap), 256, smemT, stream>>>(n, n, n, 1.0, N, n, sN, N, n, sN, 0.0, P, n, sP, tm, tm, batch);
    return true;
}

} // namespace

// C_i = blockdiag(Ublk_i, I) + alpha A_i B_i (fp64, m, n multiples of 64, k a multiple of 16): used by the U assembly in svd.cu
int gpub_internal_gemm_plus_e_f64(gpub_ctx_t ctx, int sidx, size_t m, size_t n, size_t k, double alpha, const double *A, size_t lda, size_t sA,
                                  const double *B, size_t ldb, size_t sB, const double *Ublk, size_t ne, size_t sUblk, double *C, size_t ldc,
                                  size_t sC, size_t batch) {
    if (m % 64 || n % 64 || k % DK || k < DK || (lda & 1) || (ldb & 1) || (sA & 1) || (sB & 1) || ((((uintptr_t) A) | ((uintptr_t) B)) & 15u))
        return GPUB_ENOTSUP;
    GPUB_ENTER(ctx, sidx);
    const size_t tm = m / 64, tn = n / 64, total = tm * tn * batch, cap = (size_t) ctx->sm_count * 8;
    constexpr size_t smem16 = sizeof(double) * (2 * 16 * DLD + 2 * 64 * (16 + 4));
    k_gemm_dmma<false, 16, 4, false, false, true><<<(unsigned) (total < cap ? total : cap), 128, smem16, stream>>>(
        m, n, k, alpha, A, lda, sA, B, ldb, sB, 0.0, C, ldc, sC, tm, tn, batch, nullptr, Ublk, sUblk, ne);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

namespace {

// true when the projector is computed from U alone (k_gemm_dmma PROJ): then the packed basis N is not an input of it
template<typename T> bool projector_reads_u_only(size_t, const T *, size_t) { return false; }
template<> bool projector_reads_u_only<double>(size_t n, const double *U, size_t sU) { return n % 64 == 0 && !(sU & 1) && !(((uintptr_t) U) & 15u); }

template<typename T>
bool try_projector_dmma(gpub_ctx_t, cudaStream_t, size_t, const T *, size_t, const unsigned *, const T *, size_t, T *, size_t, size_t) { return false; }
template<>
bool try_projector_dmma<double>(gpub_ctx_t ctx, cudaStream_t stream, size_t n, const double *U, size_t sU, const unsigned *rank, const double *N,
                                size_t sN, double *P, size_t sP, size_t batch) {
    if (!projector_reads_u_only<double>(n, U, sU)) return false;
    const size_t tm = n / 64, total = tm * (tm + 1) / 2 * batch, cap = (size_t) ctx->sm_count * 8;
    constexpr size_t smemT = sizeof(double) * (2 * 16 * DLD + 2 * 16 * DLD);
    k_gemm_dmma<true, 16, 4, true, true><<<(unsigned) (total < cap ? total : cap), 128, smemT, stream>>>(n, n, n, 1.0, U, n, sU, U, n, sU, 0.0, P, n, sP,
                                                                                                         tm, tm, batch, rank, U, sU);
    return true;
}

} // namespace

extern "C" {

int gpub_gemm_batched_f64(gpub_ctx_t ctx, int sidx, size_t m, size_t n, size_t k, double alpha, const double *A, size_t lda,
                          size_t sA, const double *B, size_t ldb, size_t sB, double beta, double *C, size_t ldc, size_t sC,
                          size_t batch) {
    return gemm_batched<double>(ctx, sidx, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
}

int gpub_gemm_batched_f32(gpub_ctx_t ctx, int sidx, size_t m, size_t n, size_t k, float alpha, const float *A, size_t lda,
                          size_t sA, const float *B, size_t ldb, size_t sB, float beta, float *C, size_t ldc, size_t sC,
                          size_t batch) {
    return gemm_batched<float>(ctx, sidx, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
}

#define GPUB_DEF_AAT(SUF, T)                                                                                         \
    int gpub_aat_batched_##SUF(gpub_ctx_t ctx, int sidx, size_t n, const T *N, size_t sN, T *P, size_t sP, size_t batch) { \
        if (n == 0 || batch == 0) return GPUB_OK;                                                                    \
        if (!N || !P) return GPUB_EINVAL;                                                                            \
        GPUB_ENTER(ctx, sidx);                                                                                       \
        if (try_aat_dmma<T>(ctx, stream, n, N, sN, P, sP, batch)) { GPUB_LAUNCH_CHECK(); return GPUB_OK; }           \
        unsigned gx = (unsigned) (gpub_ceil_div(n * n, 256) < 4096 ? gpub_ceil_div(n * n, 256) : 4096);              \
        unsigned gy = (unsigned) (batch < 65535 ? batch : 65535);                                                    \
        k_aat<T><<<dim3(gx, gy), 256, 0, stream>>>(n, N, sN, P, sP, batch);                                          \
        GPUB_LAUNCH_CHECK();                                                                                         \
        return GPUB_OK;                                                                                              \
    }
GPUB_DEF_AAT(f64, double)
GPUB_DEF_AAT(f32, float)

#define GPUB_DEF_PROJ(SUF, T)                                                                                        \
    int gpub_nullspace_projector_batched_##SUF(gpub_ctx_t ctx, int sidx, size_t n, const T *U, size_t sU, const unsigned int *rank, \
                                               const T *N, size_t sN, T *P, size_t sP, size_t batch) {               \
        if (n == 0 || batch == 0) return GPUB_OK;                                                                    \
        if (!N || !P || !U || !rank) return GPUB_EINVAL;                                                             \
        {                                                                                                            \
            GPUB_ENTER(ctx, sidx);                                                                                   \
            if (try_projector_dmma<T>(ctx, stream, n, U, sU, rank, N, sN, P, sP, batch)) { GPUB_LAUNCH_CHECK(); return GPUB_OK; } \
        }                                                                                                            \
        return gpub_aat_batched_##SUF(ctx, sidx, n, N, sN, P, sP, batch);                                            \
    }
GPUB_DEF_PROJ(f64, double)
GPUB_DEF_PROJ(f32, float)

/* Nullspace: the packed basis N and the projector N N' from the orthogonal factor in one call. The two kernels are independent
 * (both read U and the ranks; the projector multiplies out U1 or U2 straight from U), one is a copy and the other a contraction, so
 * the packing runs on a private stream beside the projector and the call's stream joins it at the end. */
#define GPUB_DEF_NSBUILD(SUF, T)                                                                                     \
    int gpub_nullspace_build_batched_##SUF(gpub_ctx_t ctx, int sidx, size_t n, const T *U, size_t sU, const unsigned int *rank, \
                                           T *N, size_t sN, T *P, size_t sP, size_t batch) {                          \
        if (n == 0 || batch == 0) return GPUB_OK;                                                                    \
        if (!N || !P || !U || !rank) return GPUB_EINVAL;                                                             \
        if (!projector_reads_u_only<T>(n, U, sU)) {   /* N N' is multiplied out from N itself: one after the other */  \
            const int e = gpub_nullspace_pack_batched_##SUF(ctx, sidx, n, U, sU, rank, N, sN, batch);                \
            return e ? e : gpub_nullspace_projector_batched_##SUF(ctx, sidx, n, U, sU, rank, N, sN, P, sP, batch);   \
        }                                                                                                            \
        GPUB_ENTER(ctx, sidx);                                                                                       \
        cudaStream_t side = nullptr;                                                                                 \
        cudaEvent_t ev[2] = {nullptr, nullptr};                                                                      \
        int e = gpub_ctx_fork(ctx, &side, ev);                                                                       \
        if (e) return e;                                                                                             \
        /* the projector is launched FIRST: its persistent CTAs (six per SM, limited by shared memory) leave room for two CTAs */ \
        /* of the copy, which then streams beside it; launched first, the copy's 16 k CTAs would fill the GPU on their own */    \
        cudaError_t c = cudaEventRecord(ev[0], stream);                                                              \
        if (c == cudaSuccess) c = cudaStreamWaitEvent(side, ev[0], 0);                                               \
        if (c == cudaSuccess) e = gpub_nullspace_projector_batched_##SUF(ctx, sidx, n, U, sU, rank, N, sN, P, sP, batch); \
        if (c == cudaSuccess && !e) e = gpub_internal_nullspace_pack_##SUF(ctx, side, n, U, sU, rank, N, sN, batch);  \
        if (c == cudaSuccess) c = cudaEventRecord(ev[1], side);                                                      \
        if (c == cudaSuccess) c = cudaStreamWaitEvent(stream, ev[1], 0);                                             \
        cudaEventDestroy(ev[0]);                                                                                     \
        cudaEventDestroy(ev[1]);                                                                                     \
        return e ? e : (int) c;                                                                                      \
    }
GPUB_DEF_NSBUILD(f64, double)
GPUB_DEF_NSBUILD(f32, float)

} // extern "C"
