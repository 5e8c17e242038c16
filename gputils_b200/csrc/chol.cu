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
#ifndef GPUB_GRID_WAVES
#define GPUB_GRID_WAVES 2   // persistent grids: resident CTAs per SM x SM count x this
#endif
#include <type_traits>

namespace {

template<typename T> __device__ __forceinline__ T fast_rsqrt(T d);
template<> __device__ __forceinline__ double fast_rsqrt<double>(double d) {
    // MUFU.RSQ64H seed (~2^-22) + two Newton steps y <- y + y (1/2 - (d/2) y^2): ~1 ulp, no special-case
    // branch or slow-path call (d <= 0 is reported through info; the result is then NaN / inf by design)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
    y = fma(y, fma(-h, y * y, 0.5), y);
    y = fma(y, fma(-h, y * y, 0.5), y);
    return y;
}
template<> __device__ __forceinline__ float fast_rsqrt<float>(float d) {
    float r = rsqrtf(d);
    return fmaf(0.5f * r, fmaf(-d * r, r, 1.0f), r); // one Newton step: rsqrtf alone is ~2 ulp
}

template<typename T> __device__ __forceinline__ T fast_rcp(T d);
template<> __device__ __forceinline__ double fast_rcp<double>(double d) {
    // MUFU.RCP64H seed + two Newton steps y <- y + y (1 - d y)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    y = fma(y, fma(-d, y, 1.0), y);
    y = fma(y, fma(-d, y, 1.0), y);
    return y;
}
template<> __device__ __forceinline__ float fast_rcp<float>(float d) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    return fmaf(y, fmaf(-d, y, 1.0f), y);
}
// (Measured and rejected for the diagonal blocks of the blocked kernel: 4-column steps with the 4 x 4 diagonal tile factorised
// redundantly in every lane -- n = 128: fp64 0.86 -> 1.06 ms, fp32 1.17 -> 1.36 ms. The chain per column is shorter, but the diagonal
// warp runs alone and its time is set by its instruction count: 2180 instead of 1440 per 32 x 32 block, ~5 cycles each.)

// ------------------------------------------------------------------------------------------
// potrf, n <= 32: NP lanes per matrix
// Control flow is uniform across the CTA (tail groups redo the last matrix with stores masked) and the
// padding rows/columns of a group with n < NP are an identity block, so the NP column steps run
// unconditionally: the whole factorisation is one basic block that ptxas can software-pipeline
// (the pivot of column j+1 is broadcast and its rsqrt started while column j's update is still issuing).
// ------------------------------------------------------------------------------------------
#ifndef GPUB_POTRF_THREADS
#define GPUB_POTRF_THREADS 128
#endif
#ifndef GPUB_POTRF_MINB
#define GPUB_POTRF_MINB 5
#endif

// DENSE: n == NP and lda == NP, so every load/store offset is an immediate.
template<typename T, int NP, bool DENSE>
__global__ void __launch_bounds__(GPUB_POTRF_THREADS, GPUB_POTRF_MINB)
k_potrf_group(int n, T *A, size_t lda_rt, size_t strideA, int *info, size_t batch) {
    const size_t lda = DENSE ? (size_t) NP : lda_rt;
    constexpr int GROUPS = GPUB_POTRF_THREADS / NP;
    __shared__ __align__(16) T s_col[2][GROUPS][NP];
    const int grp = threadIdx.x / NP;
    const int i = threadIdx.x % NP; // row owned by this lane
    const size_t ngroups = (size_t) gridDim.x * GROUPS;
    const size_t iters = (batch + ngroups - 1) / ngroups;
    const bool row_ok = DENSE || i < n;

    for (size_t it = 0; it < iters; it++) {
        size_t mat = it * ngroups + (size_t) blockIdx.x * GROUPS + grp;
        const bool live = mat < batch;
        if (!live) mat = batch - 1;
        T *a_g = A + mat * strideA;
        T a[NP];
        // lower triangle only: lane i needs columns 0..i; everything else is an identity pattern
#pragma unroll
        for (int c = 0; c < NP; c++) a[c] = (row_ok && c <= i) ? a_g[i + (size_t) c * lda] : T(c == i ? 1 : 0);
        int bad = 0;
#pragma unroll
        for (int j = 0; j < NP; j++) {
            const T d = __shfl_sync(0xffffffffu, a[j], j, NP);
            if (!(d > T(0)) && bad == 0) bad = j + 1;
            const T r = fast_rsqrt<T>(d);
            const T l = a[j] * r; // lane j: sqrt(d); lanes below: L(i,j); lanes above: 0
            a[j] = l;
            T *col = s_col[j & 1][grp];
            col[i] = l;
            __syncwarp();
#pragma unroll
            for (int c = j + 1; c < NP; c++) a[c] = fma(-l, col[c], a[c]);
        }
        if (row_ok && live) {
#pragma unroll
            for (int c = 0; c < NP; c++)
                if (c <= i) a_g[i + (size_t) c * lda] = a[c];
        }
        if (i == 0 && live) info[mat] = bad;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// n == 4, dense: one THREAD per matrix (k_potrf4 / k_potrs4). The lane-per-row kernels keep 8 matrices (768 B) per warp in
// flight, which at 4 x 4 is not enough outstanding memory to cover the HBM latency (fp32 potrs: 55 % of the roofline with
// neither the LSU nor the issue slots busy). A thread that owns a whole matrix issues its 64 / 128 bytes as back-to-back 128-bit
// loads: 4x the bytes in flight per warp, a quarter of the memory instructions per matrix, and no shuffles at all.
// ------------------------------------------------------------------------------------------
template<typename T> __device__ __forceinline__ void load16(const T *p, T (&a)[16]);
template<> __device__ __forceinline__ void load16<float>(const float *p, float (&a)[16]) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float4 v = reinterpret_cast<const float4 *>(p)[c];
        a[4 * c] = v.x; a[4 * c + 1] = v.y; a[4 * c + 2] = v.z; a[4 * c + 3] = v.w;
    }
}
template<> __device__ __forceinline__ void load16<double>(const double *p, double (&a)[16]) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const double2 v = reinterpret_cast<const double2 *>(p)[c];
        a[2 * c] = v.x; a[2 * c + 1] = v.y;
    }
}
template<typename T> __device__ __forceinline__ void store4(T *p, T x0, T x1, T x2, T x3);
template<> __device__ __forceinline__ void store4<float>(float *p, float x0, float x1, float x2, float x3) {
    *reinterpret_cast<float4 *>(p) = make_float4(x0, x1, x2, x3);
}
template<> __device__ __forceinline__ void store4<double>(double *p, double x0, double x1, double x2, double x3) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(x0, x1);
    reinterpret_cast<double2 *>(p)[1] = make_double2(x2, x3);
}

template<typename T>
__global__ void __launch_bounds__(128) k_potrf4(T *A, int *info, size_t batch) {
    for (size_t mat = (size_t) blockIdx.x * 128 + threadIdx.x; mat < batch; mat += (size_t) gridDim.x * 128) {
        T *a_g = A + mat * 16;
        T a[16];   // a[r + 4 c]
        load16<T>(a_g, a);
        int bad = 0;
        // column 0
        if (!(a[0] > T(0))) bad = 1;
        T r = fast_rsqrt<T>(a[0]);
        const T l00 = a[0] * r, l10 = a[1] * r, l20 = a[2] * r, l30 = a[3] * r;
        // column 1
        T d = fma(-l10, l10, a[5]);
        if (!(d > T(0)) && bad == 0) bad = 2;
        r = fast_rsqrt<T>(d);
        const T l11 = d * r, l21 = fma(-l20, l10, a[6]) * r, l31 = fma(-l30, l10, a[7]) * r;
        // column 2
        d = fma(-l21, l21, fma(-l20, l20, a[10]));
        if (!(d > T(0)) && bad == 0) bad = 3;
        r = fast_rsqrt<T>(d);
        const T l22 = d * r, l32 = fma(-l31, l21, fma(-l30, l20, a[11])) * r;
        // column 3
        d = fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, a[15])));
        if (!(d > T(0)) && bad == 0) bad = 4;
        r = fast_rsqrt<T>(d);
        const T l33 = d * r;
        // the strict upper triangle is written back with the values that were read
        store4<T>(a_g, l00, l10, l20, l30);
        store4<T>(a_g + 4, a[4], l11, l21, l31);
        store4<T>(a_g + 8, a[8], a[9], l22, l32);
        store4<T>(a_g + 12, a[12], a[13], a[14], l33);
        info[mat] = bad;
    }
}

template<typename T>
__global__ void __launch_bounds__(128) k_potrs4(const T *__restrict__ L, T *b, size_t batch) {
    for (size_t mat = (size_t) blockIdx.x * 128 + threadIdx.x; mat < batch; mat += (size_t) gridDim.x * 128) {
        T l[16], x[16];
        load16<T>(L + mat * 16, l);
        T *b_g = b + mat * 4;
        if constexpr (sizeof(T) == 4) {
            const float4 v = *reinterpret_cast<const float4 *>(b_g);
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        } else {
            const double2 v0 = reinterpret_cast<const double2 *>(b_g)[0], v1 = reinterpret_cast<const double2 *>(b_g)[1];
            x[0] = v0.x; x[1] = v0.y; x[2] = v1.x; x[3] = v1.y;
        }
        const T i0 = T(1) / l[0], i1 = T(1) / l[5], i2 = T(1) / l[10], i3 = T(1) / l[15];
        // L y = b
        const T y0 = x[0] * i0;
        const T y1 = fma(-l[1], y0, x[1]) * i1;
        const T y2 = fma(-l[6], y1, fma(-l[2], y0, x[2])) * i2;
        const T y3 = fma(-l[11], y2, fma(-l[7], y1, fma(-l[3], y0, x[3]))) * i3;
        // L^T x = y
        const T x3 = y3 * i3;
        const T x2 = fma(-l[11], x3, y2) * i2;
        const T x1 = fma(-l[6], x2, fma(-l[7], x3, y1)) * i1;
        const T x0 = fma(-l[1], x1, fma(-l[2], x2, fma(-l[3], x3, y0))) * i0;
        store4<T>(b_g, x0, x1, x2, x3);
    }
}

// the same for n == 8, fp32 (fp64 is at 95 % lane-per-row): 16 back-to-back LDG.128 per thread, substitution fully unrolled
__global__ void __launch_bounds__(128) k_potrs8_f32(const float *__restrict__ L, float *b, size_t batch) {
    for (size_t mat = (size_t) blockIdx.x * 128 + threadIdx.x; mat < batch; mat += (size_t) gridDim.x * 128) {
        float l[64], x[8], inv[8];
        const float4 *lp = reinterpret_cast<const float4 *>(L + mat * 64);
#pragma unroll
        for (int c = 0; c < 8; c++) {
            // column c: rows 0..3 are needed only for c < 4 (lower triangle)
            if (c < 4) {
                const float4 v = lp[2 * c];
                l[8 * c] = v.x; l[8 * c + 1] = v.y; l[8 * c + 2] = v.z; l[8 * c + 3] = v.w;
            }
            const float4 w = lp[2 * c + 1];
            l[8 * c + 4] = w.x; l[8 * c + 5] = w.y; l[8 * c + 6] = w.z; l[8 * c + 7] = w.w;
        }
        float4 *bp = reinterpret_cast<float4 *>(b + mat * 8);
        const float4 b0 = bp[0], b1 = bp[1];
        x[0] = b0.x; x[1] = b0.y; x[2] = b0.z; x[3] = b0.w; x[4] = b1.x; x[5] = b1.y; x[6] = b1.z; x[7] = b1.w;
#pragma unroll
        for (int j = 0; j < 8; j++) inv[j] = 1.0f / l[9 * j];
#pragma unroll
        for (int j = 0; j < 8; j++) {       // L y = b
            x[j] *= inv[j];
#pragma unroll
            for (int i = j + 1; i < 8; i++) x[i] = fmaf(-l[i + 8 * j], x[j], x[i]);
        }
#pragma unroll
        for (int j = 7; j >= 0; j--) {      // L^T x = y
#pragma unroll
            for (int i = j + 1; i < 8; i++) x[j] = fmaf(-l[i + 8 * j], x[i], x[j]);
            x[j] *= inv[j];
        }
        bp[0] = make_float4(x[0], x[1], x[2], x[3]);
        bp[1] = make_float4(x[4], x[5], x[6], x[7]);
    }
}

// Fused solve + all-gather (GATHER): every solution is also stored straight into the gathered tensor of up to 8 devices
// (peers.x[p] + (peers.offset + matrix) * peers.stride, NVLink peer stores from inside the kernel), so the all-gather that would
// follow the solve of a sharded batch costs no second pass and no second launch.
template<typename T>
struct PotrsPeers {
    T *x[8];
    int count;
    size_t offset, stride;
};

// ------------------------------------------------------------------------------------------
// potrf, n == 32 (BASELINE config 2) or 16, dense: k_potrf_pair<T, N>. N / 2 lanes per matrix, lane p owns rows p AND p + N / 2.
// k_potrf_group<T, 32> (lane = row) executes 31 - j FMA instructions per column step for every row, although row i only
// needs columns <= i: n^3/2 lane-FMAs for n^3/6 useful ones, and ncu shows it FP64-issue bound (pipe 48 %, issue 49 %) at
// 55 % of DRAM throughput. Pairing a short row with a long one skips the columns >= 16 of the short row statically
// (15 - j + 31 - j FMAs per step for TWO rows, and one instruction stream serves two matrices per warp): 308 instead of 496
// warp-FMAs per matrix, and the pivot / rsqrt / broadcast overhead per matrix halves as well.
// ------------------------------------------------------------------------------------------
#ifndef GPUB_POTRF32_MINB
#define GPUB_POTRF32_MINB 4
#endif
// N = 64, 32 or 16: H = N / 2 lanes per matrix, lane p owns rows p and p + H. N = 64 is a whole warp per matrix with the
// lower triangle in registers (96 entries per lane): no CTA barrier at all (a 32 x 32-block kernel with one warp per block lost its time there).
// DENSE: n == N and lda == N; otherwise rows / columns beyond n are an identity pad and lda is a run-time value.
#ifndef GPUB_CHOL4
#define GPUB_CHOL4 1
#endif
#ifndef GPUB_QUAD128_F32_MINB
#define GPUB_QUAD128_F32_MINB 4
#endif
#ifndef GPUB_PAIR64_F32_MINB
#define GPUB_PAIR64_F32_MINB 4
#endif
template<typename T, int N, bool DENSE = true> struct PotrfPairMinB { static constexpr int value = N == 64 ? (sizeof(T) == 8 ? 2 : (DENSE ? GPUB_PAIR64_F32_MINB : 3)) : GPUB_POTRF32_MINB; };

template<typename T, int N, bool DENSE>
__global__ void __launch_bounds__(128, PotrfPairMinB<T, N, DENSE>::value) k_potrf_pair(int n_rt, T *A, size_t lda_rt, size_t strideA, int *info, size_t batch) {
    constexpr int H = N / 2;
    constexpr int GROUPS = 128 / H;
    __shared__ __align__(16) T s_col[2][GROUPS][N];
    const int n = DENSE ? N : n_rt;
    const size_t lda = DENSE ? (size_t) N : lda_rt;
    const int grp = threadIdx.x / H, p = threadIdx.x % H;
    const size_t ngroups = (size_t) gridDim.x * GROUPS;
    const size_t iters = (batch + ngroups - 1) / ngroups;
    for (size_t it = 0; it < iters; it++) {
        size_t mat = it * ngroups + (size_t) blockIdx.x * GROUPS + grp;
        const bool live = mat < batch;
        if (!live) mat = batch - 1;
        T *a_g = A + mat * strideA;
        T lo[H], hi[N];   // row p (columns 0..H-1), row p + H (all columns); only the lower triangle is read
#pragma unroll
        for (int c = 0; c < H; c++) lo[c] = (c <= p && (DENSE || p < n)) ? a_g[p + c * lda] : T((!DENSE && c == p) ? 1 : 0);
#pragma unroll
        for (int c = 0; c < N; c++) hi[c] = (c <= p + H && (DENSE || p + H < n)) ? a_g[p + H + c * lda] : T((!DENSE && c == p + H) ? 1 : 0);
        int bad = 0;
#pragma unroll
        for (int j = 0; j < H; j++) {          // pivots in the short rows
            const T d = __shfl_sync(0xffffffffu, lo[j], j, H);
            if (!(d > T(0)) && bad == 0) bad = j + 1;
            const T r = fast_rsqrt<T>(d);
            T *col = s_col[j & 1][grp];
            const T llo = lo[j] * r, lhi = hi[j] * r;
            lo[j] = llo;
            hi[j] = lhi;
            col[p] = llo;
            col[p + H] = lhi;
            __syncwarp();
#pragma unroll
            for (int c = j + 1; c < H; c++) {
                const T cc = col[c];
                lo[c] = fma(-llo, cc, lo[c]);
                hi[c] = fma(-lhi, cc, hi[c]);
            }
#pragma unroll
            for (int c = H; c < N; c++) hi[c] = fma(-lhi, col[c], hi[c]);
        }
#pragma unroll
        for (int j = H; j < N; j++) {         // pivots in the long rows; the short rows are finished
            const T d = __shfl_sync(0xffffffffu, hi[j], j - H, H);
            if (!(d > T(0)) && bad == 0) bad = j + 1;
            const T r = fast_rsqrt<T>(d);
            T *col = s_col[j & 1][grp];
            const T lhi = hi[j] * r;
            hi[j] = lhi;
            col[p + H] = lhi;
            __syncwarp();
#pragma unroll
            for (int c = j + 1; c < N; c++) hi[c] = fma(-lhi, col[c], hi[c]);
        }
        if (live) {
#pragma unroll
            for (int c = 0; c < H; c++)
                if (c <= p && (DENSE || p < n)) a_g[p + c * lda] = lo[c];
#pragma unroll
            for (int c = 0; c < N; c++)
                if (c <= p + H && (DENSE || p + H < n)) a_g[p + H + c * lda] = hi[c];
            if (p == 0) info[mat] = bad;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// potrf, 64 < n <= 32*NB (NB = 3, 4): blocked right-looking factorisation, one matrix per CTA (k_potrf_pipe below).
// The matrix is cut into 32 x 32 blocks; one warp owns one block of the lower triangle for the whole factorisation. For each
// block column hb: the diagonal warp factorises its block (lane = row, shuffled pivot, rsqrt, column broadcast through shared
// memory), the warps below solve L21 = A21 L11^-T with no inter-lane dependency (every lane owns a row), and every trailing warp
// applies its rank-32 update. Rows / columns beyond n are an identity pad.
// ------------------------------------------------------------------------------------------
// resident CTAs per SM the register allocation is tuned for
#ifndef GPUB_BLK4_F32_MINB
#define GPUB_BLK4_F32_MINB 3
#endif
template<typename T, int NB> struct PotrfBlkMinB { static constexpr int value = NB == 2 ? 5 : (NB == 3 ? 3 : (sizeof(T) == 4 ? GPUB_BLK4_F32_MINB : 2)); };

// fp64: the trailing update runs on the FP64 tensor pipe (GPUB_BLK_DMMA). A block that still waits for its block column keeps
// its 32 x 32 entries as 4 x 4 DMMA accumulator tiles (m8n8k4: lane (g, q) = (lane / 4, lane % 4) holds rows 8i + g, columns
// 8j + 2q, 8j + 2q + 1); the rank-32 update is then 8 k-steps of (4 + 4) LDS.64 fragment loads and 16 DMMA instead of 32 k-steps
// of 16 broadcast LDS.128 and 32 DFMA -- the lane = row form saturated the shared-memory writeback path (ncu / phase timers:
// trailing 34 % of a matrix). When the block's own column comes up it is turned into the lane = row form once, through its
// (free) panel slot. The panel slots have a row stride of 36 doubles so that the fragment loads are bank-conflict free.
#ifndef GPUB_BLK_DMMA
#define GPUB_BLK_DMMA 1
#endif
__device__ __forceinline__ void chol_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// k_potrf_pipe<T, NB>: ONE CTA barrier per block column (its predecessor had a barrier after each of the three phases). In the
// interval of block column hb
//   * every warp with h >= hb first applies the rank-32 update of block column hb - 1 to its block,
//   * the diagonal warp (hb, hb) then factorises its block column by column and publishes its progress (a counter in shared
//     memory, written after the column and its reciprocal pivot are in place),
//   * the panel warps (rb > hb, hb) follow it one column behind: column j of the panel solve needs only column j of L11, so the
//     panel solves finish ~100 cycles after the diagonal block instead of a whole phase later.
// The three-barrier predecessor's stall samples were 72 % barrier waits, most of them nine warps waiting for the diagonal warp; here the panel
// phase (18 % of a matrix) disappears from the critical path: per block column it is update + diagonal block.
// The panel slots are double-buffered by block-column parity (the trailing warps of column hb - 1 still read L(., hb - 1) while
// the panel warps of column hb write L(., hb)); the only spinning is the panel warps' poll of the progress counter, inside one
// CTA, on a warp that never waits for them.
// ------------------------------------------------------------------------------------------
// The progress signal is an mbarrier per chunk of GPUB_PIPE_CHUNK columns (count 1): the diagonal warp's lane 0 arrives (release)
// after the __syncwarp that follows the chunk's last column store, a panel warp sleeps in try_wait (acquire) -- no polling loop
// competing for issue slots and, unlike a flag + __threadfence_block (MEMBAR.CTA in front of the next LDS of the chain, measured
// 0.86 -> 1.09 ms), nothing on the diagonal warp's dependency chain.
// Measured and rejected in the diagonal loop (n = 128 fp64 0.78 -> 0.84 ms): pivots carried through a reciprocal so that the rsqrt
// leaves the pivot-to-pivot chain -- the diagonal warp is bound by its instruction count, not by the chain latency.
#ifndef GPUB_DIAG_SPLIT
#define GPUB_DIAG_SPLIT 1
#endif
#ifndef GPUB_PIPE_CHUNK
#define GPUB_PIPE_CHUNK 4
#endif
__device__ __forceinline__ void chol_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned) __cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void chol_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned) __cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void chol_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "CHOL_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CHOL_DONE_%=;\n\t"
        "bra CHOL_WAIT_%=;\n\t"
        "CHOL_DONE_%=:\n\t}" ::"r"((unsigned) __cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

template<typename T, int NB>
__global__ void __launch_bounds__(32 * (NB * (NB + 1) / 2), PotrfBlkMinB<T, NB>::value) k_potrf_pipe(int n, T *A, size_t lda, size_t strideA, int *info, size_t batch) {
    constexpr bool FRAG = GPUB_BLK_DMMA && sizeof(T) == 8;   // trailing blocks in DMMA accumulator layout
    constexpr int LDP = FRAG ? 36 : 32;                      // row stride of a panel slot
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T (*s_p)[NB][32][LDP] = reinterpret_cast<T (*)[NB][32][LDP]>(smem_raw);                              // [parity][block row][k][row]
    T (*s_d)[LDP] = reinterpret_cast<T (*)[LDP]>(smem_raw + sizeof(T) * 2 * NB * 32 * LDP);              // current diagonal factor [k][row]
    T *s_rinv = reinterpret_cast<T *>(smem_raw + sizeof(T) * (2 * NB + 1) * 32 * LDP);
    __shared__ int s_bad;
    __shared__ __align__(8) uint64_t s_bar[32 / GPUB_PIPE_CHUNK];   // chunk c of the current diagonal block is published
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    int rb = 0;
    while ((rb + 1) * (rb + 2) / 2 <= warp) rb++;
    const int h = warp - rb * (rb + 1) / 2;
    const int row = 32 * rb + lane;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < 32 / GPUB_PIPE_CHUNK; c++) chol_mbar_init(&s_bar[c], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned use = 0;        // diagonal blocks factorised so far by this CTA: every barrier completes one phase per block
    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        T *a_g = A + mat * strideA;
        T a[32];
        if (FRAG && h > 0) {
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int r_ = 32 * rb + 8 * i + g, c_ = 32 * h + 8 * j + 2 * q + e;
                        a[(4 * i + j) * 2 + e] = (r_ < n && c_ <= r_) ? a_g[r_ + (size_t) c_ * lda] : T(r_ == c_ ? 1 : 0);
                    }
        } else {
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const int col = 32 * h + c;
                a[c] = (row < n && col <= row) ? a_g[row + (size_t) col * lda] : T(row == col ? 1 : 0);
            }
        }
        if (threadIdx.x == 0) s_bad = 0;
        __syncthreads();
#pragma unroll 1
        for (int hb = 0; hb < NB; hb++, use++) {
            if (h >= hb) {
                if (hb > 0) { // rank-32 update with block column hb - 1
                    const int pp = (hb - 1) & 1;
                    if constexpr (FRAG) {
#pragma unroll 2
                        for (int k4 = 0; k4 < 32; k4 += 4) {
                            double af[4], bf[4];
#pragma unroll
                            for (int i = 0; i < 4; i++) af[i] = -(double) s_p[pp][rb][k4 + q][8 * i + g];
#pragma unroll
                            for (int j = 0; j < 4; j++) bf[j] = (double) s_p[pp][h][k4 + q][8 * j + g];
#pragma unroll
                            for (int i = 0; i < 4; i++)
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    double c0 = (double) a[(4 * i + j) * 2], c1 = (double) a[(4 * i + j) * 2 + 1];
                                    chol_dmma(c0, c1, af[i], bf[j]);
                                    a[(4 * i + j) * 2] = (T) c0;
                                    a[(4 * i + j) * 2 + 1] = (T) c1;
                                }
                        }
                    } else {
#pragma unroll 8
                        for (int k = 0; k < 32; k++) {
                            const T lk = s_p[pp][rb][k][lane];
#pragma unroll
                            for (int c = 0; c < 32; c++) a[c] = fma(-lk, s_p[pp][h][k][c], a[c]);
                        }
                    }
                }
                if (h == hb) {
                    if (FRAG && hb > 0) {
                        // accumulator tiles -> lane = row through a slot nobody reads now: the diagonal warp uses s_d (its readers wait on
                        // the chunk barriers first), a panel warp its own slot of this column's parity
                        T *scr = rb == hb ? &s_d[0][0] : &s_p[hb & 1][rb][0][0];
#pragma unroll
                        for (int i = 0; i < 4; i++)
#pragma unroll
                            for (int j = 0; j < 4; j++)
#pragma unroll
                                for (int e = 0; e < 2; e++) scr[(8 * j + 2 * q + e) * LDP + 8 * i + g] = a[(4 * i + j) * 2 + e];
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 32; c++) a[c] = scr[c * LDP + lane];
                        __syncwarp();
                    }
                    if (rb == hb) { // diagonal block
                        int bad = 0;
                        T d = __shfl_sync(0xffffffffu, a[0], 0);
                        // One column: pivot -> rsqrt -> scale -> publish -> update of the columns j < c < CEND. The diagonal warp runs alone
                        // and is bound by its instruction count, so with SPLIT (fp64) the columns 0..15 update only the columns up to 15; the
                        // rank-16 update they owe to the lower-right 16 x 16 corner is then done at once on the tensor pipe (16 DMMA + the
                        // two layout changes: ~80 instructions instead of 256 DFMA + 128 LDS), and the columns 16..31 follow as before.
                        constexpr bool SPLIT = FRAG && GPUB_DIAG_SPLIT;
                        auto column = [&](auto jt, auto cendt) {
                            constexpr int j = decltype(jt)::value, CEND = decltype(cendt)::value;
                            if (!(d > T(0)) && bad == 0) bad = j + 1;
                            const T r = fast_rsqrt<T>(d);
                            const T l = a[j] * r;
                            a[j] = l;
                            if (j + 1 < CEND) d = __shfl_sync(0xffffffffu, fma(-l, l, a[j + 1 < 32 ? j + 1 : j]), j + 1);
                            s_d[j][lane] = l;
                            if (lane == j) s_rinv[j] = r;
                            __syncwarp();
                            if ((j + 1) % GPUB_PIPE_CHUNK == 0 && lane == 0) chol_mbar_arrive(&s_bar[j / GPUB_PIPE_CHUNK]);
#pragma unroll
                            for (int c = j + 1; c < CEND; c++) a[c] = fma(-l, s_d[j][c], a[c]);
                        };
                        auto columns = [&](auto j0t, auto cendt) {   // 16 columns starting at j0, explicit constants keep a[] in registers
                            constexpr int J0 = decltype(j0t)::value;
#define GPUB_COL(K) column(std::integral_constant<int, J0 + K>{}, cendt);
                            GPUB_COL(0) GPUB_COL(1) GPUB_COL(2) GPUB_COL(3) GPUB_COL(4) GPUB_COL(5) GPUB_COL(6) GPUB_COL(7)
                            GPUB_COL(8) GPUB_COL(9) GPUB_COL(10) GPUB_COL(11) GPUB_COL(12) GPUB_COL(13) GPUB_COL(14) GPUB_COL(15)
#undef GPUB_COL
                        };
                        if constexpr (SPLIT) {
                            columns(std::integral_constant<int, 0>{}, std::integral_constant<int, 16>{});
                            // corner (rows, columns 16..31) -= L21 L21^T with L21 = rows 16..31 of the columns just published in s_d
                            T *scr = &s_p[hb & 1][hb][0][0];   // the diagonal block row has no panel slot of its own: free scratch
                            if (lane >= 16) {
#pragma unroll
                                for (int c = 0; c < 16; c++) scr[c * LDP + lane - 16] = a[16 + c];
                            }
                            __syncwarp();
                            double cf[2][2][2];
#pragma unroll
                            for (int i = 0; i < 2; i++)
#pragma unroll
                                for (int jj = 0; jj < 2; jj++)
#pragma unroll
                                    for (int e = 0; e < 2; e++) cf[i][jj][e] = (double) scr[(8 * jj + 2 * q + e) * LDP + 8 * i + g];
#pragma unroll
                            for (int k4 = 0; k4 < 16; k4 += 4) {
                                const double f0 = (double) s_d[k4 + q][16 + g], f1 = (double) s_d[k4 + q][24 + g];
                                chol_dmma(cf[0][0][0], cf[0][0][1], -f0, f0);
                                chol_dmma(cf[0][1][0], cf[0][1][1], -f0, f1);
                                chol_dmma(cf[1][0][0], cf[1][0][1], -f1, f0);
                                chol_dmma(cf[1][1][0], cf[1][1][1], -f1, f1);
                            }
                            __syncwarp();
#pragma unroll
                            for (int i = 0; i < 2; i++)
#pragma unroll
                                for (int jj = 0; jj < 2; jj++)
#pragma unroll
                                    for (int e = 0; e < 2; e++) scr[(8 * jj + 2 * q + e) * LDP + 8 * i + g] = (T) cf[i][jj][e];
                            __syncwarp();
                            if (lane >= 16) {
#pragma unroll
                                for (int c = 0; c < 16; c++) a[16 + c] = scr[c * LDP + lane - 16];
                            }
                            d = __shfl_sync(0xffffffffu, a[16], 16);
                            columns(std::integral_constant<int, 16>{}, std::integral_constant<int, 32>{});
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                if (!(d > T(0)) && bad == 0) bad = j + 1;
                                const T r = fast_rsqrt<T>(d);
                                const T l = a[j] * r;
                                a[j] = l;
                                if (j + 1 < 32) d = __shfl_sync(0xffffffffu, fma(-l, l, a[j + 1 < 32 ? j + 1 : j]), j + 1);
                                s_d[j][lane] = l;
                                if (lane == j) s_rinv[j] = r;
                                __syncwarp();
                                if ((j + 1) % GPUB_PIPE_CHUNK == 0 && lane == 0) chol_mbar_arrive(&s_bar[j / GPUB_PIPE_CHUNK]);
#pragma unroll
                                for (int c = j + 1; c < 32; c++) a[c] = fma(-l, s_d[j][c], a[c]);
                            }
                        }
                        if (lane == 0 && bad != 0) atomicCAS(&s_bad, 0, 32 * hb + bad);
                    } else { // panel block: one column behind the diagonal warp
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            if (j % GPUB_PIPE_CHUNK == 0) chol_mbar_wait(&s_bar[j / GPUB_PIPE_CHUNK], use & 1u);
                            const T l = a[j] * s_rinv[j];
                            a[j] = l;
                            s_p[hb & 1][rb][j][lane] = l;
#pragma unroll
                            for (int c = j + 1; c < 32; c++) a[c] = fma(-l, s_d[j][c], a[c]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (row < n) {
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const int col = 32 * h + c;
                if (col <= row) a_g[row + (size_t) col * lda] = a[c];
            }
        }
        if (threadIdx.x == 0) info[mat] = s_bad;
        __syncthreads();
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
// potrs, n <= 32: NP lanes per matrix.
// With D = diag(L) and the row-scaled factor Lb = D^-1 L (unit diagonal, Lb(i,j) = L(i,j) / L(i,i)):
//     L y = b      <=>  Lb y = D^-1 b          (forward, row layout:    lane i owns row i of Lb)
//     L^T z = y    <=>  Lb^T (D z) = y         (backward, column layout: lane i owns column i of Lb)
// so every lane scales by its OWN reciprocal diagonal (computed once, in parallel) and each substitution
// step on the critical path is one shuffle + one FMA. The transposition goes through padded shared memory
// (row writes and column reads both conflict-free). Only the lower triangle of L is read.
// ------------------------------------------------------------------------------------------
#ifndef GPUB_POTRS_THREADS
#define GPUB_POTRS_THREADS 128
#endif
#ifndef GPUB_POTRS_MINB
#define GPUB_POTRS_MINB 5
#endif

template<typename T, int NP, bool DENSE, bool GATHER = false>
__global__ void __launch_bounds__(GPUB_POTRS_THREADS, GPUB_POTRS_MINB)
k_potrs_group(int n, const T *__restrict__ L, size_t ldl_rt, size_t strideL, T *b, size_t strideB, size_t batch, PotrsPeers<T> peers) {
    constexpr int GROUPS = GPUB_POTRS_THREADS / NP;
    constexpr int LDP = NP + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / NP;
    const int i = threadIdx.x % NP;
    T *s_t = reinterpret_cast<T *>(smem_raw) + (size_t) grp * NP * LDP;
    const size_t ldl = DENSE ? (size_t) NP : ldl_rt;
    const size_t ngroups = (size_t) gridDim.x * GROUPS;
    const size_t iters = (batch + ngroups - 1) / ngroups;
    const bool row_ok = DENSE || i < n;

    for (size_t it = 0; it < iters; it++) {
        size_t mat = it * ngroups + (size_t) blockIdx.x * GROUPS + grp;
        const bool live = mat < batch;
        if (!live) mat = batch - 1;
        const T *l_g = L + mat * strideL;
        T *b_g = b + mat * strideB;
        T l[NP]; // strictly lower part of row i
#pragma unroll
        for (int c = 0; c < NP; c++) l[c] = (row_ok && c < i) ? l_g[i + (size_t) c * ldl] : T(0);
        const T diag = row_ok ? l_g[i + (size_t) i * ldl] : T(1);
        T x = row_ok ? b_g[i] : T(0);
        const T dinv = T(1) / diag;
#pragma unroll
        for (int c = 0; c < NP; c++) {
            l[c] *= dinv;
            s_t[i * LDP + c] = l[c];
        }
        x *= dinv;
        // forward: Lb y = D^-1 b
#pragma unroll
        for (int j = 0; j < NP - 1; j++) {
            const T yj = __shfl_sync(0xffffffffu, x, j, NP);
            x = fma(-l[j], yj, x);
        }
        __syncwarp();
        // column i of Lb: Lb(r, i) for r > i, zero elsewhere
#pragma unroll
        for (int r = 0; r < NP; r++) l[r] = s_t[r * LDP + i];
        __syncwarp();
        // backward: Lb^T v = y, z = D^-1 v
#pragma unroll
        for (int jj = NP - 1; jj > 0; jj--) {
            const T vj = __shfl_sync(0xffffffffu, x, jj, NP);
            x = fma(-l[jj], vj, x);
        }
        x *= dinv;
        if (row_ok && live) {
            b_g[i] = x;
            if (GATHER) {
#pragma unroll
                for (int pr = 0; pr < 8; pr++)
                    if (pr < peers.count) peers.x[pr][(peers.offset + mat) * peers.stride + i] = x;
            }
        }
    }
}

// p[c] summed over the 32 lanes, result for c == lane
template<typename T>
__device__ __forceinline__ T transpose_reduce32(T (&p)[32], int lane) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; k++) {
            const T keep = up ? p[o + k] : p[k];
            const T send = up ? p[k] : p[o + k];
            p[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return p[0];
}

// ------------------------------------------------------------------------------------------
// potrs, 32 < n <= 64: k_potrs_pair64<T, DENSE>, a WARP per matrix, lane p owns rows p and p + 32 of L in registers
// (the layout of k_potrf_pair<T, 64>), no CTA barrier. With the row-scaled factor Lb = D^-1 L:
//   forward  Lb y = D^-1 b : 64 steps of one shuffle + two FMAs (column j of Lb is spread over the lanes' registers);
//   backward Lb^T v = y, x = D^-1 v, by 32 x 32 blocks: the two diagonal blocks are transposed once through a padded
//            shared tile so that lane p owns COLUMN p (31 steps of one shuffle + one FMA each); the coupling block
//            Lb(32:64, 0:32)^T v_hi is 32 per-lane products reduced by a transpose-reduce butterfly (31 shuffles).
// L is read exactly once (lower triangle). Rows / columns beyond n are an identity pad when !DENSE.
// ------------------------------------------------------------------------------------------
// Pieces shared by k_potrs_pair64 and k_potrs_quad128: a 64 x 64 lower-triangular block held by one warp as row pairs
// (lane p: rows p and p + 32), already row-scaled to the strictly lower part of Lb = D^-1 L.
template<typename T, int H>
__device__ __forceinline__ void pair64_load_scaled(const T *l_g, size_t ldl, int nloc, int p, T (&lo)[H], T (&hi)[2 * H], T &ilo, T &ihi) {
    // l_g points at the block's (0, 0) entry; rows >= nloc are an identity pad. H lanes per block (H = 32: a warp per 64 x 64 block;
    // H = 16 / 8: two / four 32 x 32 / 16 x 16 matrices per warp), lane p of the group owns rows p and p + H
    const bool lo_ok = p < nloc, hi_ok = p + H < nloc;
#pragma unroll
    for (int c = 0; c < H; c++) lo[c] = (c <= p && lo_ok) ? l_g[p + c * ldl] : T(c == p ? 1 : 0);
#pragma unroll
    for (int c = 0; c < 2 * H; c++) hi[c] = (c <= p + H && hi_ok) ? l_g[p + H + c * ldl] : T(c == p + H ? 1 : 0);
    T dlo = T(1), dhi = T(1);
#pragma unroll
    for (int c = 0; c < H; c++) {
        if (c == p) dlo = lo[c];
        if (c == p) dhi = hi[c + H];
    }
    ilo = T(1) / dlo;
    ihi = T(1) / dhi;
#pragma unroll
    for (int c = 0; c < H; c++) lo[c] = c < p ? lo[c] * ilo : T(0);
#pragma unroll
    for (int c = 0; c < 2 * H; c++) hi[c] = c < p + H ? hi[c] * ihi : T(0);
}

// Lb y = x (x already scaled by D^-1); on exit xlo, xhi hold y. Shuffles stay inside the group of H lanes.
template<typename T, int H>
__device__ __forceinline__ void pair64_forward(const T (&lo)[H], const T (&hi)[2 * H], T &xlo, T &xhi) {
#pragma unroll
    for (int j = 0; j < H; j++) {
        const T yj = __shfl_sync(0xffffffffu, xlo, j, H);
        xlo = fma(-lo[j], yj, xlo);
        xhi = fma(-hi[j], yj, xhi);
    }
#pragma unroll
    for (int j = H; j < 2 * H - 1; j++) {
        const T yj = __shfl_sync(0xffffffffu, xhi, j - H, H);
        xhi = fma(-hi[j], yj, xhi);
    }
}

// Lb^T v = x; on exit xlo, xhi hold v (the caller multiplies by D^-1). tile: the group's [H][H + 1] scratch
template<typename T, int H>
__device__ __forceinline__ void pair64_backward(const T (&lo)[H], const T (&hi)[2 * H], T &xlo, T &xhi, T (*tile)[H + 1], int p) {
    __syncwarp();
#pragma unroll
    for (int c = 0; c < H; c++) tile[p][c] = hi[H + c];
    __syncwarp();
    {
        T cb[H];
#pragma unroll
        for (int r = 0; r < H; r++) cb[r] = tile[r][p];
#pragma unroll
        for (int jj = H - 1; jj > 0; jj--) {
            const T vj = __shfl_sync(0xffffffffu, xhi, jj, H);
            xhi = fma(-cb[jj], vj, xhi);
        }
    }
    {
        T pr[H];
#pragma unroll
        for (int c = 0; c < H; c++) pr[c] = hi[c] * xhi;
        TReduce<T, H, H / 2>::run(pr, p);
        xlo -= pr[0];
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < H; c++) tile[p][c] = lo[c];
    __syncwarp();
    {
        T cb[H];
#pragma unroll
        for (int r = 0; r < H; r++) cb[r] = tile[r][p];
#pragma unroll
        for (int jj = H - 1; jj > 0; jj--) {
            const T vj = __shfl_sync(0xffffffffu, xlo, jj, H);
            xlo = fma(-cb[jj], vj, xlo);
        }
    }
}

// potrs, n == 32 or 16, dense: k_potrs_pair<T, N>, N / 2 lanes per matrix (two / four matrices per warp), the layout and the pieces of
// k_potrs_pair64. Against the lane = row kernel (k_potrs_group, a warp per 32 x 32 matrix) every shuffle, shared-memory store and
// load of the substitution serves two (four) matrices: the kernel is bound by LSU wavefronts (ncu: 84 % busy), 124 of the ~320 per
// matrix being the 62 64-bit shuffles of the two sweeps.
#ifndef GPUB_POTRS_PAIR
#define GPUB_POTRS_PAIR 1
#endif
#ifndef GPUB_POTRS_PAIR16
#define GPUB_POTRS_PAIR16 1
#endif
template<typename T, int N, bool GATHER>
__global__ void __launch_bounds__(128, 4) k_potrs_pair(const T *__restrict__ L, size_t strideL, T *b, size_t strideB, size_t batch, PotrsPeers<T> peers) {
    constexpr int H = N / 2, GROUPS = 128 / H;
    __shared__ T s_t[GROUPS][H][H + 1];
    const int grp = threadIdx.x / H, p = threadIdx.x % H;
    const size_t ngroups = (size_t) gridDim.x * GROUPS;
    const size_t iters = (batch + ngroups - 1) / ngroups;
    for (size_t it = 0; it < iters; it++) {
        size_t mat = it * ngroups + (size_t) blockIdx.x * GROUPS + grp;
        const bool live = mat < batch;
        if (!live) mat = batch - 1;
        T *b_g = b + mat * strideB;
        T lo[H], hi[N], ilo, ihi;
        pair64_load_scaled<T>(L + mat * strideL, (size_t) N, N, p, lo, hi, ilo, ihi);
        T xlo = b_g[p] * ilo, xhi = b_g[p + H] * ihi;
        pair64_forward<T>(lo, hi, xlo, xhi);
        pair64_backward<T>(lo, hi, xlo, xhi, s_t[grp], p);
        xlo *= ilo;
        xhi *= ihi;
        if (live) {
            b_g[p] = xlo;
            b_g[p + H] = xhi;
            if (GATHER) {
#pragma unroll
                for (int pr = 0; pr < 8; pr++)
                    if (pr < peers.count) {
                        T *o = peers.x[pr] + (peers.offset + mat) * peers.stride;
                        o[p] = xlo;
                        o[p + H] = xhi;
                    }
            }
        }
    }
}

template<typename T, bool DENSE>
__global__ void __launch_bounds__(128, sizeof(T) == 8 ? 2 : 4) k_potrs_pair64(int n_rt, const T *__restrict__ L, size_t ldl_rt, size_t strideL, T *b,
                                                                               size_t strideB, size_t batch) {
    __shared__ T s_t[4][32][33];
    const int n = DENSE ? 64 : n_rt;
    const size_t ldl = DENSE ? (size_t) 64 : ldl_rt;
    const int warp = threadIdx.x >> 5, p = threadIdx.x & 31;
    const size_t nwarps = (size_t) gridDim.x * 4;
    const size_t iters = (batch + nwarps - 1) / nwarps;
    for (size_t it = 0; it < iters; it++) {
        size_t mat = it * nwarps + (size_t) blockIdx.x * 4 + warp;
        const bool live = mat < batch;
        if (!live) mat = batch - 1;
        T *b_g = b + mat * strideB;
        T lo[32], hi[64], ilo, ihi;
        pair64_load_scaled<T>(L + mat * strideL, ldl, n, p, lo, hi, ilo, ihi);
        T xlo = p < n ? b_g[p] * ilo : T(0), xhi = p + 32 < n ? b_g[p + 32] * ihi : T(0);
        pair64_forward<T>(lo, hi, xlo, xhi);
        pair64_backward<T>(lo, hi, xlo, xhi, s_t[warp], p);
        if (live) {
            if (p < n) b_g[p] = xlo * ilo;
            if (p + 32 < n) b_g[p + 32] = xhi * ihi;
        }
    }
}

// ------------------------------------------------------------------------------------------
// potrs, 64 < n <= 128: k_potrs_quad128<T>, one matrix per CTA of four warps, L = [L11 0; L21 L22] in 64 x 64 blocks:
//   warp 0 holds L11 and warp 3 holds L22 as row pairs (pair64 pieces above), warps 1 and 2 hold the 64 rows of L21 (one full
//   row of 64 entries per lane). Everything is row-scaled (Lb = D^-1 L) when loaded; L is read exactly once.
//   forward : warp 0 solves block 1 -> warps 1, 2: rhs2 -= Lb21 y1 (a 64-term dot product per lane against the broadcast y1)
//             -> warp 3 solves block 2 forward AND backward;
//   backward: warps 1, 2: Lb21^T v2 (64 per-lane products, two 31-shuffle transpose-reduce butterflies) -> warp 0 finishes.
// Five CTA barriers per matrix instead of the sixteen of a 32 x 32-block kernel, and no warp ever waits inside a 32-column block.
// ------------------------------------------------------------------------------------------
// The four warps of k_potrs_quad128 play different roles and meet at the CTA barrier from their own branches (every thread executes
// the same number of barriers). That is what named barriers are for at the PTX level (bar.sync id, count: warps may arrive from
// different instructions); __syncthreads() in role-dependent code is outside the CUDA C++ rules and compute-sanitizer synccheck
// reports it, so the kernel uses barrier 1 with an explicit thread count.
// synccheck additionally wants every warp to arrive from the SAME barrier instruction: built with -DGPUB_SYNCCHECK_CLEAN the barrier is
// one non-inlined function (synccheck: 0 errors on scripts/dev_sanitize.py); the call costs 7-10 % at n = 128, so the default inlines it.
#ifdef GPUB_SYNCCHECK_CLEAN
__device__ __noinline__ void quad_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
#else
__device__ __forceinline__ void quad_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
#endif

template<typename T>
__global__ void __launch_bounds__(128, sizeof(T) == 8 ? 2 : GPUB_QUAD128_F32_MINB) k_potrs_quad128(int n, const T *__restrict__ L, size_t ldl, size_t strideL, T *b,
                                                                                size_t strideB, size_t batch) {
    __shared__ T s_t[2][32][33];
    __shared__ __align__(16) T s_x[128];       // scaled right-hand side -> y -> v
    __shared__ T s_part[2][64];
    const int warp = threadIdx.x >> 5, p = threadIdx.x & 31;
    const int n2 = n - 64;                      // rows of the second block (1..64)
    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        const T *l_g = L + mat * strideL;
        T *b_g = b + mat * strideB;
        if (warp == 0 || warp == 3) {
            const int o = warp == 0 ? 0 : 64, nloc = warp == 0 ? 64 : n2;
            T lo[32], hi[64], ilo, ihi;
            pair64_load_scaled<T>(l_g + o + (size_t) o * ldl, ldl, nloc, p, lo, hi, ilo, ihi);
            T xlo = p < nloc ? b_g[o + p] * ilo : T(0), xhi = p + 32 < nloc ? b_g[o + p + 32] * ihi : T(0);
            if (warp == 0) {
                pair64_forward<T>(lo, hi, xlo, xhi);
                s_x[p] = xlo;
                s_x[p + 32] = xhi;
                quad_sync();                  // (1) y1 published
                quad_sync();                  // (2) warps 1, 2 have updated the second right-hand side
                quad_sync();                  // (3) v2 published
                quad_sync();                  // (4) coupling partials published
                xlo -= s_part[0][p] + s_part[1][p];
                xhi -= s_part[0][p + 32] + s_part[1][p + 32];
                pair64_backward<T>(lo, hi, xlo, xhi, s_t[0], p);
                b_g[p] = xlo * ilo;
                b_g[p + 32] = xhi * ihi;
            } else {
                s_x[64 + p] = xlo;             // scaled rhs of block 2, updated by warps 1, 2 after barrier (1)
                s_x[96 + p] = xhi;
                quad_sync();                  // (1)
                quad_sync();                  // (2)
                xlo = s_x[64 + p];
                xhi = s_x[96 + p];
                pair64_forward<T>(lo, hi, xlo, xhi);
                pair64_backward<T>(lo, hi, xlo, xhi, s_t[1], p);
                s_x[64 + p] = xlo;             // v2
                s_x[96 + p] = xhi;
                if (p < nloc) b_g[64 + p] = xlo * ilo;
                if (p + 32 < nloc) b_g[96 + p] = xhi * ihi;
                quad_sync();                  // (3)
                quad_sync();                  // (4)
            }
        } else {
            // one row of L21 per lane: row 64 + 32 (warp - 1) + p, scaled by the reciprocal diagonal of ITS row
            const int rloc = 32 * (warp - 1) + p, row = 64 + rloc;
            const bool ok = rloc < n2;
            T r[64];
#pragma unroll
            for (int c = 0; c < 64; c++) r[c] = ok ? l_g[row + c * ldl] : T(0);
            const T idg = ok ? T(1) / l_g[row + (size_t) row * ldl] : T(1);
#pragma unroll
            for (int c = 0; c < 64; c++) r[c] *= idg;
            quad_sync();                      // (1) y1 published
            T t0 = 0, t1 = 0;
#pragma unroll
            for (int c = 0; c < 64; c += 2) {
                t0 = fma(r[c], s_x[c], t0);
                t1 = fma(r[c + 1], s_x[c + 1], t1);
            }
            s_x[row] -= t0 + t1;
            quad_sync();                      // (2)
            quad_sync();                      // (3) v2 published
            const T v = s_x[row];
            T pr[32];
#pragma unroll
            for (int c = 0; c < 32; c++) pr[c] = r[c] * v;
            TReduce<T, 32, 16>::run(pr, p);
            s_part[warp - 1][p] = pr[0];
#pragma unroll
            for (int c = 0; c < 32; c++) pr[c] = r[32 + c] * v;
            TReduce<T, 32, 16>::run(pr, p);
            s_part[warp - 1][p + 32] = pr[0];
            quad_sync();                      // (4)
        }
        quad_sync();                          // (5) the shared vectors are free for the next matrix
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
    // fp32 only: the fp64 factorisation measured 0.78 ms thread-per-matrix against 0.37 ms lane-per-row (8388608 matrices)
    if (GPUB_CHOL4 && sizeof(T) == 4 && n == 4 && lda == 4 && strideA == 16 && (((uintptr_t) A) & 15u) == 0) {
        const size_t want = gpub_ceil_div(batch, (size_t) 128), cap = (size_t) ctx->sm_count * 16;
        k_potrf4<T><<<(unsigned) (want < cap ? want : cap), 128, 0, stream>>>(A, info, batch);
        GPUB_LAUNCH_CHECK();
        return GPUB_OK;
    }
    if (n <= 32) {
        const int np = n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32;
        const size_t groups = GPUB_POTRF_THREADS / np;
        const size_t want = gpub_ceil_div(batch, groups);
        // persistent-style grid: resident CTAs per SM x SM count, every CTA walks the batch with a grid stride
        const size_t cap = (size_t) ctx->sm_count * GPUB_POTRF_MINB * GPUB_GRID_WAVES;
        const unsigned grid = (unsigned) (want < cap ? want : cap);
        constexpr int TH = GPUB_POTRF_THREADS;
        const bool dense = (n == (size_t) np) && lda == n;
#define GPUB_POTRF_CASE(NPV)                                                                              \
    if (dense) k_potrf_group<T, NPV, true><<<grid, TH, 0, stream>>>((int) n, A, lda, strideA, info, batch); \
    else k_potrf_group<T, NPV, false><<<grid, TH, 0, stream>>>((int) n, A, lda, strideA, info, batch);
        if (dense && (n == 32 || n == 16)) {
            const size_t want32 = gpub_ceil_div(batch, (size_t) (256 / n)), cap32 = (size_t) ctx->sm_count * GPUB_POTRF32_MINB * GPUB_GRID_WAVES;
            const unsigned g32 = (unsigned) (want32 < cap32 ? want32 : cap32);
            if (n == 32) k_potrf_pair<T, 32, true><<<g32, 128, 0, stream>>>((int) n, A, lda, strideA, info, batch);
            else k_potrf_pair<T, 16, true><<<g32, 128, 0, stream>>>((int) n, A, lda, strideA, info, batch);
            GPUB_LAUNCH_CHECK();
            return GPUB_OK;
        }
        switch (np) {
            case 4: GPUB_POTRF_CASE(4) break;
            case 8: GPUB_POTRF_CASE(8) break;
            case 16: GPUB_POTRF_CASE(16) break;
            default: GPUB_POTRF_CASE(32) break;
        }
#undef GPUB_POTRF_CASE
    } else if (n <= 64) {
        // a warp per matrix, the triangle in registers (k_potrf_pair<T, 64>): 4 matrices per CTA
        const size_t want = gpub_ceil_div(batch, (size_t) 4), cap = (size_t) ctx->sm_count * PotrfPairMinB<T, 64>::value * 4;   // measured: 4 waves of CTAs beat 2 by 4-7 % here (shorter tail)
        const unsigned grid = (unsigned) (want < cap ? want : cap);
        if (n == 64 && lda == 64) k_potrf_pair<T, 64, true><<<grid, 128, 0, stream>>>((int) n, A, lda, strideA, info, batch);
        else k_potrf_pair<T, 64, false><<<grid, 128, 0, stream>>>((int) n, A, lda, strideA, info, batch);
    } else if (n <= 128) {
        const size_t cap = (size_t) ctx->sm_count * 8;
        const unsigned grid = (unsigned) (batch < cap ? batch : cap);
        if (n <= 96) {
            constexpr int LDPH = (GPUB_BLK_DMMA && sizeof(T) == 8) ? 36 : 32;
            const size_t smem = sizeof(T) * ((size_t) (2 * 3 + 1) * 32 * LDPH + 32);
            GPUB_CUDA(cudaFuncSetAttribute(k_potrf_pipe<T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            k_potrf_pipe<T, 3><<<grid, 32 * 6, smem, stream>>>((int) n, A, lda, strideA, info, batch);
        } else {
            constexpr int LDPH = (GPUB_BLK_DMMA && sizeof(T) == 8) ? 36 : 32;
            const size_t smem = sizeof(T) * ((size_t) (2 * 4 + 1) * 32 * LDPH + 32);
            GPUB_CUDA(cudaFuncSetAttribute(k_potrf_pipe<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            k_potrf_pipe<T, 4><<<grid, 32 * 10, smem, stream>>>((int) n, A, lda, strideA, info, batch);
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
int potrs_batched(gpub_ctx_t ctx, int sidx, size_t n, const T *L, size_t ldl, size_t strideL, T *b, size_t strideB, size_t batch,
                  const PotrsPeers<T> *peers = nullptr) {
    if (n == 0 || batch == 0) return GPUB_OK;
    if (!L || !b || ldl < n) return GPUB_EINVAL;
    if (n > 8192) return GPUB_ENOTSUP;
    GPUB_ENTER(ctx, sidx);
    if (peers && n > 32) {
        // shapes without a fused kernel: plain solve, then one peer copy of the shard per destination (same stream)
        int e = potrs_batched<T>(ctx, sidx, n, L, ldl, strideL, b, strideB, batch, nullptr);
        if (e) return e;
        for (int p = 0; p < peers->count; p++)
            GPUB_CUDA(cudaMemcpy2DAsync(peers->x[p] + peers->offset * peers->stride, peers->stride * sizeof(T), b, strideB * sizeof(T), n * sizeof(T),
                                        batch, cudaMemcpyDefault, stream));
        return GPUB_OK;
    }
    if (!peers && GPUB_CHOL4 && n == 4 && ldl == 4 && strideL == 16 && strideB == 4 && ((((uintptr_t) L) | ((uintptr_t) b)) & 15u) == 0) {
        const size_t want = gpub_ceil_div(batch, (size_t) 128), cap = (size_t) ctx->sm_count * 16;
        k_potrs4<T><<<(unsigned) (want < cap ? want : cap), 128, 0, stream>>>(L, b, batch);
        GPUB_LAUNCH_CHECK();
        return GPUB_OK;
    }
    if (!peers && GPUB_CHOL4 && sizeof(T) == 4 && n == 8 && ldl == 8 && strideL == 64 && strideB == 8 && ((((uintptr_t) L) | ((uintptr_t) b)) & 15u) == 0) {
        const size_t want = gpub_ceil_div(batch, (size_t) 128), cap = (size_t) ctx->sm_count * 16;
        k_potrs8_f32<<<(unsigned) (want < cap ? want : cap), 128, 0, stream>>>((const float *) L, (float *) b, batch);
        GPUB_LAUNCH_CHECK();
        return GPUB_OK;
    }
#if GPUB_POTRS_PAIR
    if ((n == 32 || (GPUB_POTRS_PAIR16 && n == 16)) && ldl == n && strideB >= n) {
        // two (four) matrices per warp, row pairs in registers
        const size_t want = gpub_ceil_div(batch, (size_t) (256 / n)), cap = (size_t) ctx->sm_count * 4 * GPUB_GRID_WAVES;
        const unsigned grid = (unsigned) (want < cap ? want : cap);
        const PotrsPeers<T> none = PotrsPeers<T>();
        if (n == 32) {
            if (peers) k_potrs_pair<T, 32, true><<<grid, 128, 0, stream>>>(L, strideL, b, strideB, batch, *peers);
            else k_potrs_pair<T, 32, false><<<grid, 128, 0, stream>>>(L, strideL, b, strideB, batch, none);
        } else {
            if (peers) k_potrs_pair<T, 16, true><<<grid, 128, 0, stream>>>(L, strideL, b, strideB, batch, *peers);
            else k_potrs_pair<T, 16, false><<<grid, 128, 0, stream>>>(L, strideL, b, strideB, batch, none);
        }
        GPUB_LAUNCH_CHECK();
        return GPUB_OK;
    }
#endif
    if (n <= 32) {
        const int np = n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32;
        const size_t groups = GPUB_POTRS_THREADS / np;
        const size_t want = gpub_ceil_div(batch, groups);
        const size_t cap = (size_t) ctx->sm_count * GPUB_POTRS_MINB * GPUB_GRID_WAVES;
        const unsigned grid = (unsigned) (want < cap ? want : cap);
        const size_t smem = groups * np * (np + 1) * sizeof(T);
        const bool dense = (n == (size_t) np) && ldl == n;
#define GPUB_POTRS_LAUNCH(NPV, DN)                                                                                   \
    {                                                                                                                \
        if (peers) {                                                                                                 \
            auto kern = k_potrs_group<T, NPV, DN, true>;                                                             \
            if (smem > 48 * 1024) GPUB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
            kern<<<grid, GPUB_POTRS_THREADS, smem, stream>>>((int) n, L, ldl, strideL, b, strideB, batch, *peers);   \
        } else {                                                                                                     \
            auto kern = k_potrs_group<T, NPV, DN, false>;                                                            \
            if (smem > 48 * 1024) GPUB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
            kern<<<grid, GPUB_POTRS_THREADS, smem, stream>>>((int) n, L, ldl, strideL, b, strideB, batch, PotrsPeers<T>()); \
        }                                                                                                            \
    }
#define GPUB_POTRS_CASE(NPV)                                                                                         \
    if (dense) GPUB_POTRS_LAUNCH(NPV, true) else GPUB_POTRS_LAUNCH(NPV, false)
        switch (np) {
            case 4: GPUB_POTRS_CASE(4) break;
            case 8: GPUB_POTRS_CASE(8) break;
            case 16: GPUB_POTRS_CASE(16) break;
            default: GPUB_POTRS_CASE(32) break;
        }
#undef GPUB_POTRS_LAUNCH
#undef GPUB_POTRS_CASE
    } else if (n <= 64) {
        const size_t want = gpub_ceil_div(batch, (size_t) 4), cap = (size_t) ctx->sm_count * (sizeof(T) == 8 ? 2 : 4) * GPUB_GRID_WAVES;
        const unsigned grid = (unsigned) (want < cap ? want : cap);
        if (n == 64 && ldl == 64) k_potrs_pair64<T, true><<<grid, 128, 0, stream>>>((int) n, L, ldl, strideL, b, strideB, batch);
        else k_potrs_pair64<T, false><<<grid, 128, 0, stream>>>((int) n, L, ldl, strideL, b, strideB, batch);
    } else if (n <= 128) {
        const size_t cap = (size_t) ctx->sm_count * 8;
        const unsigned grid = (unsigned) (batch < cap ? batch : cap);
        k_potrs_quad128<T><<<grid, 128, 0, stream>>>((int) n, L, ldl, strideL, b, strideB, batch);
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

#define GPUB_DEF_POTRS_GATHER(SUF, T)                                                                                 \
    int gpub_potrs_allgather_batched_##SUF(gpub_ctx_t c, int s, size_t n, const T *L, size_t ldl, size_t sL, T *b, size_t sB, size_t bt, \
                                           T *const *peer_x, int n_peers, size_t shard_offset, size_t stride_x) {     \
        if (n_peers < 0 || n_peers > 8 || (n_peers && !peer_x) || stride_x < n) return GPUB_EINVAL;                   \
        PotrsPeers<T> pp;                                                                                             \
        for (int p = 0; p < 8; p++) pp.x[p] = p < n_peers ? peer_x[p] : nullptr;                                      \
        for (int p = 0; p < n_peers; p++) if (!pp.x[p]) return GPUB_EINVAL;                                           \
        pp.count = n_peers; pp.offset = shard_offset; pp.stride = stride_x;                                           \
        return potrs_batched<T>(c, s, n, L, ldl, sL, b, sB, bt, n_peers ? &pp : nullptr);                             \
    }
GPUB_DEF_POTRS_GATHER(f64, double)
GPUB_DEF_POTRS_GATHER(f32, float)

} // extern "C"
