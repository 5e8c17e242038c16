// Batched Householder QR, application of Q / Q^T, upper-triangular solve and fused least squares.
// Replaces cusolverDn{D,S}geqrf / ormqr, cublas{D,S}trsm and cublas{D,S}gelsBatched
// (ref: tensor.cuh:1866-1927, 1929-1995, 1340-1394).
//
// Storage follows LAPACK: R on and above the diagonal, the reflector v_j (v_j[j] = 1 implicit) below it,
// H_j = I - tau_j v_j v_j^T, Q = H_0 H_1 ... H_{n-1}. The sign rule is LAPACK's dlarfg:
// beta = -sign(alpha) * ||(alpha, x)||, tau = (beta - alpha) / beta, v = x / (alpha - beta), and tau = 0
// (H = I) when x == 0 -- the rule SVD sign parity depends on (hard part 2).
//
//   k_gels_sub<T, M, N, LPM> : fused gels for small tall systems (cfg3: 64 x 16 fp32). LPM lanes per matrix,
//        each lane owns M/LPM consecutive rows of [A | b] in registers (128-bit loads), the column norms and the
//        v^T A products are log2(LPM)-step butterflies inside the group, R x = Q^T b is solved in registers.
//   k_geqrf_cta / k_gels_cta / k_ormqr_cta / k_trsv_cta : any shape, one matrix per CTA, staged in shared
//        memory when it fits, else in place in global memory (L2 resident).
#include "common.cuh"
#ifndef GPUB_GRID_WAVES
#define GPUB_GRID_WAVES 2   // persistent grids: resident CTAs per SM x SM count x this
#endif

namespace {

constexpr int QT = 256; // threads per CTA for the generic kernels

template<typename T> __device__ __forceinline__ T t_sqrt(T x);
template<> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template<> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }

template<typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template<typename T>
__device__ __forceinline__ T cta_sum(T v, T *s_red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    T tot = 0;
#pragma unroll
    for (int w = 0; w < QT / 32; w++) tot += s_red[w];
    __syncthreads();
    return tot;
}

// dlarfg on (alpha, xnorm): returns beta, writes tau and the scale 1/(alpha-beta) for x
template<typename T>
__device__ __forceinline__ T larfg(T alpha, T xnorm2, T *tau, T *scale) {
    // a column at the bottom of the exponent range counts as already reduced, like an exact zero (squares of its entries are
    // subnormal: tau and v would no longer make an orthogonal reflector -- 1e-5 off for the all-ones matrix, whose every column
    // is rounding noise of the one before); what is dropped is below 1e-145 (fp64) / 1e-15 (fp32) in absolute value
    if (xnorm2 == T(0) || !(alpha * alpha + xnorm2 >= (sizeof(T) == 8 ? T(1e-290) : T(1e-30)))) {
        *tau = T(0);
        *scale = T(0);
        return alpha;
    }
    T nrm = t_sqrt<T>(alpha * alpha + xnorm2);
    T beta = alpha >= T(0) ? -nrm : nrm;
    *tau = (beta - alpha) / beta;
    *scale = T(1) / (alpha - beta);
    return beta;
}

// Factor the first n columns of the m x ncols panel M (leading dimension ld) in place;
// columns n..ncols-1 (right-hand sides) are only transformed by Q^T.
template<typename T>
__device__ void householder_panel(T *M, size_t ld, int m, int n, int ncols, T *tau_out, T *s_red, T *s_bc) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kmax = n < m ? n : m;
    for (int j = 0; j < kmax; j++) {
        T *cj = M + (size_t) j * ld;
        T part = 0;
        for (int r = j + 1 + tid; r < m; r += QT) part = fma(cj[r], cj[r], part);
        const T xnorm2 = cta_sum(part, s_red);
        if (tid == 0) {
            T tau, scale;
            T beta = larfg<T>(cj[j], xnorm2, &tau, &scale);
            cj[j] = beta;
            s_bc[0] = tau;
            s_bc[1] = scale;
            if (tau_out) tau_out[j] = tau;
        }
        __syncthreads();
        const T tau = s_bc[0], scale = s_bc[1];
        if (tau != T(0)) {
            for (int r = j + 1 + tid; r < m; r += QT) cj[r] *= scale;
            __syncthreads();
            for (int c = j + 1 + warp; c < ncols; c += QT / 32) {
                T *cc = M + (size_t) c * ld;
                T w = 0;
                for (int r = j + 1 + lane; r < m; r += 32) w = fma(cj[r], cc[r], w);
                w = warp_sum(w) + cc[j];
                const T tw = tau * w;
                __syncwarp();                               // every lane has read cc[j] before lane 0 overwrites it
                if (lane == 0) cc[j] -= tw;
                for (int r = j + 1 + lane; r < m; r += 32) cc[r] = fma(-tw, cj[r], cc[r]);
            }
        }
        __syncthreads();
    }
}

// back substitution R x = y on the leading n x n of M, y = column `rhs` (length >= n), in place
template<typename T>
__device__ void upper_solve(const T *M, size_t ld, int n, T *y, int *bad, T *s_bc) {
    const int tid = threadIdx.x;
    for (int j = n - 1; j >= 0; j--) {
        if (tid == 0) {
            T d = M[j + (size_t) j * ld];
            if (d == T(0) && bad && *bad == 0) *bad = j + 1;
            y[j] = y[j] / d;
            s_bc[0] = y[j];
        }
        __syncthreads();
        const T xj = s_bc[0];
        const T *cj = M + (size_t) j * ld;
        for (int r = tid; r < j; r += QT) y[r] = fma(-cj[r], xj, y[r]);
        __syncthreads();
    }
}

template<typename T>
__device__ __forceinline__ void copy_in(T *dst, size_t ldd, const T *src, size_t lds, int m, int n) {
    for (size_t e = threadIdx.x; e < (size_t) m * n; e += QT) {
        int r = (int) (e % m), c = (int) (e / m);
        dst[r + (size_t) c * ldd] = src[r + (size_t) c * lds];
    }
}

// mode 0: geqrf (tau out); mode 1: gels (b is column n, solved in place)
template<typename T>
__global__ void __launch_bounds__(QT) k_qr_cta(int m, int n, T *A, size_t lda, size_t sA, T *tau, size_t sTau, T *b, size_t sB,
                                                int *info, size_t batch, int use_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    __shared__ T s_red[QT / 32];
    __shared__ T s_bc[2];
    __shared__ int s_bad;
    const bool gels = b != nullptr;
    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        T *a_g = A + mat * sA;
        T *b_g = gels ? b + mat * sB : nullptr;
        T *tau_g = tau ? tau + mat * sTau : nullptr;
        if (threadIdx.x == 0) s_bad = 0;
        if (use_smem) {
            const size_t ld = (size_t) m | 1; // odd leading dimension: column sweeps by different warps spread over banks
            copy_in(sm, ld, a_g, lda, m, n);
            if (gels)
                for (int r = threadIdx.x; r < m; r += QT) sm[r + (size_t) n * ld] = b_g[r];
            __syncthreads();
            householder_panel<T>(sm, ld, m, n, gels ? n + 1 : n, tau_g, s_red, s_bc);
            if (gels) upper_solve<T>(sm, ld, n, sm + (size_t) n * ld, &s_bad, s_bc);
            copy_in(a_g, lda, sm, ld, m, n);
            if (gels)
                for (int r = threadIdx.x; r < m; r += QT) b_g[r] = sm[r + (size_t) n * ld];
        } else {
            // in place in global memory (L2 resident); gels never takes this branch, see gels_batched()
            __syncthreads();
            householder_panel<T>(a_g, lda, m, n, n, tau_g, s_red, s_bc);
        }
        __syncthreads();
        if (info && threadIdx.x == 0) info[mat] = s_bad;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Blocked Householder QR for tall matrices that do not fit in shared memory (cfg4: 1024 x 128 fp64):
// k_geqrf_wy<T, NBQ>, one matrix per CTA, compact-WY with panels of NBQ columns.
//   panel   : the m x NBQ panel lives in shared memory, stored [column][row] (conflict-free for lanes walking
//             rows). Per column: (A) |x|^2, (B) larfg by one thread, (C) scale v and, in the same sweep, the
//             dot products of v with the remaining panel columns and with the previous reflectors (those give
//             the column of the triangular factor T), (D) rank-1 update of the remaining panel columns fused
//             with the |x|^2 accumulation of the next column: 3 CTA barriers per column.
//   trailing: A2 <- (I - V T^T V^T) A2. Each warp takes two trailing columns at a time: one sweep over the
//             rows accumulates V^T a for both columns (NBQ shared loads feed 2*NBQ FMAs), a butterfly reduces
//             them, every lane forms W = T^T (V^T a) redundantly, a second sweep applies a -= V W.
//             V is made explicit in shared memory (unit diagonal, zeros above) so the sweeps are branch-free.
// The matrix itself stays in global memory (1 MB per matrix: L2 resident while its CTA works on it).
// ------------------------------------------------------------------------------------------
template<typename T, int NBQ>
__global__ void __launch_bounds__(QT) k_geqrf_wy(int m, int n, T *A, size_t lda, size_t sA, T *tau, size_t sTau, size_t batch, int ldv) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *Vp = reinterpret_cast<T *>(smem_raw);          // [NBQ][ldv]
    T *Tm = Vp + (size_t) NBQ * ldv;                  // [NBQ][NBQ], T(i, k) at Tm[i * NBQ + k], upper triangular
    T *s_red = Tm + NBQ * NBQ;                        // [QT/32][NBQ] per-warp partial sums
    T *s_misc = s_red + (QT / 32) * NBQ;              // tau, scale, ...
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NWARP = QT / 32;

    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        T *a_g = A + mat * sA;
        T *tau_g = tau + mat * sTau;
        const int kmax = n < m ? n : m;
        for (int j0 = 0; j0 < kmax; j0 += NBQ) {
            const int nb = (kmax - j0) < NBQ ? (kmax - j0) : NBQ;
            const int rows = m - j0;
            // ---- load the panel: rows j0.., columns j0..j0+nb-1 ----
            for (int k = 0; k < nb; k++)
                for (int r = tid; r < rows; r += QT) Vp[(size_t) k * ldv + r] = a_g[(size_t) (j0 + r) + (size_t) (j0 + k) * lda];
            for (int e = tid; e < NBQ * NBQ; e += QT) Tm[e] = T(0);
            __syncthreads();
            // |x|^2 of column 0 (rows > 0)
            T nrm_part = 0;
            for (int r = 1 + tid; r < rows; r += QT) nrm_part = fma(Vp[r], Vp[r], nrm_part);
            for (int jj = 0; jj < nb; jj++) {
                T *vj = Vp + (size_t) jj * ldv;
                // (A) finish the norm
                nrm_part = warp_sum(nrm_part);
                if (lane == 0) s_red[warp * NBQ] = nrm_part;
                __syncthreads();
                // (B) reflector parameters
                if (tid == 0) {
                    T x2 = 0;
                    for (int w = 0; w < NWARP; w++) x2 += s_red[w * NBQ];
                    T t, sc;
                    const T beta = larfg<T>(vj[jj], x2, &t, &sc);
                    vj[jj] = beta;
                    s_misc[0] = t;
                    s_misc[1] = sc;
                    tau_g[j0 + jj] = t;
                    Tm[jj * NBQ + jj] = t;
                }
                __syncthreads();
                const T tj = s_misc[0], scale = s_misc[1];
                // (C) scale v; dots with the remaining panel columns (slots jj+1..nb-1) and previous reflectors (slots 0..jj-1)
                T dots[NBQ];
#pragma unroll
                for (int c = 0; c < NBQ; c++) dots[c] = T(0);
                if (tj != T(0)) {
                    for (int r = jj + tid; r < rows; r += QT) {
                        T v = T(1);
                        if (r > jj) {
                            v = vj[r] * scale;
                            vj[r] = v;
                        }
#pragma unroll
                        for (int c = 0; c < NBQ; c++) {
                            if (c < nb && c != jj) {
                                // previous reflector i = c < jj: its entry at row r (r >= jj > i) is stored; unit / zero parts never reached
                                dots[c] = fma(v, Vp[(size_t) c * ldv + r], dots[c]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < NBQ; c++) {
                    const T d = warp_sum(dots[c]);
                    if (lane == 0) s_red[warp * NBQ + c] = d;
                }
                __syncthreads();
                // (D) update the remaining panel columns, column jj of T, and start the next norm
#pragma unroll
                for (int c = 0; c < NBQ; c++) {
                    T d = 0;
#pragma unroll
                    for (int w = 0; w < NWARP; w++) d += s_red[w * NBQ + c];
                    dots[c] = d;
                }
                nrm_part = 0;
                if (tj != T(0)) {
                    for (int r = jj + tid; r < rows; r += QT) {
                        const T v = r > jj ? vj[r] : T(1);
#pragma unroll
                        for (int c = 0; c < NBQ; c++) {
                            if (c > jj && c < nb) {
                                T *pc = Vp + (size_t) c * ldv + r;
                                const T nv = fma(-tj * dots[c], v, *pc);
                                *pc = nv;
                                if (c == jj + 1 && r > jj + 1) nrm_part = fma(nv, nv, nrm_part);
                            }
                        }
                    }
                } else if (jj + 1 < nb) {
                    for (int r = jj + 2 + tid; r < rows; r += QT) {
                        const T x = Vp[(size_t) (jj + 1) * ldv + r];
                        nrm_part = fma(x, x, nrm_part);
                    }
                }
                if (warp == 0 && jj > 0) {
                    // T(0:jj, jj) = -tau_j * T(0:jj, 0:jj) * (V(:, 0:jj)^T v_j); lane i computes row i
                    if (lane < NBQ) {
                        T d = 0;
                        for (int w = 0; w < NWARP; w++) d += s_red[w * NBQ + lane];
                        s_misc[2 + lane] = d;
                    }
                    __syncwarp();
                    if (lane < jj) {
                        T acc = 0;
                        for (int i2 = lane; i2 < jj; i2++) acc = fma(Tm[lane * NBQ + i2], s_misc[2 + i2], acc);
                        Tm[lane * NBQ + jj] = -tj * acc;
                    }
                }
                __syncthreads();
            }
            // ---- write the factored panel back (R above / on the diagonal, reflectors below) ----
            for (int k = 0; k < nb; k++)
                for (int r = tid; r < rows; r += QT) a_g[(size_t) (j0 + r) + (size_t) (j0 + k) * lda] = Vp[(size_t) k * ldv + r];
            __syncthreads();
            const int ntrail = n - (j0 + nb);
            if (ntrail > 0) {
                // explicit V: zeros above, ones on the diagonal
                for (int e = tid; e < nb * nb; e += QT) {
                    const int k = e / nb, r = e % nb;
                    if (r <= k) Vp[(size_t) k * ldv + r] = (r == k) ? T(1) : T(0);
                }
                __syncthreads();
                for (int cp = warp * 2; cp < ntrail; cp += NWARP * 2) {
                    const bool two = cp + 1 < ntrail;
                    T *c0 = a_g + (size_t) j0 + (size_t) (j0 + nb + cp) * lda;
                    T *c1 = two ? c0 + lda : c0;
                    T w0[NBQ], w1[NBQ];
#pragma unroll
                    for (int k = 0; k < NBQ; k++) w0[k] = w1[k] = T(0);
#pragma unroll 4
                    for (int r = lane; r < rows; r += 32) {
                        const T x0 = c0[r], x1 = c1[r];
#pragma unroll
                        for (int k = 0; k < NBQ; k++) {
                            if (k < nb) {
                                const T v = Vp[(size_t) k * ldv + r];
                                w0[k] = fma(v, x0, w0[k]);
                                w1[k] = fma(v, x1, w1[k]);
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < NBQ; k++) {
                        w0[k] = warp_sum(w0[k]);
                        w1[k] = warp_sum(w1[k]);
                    }
                    // W = T^T w  (W_k = sum_{i <= k} T(i, k) w_i), computed downwards in place from the last entry
#pragma unroll
                    for (int k = NBQ - 1; k >= 0; k--) {
                        T s0 = 0, s1 = 0;
#pragma unroll
                        for (int i2 = 0; i2 <= k; i2++) {
                            const T t = Tm[i2 * NBQ + k];
                            s0 = fma(t, w0[i2], s0);
                            s1 = fma(t, w1[i2], s1);
                        }
                        w0[k] = s0;
                        w1[k] = s1;
                    }
#pragma unroll 4
                    for (int r = lane; r < rows; r += 32) {
                        T x0 = c0[r], x1 = c1[r];
#pragma unroll
                        for (int k = 0; k < NBQ; k++) {
                            if (k < nb) {
                                const T v = Vp[(size_t) k * ldv + r];
                                x0 = fma(-v, w0[k], x0);
                                x1 = fma(-v, w1[k], x1);
                            }
                        }
                        c0[r] = x0;
                        if (two) c1[r] = x1;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// fp64 tall-skinny QR on the FP64 tensor pipe: k_geqrf_tc<RPT>, one matrix per CTA, 512 threads, m <= 512 * RPT.
// Compact WY with panels of 16 columns (cfg4: 1024 x 128, 8 panels):
//   panel   : every thread keeps its RPT rows of the 16 panel columns in REGISTERS. One CTA barrier per column: the
//             dots x^T a_c of the unscaled pivot column with itself and with the columns to its right are reduced
//             together (transpose-reduce butterfly per warp, 16 per-warp partials through shared memory), then every
//             warp derives beta, tau, 1/(alpha - beta) and the update coefficients tau (a_c[j] + scale x^T a_c)
//             redundantly, so no second barrier is needed to broadcast them.
//   T       : G = V^T V by DMMA (m8n8k4) tiles over 16 row ranges, then T(i, j) = -tau_j sum_k T(i, k) G(k, j), one
//             lane per row of T (no cross-lane dependency).
//   trailing: A2 <- A2 - V (T^T (V^T A2)). One warp per block of 8 trailing columns. Pass 1: W = V^T A2 with V
//             fragments from shared memory (stored [column][row], leading dimension = 4 mod 16: both fragment
//             patterns are bank-conflict free) and A2 fragments straight from global memory (every LDG.64 of the warp
//             covers whole sectors), two independent accumulator sets. W' = -T^T W in the warp's own scratch.
//             Pass 2: A2 tile (8 x 8, DMMA accumulator layout) += V W', 4 DMMA per tile, read and written in place.
//             Passes of different warps are independent: no CTA barrier inside the trailing update.
// Output is LAPACK's (R on / above the diagonal, reflectors below, tau), same dlarfg rule as the other kernels.
// ------------------------------------------------------------------------------------------
#ifdef GPUB_TCQ_PROFILE
__device__ unsigned long long g_tcq_prof[12];
#define TCQ_T(idx) do { __syncthreads(); if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&g_tcq_prof[idx], (unsigned long long) (t_ - tcq_t0)); tcq_t0 = t_; } } while (0)
#else
#define TCQ_T(idx) do { } while (0)
#endif
constexpr int TCQ_THREADS = 512;
constexpr int TCQ_WARPS = TCQ_THREADS / 32;
constexpr int TCQ_NB = 16;
#ifndef GPUB_TCQ_D1
#define GPUB_TCQ_D1 4
#endif
#ifndef GPUB_TCQ_D2
#define GPUB_TCQ_D2 1
#endif
constexpr int TCQ_D1 = GPUB_TCQ_D1;   // pass 1: depth of the register pipeline (chunks of 8 rows)
constexpr int TCQ_D2 = GPUB_TCQ_D2;   // pass 2: depth of the register pipeline (blocks of 16 rows)
constexpr int TCQ_NSL = 2;  // column blocks (slots) per warp in one chunk of the trailing update
constexpr int TCQ_LDW = 20; // scratch leading dimension, = 4 mod 16
constexpr int TCQ_PART = 4 * 16 * 128;      // doubles: partial W of 4 row groups x 16 column blocks (>= 16 x 256 for the Gram partials)
constexpr int TCQ_WP = 16 * 8 * TCQ_LDW;    // doubles: W' of 16 column blocks

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

__host__ __device__ constexpr int tcq_pow2ceil(int x) { return x <= 1 ? 1 : (x <= 2 ? 2 : (x <= 4 ? 4 : (x <= 8 ? 8 : 16))); }

// shared-memory layout: fixed-size regions first, so every address is base + compile-time offset (one live pointer)
constexpr int TCQ_OFF_TS = 0;                                  // [16][17]  T
constexpr int TCQ_OFF_GS = TCQ_OFF_TS + TCQ_NB * 17;           // [16][17]  G = V^T V
constexpr int TCQ_OFF_RED = TCQ_OFF_GS + TCQ_NB * 17;          // [2][16 warps][16] per-warp partial dots
constexpr int TCQ_OFF_PIV = TCQ_OFF_RED + 2 * TCQ_WARPS * TCQ_NB; // [2][16] pivot row
constexpr int TCQ_OFF_TAU = TCQ_OFF_PIV + 2 * TCQ_NB;          // [16]
constexpr int TCQ_OFF_PART = TCQ_OFF_TAU + TCQ_NB;             // [4 row groups][16 column blocks][8 columns][16] partial W
                                                               // (also the 16 per-warp 16 x 16 partial Gram tiles)
constexpr int TCQ_OFF_WP = TCQ_OFF_PART + TCQ_PART;            // [16 column blocks][8][TCQ_LDW]  W' = -T^T W
constexpr int TCQ_OFF_VS = TCQ_OFF_WP + TCQ_WP;                // [16][ldv] explicit V, multiple of 2 doubles (16-byte aligned)
struct TcqShared {
    double *base;
    __device__ __forceinline__ double *Ts() const { return base + TCQ_OFF_TS; }
    __device__ __forceinline__ double *Gs() const { return base + TCQ_OFF_GS; }
    __device__ __forceinline__ double *red() const { return base + TCQ_OFF_RED; }
    __device__ __forceinline__ double *piv() const { return base + TCQ_OFF_PIV; }
    __device__ __forceinline__ double *taus() const { return base + TCQ_OFF_TAU; }
    __device__ __forceinline__ double *part() const { return base + TCQ_OFF_PART; }
    __device__ __forceinline__ double *wp() const { return base + TCQ_OFF_WP; }
    __device__ __forceinline__ double *Vs() const { return base + TCQ_OFF_VS; }
};

template<int RPT, int JJ>
__device__ __forceinline__ void tcq_panel_column(double (&p)[TCQ_NB][RPT], const int (&rr)[RPT], const TcqShared &sh, int warp, int lane,
                                                 double *tau_g, int tau_ok_upto) {
    constexpr int NV = TCQ_NB - JJ;          // values reduced: column JJ with itself and with columns JJ+1..15
    constexpr int P = tcq_pow2ceil(NV);
    const int buf = JJ & 1;
    double part[P];
#pragma unroll
    for (int s = 0; s < P; s++) part[s] = 0.0;
#pragma unroll
    for (int i = 0; i < RPT; i++) {
        const double x = rr[i] > JJ ? p[JJ][i] : 0.0;
#pragma unroll
        for (int s = 0; s < NV; s++) part[s] = fma(x, p[JJ + s][i], part[s]);
    }
    TReduce<double, P, 16>::run(part, lane);
    if ((lane & (32 / P - 1)) == 0) sh.red()[(buf * TCQ_WARPS + warp) * TCQ_NB + lane / (32 / P)] = part[0];
    if (rr[0] == JJ) {                       // the owner of the pivot row publishes it
#pragma unroll
        for (int s = 0; s < NV; s++) sh.piv()[buf * TCQ_NB + s] = p[JJ + s][0];
    }
    __syncthreads();
    // lane l < NV: total of value l (column JJ + l) and the pivot-row entry of that column
    double tot = 0.0, tot2 = 0.0, pv = 0.0;
    if (lane < NV) {
#pragma unroll
        for (int w = 0; w < TCQ_WARPS; w += 2) {
            tot += sh.red()[(buf * TCQ_WARPS + w) * TCQ_NB + lane];
            tot2 += sh.red()[(buf * TCQ_WARPS + w + 1) * TCQ_NB + lane];
        }
        tot += tot2;
        pv = sh.piv()[buf * TCQ_NB + lane];
    }
    const double xnorm2 = __shfl_sync(0xffffffffu, tot, 0);
    const double alpha = __shfl_sync(0xffffffffu, pv, 0);
    double tau = 0.0, scale = 0.0, beta = alpha;
    // a column whose norm is at the bottom of the exponent range (rounding noise of rounding noise: every column of an all-ones matrix
    // is 1e-16 of the one before, and the flush-to-zero seeds below turn a subnormal sum of squares into inf * 0 = NaN) counts as
    // already reduced, like an exact zero: tau = 0. What is dropped is below 1e-145 in absolute value.
    if (xnorm2 != 0.0 && fma(alpha, alpha, xnorm2) >= 1e-290) {
        // dlarfg without the sqrt / divide slow paths (every warp evaluates this redundantly): rsqrt and rcp seeds
        // plus Newton steps, each result corrected once more against its defining residual (< 1 ulp)
        const double ss = fma(alpha, alpha, xnorm2);
        double rn;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rn) : "d"(ss));
        const double hs = 0.5 * ss;
        rn = fma(rn, fma(-hs, rn * rn, 0.5), rn);
        rn = fma(rn, fma(-hs, rn * rn, 0.5), rn);
        double nrm = ss * rn;
        nrm = fma(fma(-nrm, nrm, ss), 0.5 * rn, nrm);
        beta = alpha >= 0.0 ? -nrm : nrm;
        const double rbeta = alpha >= 0.0 ? -rn : rn;
        const double num = beta - alpha;
        tau = num * rbeta;
        tau = fma(fma(-tau, beta, num), rbeta, tau);
        const double den = alpha - beta;
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
        y = fma(fma(-den, y, 1.0), y, y);
        y = fma(fma(-den, y, 1.0), y, y);
        scale = fma(fma(-den, y, 1.0), y, y);
    }
    const double coef = tau != 0.0 ? tau * fma(scale, tot, pv) : 0.0;      // lane l >= 1: tau * v^T a_(JJ+l)
    double v[RPT];
#pragma unroll
    for (int i = 0; i < RPT; i++) v[i] = rr[i] > JJ ? p[JJ][i] * scale : (rr[i] == JJ ? 1.0 : 0.0);
#pragma unroll
    for (int s = 1; s < NV; s++) {
        const double cf = __shfl_sync(0xffffffffu, coef, s);
#pragma unroll
        for (int i = 0; i < RPT; i++) p[JJ + s][i] = fma(-cf, v[i], p[JJ + s][i]);
    }
#pragma unroll
    for (int i = 0; i < RPT; i++) p[JJ][i] = rr[i] > JJ ? v[i] : (rr[i] == JJ ? beta : p[JJ][i]);
    if (warp == 0 && lane == 0) {
        sh.taus()[JJ] = tau;
        if (JJ < tau_ok_upto) tau_g[JJ] = tau;
    }
    if constexpr (JJ + 1 < TCQ_NB) tcq_panel_column<RPT, JJ + 1>(p, rr, sh, warp, lane, tau_g, tau_ok_upto);
}

// Gs (upper triangle with diagonal) = X^T X of the 16 columns stored in the V buffer ([column][row], zero beyond the last row):
// DMMA tiles over 16 row ranges (one per warp; lane (g, q) takes rows r + 2q, r + 2q + 1 of a chunk of 8: one LDS.128 per fragment
// pair, the k order inside a chunk is free), per-warp partials through shared memory, fixed summation order
__device__ __forceinline__ void tcq_gram(const TcqShared &sh, int mj16, int ldv) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const unsigned vs_sh = (unsigned) __cvta_generic_to_shared(sh.Vs());
    const unsigned v8off = (unsigned) (8 * ldv * 8);   // byte offset of column c + 8
    {
        const int kr = ((mj16 / 8 + TCQ_WARPS - 1) / TCQ_WARPS) * 8;
        const int r0 = warp * kr, r1 = (r0 + kr) < mj16 ? (r0 + kr) : mj16;
        double g00[2] = {0, 0}, g01[2] = {0, 0}, g11[2] = {0, 0};
        unsigned va0 = vs_sh + (unsigned) ((g * ldv + 2 * q + r0) * 8);
        for (int r = r0; r < r1; r += 8, va0 += 64) {
            const double2 f0 = lds_f64x2(va0), f1 = lds_f64x2(va0 + v8off);
            dmma884(g00[0], g00[1], f0.x, f0.x); dmma884(g00[0], g00[1], f0.y, f0.y);
            dmma884(g01[0], g01[1], f0.x, f1.x); dmma884(g01[0], g01[1], f0.y, f1.y);
            dmma884(g11[0], g11[1], f1.x, f1.x); dmma884(g11[0], g11[1], f1.y, f1.y);
        }
        double *gp = sh.part() + (size_t) warp * 256;
        gp[g * 16 + 2 * q] = g00[0]; gp[g * 16 + 2 * q + 1] = g00[1];
        gp[g * 16 + 8 + 2 * q] = g01[0]; gp[g * 16 + 8 + 2 * q + 1] = g01[1];
        gp[(8 + g) * 16 + 8 + 2 * q] = g11[0]; gp[(8 + g) * 16 + 8 + 2 * q + 1] = g11[1];
    }
    __syncthreads();
    if (tid < 256) {
        const int i = tid >> 4, j = tid & 15;
        if (i <= j) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < TCQ_WARPS; w++) acc += sh.part()[(size_t) w * 256 + tid];
            sh.Gs()[i * 17 + j] = acc;
        }
    }
    __syncthreads();
}

// T of the compact WY form from G = V^T V (strict upper triangle) and tau: T(i, j) = -tau_j sum_k T(i, k) G(k, j), one lane per row
__device__ __forceinline__ void tcq_form_T(const TcqShared &sh, int mj16, int ldv) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    tcq_gram(sh, mj16, ldv);
    if (warp == 0 && lane < TCQ_NB) {
        double trow[TCQ_NB];
#pragma unroll
        for (int j = 0; j < TCQ_NB; j++) {
            const double tj = sh.taus()[j];
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < j; k++)
                if (k >= lane) acc = fma(trow[k], sh.Gs()[k * 17 + j], acc);
            trow[j] = j == lane ? tj : (j > lane ? -tj * acc : 0.0);
            sh.Ts()[lane * 17 + j] = trow[j];
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Row-split trailing update: tgt <- (I - V T' V') tgt (TRT; I - V T V' otherwise) with every column read ONCE and written ONCE.
// tgt: matrix whose columns [col0, col_end) are updated on rows j0 .. j0 + mj - 1 (leading dimension ldt).
// The 16 warps split the ROWS (8-row groups, GPW per warp); for a pair of 8-column blocks a lane keeps its piece of the columns in
// registers across both passes: lane (g, q) holds rows 8t + 2q, 8t + 2q + 1 of column g of every group t of its warp -- that is
// the B fragment of pass 1 (W = V' A2: the k order inside a group is free, k-slot q takes row 2q, then row 2q + 1) AND the
// accumulator of pass 2 computed transposed (D'(column, row) += W'' V'), so no layout change and no second trip to memory.
// Per pair of column blocks: partial W of the 16 warps -> shared memory -> summed -> W' = -T' W -> pass 2, three CTA barriers.
// (Measured and rejected: two independent halves of 8 warps on alternating column blocks, named barriers -- 1.05 instead of
// 0.92 ms: the update is bound by the bytes it moves through L2, not by exposed latency.)
// ncu of the column-split predecessor (one warp per 8 columns, the column streamed from L2 in each pass): DRAM 2.67 x the
// algorithmic bytes and an FP64 pipe 10 % busy behind global-memory fragment loads.
// ------------------------------------------------------------------------------------------
template<bool TRT, int GPW>
__device__ __forceinline__ void tcq_trailing_rs(TcqShared sh, double *tgt, size_t ldt, int col0, int col_end, int j0, int mj, int ldv
#ifdef GPUB_TCQ_PROFILE
                                                , long long &tcq_t0
#endif
) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int mj16 = (mj + 15) & ~15;
    const int ncb = (col_end - col0 + 7) / 8;
    const unsigned vs_sh = (unsigned) __cvta_generic_to_shared(sh.Vs());
    const unsigned v8off = (unsigned) (8 * ldv * 8);
    const unsigned s4off = (unsigned) (4 * ldv * 8);
    tcq_form_T(sh, mj16, ldv);
    TCQ_T(3);
    // this warp's 8-row groups [t0, t1)
    const int ng = mj16 >> 3;
    const int gpw = (ng + TCQ_WARPS - 1) / TCQ_WARPS;
    const int t0 = warp * gpw;
    const int t1 = (t0 + gpw) < ng ? (t0 + gpw) : ng;
    double *wsum = sh.wp() + 1024;                          // [2 slots][16 panel columns][8 columns]
    for (int cb0 = 0; cb0 < ncb; cb0 += 2) {
        double2 a[2][GPW];
        double *cp[2];
        unsigned okmask = 0;
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {
            const int cb = cb0 + sl, c0 = col0 + 8 * cb;
            const bool ok = cb < ncb && c0 + g < col_end;   // a lane beyond the last column reads a valid one and never stores
            okmask |= (ok ? 1u : 0u) << sl;
            cp[sl] = tgt + (size_t) j0 + (size_t) (ok ? c0 + g : col0) * ldt + 2 * q;
#pragma unroll
            for (int i = 0; i < GPW; i++) {
                const int r = 8 * (t0 + i) + 2 * q;
                if (t0 + i < t1 && r + 1 < mj) a[sl][i] = *reinterpret_cast<const double2 *>(cp[sl] + 8 * (t0 + i));
                else {
                    a[sl][i].x = (t0 + i < t1 && r < mj) ? cp[sl][8 * (t0 + i)] : 0.0;
                    a[sl][i].y = 0.0;
                }
            }
        }
        // L2 prefetch of this warp's rows of the next pair of column blocks: one lane per 128-byte line (16 rows of a column)
        if (cb0 + 2 < ncb && q == 0) {
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                const int c = col0 + 8 * (cb0 + 2 + sl) + g;
                if (c < col_end) {
                    const double *nx = tgt + (size_t) j0 + (size_t) c * ldt;
#pragma unroll
                    for (int i = 0; i < GPW; i += 2)
                        if (t0 + i < t1 && 8 * (t0 + i) < mj) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 8 * (t0 + i)));
                }
            }
        }
#ifdef GPUB_TCQ_PROFILE
#pragma unroll
        for (int sl = 0; sl < 2; sl++)
#pragma unroll
            for (int i = 0; i < GPW; i++) asm volatile("" ::"d"(a[sl][i].x), "d"(a[sl][i].y));
        TCQ_T(7);
#endif
        // ---- pass 1: partial W (16 panel columns x 8 columns per slot) over this warp's rows ----
        {
            double acc[2][2][2];
#pragma unroll
            for (int sl = 0; sl < 2; sl++) acc[sl][0][0] = acc[sl][0][1] = acc[sl][1][0] = acc[sl][1][1] = 0.0;
            const unsigned va = vs_sh + (unsigned) ((g * ldv + 2 * q + 8 * t0) * 8);
#pragma unroll
            for (int i = 0; i < GPW; i++) {
                if (t0 + i < t1) {
                    const double2 f0 = lds_f64x2(va + 64 * i), f1 = lds_f64x2(va + 64 * i + v8off);
#pragma unroll
                    for (int sl = 0; sl < 2; sl++) {
                        dmma884(acc[sl][0][0], acc[sl][0][1], f0.x, a[sl][i].x);
                        dmma884(acc[sl][1][0], acc[sl][1][1], f1.x, a[sl][i].x);
                    }
#pragma unroll
                    for (int sl = 0; sl < 2; sl++) {
                        dmma884(acc[sl][0][0], acc[sl][0][1], f0.y, a[sl][i].y);
                        dmma884(acc[sl][1][0], acc[sl][1][1], f1.y, a[sl][i].y);
                    }
                }
            }
            // part[warp][slot][panel column k][column c]: lane (g, q) holds W(8 mt + g, 2q), W(8 mt + g, 2q + 1)
#pragma unroll
            for (int sl = 0; sl < 2; sl++)
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
                    *reinterpret_cast<double2 *>(sh.part() + (size_t) ((warp * 2 + sl) * 16 + 8 * mt + g) * 8 + 2 * q) =
                        make_double2(acc[sl][mt][0], acc[sl][mt][1]);
        }
        __syncthreads();
        TCQ_T(4);
        if (tid < 256) {
            const int sl = tid >> 7, e = tid & 127;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int w = 0; w < TCQ_WARPS; w += 2) {
                s0 += sh.part()[(size_t) ((w * 2 + sl) * 128) + e];
                s1 += sh.part()[(size_t) (((w + 1) * 2 + sl) * 128) + e];
            }
            wsum[tid] = s0 + s1;
        }
        __syncthreads();
        if (tid < 256) {
            // W'(i, c) = -sum_k T(k, i) W(k, c) (T' for geqrf / Q'), -sum_k T(i, k) W(k, c) for Q
            const int sl = tid >> 7, i = (tid >> 3) & 15, c = tid & 7;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < TCQ_NB; k++) {
                const double t = TRT ? (k <= i ? sh.Ts()[k * 17 + i] : 0.0) : (k >= i ? sh.Ts()[i * 17 + k] : 0.0);
                acc = fma(t, wsum[sl * 128 + k * 8 + c], acc);
            }
            sh.wp()[(size_t) (sl * 8 + c) * TCQ_LDW + i] = -acc;
        }
        __syncthreads();
        TCQ_T(5);
        // ---- pass 2: D'(column g, rows 2q, 2q + 1) += sum_k W'(k, column) V(row, k): A = W'', B = V' from shared memory ----
        {
            double bw[2][4];
#pragma unroll
            for (int sl = 0; sl < 2; sl++)
#pragma unroll
                for (int s4 = 0; s4 < 4; s4++) bw[sl][s4] = sh.wp()[(size_t) (sl * 8 + g) * TCQ_LDW + 4 * s4 + q];
            const unsigned vb = vs_sh + (unsigned) ((q * ldv + g + 8 * t0) * 8);
#pragma unroll
            for (int i = 0; i < GPW; i++) {
                if (t0 + i < t1) {
#pragma unroll
                    for (int s4 = 0; s4 < 4; s4++) {
                        const double v = lds_f64(vb + 64 * i + s4 * s4off);
#pragma unroll
                        for (int sl = 0; sl < 2; sl++) dmma884(a[sl][i].x, a[sl][i].y, bw[sl][s4], v);
                    }
                }
            }
        }
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {
            if (okmask & (1u << sl)) {
#pragma unroll
                for (int i = 0; i < GPW; i++) {
                    const int r = 8 * (t0 + i) + 2 * q;
                    if (t0 + i < t1) {
                        if (r + 1 < mj) *reinterpret_cast<double2 *>(cp[sl] + 8 * (t0 + i)) = a[sl][i];
                        else if (r < mj) cp[sl][8 * (t0 + i)] = a[sl][i].x;
                    }
                }
            }
        }
        TCQ_T(6);
    }
}

// dlarfg from alpha and the squared norm of the entries below it, without the sqrt / divide slow paths (as tcq_panel_column)
__device__ __forceinline__ void tcq_larfg(double alpha, double xnorm2, double &beta, double &tau, double &scale) {
    tau = 0.0; scale = 0.0; beta = alpha;
    if (xnorm2 != 0.0 && fma(alpha, alpha, xnorm2) >= 1e-290) {   // (below: counts as already reduced, see tcq_panel_column)
        const double ss = fma(alpha, alpha, xnorm2);
        double rn;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rn) : "d"(ss));
        const double hs = 0.5 * ss;
        rn = fma(rn, fma(-hs, rn * rn, 0.5), rn);
        rn = fma(rn, fma(-hs, rn * rn, 0.5), rn);
        double nrm = ss * rn;
        nrm = fma(fma(-nrm, nrm, ss), 0.5 * rn, nrm);
        beta = alpha >= 0.0 ? -nrm : nrm;
        const double rbeta = alpha >= 0.0 ? -rn : rn;
        const double num = beta - alpha;
        tau = num * rbeta;
        tau = fma(fma(-tau, beta, num), rbeta, tau);
        const double den = alpha - beta;
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
        y = fma(fma(-den, y, 1.0), y, y);
        y = fma(fma(-den, y, 1.0), y, y);
        scale = fma(fma(-den, y, 1.0), y, y);
    }
}

// ------------------------------------------------------------------------------------------
// Gram panel (GPUB_TCQ_GRAMPANEL): the 16 Householder steps of a panel without a CTA barrier per column.
// A reflection is orthogonal on the rows it touches, so the Gram matrix S of the ACTIVE rows (>= j) of the panel only loses the
// pivot row from step to step: S(j+1) = S(j) - R(j, :)' R(j, :) -- the Cholesky recurrence of G = P'P. Everything a step needs
// from the 1024 rows is in S: |x|^2 = S(j, j) - alpha^2 and x'a_c = S(j, c) - alpha P(j, c), where alpha and P(j, c) belong to
// the pivot row. So ONE warp runs all 16 steps on G (one DMMA pass over the panel) and the top 16 x 16 block (lane = column),
// and publishes per step the scale 1 / (alpha - beta) and the update coefficients tau v'a_c; every other row then replays the 16
// rank-1 updates thread-locally. Two CTA barriers per panel instead of 16.
// The recurrence is as accurate as the panel is well conditioned (errors grow like eps cond^2): a step whose remaining column
// norm S(j, j) has dropped below 1e-3 of the column's own norm, or whose |x|^2 is below 1e-6 of it (a column that is already
// reduced, where LAPACK would return tau = 0), flags the panel and the whole panel is redone by the column-by-column path.
// ------------------------------------------------------------------------------------------
#ifndef GPUB_TCQ_GRAMPANEL
#define GPUB_TCQ_GRAMPANEL 1
#endif
constexpr int TCQ_OFF_TW = TCQ_OFF_WP;                  // [16][16] tau_j v_j'a_c (the W' scratch is free during the panel)
constexpr int TCQ_OFF_SC = TCQ_OFF_WP + 256;            // [16] scale_j
constexpr int TCQ_OFF_TB = TCQ_OFF_WP + 272;            // [16][16] factored top block (R on / above, V below the diagonal)
constexpr int TCQ_OFF_FLAG = TCQ_OFF_WP + 528;          // panel must be redone column by column

template<int J>
__device__ __forceinline__ void tcq_gram_step(double (&b)[TCQ_NB], double (&s)[TCQ_NB], double gd, int c, int nb, const TcqShared &sh,
                                              double *tau_g, int &bad) {
    // no branch on nb: the 16 steps are one straight line of code, so the updates of step J that the next pivot does not need
    // overlap the latency chain (rsqrt, rcp, Newton steps) of step J + 1. Columns beyond nb are zero: tau = 0, nothing changes.
    const double alpha = __shfl_sync(0xffffffffu, b[J], J);
    const double sjj = __shfl_sync(0xffffffffu, s[J], J);
    const double gjj = __shfl_sync(0xffffffffu, gd, J);
    const double xnorm2 = fma(-alpha, alpha, sjj);
    if (J < nb && (!(sjj > 1e-3 * gjj) || !(xnorm2 > 1e-6 * sjj))) bad = 1;
    double beta, tau, scale;
    tcq_larfg(alpha, xnorm2 > 0.0 ? xnorm2 : 0.0, beta, tau, scale);
    const double pjc = b[J];
    const double dot = fma(-alpha, pjc, s[J]);          // x'a_c over the rows below the pivot
    const double tw = tau * fma(scale, dot, pjc);       // tau v'a_c
    const double rjc = c > J ? pjc - tw : (c == J ? beta : 0.0);
    if (c > J && c < TCQ_NB) sh.base[TCQ_OFF_TW + J * 16 + c] = tw;
    if (c == J) {
        sh.base[TCQ_OFF_SC + J] = scale;
        sh.taus()[J] = tau;
        if (J < nb) tau_g[J] = tau;
    }
#pragma unroll
    for (int r = J + 1; r < TCQ_NB; r++) {
        const double vr = scale * __shfl_sync(0xffffffffu, b[r], J);
        b[r] = c > J ? fma(-tw, vr, b[r]) : (c == J ? vr : b[r]);
    }
    if (c >= J) b[J] = rjc;
#pragma unroll
    for (int a = J + 1; a < TCQ_NB; a++) {
        const double ra = __shfl_sync(0xffffffffu, rjc, a);
        if (c >= a) s[a] = fma(-ra, rjc, s[a]);
    }
    if constexpr (J + 1 < TCQ_NB) tcq_gram_step<J + 1>(b, s, gd, c, nb, sh, tau_g, bad);
}

// returns true when the panel was factored (p holds R / V, taus are published); false: p is untouched, use the column path
template<int RPT>
__device__ __forceinline__ bool tcq_gram_panel(double (&p)[TCQ_NB][RPT], const int (&rr)[RPT], const TcqShared &sh, int mj16, int nb, int ldv,
                                               double *tau_g
#ifdef GPUB_TCQ_PROFILE
                                               , long long &tcq_t0
#endif
) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // the raw panel into the V buffer ([column][row], zero padded), then G = P'P on the tensor pipe
#pragma unroll
    for (int i = 0; i < RPT; i++) {
        const int r = tid + TCQ_THREADS * i;
        if (r < mj16) {
#pragma unroll
            for (int c = 0; c < TCQ_NB; c++) sh.Vs()[(size_t) c * ldv + r] = p[c][i];
        }
    }
    __syncthreads();
    tcq_gram(sh, mj16, ldv);
    TCQ_T(8);
    if (warp == 0) {
        const int c = lane & 15;
        double b[TCQ_NB], s[TCQ_NB];
#pragma unroll
        for (int r = 0; r < TCQ_NB; r++) {
            b[r] = sh.Vs()[(size_t) c * ldv + r];
            s[r] = r <= c ? sh.Gs()[r * 17 + c] : 0.0;
        }
        double gd = 0.0;
#pragma unroll
        for (int r = 0; r < TCQ_NB; r++) gd = r == c ? s[r] : gd;
        int bad = 0;
        tcq_gram_step<0>(b, s, gd, lane < 16 ? c : 99, nb, sh, tau_g, bad);
        if (lane < 16) {
#pragma unroll
            for (int r = 0; r < TCQ_NB; r++) sh.base[TCQ_OFF_TB + r * 16 + c] = b[r];
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) reinterpret_cast<int *>(sh.base + TCQ_OFF_FLAG)[0] = bad;
    }
    __syncthreads();
    TCQ_T(9);
    // the panel comes back from the V buffer: its registers are free while warp 0 runs the recurrence (no spills there)
#pragma unroll
    for (int i = 0; i < RPT; i++) {
        const int r = tid + TCQ_THREADS * i;
#pragma unroll
        for (int c = 0; c < TCQ_NB; c++) p[c][i] = r < mj16 ? sh.Vs()[(size_t) c * ldv + r] : 0.0;
    }
    if (reinterpret_cast<const int *>(sh.base + TCQ_OFF_FLAG)[0]) {
        __syncthreads();                                    // everyone has its rows before the column path reuses the scratch
        return false;
    }
    // rows below the top block replay the 16 rank-1 updates (the coefficient row of a step is read when the step starts: the
    // compiler barrier keeps the 136 coefficients from being hoisted into registers all at once)
#pragma unroll
    for (int j = 0; j < TCQ_NB; j++) {
        asm volatile("" ::: "memory");
        const double sc = sh.base[TCQ_OFF_SC + j];
        double vj[RPT];
#pragma unroll
        for (int i = 0; i < RPT; i++) vj[i] = sc * p[j][i];
#pragma unroll
        for (int c = j + 1; c < TCQ_NB; c++) {
            const double tw = sh.base[TCQ_OFF_TW + j * 16 + c];
#pragma unroll
            for (int i = 0; i < RPT; i++) p[c][i] = fma(-tw, vj[i], p[c][i]);
        }
#pragma unroll
        for (int i = 0; i < RPT; i++) p[j][i] = vj[i];
    }
    // the rows of the top block come from the recurrence itself
#pragma unroll
    for (int i = 0; i < RPT; i++) {
        if (rr[i] >= 0 && rr[i] < TCQ_NB) {
#pragma unroll
            for (int c = 0; c < TCQ_NB; c++) p[c][i] = sh.base[TCQ_OFF_TB + rr[i] * 16 + c];
        }
    }
    TCQ_T(10);
    return true;
}

template<int RPT>
__global__ void __launch_bounds__(TCQ_THREADS, 1) k_geqrf_tc(int m, int n, double *A, size_t lda, size_t sA, double *tau, size_t sTau, size_t batch,
                                                              int ldv) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TcqShared sh;
    sh.base = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int kmax = n < m ? n : m;

    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        double *a_g = A + mat * sA;
        double *tau_g = tau + mat * sTau;
#ifdef GPUB_TCQ_PROFILE
        long long tcq_t0 = clock64();
#endif
        for (int j0 = 0; j0 < kmax; j0 += TCQ_NB) {
            const int nb = (kmax - j0) < TCQ_NB ? (kmax - j0) : TCQ_NB;
            const int mj = m - j0;                       // rows of the panel
            const int mj16 = (mj + 15) & ~15;             // V is zero-padded to a multiple of 16 rows
            // ---- panel into registers: thread owns panel rows tid, tid + 512, ... ----
            double p[TCQ_NB][RPT];
            int rr[RPT];
#pragma unroll
            for (int i = 0; i < RPT; i++) {
                rr[i] = tid + TCQ_THREADS * i;
                const bool ok = rr[i] < mj;
#pragma unroll
                for (int c = 0; c < TCQ_NB; c++) p[c][i] = (ok && c < nb) ? a_g[(size_t) (j0 + rr[i]) + (size_t) (j0 + c) * lda] : 0.0;
                if (!ok) rr[i] = -1;                     // never "below the pivot"
            }
            TCQ_T(0);
#if GPUB_TCQ_GRAMPANEL
            if (!tcq_gram_panel<RPT>(p, rr, sh, mj16, nb, ldv, tau_g + j0
#ifdef GPUB_TCQ_PROFILE
                                     , tcq_t0
#endif
                                     ))
#endif
                tcq_panel_column<RPT, 0>(p, rr, sh, warp, lane, tau_g + j0, nb);
            TCQ_T(1);
            // ---- factored panel back to global memory; explicit V (unit diagonal, zeros above) to shared memory ----
#pragma unroll
            for (int i = 0; i < RPT; i++) {
                const int r = tid + TCQ_THREADS * i;
                if (r < mj16) {
#pragma unroll
                    for (int c = 0; c < TCQ_NB; c++) {
                        if (r < mj && c < nb) a_g[(size_t) (j0 + r) + (size_t) (j0 + c) * lda] = p[c][i];
                        sh.Vs()[(size_t) c * ldv + r] = r < mj ? (r > c ? p[c][i] : (r == c ? 1.0 : 0.0)) : 0.0;
                    }
                }
            }
            const int ntrail = n - (j0 + TCQ_NB);
            if (ntrail <= 0) { __syncthreads(); continue; }
            __syncthreads();
            TCQ_T(2);
            tcq_trailing_rs<true, 4 * RPT>(sh, a_g, lda, j0 + TCQ_NB, n, j0, mj, ldv
#ifdef GPUB_TCQ_PROFILE
                                           , tcq_t0
#endif
            );
            __syncthreads();
            TCQ_T(6);
        }
    }
}

// ------------------------------------------------------------------------------------------
// C <- Q C or Q^T C with Q = H_0 ... H_{k-1} from geqrf, blocked: k_ormqr_tc<TRANS>, fp64, m <= 1024.
// One CTA per (matrix, chunk of 64 columns of C). For every panel of 16 reflectors (last to first for Q, first to last for
// Q^T) the CTA rebuilds the explicit V panel in shared memory from the factored matrix, recomputes T from V^T V and tau
// (cheaper than storing 16 x 16 factors per panel), and applies I - V T V^T (or T^T) with the same two DMMA passes as the
// trailing update of k_geqrf_tc. Replaces the column-by-column k_ormqr_cta for the U assembly of Svd / Nullspace and getQR.
// ------------------------------------------------------------------------------------------
template<bool TRANS>
__global__ void __launch_bounds__(TCQ_THREADS, 1) k_ormqr_tc(int m, int ncols, int k, const double *__restrict__ A, size_t lda, size_t sA,
                                                              const double *__restrict__ tau, size_t sTau, double *C, size_t ldc, size_t sC,
                                                              size_t batch, int chunks, int cols_per_chunk, int ldv) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TcqShared sh;
    sh.base = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x;
    const int npan = (k + TCQ_NB - 1) / TCQ_NB;
    for (size_t t = blockIdx.x; t < batch * (size_t) chunks; t += gridDim.x) {
        const size_t mat = t / chunks;
        const int ch = (int) (t - mat * chunks);
        const double *a_g = A + mat * sA;
        const double *tau_g = tau + mat * sTau;
        double *c_g = C + mat * sC;
        const int col0 = ch * cols_per_chunk;
        const int col_end = (col0 + cols_per_chunk) < ncols ? (col0 + cols_per_chunk) : ncols;
#ifdef GPUB_TCQ_PROFILE
        long long tcq_t0 = clock64();
#endif
        for (int pi = 0; pi < npan; pi++) {
            const int p = TRANS ? pi : npan - 1 - pi;
            const int j0 = p * TCQ_NB;
            const int nb = (k - j0) < TCQ_NB ? (k - j0) : TCQ_NB;
            const int mj = m - j0, mj16 = (mj + 15) & ~15;
            for (int r = tid; r < mj16; r += TCQ_THREADS) {
#pragma unroll
                for (int c = 0; c < TCQ_NB; c++) {
                    double v = 0.0;
                    if (r < mj && c < nb) v = r > c ? a_g[(size_t) (j0 + r) + (size_t) (j0 + c) * lda] : (r == c ? 1.0 : 0.0);
                    sh.Vs()[(size_t) c * ldv + r] = v;
                }
            }
            if (tid < TCQ_NB) sh.taus()[tid] = tid < nb ? tau_g[j0 + tid] : 0.0;
            __syncthreads();
            tcq_trailing_rs<TRANS, 8>(sh, c_g, ldc, col0, col_end, j0, mj, ldv
#ifdef GPUB_TCQ_PROFILE
                                      , tcq_t0
#endif
            );
            __syncthreads();
        }
    }
}

template<typename T>
bool try_ormqr_tc(gpub_ctx_t, cudaStream_t, int, size_t, size_t, size_t, const T *, size_t, size_t, const T *, size_t, T *, size_t, size_t, size_t, int *) {
    return false;
}
template<>
bool try_ormqr_tc<double>(gpub_ctx_t ctx, cudaStream_t stream, int trans, size_t m, size_t ncols, size_t k, const double *A, size_t lda, size_t sA,
                          const double *tau, size_t sTau, double *C, size_t ldc, size_t sC, size_t batch, int *err) {
    if (m > 1024 || m < 64 || ncols < 16 || (ncols & 1) || k < 8) return false;
    if ((ldc & 1) || (sC & 1) || (((uintptr_t) C) & 15u)) return false;
    const size_t ldv = (m + 15) / 16 * 16 + 8;
    const size_t smem = ((size_t) TCQ_OFF_VS + TCQ_NB * ldv) * sizeof(double);
    if (smem > (size_t) ctx->max_smem_optin) return false;
    const int cols_per_chunk = 8 * 4 * TCQ_NSL;
    const size_t chunks = gpub_ceil_div(ncols, (size_t) cols_per_chunk), total = chunks * batch, cap = (size_t) ctx->sm_count * 8;
    const unsigned grid = (unsigned) (total < cap ? total : cap);
    cudaError_t e;
    if (trans) {
        e = cudaFuncSetAttribute(k_ormqr_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e == cudaSuccess)
            k_ormqr_tc<true><<<grid, TCQ_THREADS, smem, stream>>>((int) m, (int) ncols, (int) k, A, lda, sA, tau, sTau, C, ldc, sC, batch, (int) chunks,
                                                                  cols_per_chunk, (int) ldv);
    } else {
        e = cudaFuncSetAttribute(k_ormqr_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e == cudaSuccess)
            k_ormqr_tc<false><<<grid, TCQ_THREADS, smem, stream>>>((int) m, (int) ncols, (int) k, A, lda, sA, tau, sTau, C, ldc, sC, batch, (int) chunks,
                                                                   cols_per_chunk, (int) ldv);
    }
    *err = e == cudaSuccess ? GPUB_OK : (int) e;
    return true;
}

template<typename T>
bool try_geqrf_tc(gpub_ctx_t, cudaStream_t, size_t, size_t, T *, size_t, size_t, T *, size_t, size_t, int *) { return false; }
template<>
bool try_geqrf_tc<double>(gpub_ctx_t ctx, cudaStream_t stream, size_t m, size_t n, double *A, size_t lda, size_t sA, double *tau, size_t sTau,
                          size_t batch, int *err) {
    if (m > 1024 || m < 64 || n < 16 || (n & 1)) return false;   // odd n: a lane's two tile columns would straddle the edge
    if ((lda & 1) || (sA & 1) || (((uintptr_t) A) & 15u)) return false;   // 128-bit accesses to column pairs of rows
    const size_t ldv = (m + 15) / 16 * 16 + 8;   // = 8 (mod 16): the 128-bit fragment loads of pass 1 are bank-conflict free
    const size_t smem = ((size_t) TCQ_OFF_VS + TCQ_NB * ldv) * sizeof(double);
    if (smem > (size_t) ctx->max_smem_optin) return false;
    // one CTA per SM, every CTA the same number of matrices: 256 matrices run as 2 x 128 rather than 148 + 108 -- the makespan in
    // matrices is the same, but the matrices in flight (1 MB each) then fit the 126 MB L2, which the trailing passes re-read
    const size_t waves = gpub_ceil_div(batch, (size_t) ctx->sm_count);
#ifdef GPUB_TCQ_GRID
    const unsigned grid = (unsigned) (batch < (size_t) GPUB_TCQ_GRID ? batch : (size_t) GPUB_TCQ_GRID);
#else
    const unsigned grid = (unsigned) gpub_ceil_div(batch, waves);
#endif
    cudaError_t e;
    if (m <= 512) {
        e = cudaFuncSetAttribute(k_geqrf_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e == cudaSuccess) k_geqrf_tc<1><<<grid, TCQ_THREADS, smem, stream>>>((int) m, (int) n, A, lda, sA, tau, sTau, batch, (int) ldv);
    } else {
        e = cudaFuncSetAttribute(k_geqrf_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e == cudaSuccess) k_geqrf_tc<2><<<grid, TCQ_THREADS, smem, stream>>>((int) m, (int) n, A, lda, sA, tau, sTau, batch, (int) ldv);
    }
    *err = e == cudaSuccess ? GPUB_OK : (int) e;
    return true;
}

// C <- Q^T C (trans) or Q C: every warp owns whole columns of C, so no CTA-wide synchronisation
template<typename T>
__global__ void __launch_bounds__(QT) k_ormqr_cta(int trans, int m, int ncols, int k, const T *__restrict__ A, size_t lda, size_t sA,
                                                   const T *__restrict__ tau, size_t sTau, T *C, size_t ldc, size_t sC,
                                                   size_t batch, int col_blocks) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = QT / 32;
    for (size_t t = blockIdx.x; t < batch * col_blocks; t += gridDim.x) {
        const size_t mat = t / col_blocks;
        const int cb = (int) (t % col_blocks);
        const T *a_g = A + mat * sA;
        const T *tau_g = tau + mat * sTau;
        T *c_g = C + mat * sC;
        for (int c = cb * wpb + warp; c < ncols; c += col_blocks * wpb) {
            T *cc = c_g + (size_t) c * ldc;
            for (int jj = 0; jj < k; jj++) {
                const int j = trans ? jj : k - 1 - jj;
                const T tj = tau_g[j];
                if (tj == T(0)) continue;
                const T *v = a_g + (size_t) j * lda;
                T w = 0;
                for (int r = j + 1 + lane; r < m; r += 32) w = fma(v[r], cc[r], w);
                w = warp_sum(w) + cc[j];
                const T tw = tj * w;
                __syncwarp();
                if (lane == 0) cc[j] -= tw;
                for (int r = j + 1 + lane; r < m; r += 32) cc[r] = fma(-tw, v[r], cc[r]);
                __syncwarp();
            }
        }
    }
}

template<typename T>
__global__ void __launch_bounds__(QT) k_trsv_cta(int n, const T *__restrict__ R, size_t ldr, size_t sR, T *b, size_t sB, size_t batch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *x = reinterpret_cast<T *>(smem_raw);
    __shared__ T s_bc[2];
    for (size_t mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        T *b_g = b + mat * sB;
        for (int r = threadIdx.x; r < n; r += QT) x[r] = b_g[r];
        __syncthreads();
        upper_solve<T>(R + mat * sR, ldr, n, x, nullptr, s_bc);
        for (int r = threadIdx.x; r < n; r += QT) b_g[r] = x[r];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// fused gels for small tall systems: k_gels_sub<T, M, N, LPM>
// LPM lanes per matrix (32/LPM matrices per warp), lane l owns the RPL = M/LPM consecutive rows
// l*RPL .. l*RPL+RPL-1 of [A | b] in registers, so loads/stores are 128-bit and a column of one matrix is one
// contiguous, fully used run of sectors. Column norms and the v^T a_c products are log2(LPM)-step xor
// butterflies inside the group (3 steps for LPM = 8 instead of 5 for a full warp, and one shuffle instruction
// serves all 32/LPM matrices of the warp); the (N-j) products of a column step are independent, so their
// butterflies pipeline. R x = Q^T b is solved in registers. Control flow is uniform across the CTA.
// ------------------------------------------------------------------------------------------
template<typename T> struct VecQ;
template<> struct VecQ<double> { using type = double2; static constexpr int N = 2; };
template<> struct VecQ<float> { using type = float4; static constexpr int N = 4; };
__device__ __forceinline__ void unpack_q(const double2 &v, double *o) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void unpack_q(const float4 &v, float *o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ double2 pack_q(const double *o) { return make_double2(o[0], o[1]); }
__device__ __forceinline__ float4 pack_q(const float *o) { return make_float4(o[0], o[1], o[2], o[3]); }

template<typename T, int LPM>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int o = LPM / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#ifndef GPUB_GELS_MINB
#define GPUB_GELS_MINB 2
#endif

template<typename T, int M, int N, int LPM>
__global__ void __launch_bounds__(128, GPUB_GELS_MINB) k_gels_sub(T *A, size_t sA, T *b, size_t sB, int *info, size_t batch) {
    constexpr int RPL = M / LPM;           // rows per lane
    constexpr int VN = VecQ<T>::N;
    constexpr int NV = RPL / VN;           // 128-bit accesses per lane per column
    using V = typename VecQ<T>::type;
    static_assert(RPL % VN == 0, "rows per lane must be a multiple of the vector width");
    constexpr int MPC = 128 / LPM;         // matrices per CTA iteration
    const int l = threadIdx.x % LPM;       // lane inside the group
    const int gl = (threadIdx.x & 31) - l; // first lane of the group inside the warp
    const int row0 = l * RPL;
    const size_t ngroups = (size_t) gridDim.x * MPC;
    const size_t iters = (batch + ngroups - 1) / ngroups;

    for (size_t it = 0; it < iters; it++) {
        size_t mat = it * ngroups + (size_t) blockIdx.x * MPC + threadIdx.x / LPM;
        const bool live = mat < batch;
        if (!live) mat = batch - 1;
        T *a_g = A + mat * sA + row0;
        T *b_g = b + mat * sB + row0;
        T a[N + 1][RPL];
#pragma unroll
        for (int c = 0; c < N; c++)
#pragma unroll
            for (int v = 0; v < NV; v++) unpack_q(reinterpret_cast<const V *>(a_g + (size_t) c * M)[v], &a[c][v * VN]);
#pragma unroll
        for (int v = 0; v < NV; v++) unpack_q(reinterpret_cast<const V *>(b_g)[v], &a[N][v * VN]);

        int bad = 0;
#pragma unroll
        for (int j = 0; j < N; j++) {
            // row j lives in lane j / RPL of the group, slot j % RPL (both compile-time)
            T part = 0;
#pragma unroll
            for (int t = 0; t < RPL; t++) {
                const T x = (row0 + t > j) ? a[j][t] : T(0);
                part = fma(x, x, part);
            }
            const T xnorm2 = group_sum<T, LPM>(part);
            const T alpha = __shfl_sync(0xffffffffu, a[j][j % RPL], gl + j / RPL);
            T tau, scale;
            const T beta = larfg<T>(alpha, xnorm2, &tau, &scale);
            if (beta == T(0) && bad == 0) bad = j + 1;
            T v[RPL];
#pragma unroll
            for (int t = 0; t < RPL; t++) {
                const int row = row0 + t;
                v[t] = row > j ? a[j][t] * scale : (row == j ? T(1) : T(0));
                a[j][t] = row > j ? v[t] : (row == j ? beta : a[j][t]);
            }
            T w[N + 1];
#pragma unroll
            for (int c = j + 1; c <= N; c++) {
                T acc = 0;
#pragma unroll
                for (int t = 0; t < RPL; t++) acc = fma(v[t], a[c][t], acc);
                w[c] = acc;
            }
#pragma unroll
            for (int c = j + 1; c <= N; c++) w[c] = tau * group_sum<T, LPM>(w[c]);
#pragma unroll
            for (int c = j + 1; c <= N; c++)
#pragma unroll
                for (int t = 0; t < RPL; t++) a[c][t] = fma(-w[c], v[t], a[c][t]);
        }
        // back substitution on R: row j in lane j / RPL, slot j % RPL
#pragma unroll
        for (int j = N - 1; j >= 0; j--) {
            const T xj = __shfl_sync(0xffffffffu, a[N][j % RPL], gl + j / RPL) / __shfl_sync(0xffffffffu, a[j][j % RPL], gl + j / RPL);
#pragma unroll
            for (int t = 0; t < RPL; t++) {
                const int row = row0 + t;
                if (t < RPL && row0 < N) // only the lanes that hold rows of R
                    a[N][t] = row == j ? xj : (row < j ? fma(-a[j][t], xj, a[N][t]) : a[N][t]);
            }
        }
        if (live) {
#pragma unroll
            for (int c = 0; c < N; c++)
#pragma unroll
                for (int v = 0; v < NV; v++) reinterpret_cast<V *>(a_g + (size_t) c * M)[v] = pack_q(&a[c][v * VN]);
#pragma unroll
            for (int v = 0; v < NV; v++) reinterpret_cast<V *>(b_g)[v] = pack_q(&a[N][v * VN]);
            if (info && l == 0) info[mat] = bad;
        }
    }
}

// ------------------------------------------------------------------------------------------
// fp32 variant on packed arithmetic: k_gels_f2<M, N, LPM>. Same mapping and the same operations as k_gels_sub, but the
// RPL = 8 rows of a lane are held as float2 pairs and every dot product / rank-1 update goes through FFMA2
// (fma.rn.f32x2, two IEEE FMAs per issue slot on sm_100): the kernel is issue-bound, so halving the FMA instruction
// count is what moves it. larfg and the back substitution use rsqrt / rcp seeds + one Newton step instead of the
// sqrt / divide slow-path calls (every lane evaluates them redundantly).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float rcp_nr(float d) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    return fmaf(fmaf(-d, y, 1.0f), y, y);
}

#ifndef GPUB_GELS_STAGE
#define GPUB_GELS_STAGE 0
#endif
#ifndef GPUB_GELS_EARLY
#define GPUB_GELS_EARLY 1
#endif
#ifndef GPUB_GELS_L2PF
#define GPUB_GELS_L2PF 0
#endif
__device__ __forceinline__ void gels_cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}

// Row ownership (GPUB_GELS_ROWMAP = 1): lane l of a group owns the 4-row pieces {v * 4 LPM + 4 l .. + 3}, v = 0 .. RPL/4 - 1,
// so the LPM lanes of a group touch one contiguous 16 LPM-byte run per 128-bit access (a full 128-byte line for LPM = 8) instead of
// every other 16 bytes of a 32 LPM-byte run: half the L2 requests for the same data, on loads and on stores.
// Staging (TMA = true, dense batches): the WG = 32 / LPM systems of a warp are contiguous in global memory, so ONE lane moves
// them with two bulk copies (cp.async.bulk, SASS UBLKCP: [A of WG systems], [b of WG systems]) into the warp's shared-memory
// slot, completion counted on the warp's mbarrier. The copy of iteration it + 1 is issued as soon as the lanes have pulled
// iteration it into registers, so the HBM latency of the next systems hides behind the factorisation of the current ones
// (ncu before: 22 % of the warp stalls were the long-scoreboard wait on the 34 LDG.128 at the top of every iteration).
// GPUB_GELS_STAGE = 1 keeps the older per-lane cp.async staging for comparison. Column j is stored as soon as step j has
// finished it, so the 34 STG.128 are not back to back at the end either.
#ifndef GPUB_GELS_ROWMAP
#define GPUB_GELS_ROWMAP 1
#endif
#ifndef GPUB_GELS_TMA
#define GPUB_GELS_TMA 1
#endif
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned) __cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned) __cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "GELS_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GELS_DONE_%=;\n\t"
        "bra GELS_WAIT_%=;\n\t"
        "GELS_DONE_%=:\n\t}" ::"r"((unsigned) __cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned) __cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned) __cvta_generic_to_shared(bar)) : "memory");
}

template<int M, int N, int LPM, bool TMA>
__global__ void __launch_bounds__(128, GPUB_GELS_MINB) k_gels_f2(float *A, size_t sA, float *b, size_t sB, int *info, size_t batch) {
    constexpr int RPL = M / LPM;           // rows per lane
    constexpr int NP = RPL / 2;            // float2 pairs per lane per column
    constexpr int NV = RPL / 4;            // 128-bit accesses per lane per column
    static_assert(RPL % 4 == 0, "rows per lane must be a multiple of 4");
    constexpr int MPC = 128 / LPM;
    constexpr int WG = 32 / LPM;           // systems per warp
    constexpr int SYS4 = (N + 1) * (M / 4); // float4 per system [A | b]
    constexpr int PS = GPUB_GELS_ROWMAP ? 4 * LPM : 4;          // rows between consecutive pieces of a lane
    constexpr int PS4 = GPUB_GELS_ROWMAP ? LPM : 1;             // the same in float4 units
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int l = threadIdx.x % LPM;
    const int g = (threadIdx.x & 31) / LPM;
    const int warp = threadIdx.x >> 5;
    const int gl = (threadIdx.x & 31) - l;
    const int row0 = GPUB_GELS_ROWMAP ? 4 * l : l * RPL;        // first row of piece 0
    const int l4 = GPUB_GELS_ROWMAP ? l : l * NV;               // float4 index of piece 0 inside a column
    // row of pair t: piece t / 2, pair t % 2 inside the piece
    auto row_of = [&](int t) { return row0 + (t >> 1) * PS + 2 * (t & 1); };
    // where the diagonal entry (j, j) lives
    auto diag_lane = [](int j) { return GPUB_GELS_ROWMAP ? (j % (4 * LPM)) / 4 : j / RPL; };
    auto diag_pair = [](int j) { return GPUB_GELS_ROWMAP ? 2 * (j / (4 * LPM)) + (j % 4) / 2 : (j % RPL) / 2; };
    const size_t ngroups = (size_t) gridDim.x * MPC;
    const size_t iters = (batch + ngroups - 1) / ngroups;

    // ---- staging ----
    // TMA: per warp [WG systems of A][WG right-hand sides], then the 4 mbarriers of the CTA
    float4 *wstage = reinterpret_cast<float4 *>(smem_raw) + (size_t) warp * WG * SYS4;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t) 4 * WG * SYS4 * sizeof(float4));
    float4 *stage = reinterpret_cast<float4 *>(smem_raw) + (size_t) (threadIdx.x / LPM) * SYS4;   // cp.async staging: this group's system
    auto tma_issue = [&](size_t it_) {       // one lane of the warp
        const size_t m0 = it_ * ngroups + (size_t) blockIdx.x * MPC + (size_t) warp * WG;
        if (m0 >= batch) return;
        const unsigned nlive = (unsigned) (batch - m0 < (size_t) WG ? batch - m0 : (size_t) WG);
        mbar_expect_tx(&bars[warp], nlive * (unsigned) (SYS4 * sizeof(float4)));
        bulk_g2s(wstage, A + m0 * (size_t) (M * N), nlive * (unsigned) (M * N * sizeof(float)), &bars[warp]);
        bulk_g2s(wstage + WG * (M * N / 4), b + m0 * (size_t) M, nlive * (unsigned) (M * sizeof(float)), &bars[warp]);
    };
    auto prefetch = [&](size_t it_) {
        size_t mat = it_ * ngroups + (size_t) blockIdx.x * MPC + threadIdx.x / LPM;
        if (mat >= batch) mat = batch - 1;
        const float *a_n = A + mat * sA + row0, *b_n = b + mat * sB + row0;
#pragma unroll
        for (int c = 0; c <= N; c++)
#pragma unroll
            for (int v = 0; v < NV; v++) gels_cp_async16(&stage[c * (M / 4) + v * LPM + l], (c < N ? a_n + (size_t) c * M : b_n) + v * PS);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if constexpr (TMA) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int w = 0; w < 4; w++) mbar_init(&bars[w], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0 && iters > 0) tma_issue(0);
    } else {
#if GPUB_GELS_STAGE
        if (iters > 0) prefetch(0);
#else
        (void) stage; (void) prefetch;
#endif
    }
    for (size_t it = 0; it < iters; it++) {
        size_t mat = it * ngroups + (size_t) blockIdx.x * MPC + threadIdx.x / LPM;
        const bool live = mat < batch;
        if (!live) mat = batch - 1;
        float *a_g = A + mat * sA + row0;
        float *b_g = b + mat * sB + row0;
        float2 a[N + 1][NP];
        if constexpr (TMA) {
            const size_t m0 = it * ngroups + (size_t) blockIdx.x * MPC + (size_t) warp * WG;
            if (m0 < batch) mbar_wait(&bars[warp], (unsigned) (it & 1));
#pragma unroll
            for (int c = 0; c <= N; c++) {
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    const float4 x = c < N ? wstage[g * (M * N / 4) + c * (M / 4) + l4 + v * PS4]
                                           : wstage[WG * (M * N / 4) + g * (M / 4) + l4 + v * PS4];
                    a[c][2 * v] = make_float2(x.x, x.y);
                    a[c][2 * v + 1] = make_float2(x.z, x.w);
                }
            }
            // every lane has pulled its pieces: order the generic-proxy reads before the async-proxy refill of the slot
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if ((threadIdx.x & 31) == 0 && it + 1 < iters) tma_issue(it + 1);
        } else {
#if GPUB_GELS_STAGE
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
#pragma unroll
            for (int c = 0; c <= N; c++) {
#pragma unroll
                for (int v = 0; v < NV; v++) {
#if GPUB_GELS_STAGE
                    const float4 x = stage[c * (M / 4) + v * LPM + l];
#else
                    const float4 x = *reinterpret_cast<const float4 *>((c < N ? a_g + (size_t) c * M : b_g) + v * PS);
#endif
                    a[c][2 * v] = make_float2(x.x, x.y);
                    a[c][2 * v + 1] = make_float2(x.z, x.w);
                }
            }
#if GPUB_GELS_STAGE
            if (it + 1 < iters) prefetch(it + 1);   // the lane has consumed its own pieces: the slot can be refilled
#endif
        }
        int bad = 0;
#pragma unroll
        for (int j = 0; j < N; j++) {
            // |x|^2 over the rows below j
            float2 part = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < NP; t++) {
                const float2 x = make_float2(row_of(t) > j ? a[j][t].x : 0.f, row_of(t) + 1 > j ? a[j][t].y : 0.f);
                part = __ffma2_rn(x, x, part);
            }
            const float xnorm2 = group_sum<float, LPM>(part.x + part.y);
            const float ajj = (j & 1) ? a[j][diag_pair(j)].y : a[j][diag_pair(j)].x;
            const float alpha = __shfl_sync(0xffffffffu, ajj, gl + diag_lane(j));
            float tau = 0.f, scale = 0.f, beta = alpha;
            if (xnorm2 != 0.f && fmaf(alpha, alpha, xnorm2) >= 1e-30f) {   // (below: counts as already reduced, see larfg)
                const float ss = fmaf(alpha, alpha, xnorm2);
                float rn = rsqrtf(ss);
                rn = fmaf(0.5f * rn, fmaf(-ss * rn, rn, 1.0f), rn);
                float nrm = ss * rn;
                nrm = fmaf(fmaf(-nrm, nrm, ss), 0.5f * rn, nrm);
                beta = alpha >= 0.f ? -nrm : nrm;
                const float rbeta = alpha >= 0.f ? -rn : rn;
                const float num = beta - alpha;
                tau = num * rbeta;
                tau = fmaf(fmaf(-tau, beta, num), rbeta, tau);
                scale = rcp_nr(alpha - beta);
            }
            if (beta == 0.f && bad == 0) bad = j + 1;
            float2 v[NP];
#pragma unroll
            for (int t = 0; t < NP; t++) {
                const int r0 = row_of(t), r1 = r0 + 1;
                v[t] = make_float2(r0 > j ? a[j][t].x * scale : (r0 == j ? 1.f : 0.f), r1 > j ? a[j][t].y * scale : (r1 == j ? 1.f : 0.f));
                a[j][t] = make_float2(r0 > j ? v[t].x : (r0 == j ? beta : a[j][t].x), r1 > j ? v[t].y : (r1 == j ? beta : a[j][t].y));
            }
            if (GPUB_GELS_EARLY && live) {                     // column j is final
#pragma unroll
                for (int vv = 0; vv < NV; vv++)
                    *reinterpret_cast<float4 *>(a_g + (size_t) j * M + vv * PS) = make_float4(a[j][2 * vv].x, a[j][2 * vv].y, a[j][2 * vv + 1].x, a[j][2 * vv + 1].y);
            }
            float w[N + 1];
#pragma unroll
            for (int c = j + 1; c <= N; c++) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int t = 0; t < NP; t++) acc = __ffma2_rn(v[t], a[c][t], acc);
                w[c] = acc.x + acc.y;
            }
#pragma unroll
            for (int c = j + 1; c <= N; c++) w[c] = tau * group_sum<float, LPM>(w[c]);
#pragma unroll
            for (int c = j + 1; c <= N; c++) {
                const float2 nw = make_float2(-w[c], -w[c]);
#pragma unroll
                for (int t = 0; t < NP; t++) a[c][t] = __ffma2_rn(nw, v[t], a[c][t]);
            }
        }
        // back substitution on R: row j lives in lane diag_lane(j), pair diag_pair(j); only pairs holding rows < N take part
        constexpr int NPB = (GPUB_GELS_ROWMAP && N <= 4 * LPM) ? (NP < 2 ? NP : 2) : NP;
#pragma unroll
        for (int j = N - 1; j >= 0; j--) {
            const float bj = (j & 1) ? a[N][diag_pair(j)].y : a[N][diag_pair(j)].x;
            const float rjj = (j & 1) ? a[j][diag_pair(j)].y : a[j][diag_pair(j)].x;
            const float num = __shfl_sync(0xffffffffu, bj, gl + diag_lane(j)), den = __shfl_sync(0xffffffffu, rjj, gl + diag_lane(j));
            const float rden = rcp_nr(den);
            float xj = num * rden;
            xj = fmaf(fmaf(-xj, den, num), rden, xj);
            if (row0 < N) {            // only the lanes that hold rows of R
#pragma unroll
                for (int t = 0; t < NPB; t++) {
                    const int r0 = row_of(t), r1 = r0 + 1;
                    a[N][t].x = r0 == j ? xj : (r0 < j ? fmaf(-a[j][t].x, xj, a[N][t].x) : a[N][t].x);
                    a[N][t].y = r1 == j ? xj : (r1 < j ? fmaf(-a[j][t].y, xj, a[N][t].y) : a[N][t].y);
                }
            }
        }
        if (live) {
            if (!GPUB_GELS_EARLY) {
#pragma unroll
                for (int c = 0; c < N; c++) {
#pragma unroll
                    for (int vv = 0; vv < NV; vv++)
                        *reinterpret_cast<float4 *>(a_g + (size_t) c * M + vv * PS) = make_float4(a[c][2 * vv].x, a[c][2 * vv].y, a[c][2 * vv + 1].x, a[c][2 * vv + 1].y);
                }
            }
#pragma unroll
            for (int vv = 0; vv < NV; vv++)
                *reinterpret_cast<float4 *>(b_g + vv * PS) = make_float4(a[N][2 * vv].x, a[N][2 * vv].y, a[N][2 * vv + 1].x, a[N][2 * vv + 1].y);
            if (info && l == 0) info[mat] = bad;
        }
    }
}

template<typename T, int M, int N, int LPM> struct GelsKernel {
    static void launch(unsigned grid, cudaStream_t stream, T *A, size_t sA, T *b, size_t sB, int *info, size_t batch) {
        k_gels_sub<T, M, N, LPM><<<grid, 128, 0, stream>>>(A, sA, b, sB, info, batch);
    }
};
template<int M, int N, int LPM> struct GelsKernel<float, M, N, LPM> {
    static void launch(unsigned grid, cudaStream_t stream, float *A, size_t sA, float *b, size_t sB, int *info, size_t batch) {
        if (GPUB_GELS_TMA && sA == (size_t) M * N && sB == (size_t) M) {
            // dense batch: bulk-copy staging, one resident wave that grid-strides (grid is a multiple of the SM count)
            const size_t smem = (size_t) (128 / LPM) * (N + 1) * M * sizeof(float) + 64;
            cudaFuncSetAttribute(k_gels_f2<M, N, LPM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            k_gels_f2<M, N, LPM, true><<<grid, 128, smem, stream>>>(A, sA, b, sB, info, batch);
            return;
        }
        const size_t smem = GPUB_GELS_STAGE ? (size_t) (128 / LPM) * (N + 1) * M * sizeof(float) : 0;   // one staged system per lane group
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_gels_f2<M, N, LPM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        k_gels_f2<M, N, LPM, false><<<grid, 128, smem, stream>>>(A, sA, b, sB, info, batch);
    }
};

template<typename T>
size_t qr_smem_bytes(size_t m, size_t n, bool gels) {
    return ((m | 1) * (n + (gels ? 1 : 0))) * sizeof(T);
}

template<typename T>
int geqrf_batched(gpub_ctx_t ctx, int sidx, size_t m, size_t n, T *A, size_t lda, size_t sA, T *tau, size_t sTau, size_t batch) {
    if (m == 0 || n == 0 || batch == 0) return GPUB_OK;
    if (!A || !tau || lda < m) return GPUB_EINVAL;
    if (m > (size_t) INT32_MAX || n > (size_t) INT32_MAX) return GPUB_ENOTSUP;
    GPUB_ENTER(ctx, sidx);
    const size_t bytes = qr_smem_bytes<T>(m, n, false);
    const int use_smem = bytes <= (size_t) ctx->max_smem_optin - 2048 ? 1 : 0;
    if (m >= 256 && n >= 32) {
        int etc = GPUB_OK;
        if (try_geqrf_tc<T>(ctx, stream, m, n, A, lda, sA, tau, sTau, batch, &etc)) {
            if (etc != GPUB_OK) return etc;
            GPUB_LAUNCH_CHECK();
            return GPUB_OK;
        }
    }
    if (!use_smem || (m >= 256 && n >= 32)) {
        // blocked compact-WY: the widest panel whose [NBQ][m] block fits in shared memory
        const size_t ldv = (m + 3) / 4 * 4 + 4;
        const size_t avail = (size_t) ctx->max_smem_optin - 4096;
        auto need = [&](size_t nbq) { return (nbq * ldv + nbq * nbq + (QT / 32) * nbq + 8 + nbq) * sizeof(T); };
        const size_t cap = (size_t) ctx->sm_count * 2;
        const unsigned grid = (unsigned) (batch < cap ? batch : cap);
#define GPUB_WY_LAUNCH(NBQV)                                                                                        \
    {                                                                                                                \
        const size_t smem = need(NBQV);                                                                              \
        if (smem > 48 * 1024)                                                                                        \
            GPUB_CUDA(cudaFuncSetAttribute(k_geqrf_wy<T, NBQV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
        k_geqrf_wy<T, NBQV><<<grid, QT, smem, stream>>>((int) m, (int) n, A, lda, sA, tau, sTau, batch, (int) ldv);  \
        GPUB_LAUNCH_CHECK();                                                                                         \
        return GPUB_OK;                                                                                              \
    }
        if (need(16) <= avail && n >= 16) GPUB_WY_LAUNCH(16)
        if (need(8) <= avail) GPUB_WY_LAUNCH(8)
        if (need(4) <= avail) GPUB_WY_LAUNCH(4)
#undef GPUB_WY_LAUNCH
    }
    const size_t smem = use_smem ? bytes : 0;
    if (smem > 48 * 1024) GPUB_CUDA(cudaFuncSetAttribute(k_qr_cta<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    const size_t cap = (size_t) ctx->sm_count * 4;
    const unsigned grid = (unsigned) (batch < cap ? batch : cap);
    k_qr_cta<T><<<grid, QT, smem, stream>>>((int) m, (int) n, A, lda, sA, tau, sTau, nullptr, 0, nullptr, batch, use_smem);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

template<typename T>
int ormqr_batched(gpub_ctx_t ctx, int sidx, int trans, size_t m, size_t ncols, size_t k, const T *A, size_t lda, size_t sA,
                  const T *tau, size_t sTau, T *C, size_t ldc, size_t sC, size_t batch) {
    if (m == 0 || ncols == 0 || k == 0 || batch == 0) return GPUB_OK;
    if (!A || !tau || !C || lda < m || ldc < m || k > m) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    {
        int etc = GPUB_OK;
        if (try_ormqr_tc<T>(ctx, stream, trans, m, ncols, k, A, lda, sA, tau, sTau, C, ldc, sC, batch, &etc)) {
            if (etc != GPUB_OK) return etc;
            GPUB_LAUNCH_CHECK();
            return GPUB_OK;
        }
    }
    // split the columns of C over several CTAs when the batch alone cannot fill the GPU
    size_t col_blocks = 1;
    const size_t wpb = QT / 32;
    const size_t target = (size_t) ctx->sm_count * 4;
    if (batch < target) {
        col_blocks = gpub_ceil_div(target, batch);
        const size_t maxb = gpub_ceil_div(ncols, wpb);
        if (col_blocks > maxb) col_blocks = maxb;
    }
    const size_t total = batch * col_blocks;
    const size_t cap = (size_t) ctx->sm_count * 8;
    const unsigned grid = (unsigned) (total < cap ? total : cap);
    k_ormqr_cta<T><<<grid, QT, 0, stream>>>(trans, (int) m, (int) ncols, (int) k, A, lda, sA, tau, sTau, C, ldc, sC, batch,
                                            (int) col_blocks);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

template<typename T>
int trsv_batched(gpub_ctx_t ctx, int sidx, size_t n, const T *R, size_t ldr, size_t sR, T *b, size_t sB, size_t batch) {
    if (n == 0 || batch == 0) return GPUB_OK;
    if (!R || !b || ldr < n) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    const size_t smem = n * sizeof(T);
    if (smem > (size_t) ctx->max_smem_optin - 1024) return GPUB_ENOTSUP;
    if (smem > 48 * 1024) GPUB_CUDA(cudaFuncSetAttribute(k_trsv_cta<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    const size_t cap = (size_t) ctx->sm_count * 8;
    const unsigned grid = (unsigned) (batch < cap ? batch : cap);
    k_trsv_cta<T><<<grid, QT, smem, stream>>>((int) n, R, ldr, sR, b, sB, batch);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

template<typename T, int M, int N, int LPM>
int launch_gels_sub(gpub_ctx_t ctx, cudaStream_t stream, T *A, size_t sA, T *b, size_t sB, int *info, size_t batch) {
    const size_t want = gpub_ceil_div(batch, (size_t) (128 / LPM));
    const size_t cap = (size_t) ctx->sm_count * GPUB_GELS_MINB * GPUB_GRID_WAVES;
    const unsigned grid = (unsigned) (want < cap ? want : cap);
    GelsKernel<T, M, N, LPM>::launch(grid, stream, A, sA, b, sB, info, batch);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

// lanes per matrix: 8 consecutive fp32 rows (or 4 fp64 rows) per lane keeps [A | b] within ~140 registers
template<typename T> struct GelsLpm;
#ifndef GPUB_GELS_RPL_F32
#define GPUB_GELS_RPL_F32 8
#endif
template<> struct GelsLpm<float> { static constexpr int rows_per_lane = GPUB_GELS_RPL_F32; };
template<> struct GelsLpm<double> { static constexpr int rows_per_lane = 4; };

template<typename T>
int gels_batched(gpub_ctx_t ctx, int sidx, size_t m, size_t n, T *A, size_t lda, size_t sA, T *b, size_t sB, int *info,
                 size_t batch) {
    if (m == 0 || batch == 0) return GPUB_OK;
    if (!A || !b || lda < m || n > m) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    if (n == 0) return GPUB_OK;
    const bool vec_ok = lda == m && ((((uintptr_t) A) | ((uintptr_t) b)) & 15u) == 0 && (sA * sizeof(T)) % 16 == 0 &&
                        (sB * sizeof(T)) % 16 == 0;
    if (vec_ok) {
        constexpr int R = GelsLpm<T>::rows_per_lane;
        if (m == 64 && n == 16) return launch_gels_sub<T, 64, 16, 64 / R>(ctx, stream, A, sA, b, sB, info, batch);
        if (m == 32 && n == 16) return launch_gels_sub<T, 32, 16, 32 / R>(ctx, stream, A, sA, b, sB, info, batch);
        if (m == 32 && n == 8) return launch_gels_sub<T, 32, 8, 32 / R>(ctx, stream, A, sA, b, sB, info, batch);
        if (m == 16 && n == 8) return launch_gels_sub<T, 16, 8, 16 / R>(ctx, stream, A, sA, b, sB, info, batch);
    }
    const size_t bytes = qr_smem_bytes<T>(m, n, true);
    if (bytes > (size_t) ctx->max_smem_optin - 2048) {
        // too large for one CTA's shared memory: geqrf in place, then Q^T b, then the triangular solve,
        // with tau kept in a stream-ordered scratch buffer
        T *tau = nullptr;
        GPUB_CUDA(cudaMallocAsync((void **) &tau, n * batch * sizeof(T), stream));
        const size_t cap = (size_t) ctx->sm_count * 4;
        const unsigned grid = (unsigned) (batch < cap ? batch : cap);
        k_qr_cta<T><<<grid, QT, 0, stream>>>((int) m, (int) n, A, lda, sA, tau, n, nullptr, 0, nullptr, batch, 0);
        GPUB_LAUNCH_CHECK();
        int e = ormqr_batched<T>(ctx, sidx, 1, m, 1, n, A, lda, sA, tau, n, b, m, sB, batch);
        if (e) return e;
        e = trsv_batched<T>(ctx, sidx, n, A, lda, sA, b, sB, batch);
        if (e) return e;
        if (info) GPUB_CUDA(cudaMemsetAsync(info, 0, batch * sizeof(int), stream));
        GPUB_CUDA(cudaFreeAsync(tau, stream));
        return GPUB_OK;
    }
    if (bytes > 48 * 1024) GPUB_CUDA(cudaFuncSetAttribute(k_qr_cta<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    const size_t cap = (size_t) ctx->sm_count * 4;
    const unsigned grid = (unsigned) (batch < cap ? batch : cap);
    k_qr_cta<T><<<grid, QT, bytes, stream>>>((int) m, (int) n, A, lda, sA, nullptr, 0, b, sB, info, batch, 1);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

} // namespace

extern "C" {

#ifdef GPUB_TCQ_PROFILE
int gpub_debug_tcq_profile(unsigned long long *out8, int reset) {   /* out8: 12 counters */
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, g_tcq_prof, sizeof(unsigned long long) * 12);
    if (reset) { unsigned long long z[12] = {0}; cudaMemcpyToSymbol(g_tcq_prof, z, sizeof(z)); }
    return 0;
}
#endif

int gpub_geqrf_batched_f64(gpub_ctx_t c, int s, size_t m, size_t n, double *A, size_t lda, size_t sA, double *tau, size_t sT, size_t b) { return geqrf_batched<double>(c, s, m, n, A, lda, sA, tau, sT, b); }
int gpub_geqrf_batched_f32(gpub_ctx_t c, int s, size_t m, size_t n, float *A, size_t lda, size_t sA, float *tau, size_t sT, size_t b) { return geqrf_batched<float>(c, s, m, n, A, lda, sA, tau, sT, b); }

int gpub_ormqr_batched_f64(gpub_ctx_t c, int s, int trans, size_t m, size_t nc, size_t k, const double *A, size_t lda, size_t sA, const double *tau, size_t sT, double *C, size_t ldc, size_t sC, size_t b) { return ormqr_batched<double>(c, s, trans, m, nc, k, A, lda, sA, tau, sT, C, ldc, sC, b); }
int gpub_ormqr_batched_f32(gpub_ctx_t c, int s, int trans, size_t m, size_t nc, size_t k, const float *A, size_t lda, size_t sA, const float *tau, size_t sT, float *C, size_t ldc, size_t sC, size_t b) { return ormqr_batched<float>(c, s, trans, m, nc, k, A, lda, sA, tau, sT, C, ldc, sC, b); }

int gpub_trsv_upper_batched_f64(gpub_ctx_t c, int s, size_t n, const double *R, size_t ldr, size_t sR, double *b, size_t sB, size_t bt) { return trsv_batched<double>(c, s, n, R, ldr, sR, b, sB, bt); }
int gpub_trsv_upper_batched_f32(gpub_ctx_t c, int s, size_t n, const float *R, size_t ldr, size_t sR, float *b, size_t sB, size_t bt) { return trsv_batched<float>(c, s, n, R, ldr, sR, b, sB, bt); }

int gpub_gels_batched_f64(gpub_ctx_t c, int s, size_t m, size_t n, double *A, size_t lda, size_t sA, double *b, size_t sB, int *info, size_t bt) { return gels_batched<double>(c, s, m, n, A, lda, sA, b, sB, info, bt); }
int gpub_gels_batched_f32(gpub_ctx_t c, int s, size_t m, size_t n, float *A, size_t lda, size_t sA, float *b, size_t sB, int *info, size_t bt) { return gels_batched<float>(c, s, m, n, A, lda, sA, b, sB, info, bt); }

} // extern "C"
