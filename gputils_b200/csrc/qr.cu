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
    if (xnorm2 == T(0)) {
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
    const size_t cap = (size_t) ctx->sm_count * GPUB_GELS_MINB * 2;
    const unsigned grid = (unsigned) (want < cap ? want : cap);
    k_gels_sub<T, M, N, LPM><<<grid, 128, 0, stream>>>(A, sA, b, sB, info, batch);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

// lanes per matrix: 8 consecutive fp32 rows (or 4 fp64 rows) per lane keeps [A | b] within ~140 registers
template<typename T> struct GelsLpm;
template<> struct GelsLpm<float> { static constexpr int rows_per_lane = 8; };
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

int gpub_geqrf_batched_f64(gpub_ctx_t c, int s, size_t m, size_t n, double *A, size_t lda, size_t sA, double *tau, size_t sT, size_t b) { return geqrf_batched<double>(c, s, m, n, A, lda, sA, tau, sT, b); }
int gpub_geqrf_batched_f32(gpub_ctx_t c, int s, size_t m, size_t n, float *A, size_t lda, size_t sA, float *tau, size_t sT, size_t b) { return geqrf_batched<float>(c, s, m, n, A, lda, sA, tau, sT, b); }

int gpub_ormqr_batched_f64(gpub_ctx_t c, int s, int trans, size_t m, size_t nc, size_t k, const double *A, size_t lda, size_t sA, const double *tau, size_t sT, double *C, size_t ldc, size_t sC, size_t b) { return ormqr_batched<double>(c, s, trans, m, nc, k, A, lda, sA, tau, sT, C, ldc, sC, b); }
int gpub_ormqr_batched_f32(gpub_ctx_t c, int s, int trans, size_t m, size_t nc, size_t k, const float *A, size_t lda, size_t sA, const float *tau, size_t sT, float *C, size_t ldc, size_t sC, size_t b) { return ormqr_batched<float>(c, s, trans, m, nc, k, A, lda, sA, tau, sT, C, ldc, sC, b); }

int gpub_trsv_upper_batched_f64(gpub_ctx_t c, int s, size_t n, const double *R, size_t ldr, size_t sR, double *b, size_t sB, size_t bt) { return trsv_batched<double>(c, s, n, R, ldr, sR, b, sB, bt); }
int gpub_trsv_upper_batched_f32(gpub_ctx_t c, int s, size_t n, const float *R, size_t ldr, size_t sR, float *b, size_t sB, size_t bt) { return trsv_batched<float>(c, s, n, R, ldr, sR, b, sB, bt); }

int gpub_gels_batched_f64(gpub_ctx_t c, int s, size_t m, size_t n, double *A, size_t lda, size_t sA, double *b, size_t sB, int *info, size_t bt) { return gels_batched<double>(c, s, m, n, A, lda, sA, b, sB, info, bt); }
int gpub_gels_batched_f32(gpub_ctx_t c, int s, size_t m, size_t n, float *A, size_t lda, size_t sA, float *b, size_t sB, int *info, size_t bt) { return gels_batched<float>(c, s, m, n, A, lda, sA, b, sB, info, bt); }

} // extern "C"
