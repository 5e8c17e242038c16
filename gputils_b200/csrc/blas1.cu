// Flat (whole-tensor) kernels: reductions, scal/axpy/rot, transpose, row gather, pointer table,
// rank counting, nullspace packing and the synthetic-data generators.
// All are HBM-bound streaming kernels: 128-bit loads, grid sized as a multiple of the SM count,
// deterministic two-level reductions (per-CTA partials, last CTA folds them in a fixed order and
// writes the scalar straight into mapped pinned host memory -- no D2H memcpy).
#include "common.cuh"

namespace {

constexpr int kThreads = 512;

template<typename T> struct VecOf;
template<> struct VecOf<double> { using type = double2; static constexpr int N = 2; };
template<> struct VecOf<float> { using type = float4; static constexpr int N = 4; };

template<typename T> __device__ __forceinline__ void unpack(const double2 &v, T *o) { o[0] = v.x; o[1] = v.y; }
template<typename T> __device__ __forceinline__ void unpack(const float4 &v, T *o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ double2 pack(const double *o) { return make_double2(o[0], o[1]); }
__device__ __forceinline__ float4 pack(const float *o) { return make_float4(o[0], o[1], o[2], o[3]); }

static inline bool aligned16(const void *p) { return (((uintptr_t) p) & 15u) == 0; }

inline int stream_grid(gpub_ctx_t ctx, size_t n, int per_thread) {
    size_t want = gpub_ceil_div(n, (size_t) kThreads * per_thread);
    size_t cap = (size_t) ctx->sm_count * 4;
    if (want < 1) want = 1;
    return (int) (want < cap ? want : cap);
}

// ------------------------------------------------------------------------------------------
// sum-type reductions (dot, sum of squares, sum of |x|), accumulated in double
// ------------------------------------------------------------------------------------------
enum { OP_DOT = 0, OP_SUMSQ = 1, OP_ASUM = 2 };

template<int OP> __device__ __forceinline__ double term(double a, double b) {
    if (OP == OP_DOT) return a * b;
    if (OP == OP_SUMSQ) return a * a;
    return fabs(a);
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double s_part[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane_id() == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0;
    if (threadIdx.x < 32) {
        r = (threadIdx.x < (blockDim.x >> 5)) ? s_part[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r; // valid in thread 0
}

template<typename T, int OP, bool VEC>
__global__ void __launch_bounds__(kThreads) k_reduce_sum(size_t n, const T *__restrict__ x, const T *__restrict__ y,
                                                         double *partials, unsigned int *ticket, T *out, bool take_sqrt, double prescale) {
    using V = typename VecOf<T>::type;
    constexpr int VN = VecOf<T>::N;
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t) gridDim.x * blockDim.x;
    double acc = 0;
    size_t done = 0;
    if (VEC) {
        const size_t nv = n / VN;
        const V *xv = reinterpret_cast<const V *>(x);
        const V *yv = reinterpret_cast<const V *>(y);
#pragma unroll 4
        for (size_t i = tid; i < nv; i += nth) {
            T a[VN], b[VN];
            unpack<T>(xv[i], a);
            if (OP == OP_DOT) unpack<T>(yv[i], b);
#pragma unroll
            for (int j = 0; j < VN; j++) acc += term<OP>((double) a[j] * prescale, OP == OP_DOT ? (double) b[j] : 0.0);
        }
        done = nv * VN;
    }
    for (size_t i = done + tid; i < n; i += nth)
        acc += term<OP>((double) x[i] * prescale, OP == OP_DOT ? (double) y[i] : 0.0);

    double bs = block_sum(acc);
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = bs;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double v = 0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += ((volatile double *) partials)[i];
        double tot = block_sum(v);
        if (threadIdx.x == 0) {
            *out = (T) (take_sqrt ? sqrt(tot) : tot);
            *ticket = 0;
            __threadfence_system();
        }
    }
}

template<typename T, int OP>
int reduce_sum(gpub_ctx_t ctx, int sidx, size_t n, const T *x, const T *y, T *result_host, bool take_sqrt, double prescale = 1.0) {
    if (!result_host) return GPUB_EINVAL;
    if (n == 0) {
        *result_host = T(0);
        return GPUB_OK;
    }
    if (!x || (OP == OP_DOT && !y)) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    int grid = stream_grid(ctx, n, 16);
    bool vec = aligned16(x) && (OP != OP_DOT || aligned16(y));
    T *out = reinterpret_cast<T *>(slot->h_result);
    double *partials = reinterpret_cast<double *>(slot->d_scratch);
    if (vec)
        k_reduce_sum<T, OP, true><<<grid, kThreads, 0, stream>>>(n, x, y, partials, slot->d_counter, out, take_sqrt, prescale);
    else
        k_reduce_sum<T, OP, false><<<grid, kThreads, 0, stream>>>(n, x, y, partials, slot->d_counter, out, take_sqrt, prescale);
    GPUB_LAUNCH_CHECK();
    GPUB_CUDA(cudaStreamSynchronize(stream));
    *result_host = *out;
    return GPUB_OK;
}

// ------------------------------------------------------------------------------------------
// max / min of |x| with the first index attaining it
// ------------------------------------------------------------------------------------------
struct AbsIdx {
    double v;
    long long i;
};

template<bool MAX> __device__ __forceinline__ AbsIdx better(AbsIdx a, AbsIdx b) {
    if (a.i < 0) return b;
    if (b.i < 0) return a;
    bool take_b = MAX ? (b.v > a.v || (b.v == a.v && b.i < a.i)) : (b.v < a.v || (b.v == a.v && b.i < a.i));
    return take_b ? b : a;
}

template<bool MAX> __device__ __forceinline__ AbsIdx block_best(AbsIdx a) {
    __shared__ double s_v[32];
    __shared__ long long s_i[32];
    for (int o = 16; o > 0; o >>= 1) {
        AbsIdx b;
        b.v = __shfl_down_sync(0xffffffffu, a.v, o);
        b.i = __shfl_down_sync(0xffffffffu, a.i, o);
        a = better<MAX>(a, b);
    }
    if (lane_id() == 0) {
        s_v[threadIdx.x >> 5] = a.v;
        s_i[threadIdx.x >> 5] = a.i;
    }
    __syncthreads();
    AbsIdx r{0.0, -1};
    if (threadIdx.x < 32) {
        if (threadIdx.x < (blockDim.x >> 5)) r = AbsIdx{s_v[threadIdx.x], s_i[threadIdx.x]};
        for (int o = 16; o > 0; o >>= 1) {
            AbsIdx b;
            b.v = __shfl_down_sync(0xffffffffu, r.v, o);
            b.i = __shfl_down_sync(0xffffffffu, r.i, o);
            r = better<MAX>(r, b);
        }
    }
    __syncthreads();
    return r;
}

template<typename T, bool MAX>
__global__ void __launch_bounds__(kThreads) k_reduce_abs(size_t n, const T *__restrict__ x, AbsIdx *partials,
                                                         unsigned int *ticket, AbsIdx *out) {
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t) gridDim.x * blockDim.x;
    AbsIdx best{0.0, -1};
#pragma unroll 8
    for (size_t i = tid; i < n; i += nth) {
        AbsIdx c{fabs((double) x[i]), (long long) i};
        best = better<MAX>(best, c);
    }
    AbsIdx bb = block_best<MAX>(best);
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = bb;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        AbsIdx v{0.0, -1};
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
            AbsIdx c;
            c.v = ((volatile AbsIdx *) partials)[i].v;
            c.i = ((volatile AbsIdx *) partials)[i].i;
            v = better<MAX>(v, c);
        }
        AbsIdx tot = block_best<MAX>(v);
        if (threadIdx.x == 0) {
            *out = tot;
            *ticket = 0;
            __threadfence_system();
        }
    }
}

template<typename T, bool MAX>
int reduce_abs(gpub_ctx_t ctx, int sidx, size_t n, const T *x, T *result_host, long long *index_host) {
    if (!result_host) return GPUB_EINVAL;
    if (n == 0) {
        *result_host = T(0);
        if (index_host) *index_host = -1;
        return GPUB_OK;
    }
    if (!x) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    int grid = stream_grid(ctx, n, 8);
    AbsIdx *out = reinterpret_cast<AbsIdx *>(slot->h_result);
    AbsIdx *partials = reinterpret_cast<AbsIdx *>(slot->d_scratch);
    k_reduce_abs<T, MAX><<<grid, kThreads, 0, stream>>>(n, x, partials, slot->d_counter, out);
    GPUB_LAUNCH_CHECK();
    GPUB_CUDA(cudaStreamSynchronize(stream));
    *result_host = (T) out->v;
    if (index_host) *index_host = out->i;
    return GPUB_OK;
}

// ------------------------------------------------------------------------------------------
// scal / axpy
// ------------------------------------------------------------------------------------------
template<typename T, bool VEC, bool AXPY>
__global__ void __launch_bounds__(kThreads) k_scal_axpy(size_t n, T alpha, const T *__restrict__ x, T *y) {
    using V = typename VecOf<T>::type;
    constexpr int VN = VecOf<T>::N;
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t) gridDim.x * blockDim.x;
    size_t done = 0;
    if (VEC) {
        const size_t nv = n / VN;
        const V *xv = reinterpret_cast<const V *>(x);
        V *yv = reinterpret_cast<V *>(y);
#pragma unroll 4
        for (size_t i = tid; i < nv; i += nth) {
            T a[VN], b[VN];
            unpack<T>(yv[i], b);
            if (AXPY) {
                unpack<T>(xv[i], a);
#pragma unroll
                for (int j = 0; j < VN; j++) b[j] = alpha * a[j] + b[j];
            } else {
#pragma unroll
                for (int j = 0; j < VN; j++) b[j] = alpha * b[j];
            }
            yv[i] = pack(b);
        }
        done = nv * VN;
    }
    for (size_t i = done + tid; i < n; i += nth) y[i] = AXPY ? alpha * x[i] + y[i] : alpha * y[i];
}

template<typename T, bool AXPY>
int scal_axpy(gpub_ctx_t ctx, int sidx, size_t n, T alpha, const T *x, T *y) {
    if (n == 0) return GPUB_OK;
    if (!y || (AXPY && !x)) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    int grid = stream_grid(ctx, n, 16);
    bool vec = aligned16(y) && (!AXPY || aligned16(x));
    if (vec)
        k_scal_axpy<T, true, AXPY><<<grid, kThreads, 0, stream>>>(n, alpha, x, y);
    else
        k_scal_axpy<T, false, AXPY><<<grid, kThreads, 0, stream>>>(n, alpha, x, y);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

// ------------------------------------------------------------------------------------------
// Givens helpers
// ------------------------------------------------------------------------------------------
template<typename T>
__global__ void k_rot(size_t n, T *x, size_t incx, T *y, size_t incy, const T *dc, const T *ds, T hc, T hs, bool on_dev) {
    const T c = on_dev ? *dc : hc;
    const T s = on_dev ? *ds : hs;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        T xi = x[i * incx], yi = y[i * incy];
        x[i * incx] = c * xi + s * yi;
        y[i * incy] = c * yi - s * xi;
    }
}

// Batched Givens (additive; the reference's rot / GivensAnnihilator take one matrix, tensor.cuh:1074-1104, 2211-2312).
// k_rot_batched: the same plane rotation of two strided vectors in EVERY matrix of a batch, (c, s) per matrix from device arrays.
// k_givens_annihilate_batched: per matrix, G(i, k) that zeroes element (k, j): {rhypot, cos, -sin} from elements (i, j), (k, j)
// exactly as k_givensAnnihilateRHypot (tensor.cuh:2272-2281), then rows i and k rotated -- one warp per matrix, one launch per
// batch instead of a 1-thread kernel plus a cuBLAS call per matrix.
template<typename T>
__global__ void k_rot_batched(size_t n, T *x, size_t incx, T *y, size_t incy, size_t stride, const T *__restrict__ c, const T *__restrict__ s,
                              size_t batch) {
    const size_t total = n * batch;
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t) gridDim.x * blockDim.x) {
        const size_t b = e / n, i = e - b * n;
        const T cc = c[b], ss = s[b];
        T *xp = x + b * stride + i * incx, *yp = y + b * stride + i * incy;
        const T xi = *xp, yi = *yp;
        *xp = cc * xi + ss * yi;
        *yp = cc * yi - ss * xi;
    }
}

template<typename T> __device__ __forceinline__ T dev_rhypot(T a, T b);
template<> __device__ __forceinline__ double dev_rhypot<double>(double a, double b) { return rhypot(a, b); }
template<> __device__ __forceinline__ float dev_rhypot<float>(float a, float b) { return rhypotf(a, b); }

template<typename T>
__global__ void k_givens_annihilate_batched(T *A, size_t nrows, size_t ncols, size_t stride, size_t i, size_t k, size_t j, size_t batch) {
    const size_t warp = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t) gridDim.x * blockDim.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    for (size_t b = warp; b < batch; b += nwarps) {
        T *a = A + b * stride;
        const T xij = a[i + j * nrows], xkj = a[k + j * nrows];
        const T rh = dev_rhypot<T>(xij, xkj);
        const T c = xij * rh, s = xkj * rh;
        __syncwarp();                                      // every lane has read the pivot pair before column j is rotated
        for (size_t col = lane; col < ncols; col += 32) {
            const T xi = a[i + col * nrows], xk = a[k + col * nrows];
            a[i + col * nrows] = c * xi + s * xk;
            a[k + col * nrows] = c * xk - s * xi;
        }
    }
}

template<typename T>
int rot(gpub_ctx_t ctx, int sidx, size_t n, T *x, size_t incx, T *y, size_t incy, const T *c, const T *s, int on_dev) {
    if (n == 0) return GPUB_OK;
    if (!x || !y || !c || !s) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    int grid = (int) (gpub_ceil_div(n, 256) < 1024 ? gpub_ceil_div(n, 256) : 1024);
    T hc = on_dev ? T(0) : *c, hs = on_dev ? T(0) : *s;
    k_rot<T><<<grid, 256, 0, stream>>>(n, x, incx, y, incy, c, s, hc, hs, on_dev != 0);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

template<typename T>
__global__ void k_givens_rhypot(const T *data, T *res, size_t i, size_t k, size_t j, size_t nrows) {
    T xij = data[i + j * nrows];
    T xkj = data[k + j * nrows];
    T r = rhypot(xij, xkj);
    res[0] = r;
    res[1] = xij * r;
    res[2] = xkj * r;
}

// ------------------------------------------------------------------------------------------
// transpose / gather / tables / counting
// ------------------------------------------------------------------------------------------
template<typename T>
__global__ void k_transpose_small(size_t m, size_t n, const T *__restrict__ A, size_t sA, T *__restrict__ At, size_t sAt,
                                  size_t batch) {
    const size_t mn = m * n;
    const size_t total = mn * batch;
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t) gridDim.x * blockDim.x) {
        size_t b = e / mn, r = e - b * mn;
        size_t j = r % n, i = r / n; // At is n x m: At(j, i) at j + i*n
        At[b * sAt + r] = A[b * sA + i + j * m];
    }
}

template<typename T>
__global__ void k_transpose_tiled(size_t m, size_t n, const T *__restrict__ A, size_t sA, T *__restrict__ At, size_t sAt,
                                  size_t tiles_m, size_t tiles_n, size_t batch) {
    __shared__ T tile[32][33];
    const size_t tiles = tiles_m * tiles_n;
    for (size_t t = blockIdx.x; t < tiles * batch; t += gridDim.x) {
        size_t b = t / tiles, r = t - b * tiles;
        size_t ti = r % tiles_m, tj = r / tiles_m;
        const T *a = A + b * sA;
        T *at = At + b * sAt;
        for (int c = threadIdx.y; c < 32; c += blockDim.y) {
            size_t i = ti * 32 + threadIdx.x, j = tj * 32 + c;
            if (i < m && j < n) tile[c][threadIdx.x] = a[i + j * m];
        }
        __syncthreads();
        for (int c = threadIdx.y; c < 32; c += blockDim.y) {
            size_t j = tj * 32 + threadIdx.x, i = ti * 32 + c;
            if (i < m && j < n) at[j + i * n] = tile[threadIdx.x][c];
        }
        __syncthreads();
    }
}

template<typename T>
int transpose_batched(gpub_ctx_t ctx, int sidx, size_t m, size_t n, const T *A, size_t sA, T *At, size_t sAt, size_t batch) {
    if (m == 0 || n == 0 || batch == 0) return GPUB_OK;
    if (!A || !At) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    if (m < 16 || n < 16) {
        size_t total = m * n * batch;
        int grid = (int) (gpub_ceil_div(total, 256) < (size_t) ctx->sm_count * 16 ? gpub_ceil_div(total, 256)
                                                                                    : (size_t) ctx->sm_count * 16);
        k_transpose_small<T><<<grid, 256, 0, stream>>>(m, n, A, sA, At, sAt, batch);
    } else {
        size_t tm = gpub_ceil_div(m, 32), tn = gpub_ceil_div(n, 32);
        size_t total = tm * tn * batch;
        int grid = (int) (total < (size_t) ctx->sm_count * 16 ? total : (size_t) ctx->sm_count * 16);
        k_transpose_tiled<T><<<grid, dim3(32, 8), 0, stream>>>(m, n, A, sA, At, sAt, tm, tn, batch);
    }
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

template<typename T>
__global__ void k_gather_rows(const T *__restrict__ src, size_t ld, size_t row_from, size_t nr, size_t nc, T *__restrict__ dst) {
    const size_t total = nr * nc;
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t) gridDim.x * blockDim.x) {
        size_t c = e / nr, r = e - c * nr;
        dst[e] = src[row_from + r + c * ld];
    }
}

__global__ void k_fill_ptr_table(char *base, size_t stride_bytes, size_t count, void **table) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x)
        table[i] = base + i * stride_bytes;
}

template<typename T>
__global__ void k_count_gt(const T *__restrict__ S, size_t len, size_t sS, T eps, unsigned int *count, size_t batch) {
    // one warp per matrix: ballot-count, one read-modify-write by lane 0 (accumulating, like the reference)
    const size_t w = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= batch) return;
    unsigned c = 0;
    for (size_t j = lane_id(); j < len; j += 32) c += (S[w * sS + j] > eps) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if (lane_id() == 0) count[w] += c;
}

template<typename T>
__global__ void k_nullspace_pack(size_t n, const T *__restrict__ U, size_t sU, const unsigned int *__restrict__ rank,
                                 T *__restrict__ N, size_t sN, size_t batch) {
    const size_t nn = n * n;
    for (size_t b = blockIdx.y; b < batch; b += gridDim.y) {
        const unsigned r = rank[b] > n ? (unsigned) n : rank[b];
        const size_t keep = (n - r) * n; // the last n-r columns of U move to the front; the rest is zero
        for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t) gridDim.x * blockDim.x)
            N[b * sN + e] = e < keep ? U[b * sU + (size_t) r * n + e] : T(0);
    }
}

template<typename T>
int nullspace_pack_on(cudaStream_t stream, size_t n, const T *U, size_t sU, const unsigned int *rank, T *N, size_t sN, size_t batch) {
    if (n == 0 || batch == 0) return GPUB_OK;
    if (!U || !rank || !N) return GPUB_EINVAL;
    unsigned gx = (unsigned) (gpub_ceil_div(n * n, 256) < 64 ? gpub_ceil_div(n * n, 256) : 64);
    unsigned gy = (unsigned) (batch < 65535 ? batch : 65535);
    k_nullspace_pack<T><<<dim3(gx, gy), 256, 0, stream>>>(n, U, sU, rank, N, sN, batch);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

template<typename T>
int nullspace_pack(gpub_ctx_t ctx, int sidx, size_t n, const T *U, size_t sU, const unsigned int *rank, T *N, size_t sN,
                   size_t batch) {
    if (n == 0 || batch == 0) return GPUB_OK;
    GPUB_ENTER(ctx, sidx);
    return nullspace_pack_on<T>(stream, n, U, sU, rank, N, sN, batch);
}

// ------------------------------------------------------------------------------------------
// synthetic data
// ------------------------------------------------------------------------------------------
template<typename T>
__global__ void k_fill_uniform(size_t n, T *x, double lo, double hi, uint64_t seed) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        x[i] = (T) (lo + (hi - lo) * gpub_u01(seed, i));
}

template<typename T>
__global__ void k_fill_spd(size_t n, T *A, size_t sA, double shift, uint64_t seed, size_t batch) {
    const size_t nn = n * n;
    const size_t total = nn * batch;
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t) gridDim.x * blockDim.x) {
        size_t b = e / nn, r = e - b * nn;
        size_t i = r % n, j = r / n;
        double acc = 0;
        for (size_t k = 0; k < n; k++) {
            double gik = 2.0 * gpub_u01(seed, b * nn + i + k * n) - 1.0;
            double gjk = 2.0 * gpub_u01(seed, b * nn + j + k * n) - 1.0;
            acc += gik * gjk;
        }
        if (i == j) acc += shift;
        A[b * sA + r] = (T) acc;
    }
}

} // namespace

int gpub_internal_nullspace_pack_f64(gpub_ctx_t ctx, cudaStream_t stream, size_t n, const double *U, size_t sU, const unsigned int *rank, double *N, size_t sN, size_t batch) {
    if (!ctx) return GPUB_EINVAL;
    gpub_device_guard guard(ctx->device);
    return nullspace_pack_on<double>(stream, n, U, sU, rank, N, sN, batch);
}
int gpub_internal_nullspace_pack_f32(gpub_ctx_t ctx, cudaStream_t stream, size_t n, const float *U, size_t sU, const unsigned int *rank, float *N, size_t sN, size_t batch) {
    if (!ctx) return GPUB_EINVAL;
    gpub_device_guard guard(ctx->device);
    return nullspace_pack_on<float>(stream, n, U, sU, rank, N, sN, batch);
}

extern "C" {

int gpub_dot_f64(gpub_ctx_t c, int s, size_t n, const double *x, const double *y, double *r) { return reduce_sum<double, OP_DOT>(c, s, n, x, y, r, false); }
int gpub_dot_f32(gpub_ctx_t c, int s, size_t n, const float *x, const float *y, float *r) { return reduce_sum<float, OP_DOT>(c, s, n, x, y, r, false); }
// nrm2, fp64: the squares are summed unscaled (one pass at HBM speed); only when that sum overflowed or underflowed -- |x| beyond
// ~1e154 or below ~1e-154, where cublasDnrm2's scaled algorithm still returns a finite value -- a second pass sums (x / max|x|)^2.
// fp32 data cannot leave the range of the fp64 accumulator.
int gpub_nrm2_f64(gpub_ctx_t c, int s, size_t n, const double *x, double *r) {
    int e = reduce_sum<double, OP_SUMSQ>(c, s, n, x, nullptr, r, true);
    if (e != GPUB_OK || n == 0) return e;
    if (*r > 1e-140 && *r < 1e140) return GPUB_OK;      // sums of squares in [1e-280, 1e280] are exact enough: no second pass
    double amax = 0.0;
    e = reduce_abs<double, true>(c, s, n, x, &amax, nullptr);
    if (e != GPUB_OK) return e;
    if (!(amax > 0.0) || !(amax < 1.7976931348623157e308)) {   // all zero, or inf / NaN in the data: the plain result stands
        if (amax == 0.0) *r = 0.0;
        return GPUB_OK;
    }
    double scaled = 0.0;
    e = reduce_sum<double, OP_SUMSQ>(c, s, n, x, nullptr, &scaled, true, 1.0 / amax);
    if (e == GPUB_OK) *r = amax * scaled;
    return e;
}
int gpub_nrm2_f32(gpub_ctx_t c, int s, size_t n, const float *x, float *r) { return reduce_sum<float, OP_SUMSQ>(c, s, n, x, nullptr, r, true); }
int gpub_asum_f64(gpub_ctx_t c, int s, size_t n, const double *x, double *r) { return reduce_sum<double, OP_ASUM>(c, s, n, x, nullptr, r, false); }
int gpub_asum_f32(gpub_ctx_t c, int s, size_t n, const float *x, float *r) { return reduce_sum<float, OP_ASUM>(c, s, n, x, nullptr, r, false); }

int gpub_amax_abs_f64(gpub_ctx_t c, int s, size_t n, const double *x, double *r, long long *i) { return reduce_abs<double, true>(c, s, n, x, r, i); }
int gpub_amax_abs_f32(gpub_ctx_t c, int s, size_t n, const float *x, float *r, long long *i) { return reduce_abs<float, true>(c, s, n, x, r, i); }
int gpub_amin_abs_f64(gpub_ctx_t c, int s, size_t n, const double *x, double *r, long long *i) { return reduce_abs<double, false>(c, s, n, x, r, i); }
int gpub_amin_abs_f32(gpub_ctx_t c, int s, size_t n, const float *x, float *r, long long *i) { return reduce_abs<float, false>(c, s, n, x, r, i); }

int gpub_scal_f64(gpub_ctx_t c, int s, size_t n, double a, double *x) { return scal_axpy<double, false>(c, s, n, a, nullptr, x); }
int gpub_scal_f32(gpub_ctx_t c, int s, size_t n, float a, float *x) { return scal_axpy<float, false>(c, s, n, a, nullptr, x); }
int gpub_axpy_f64(gpub_ctx_t c, int s, size_t n, double a, const double *x, double *y) { return scal_axpy<double, true>(c, s, n, a, x, y); }
int gpub_axpy_f32(gpub_ctx_t c, int s, size_t n, float a, const float *x, float *y) { return scal_axpy<float, true>(c, s, n, a, x, y); }

int gpub_rot_f64(gpub_ctx_t c, int s, size_t n, double *x, size_t ix, double *y, size_t iy, const double *cc, const double *ss, int d) { return rot<double>(c, s, n, x, ix, y, iy, cc, ss, d); }
int gpub_rot_f32(gpub_ctx_t c, int s, size_t n, float *x, size_t ix, float *y, size_t iy, const float *cc, const float *ss, int d) { return rot<float>(c, s, n, x, ix, y, iy, cc, ss, d); }

int gpub_givens_rhypot_f64(gpub_ctx_t ctx, int sidx, const double *data, double *res, size_t i, size_t k, size_t j, size_t nrows) {
    if (!data || !res) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    k_givens_rhypot<double><<<1, 1, 0, stream>>>(data, res, i, k, j, nrows);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}
int gpub_givens_rhypot_f32(gpub_ctx_t ctx, int sidx, const float *data, float *res, size_t i, size_t k, size_t j, size_t nrows) {
    if (!data || !res) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    k_givens_rhypot<float><<<1, 1, 0, stream>>>(data, res, i, k, j, nrows);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

#define GPUB_DEF_GIVENS_BATCHED(SUF, T)                                                                                \
    int gpub_rot_batched_##SUF(gpub_ctx_t ctx, int sidx, size_t n, T *x, size_t incx, T *y, size_t incy, size_t stride, \
                               const T *c, const T *s, size_t batch) {                                                 \
        if (n == 0 || batch == 0) return GPUB_OK;                                                                      \
        if (!x || !y || !c || !s) return GPUB_EINVAL;                                                                  \
        GPUB_ENTER(ctx, sidx);                                                                                         \
        const size_t want = gpub_ceil_div(n * batch, 256);                                                             \
        k_rot_batched<T><<<(unsigned) (want < 4096 ? want : 4096), 256, 0, stream>>>(n, x, incx, y, incy, stride, c, s, batch); \
        GPUB_LAUNCH_CHECK();                                                                                           \
        return GPUB_OK;                                                                                                \
    }                                                                                                                  \
    int gpub_givens_annihilate_batched_##SUF(gpub_ctx_t ctx, int sidx, T *A, size_t nrows, size_t ncols, size_t stride, \
                                             size_t i, size_t k, size_t j, size_t batch) {                             \
        if (batch == 0 || ncols == 0) return GPUB_OK;                                                                  \
        if (!A || i >= nrows || k >= nrows || j >= ncols || i == k) return GPUB_EINVAL;                                \
        GPUB_ENTER(ctx, sidx);                                                                                         \
        const size_t want = gpub_ceil_div(batch, 8);                                                                   \
        const size_t cap = (size_t) ctx->sm_count * 8;                                                                 \
        k_givens_annihilate_batched<T><<<(unsigned) (want < cap ? want : cap), 256, 0, stream>>>(A, nrows, ncols, stride, i, k, j, batch); \
        GPUB_LAUNCH_CHECK();                                                                                           \
        return GPUB_OK;                                                                                                \
    }
GPUB_DEF_GIVENS_BATCHED(f64, double)
GPUB_DEF_GIVENS_BATCHED(f32, float)

#define GPUB_DEF_GATHER(SUF, T)                                                                                        \
    int gpub_gather_rows_##SUF(gpub_ctx_t ctx, int sidx, const T *src, size_t ld, size_t row_from, size_t nr, size_t nc, \
                               T *dst) {                                                                               \
        if (nr == 0 || nc == 0) return GPUB_OK;                                                                        \
        if (!src || !dst) return GPUB_EINVAL;                                                                          \
        GPUB_ENTER(ctx, sidx);                                                                                         \
        size_t total = nr * nc;                                                                                        \
        int grid = (int) (gpub_ceil_div(total, 256) < 2048 ? gpub_ceil_div(total, 256) : 2048);                        \
        k_gather_rows<T><<<grid, 256, 0, stream>>>(src, ld, row_from, nr, nc, dst);                                    \
        GPUB_LAUNCH_CHECK();                                                                                           \
        return GPUB_OK;                                                                                                \
    }
GPUB_DEF_GATHER(f64, double)
GPUB_DEF_GATHER(f32, float)

int gpub_transpose_batched_f64(gpub_ctx_t c, int s, size_t m, size_t n, const double *A, size_t sA, double *At, size_t sAt, size_t b) { return transpose_batched<double>(c, s, m, n, A, sA, At, sAt, b); }
int gpub_transpose_batched_f32(gpub_ctx_t c, int s, size_t m, size_t n, const float *A, size_t sA, float *At, size_t sAt, size_t b) { return transpose_batched<float>(c, s, m, n, A, sA, At, sAt, b); }

int gpub_fill_ptr_table(gpub_ctx_t ctx, int sidx, void *base, size_t stride_bytes, size_t count, void **table) {
    if (count == 0) return GPUB_OK;
    if (!table) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    int grid = (int) (gpub_ceil_div(count, 256) < 2048 ? gpub_ceil_div(count, 256) : 2048);
    k_fill_ptr_table<<<grid, 256, 0, stream>>>((char *) base, stride_bytes, count, table);
    GPUB_LAUNCH_CHECK();
    return GPUB_OK;
}

#define GPUB_DEF_COUNT(SUF, T)                                                                                     \
    int gpub_count_gt_batched_##SUF(gpub_ctx_t ctx, int sidx, const T *S, size_t len, size_t sS, T eps,             \
                                    unsigned int *count, size_t batch) {                                           \
        if (batch == 0) return GPUB_OK;                                                                            \
        if (!S || !count) return GPUB_EINVAL;                                                                      \
        GPUB_ENTER(ctx, sidx);                                                                                     \
        size_t grid = gpub_ceil_div(batch * 32, 256);                                                              \
        k_count_gt<T><<<(unsigned) grid, 256, 0, stream>>>(S, len, sS, eps, count, batch);                         \
        GPUB_LAUNCH_CHECK();                                                                                       \
        return GPUB_OK;                                                                                            \
    }
GPUB_DEF_COUNT(f64, double)
GPUB_DEF_COUNT(f32, float)

int gpub_nullspace_pack_batched_f64(gpub_ctx_t c, int s, size_t n, const double *U, size_t sU, const unsigned int *r, double *N, size_t sN, size_t b) { return nullspace_pack<double>(c, s, n, U, sU, r, N, sN, b); }
int gpub_nullspace_pack_batched_f32(gpub_ctx_t c, int s, size_t n, const float *U, size_t sU, const unsigned int *r, float *N, size_t sN, size_t b) { return nullspace_pack<float>(c, s, n, U, sU, r, N, sN, b); }

#define GPUB_DEF_FILL(SUF, T)                                                                                       \
    int gpub_fill_uniform_##SUF(gpub_ctx_t ctx, int sidx, size_t n, T *x, T lo, T hi, uint64_t seed) {               \
        if (n == 0) return GPUB_OK;                                                                                 \
        if (!x) return GPUB_EINVAL;                                                                                 \
        GPUB_ENTER(ctx, sidx);                                                                                      \
        int grid = stream_grid(ctx, n, 4);                                                                          \
        k_fill_uniform<T><<<grid, kThreads, 0, stream>>>(n, x, (double) lo, (double) hi, seed);                     \
        GPUB_LAUNCH_CHECK();                                                                                        \
        return GPUB_OK;                                                                                             \
    }                                                                                                               \
    int gpub_fill_spd_batched_##SUF(gpub_ctx_t ctx, int sidx, size_t n, T *A, size_t sA, T shift, uint64_t seed,     \
                                    size_t batch) {                                                                 \
        if (n == 0 || batch == 0) return GPUB_OK;                                                                   \
        if (!A) return GPUB_EINVAL;                                                                                 \
        GPUB_ENTER(ctx, sidx);                                                                                      \
        int grid = stream_grid(ctx, n * n * batch, 1);                                                              \
        k_fill_spd<T><<<grid, kThreads, 0, stream>>>(n, A, sA, (double) shift, seed, batch);                        \
        GPUB_LAUNCH_CHECK();                                                                                        \
        return GPUB_OK;                                                                                             \
    }
GPUB_DEF_FILL(f64, double)
GPUB_DEF_FILL(f32, float)

} // extern "C"
