// ref_harness.cu -- C entry points around the UNMODIFIED reference header (GPUtils include/tensor.cuh, found
// with -I $(REF)/include, compiled where it lies; see oracle/Makefile target `ref`).
//
// TEST INFRASTRUCTURE ONLY: built into oracle/_ref/libgputils_ref.so (git-ignored), linked against cuBLAS /
// cuSOLVER 12.9. It lets the -m gpu tests and bench.py run the reference's own public API (DTensor::addAB,
// CholeskyBatchFactoriser, leastSquaresBatched, QRFactoriser, Svd, Nullspace) on the same device buffers as the
// new kernels, and time it with CUDA events. The product never links or loads this library.
//
// All pointers are DEVICE pointers in the reference layout (column-major, mats axis slowest). Each call copies the
// inputs into reference-owned DTensors outside the timed region, runs the reference method `reps` times (inputs
// restored before every repetition for in-place operations), and reports the mean milliseconds per repetition.
#include <tensor.cuh>

#include <algorithm>
#include <chrono>
#include <vector>

namespace {

struct Timer {
    cudaEvent_t a, b;
    float total = 0;
    Timer() {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    ~Timer() {
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    // events on the legacy default stream order against the reference's blocking streams
    void start() { cudaEventRecord(a, 0); }
    void stop() {
        cudaEventRecord(b, 0);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        total += ms;
    }
};

template<typename T>
void d2d(T *dst, const T *src, size_t count) {
    gpuErrChk(cudaMemcpy(dst, src, count * sizeof(T), cudaMemcpyDeviceToDevice));
}

template<typename T>
int ref_addAB(size_t m, size_t n, size_t k, size_t batch, const T *A, const T *B, T *C, T alpha, T beta, int reps, float *ms) {
    DTensor<T> dA(m, k, batch), dB(k, n, batch), dC(m, n, batch), dC0(m, n, batch);
    d2d(dA.raw(), A, m * k * batch);
    d2d(dB.raw(), B, k * n * batch);
    d2d(dC0.raw(), C, m * n * batch);
    Timer t;
    for (int r = 0; r < std::max(reps, 1); r++) {
        d2d(dC.raw(), dC0.raw(), m * n * batch);
        t.start();
        dC.addAB(dA, dB, alpha, beta);
        t.stop();
    }
    d2d(C, dC.raw(), m * n * batch);
    if (ms) *ms = t.total / std::max(reps, 1);
    return 0;
}

template<typename T>
int ref_chol_batch(size_t n, size_t batch, const T *A, T *L, const T *b, T *x, int *info, int reps, float *ms_factor,
                   float *ms_solve) {
    DTensor<T> dA(n, n, batch), dB(n, 1, batch);
    Timer tf, ts;
    for (int r = 0; r < std::max(reps, 1); r++) {
        d2d(dA.raw(), A, n * n * batch);
        if (b) d2d(dB.raw(), b, n * batch);
        CholeskyBatchFactoriser<T> chol(dA);
        tf.start();
        chol.factorise();
        tf.stop();
        if (b) {
            ts.start();
            chol.solve(dB);
            ts.stop();
        }
        if (r == std::max(reps, 1) - 1 && info)
            gpuErrChk(cudaMemcpy(info, chol.info().raw(), batch * sizeof(int), cudaMemcpyDeviceToDevice));
    }
    d2d(L, dA.raw(), n * n * batch);
    if (b && x) d2d(x, dB.raw(), n * batch);
    if (ms_factor) *ms_factor = tf.total / std::max(reps, 1);
    if (ms_solve) *ms_solve = ts.total / std::max(reps, 1);
    return 0;
}

// The reference's host path, as its API prescribes it: upload (tensor.cuh:1128-1145), factorise, solve (2135-2197), download
// (1147-1154) of whole tensors from / to host vectors. Wall-clock seconds per repetition (the calls block); the tensors are
// constructed once, outside the timed region.
template<typename T>
double ref_chol_batch_host(size_t n, size_t batch, const T *hA, const T *hb, T *hx, int *hinfo, int reps) {
    std::vector<T> vA(hA, hA + n * n * batch), vb(hb, hb + n * batch), vx;
    std::vector<int> vi;
    DTensor<T> dA(n, n, batch), dB(n, 1, batch);
    double total = 0;
    for (int r = 0; r < std::max(reps, 1); r++) {
        gpuErrChk(cudaDeviceSynchronize());
        auto t0 = std::chrono::steady_clock::now();
        dA.upload(vA);
        dB.upload(vb);
        CholeskyBatchFactoriser<T> chol(dA);
        chol.factorise();
        chol.solve(dB);
        dB.download(vx);
        chol.info().download(vi);
        gpuErrChk(cudaDeviceSynchronize());
        total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    if (hx) std::copy(vx.begin(), vx.end(), hx);
    if (hinfo) std::copy(vi.begin(), vi.end(), hinfo);
    return total / std::max(reps, 1);
}

template<typename T>
int ref_gels(size_t m, size_t n, size_t batch, const T *A, T *Aout, const T *b, T *bout, int reps, float *ms) {
    DTensor<T> dA(m, n, batch), dB(m, 1, batch);
    Timer t;
    for (int r = 0; r < std::max(reps, 1); r++) {
        d2d(dA.raw(), A, m * n * batch);
        d2d(dB.raw(), b, m * batch);
        t.start();
        dA.leastSquaresBatched(dB);
        t.stop();
    }
    if (Aout) d2d(Aout, dA.raw(), m * n * batch);
    d2d(bout, dB.raw(), m * batch);
    if (ms) *ms = t.total / std::max(reps, 1);
    return 0;
}

// The reference's QRFactoriser is single-matrix: the batch is the caller's host loop (tensor.cuh:1811-1813).
template<typename T>
int ref_qr(size_t m, size_t n, size_t batch, const T *A, T *Aout, const T *b, T *bout, int reps, float *ms_factor,
           float *ms_ls) {
    DTensor<T> dA(m, n, 1), dB(m, 1, 1);
    QRFactoriser<T> qr(dA);
    Timer tf, ts;
    for (int r = 0; r < std::max(reps, 1); r++) {
        for (size_t i = 0; i < batch; i++) {
            d2d(dA.raw(), A + i * m * n, m * n);
            tf.start();
            qr.factorise();
            tf.stop();
            if (Aout) d2d(Aout + i * m * n, dA.raw(), m * n);
            if (b) {
                d2d(dB.raw(), b + i * m, m);
                ts.start();
                qr.leastSquares(dB);
                ts.stop();
                if (bout) d2d(bout + i * m, dB.raw(), m);
            }
        }
    }
    if (ms_factor) *ms_factor = tf.total / std::max(reps, 1);
    if (ms_ls) *ms_ls = ts.total / std::max(reps, 1);
    return 0;
}

template<typename T>
int ref_svd(size_t m, size_t n, size_t batch, const T *A, T *S, T *Vt, T *U, int *info, unsigned int *rank, T eps, int reps,
            float *ms) {
    DTensor<T> dA(m, n, batch);
    Timer t;
    for (int r = 0; r < std::max(reps, 1); r++) {
        d2d(dA.raw(), A, m * n * batch);
        Svd<T> svd(dA, U != nullptr, true);
        t.start();
        svd.factorise();
        t.stop();
        if (r == std::max(reps, 1) - 1) {
            d2d(S, svd.singularValues().raw(), n * batch);
            d2d(Vt, svd.rightSingularVectors().raw(), n * n * batch);
            if (U) d2d(U, svd.leftSingularVectors().value()->raw(), m * m * batch);
            if (info) gpuErrChk(cudaMemcpy(info, svd.info().raw(), batch * sizeof(int), cudaMemcpyDeviceToDevice));
            if (rank) {
                auto const &rk = svd.rank(eps);
                gpuErrChk(cudaDeviceSynchronize());
                gpuErrChk(cudaMemcpy(rank, rk.raw(), batch * sizeof(unsigned int), cudaMemcpyDeviceToDevice));
            }
        }
    }
    if (ms) *ms = t.total / std::max(reps, 1);
    return 0;
}

template<typename T>
int ref_nullspace(size_t m, size_t n, size_t batch, const T *A, T *N, const T *b, T *proj, int reps, float *ms_build,
                  float *ms_project) {
    DTensor<T> dA(m, n, batch), dB(n, 1, batch);
    Timer tb, tp;
    for (int r = 0; r < std::max(reps, 1); r++) {
        d2d(dA.raw(), A, m * n * batch);
        tb.start();
        Nullspace<T> ns(dA);
        tb.stop();
        if (N) d2d(N, ns.nullspace().raw(), n * n * batch);
        if (b) {
            d2d(dB.raw(), b, n * batch);
            tp.start();
            ns.project(dB);
            tp.stop();
            if (proj) d2d(proj, dB.raw(), n * batch);
        }
    }
    if (ms_build) *ms_build = tb.total / std::max(reps, 1);
    if (ms_project) *ms_project = tp.total / std::max(reps, 1);
    return 0;
}

template<typename T>
int ref_reductions(size_t count, const T *x, const T *y, double *out5) {
    DTensor<T> dx(count), dy(count);
    d2d(dx.raw(), x, count);
    d2d(dy.raw(), y, count);
    out5[0] = (double) dx.normF();
    out5[1] = (double) dx.sumAbs();
    out5[2] = (double) dx.dotF(dy);
    out5[3] = (double) dx.maxAbs();
    out5[4] = (double) dx.minAbs();
    return 0;
}

} // namespace

extern "C" {

#define CUDART_VERSION_STR "12.9"
const char *ref_description(void) { return "GPUtils reference header, cuBLAS/cuSOLVER " CUDART_VERSION_STR; }

#define REF_DEFINE(T, SUF)                                                                                                 \
    int ref_addAB_##SUF(size_t m, size_t n, size_t k, size_t batch, const T *A, const T *B, T *C, T alpha, T beta, int reps, \
                        float *ms) { return ref_addAB<T>(m, n, k, batch, A, B, C, alpha, beta, reps, ms); }                \
    int ref_chol_batch_##SUF(size_t n, size_t batch, const T *A, T *L, const T *b, T *x, int *info, int reps, float *msf,  \
                             float *mss) { return ref_chol_batch<T>(n, batch, A, L, b, x, info, reps, msf, mss); }         \
    double ref_chol_batch_host_##SUF(size_t n, size_t batch, const T *hA, const T *hb, T *hx, int *hinfo, int reps) {       \
        return ref_chol_batch_host<T>(n, batch, hA, hb, hx, hinfo, reps);                                                  \
    }                                                                                                                      \
    int ref_gels_##SUF(size_t m, size_t n, size_t batch, const T *A, T *Aout, const T *b, T *bout, int reps, float *ms) {  \
        return ref_gels<T>(m, n, batch, A, Aout, b, bout, reps, ms);                                                       \
    }                                                                                                                      \
    int ref_qr_##SUF(size_t m, size_t n, size_t batch, const T *A, T *Aout, const T *b, T *bout, int reps, float *msf,     \
                     float *msl) { return ref_qr<T>(m, n, batch, A, Aout, b, bout, reps, msf, msl); }                      \
    int ref_svd_##SUF(size_t m, size_t n, size_t batch, const T *A, T *S, T *Vt, T *U, int *info, unsigned int *rank,      \
                      T eps, int reps, float *ms) { return ref_svd<T>(m, n, batch, A, S, Vt, U, info, rank, eps, reps, ms); } \
    int ref_nullspace_##SUF(size_t m, size_t n, size_t batch, const T *A, T *N, const T *b, T *proj, int reps, float *msb, \
                            float *msp) { return ref_nullspace<T>(m, n, batch, A, N, b, proj, reps, msb, msp); }           \
    int ref_reductions_##SUF(size_t count, const T *x, const T *y, double *out5) { return ref_reductions<T>(count, x, y, out5); }

REF_DEFINE(double, f64)
REF_DEFINE(float, f32)

} // extern "C"
