/*
 * core.cuh -- macros, random-vector helpers, error convention and the Session facade.
 *
 * Mirrors the public surface of the reference header's first section (ref: tensor.cuh:23-247):
 * DEFAULT_FPX, THREADS_PER_BLOCK, TEMPLATE_WITH_TYPE_T, TEMPLATE_CONSTRAINT_REQUIRES_FPX,
 * generateRealRandomVector / generateIntRandomVector, numBlocks, gpuErrChk / gpuAssert, Session.
 *
 * What changed underneath: Session no longer owns cuBLAS / cuSOLVER handles. It is a facade over the
 * per-device stream context of libgputils_b200 (gpub_ctx_*), which owns N blocking streams and the
 * per-stream reduction scratch. The byte counter stays here because it is API-level accounting that
 * the reference's tests pin (testTensor.cu:955-992).
 */
#ifndef GPUB200_CORE_CUH
#define GPUB200_CORE_CUH

#include <algorithm>
#include <cassert>
#include <concepts>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <memory>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>

#include "../gputils_b200.h"

#ifdef GPUTILS_B200_ENABLE_CUBLAS_HANDLES
#include <cublas_v2.h>
#include <cusolverDn.h>
#endif

/* ---- defaults and template helpers (same names as the reference) ---- */
#define DEFAULT_FPX double
#define THREADS_PER_BLOCK 512
#if (__cplusplus >= 201703L)
#define TEMPLATE_WITH_TYPE_T template<typename T = DEFAULT_FPX>
#else
#define TEMPLATE_WITH_TYPE_T template<typename T>
#endif
#if (__cplusplus >= 202002L)
#define TEMPLATE_CONSTRAINT_REQUIRES_FPX requires std::floating_point<T>
#else
#define TEMPLATE_CONSTRAINT_REQUIRES_FPX
#endif

static std::random_device RND_DEVICE;

/**
 * Vector of n reals drawn uniformly from [low, hi] (unseeded, like the reference).
 */
TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
std::vector<T> generateRealRandomVector(size_t n, T low, T hi) {
    std::mt19937_64 engine(RND_DEVICE());
    std::uniform_real_distribution<T> dist(low, hi);
    std::vector<T> out(n);
    for (auto &v: out) v = dist(engine);
    return out;
}

/**
 * Vector of n integers drawn uniformly from {low, ..., hi}.
 */
inline std::vector<int> generateIntRandomVector(size_t n, int low, int hi) {
    std::mt19937_64 engine(RND_DEVICE());
    std::uniform_int_distribution<int> dist(low, hi);
    std::vector<int> out(n);
    for (auto &v: out) v = dist(engine);
    return out;
}

/**
 * Number of blocks of `threads_per_block` threads that cover n tasks.
 */
constexpr size_t numBlocks(size_t n, size_t threads_per_block = THREADS_PER_BLOCK) {
    return (n + threads_per_block - 1) / threads_per_block;
}

/* ---- error convention: print file:line to stderr and exit(code) ---- */
#define gpuErrChk(status) { gpuAssert((status), __FILE__, __LINE__); } while(false)

TEMPLATE_WITH_TYPE_T
inline void gpuAssert(T code, const char *file, int line, bool abort = true) {
    if constexpr (std::is_same_v<T, cudaError_t>) {
        if (code != cudaSuccess) {
            std::cerr << "cuda error. String: " << cudaGetErrorString(code)
                      << ", file: " << file << ", line: " << line << "\n";
            if (abort) exit(code);
        }
    } else if constexpr (std::is_same_v<T, int>) {
        /* status of a libgputils_b200 launcher: >0 is a cudaError_t, <0 a GPUB_E* argument error */
        if (code != GPUB_OK) {
            if (code > 0) {
                std::cerr << "gputils_b200 launch error (cuda). String: "
                          << cudaGetErrorString(static_cast<cudaError_t>(code));
            } else {
                std::cerr << "gputils_b200 launch error. Code: " << code;
            }
            std::cerr << ", file: " << file << ", line: " << line << "\n";
            if (abort) exit(code > 0 ? code : 1);
        }
#ifdef GPUTILS_B200_ENABLE_CUBLAS_HANDLES
    } else if constexpr (std::is_same_v<T, cublasStatus_t>) {
        if (code != CUBLAS_STATUS_SUCCESS) {
            std::cerr << "cublas error. Name: " << cublasGetStatusName(code)
                      << ", file: " << file << ", line: " << line << "\n";
            if (abort) exit(code);
        }
    } else if constexpr (std::is_same_v<T, cusolverStatus_t>) {
        if (code != CUSOLVER_STATUS_SUCCESS) {
            std::cerr << "cusolver error. Status: " << code
                      << ", file: " << file << ", line: " << line << "\n";
            if (abort) exit(code);
        }
#endif
    } else {
        std::cerr << "Error: library status parser not implemented" << "\n";
    }
}

/* ================================================================================================
 *  SESSION
 * ================================================================================================ */
/** Number of streams the Session is created with; change with Session::setStreams() before first use. */
static size_t s_numStreams = 1;

/**
 * Process-wide runtime context.
 * The reference kept one cuBLAS and one cuSOLVER handle per stream here; this implementation keeps a
 * handle to the per-device stream context of libgputils_b200 instead. All public methods of the
 * reference are retained with the same meaning.
 */
class Session {
public:
    static void setStreams(size_t numStreams) { s_numStreams = numStreams; }

    static Session &getInstance() {
        static Session instance(s_numStreams);
        return instance;
    }

private:
    explicit Session(size_t numStreams) : m_numStreams(numStreams == 0 ? 1 : numStreams) {
        gpuErrChk(cudaGetDevice(&m_device));
        gpuErrChk(gpub_ctx_get(m_device, &m_ctx));
        gpuErrChk(gpub_ctx_ensure_streams(m_ctx, static_cast<int>(m_numStreams)));
    }

    ~Session() {
        /* like the reference's destructor (tensor.cuh:168-173), which destroys its handles and streams:
         * hands the streams and the per-stream scratch back, so leak checkers see a clean exit */
        gpub_multi_release();   /* NCCL communicators of sharded tensors, if any were made */
        gpub_ctx_release_all(); /* this device's context and those sharded tensors created on other devices */
#ifdef GPUTILS_B200_ENABLE_CUBLAS_HANDLES
        for (auto &h: m_cublasHandles) if (h) cublasDestroy(h);
        for (auto &h: m_cusolverHandles) if (h) cusolverDnDestroy(h);
#endif
    }

    gpub_ctx_t m_ctx = nullptr;
    int m_device = 0;
    size_t m_bytesAllocated = 0;
    size_t m_numStreams = 1;
#ifdef GPUTILS_B200_ENABLE_CUBLAS_HANDLES
    std::vector<cublasHandle_t> m_cublasHandles;
    std::vector<cusolverDnHandle_t> m_cusolverHandles;
#endif

public:
    Session(Session const &) = delete;

    void operator=(Session const &) = delete;

    /** Stream context of libgputils_b200 for the current device (additive API). */
    gpub_ctx_t context() const { return m_ctx; }

    /** Raw CUDA stream behind stream index idx (additive API). */
    cudaStream_t stream(size_t idx = 0) const {
        void *s = nullptr;
        gpuErrChk(gpub_ctx_stream(m_ctx, static_cast<int>(idx), &s));
        return static_cast<cudaStream_t>(s);
    }

    /** Device the Session was created on (the current device at first use, as in the reference). */
    int device() const { return m_device; }

    /**
     * Stream context of the CURRENT device (additive API). Equal to context() in single-GPU programs; sharded tensors
     * (gpub200/sharded.cuh) make another device current around their launches and get that device's context.
     */
    gpub_ctx_t contextOfCurrentDevice() const {
        int device = 0;
        gpuErrChk(cudaGetDevice(&device));
        if (device == m_device) return m_ctx;
        gpub_ctx_t c = nullptr;
        gpuErrChk(gpub_ctx_get(device, &c));
        return c;
    }

    /** Raw CUDA stream behind stream index idx of the current device's context (additive API). */
    cudaStream_t streamOfCurrentDevice(size_t idx = 0) const {
        void *s = nullptr;
        gpuErrChk(gpub_ctx_stream(contextOfCurrentDevice(), static_cast<int>(idx), &s));
        return static_cast<cudaStream_t>(s);
    }

    /** Number of streams (additive API). */
    size_t numStreams() const { return m_numStreams; }

#ifdef GPUTILS_B200_ENABLE_CUBLAS_HANDLES
    /* Optional: user code that still wants a cuBLAS / cuSOLVER handle bound to stream idx gets one,
     * created on first request. Nothing in this library uses them. */
    cublasHandle_t &cuBlasHandle(size_t idx = 0) {
        if (m_cublasHandles.size() <= idx) m_cublasHandles.resize(idx + 1, nullptr);
        if (!m_cublasHandles[idx]) {
            gpuErrChk(cublasCreate(&m_cublasHandles[idx]));
            gpuErrChk(cublasSetStream(m_cublasHandles[idx], stream(idx)));
        }
        return m_cublasHandles[idx];
    }

    cusolverDnHandle_t &cuSolverHandle(size_t idx = 0) {
        if (m_cusolverHandles.size() <= idx) m_cusolverHandles.resize(idx + 1, nullptr);
        if (!m_cusolverHandles[idx]) {
            gpuErrChk(cusolverDnCreate(&m_cusolverHandles[idx]));
            gpuErrChk(cusolverDnSetStream(m_cusolverHandles[idx], stream(idx)));
        }
        return m_cusolverHandles[idx];
    }
#endif

    /**
     * Device allocation that counts the bytes handed out. The reference calls cudaMalloc here (tensor.cuh:203-210); this takes
     * the block from the stream-ordered pool of the current device's context (gpub_mem_alloc: ordered on the legacy default
     * stream, so it is an ordering point for all the Session's blocking streams like cudaMalloc was, without stopping the host;
     * freed blocks stay cached). Release with Session::cudaRelease (or cudaFree, which also accepts pool memory).
     */
    cudaError_t cudaAllocate(void **d, size_t s) {
        const int st = gpub_mem_alloc(contextOfCurrentDevice(), s, d);
        if (st == GPUB_OK) m_bytesAllocated += s;
        return st >= 0 ? static_cast<cudaError_t>(st) : cudaErrorInvalidValue;
    }

    /** Returns a block obtained from cudaAllocate to its pool (additive API; the byte counter is the caller's business). */
    cudaError_t cudaRelease(void *d) {
        const int st = gpub_mem_free(d);
        return st >= 0 ? static_cast<cudaError_t>(st) : cudaErrorInvalidValue;
    }

    size_t totalAllocatedBytes() const { return m_bytesAllocated; }

    /** Adjust the byte counter (may be negative). */
    void incrementAllocatedBytes(int s) { m_bytesAllocated += s; }

    /** 64-bit variant used internally so tensors of 2 GiB and more are accounted correctly. */
    void adjustAllocatedBytes(long long s) { m_bytesAllocated = static_cast<size_t>(static_cast<long long>(m_bytesAllocated) + s); }

    void synchronizeStream(size_t idx = 0) const {
        if (idx >= m_numStreams) throw std::runtime_error("stream index out of range");
        gpuErrChk(gpub_ctx_sync(contextOfCurrentDevice(), static_cast<int>(idx)));
    }

    void synchronizeAllStreams() const {
        for (size_t i = 0; i < m_numStreams; i++) synchronizeStream(i);
    }
};

namespace gpub200 {
/* Type dispatch from T to the _f32 / _f64 entry points of the C ABI. */
inline gpub_ctx_t ctx() { return Session::getInstance().contextOfCurrentDevice(); }

template<typename T> struct Abi;

#define GPUB200_ABI(T, SUF)                                                                                         \
    template<> struct Abi<T> {                                                                                      \
        static constexpr auto dot = gpub_dot_##SUF;                                                                 \
        static constexpr auto nrm2 = gpub_nrm2_##SUF;                                                               \
        static constexpr auto asum = gpub_asum_##SUF;                                                               \
        static constexpr auto amax = gpub_amax_abs_##SUF;                                                           \
        static constexpr auto amin = gpub_amin_abs_##SUF;                                                           \
        static constexpr auto scal = gpub_scal_##SUF;                                                               \
        static constexpr auto axpy = gpub_axpy_##SUF;                                                               \
        static constexpr auto rot = gpub_rot_##SUF;                                                                 \
        static constexpr auto rhypot = gpub_givens_rhypot_##SUF;                                                    \
        static constexpr auto rot_batched = gpub_rot_batched_##SUF;                                                 \
        static constexpr auto annihilate_batched = gpub_givens_annihilate_batched_##SUF;                            \
        static constexpr auto gather_rows = gpub_gather_rows_##SUF;                                                 \
        static constexpr auto transpose = gpub_transpose_batched_##SUF;                                             \
        static constexpr auto gemm = gpub_gemm_batched_##SUF;                                                       \
        static constexpr auto potrf = gpub_potrf_batched_##SUF;                                                     \
        static constexpr auto potrs = gpub_potrs_batched_##SUF;                                                     \
        static constexpr auto potrs_allgather = gpub_potrs_allgather_batched_##SUF;                                 \
        static constexpr auto geqrf = gpub_geqrf_batched_##SUF;                                                     \
        static constexpr auto ormqr = gpub_ormqr_batched_##SUF;                                                     \
        static constexpr auto trsv = gpub_trsv_upper_batched_##SUF;                                                 \
        static constexpr auto gels = gpub_gels_batched_##SUF;                                                       \
        static constexpr auto gesvd = gpub_gesvd_batched_##SUF;                                                     \
        static constexpr auto gesvd_worksize = gpub_gesvd_batched_worksize_##SUF;                                   \
        static constexpr auto count_gt = gpub_count_gt_batched_##SUF;                                               \
        static constexpr auto nullspace_pack = gpub_nullspace_pack_batched_##SUF;                                   \
        static constexpr auto aat = gpub_aat_batched_##SUF;                                                         \
        static constexpr auto projector = gpub_nullspace_projector_batched_##SUF;                                   \
        static constexpr auto nullspace_build = gpub_nullspace_build_batched_##SUF;                                \
        static constexpr auto chol_from_host = gpub_chol_solve_from_host_##SUF;                                     \
    };
GPUB200_ABI(float, f32)
GPUB200_ABI(double, f64)
#undef GPUB200_ABI
} // namespace gpub200

#endif /* GPUB200_CORE_CUH */
