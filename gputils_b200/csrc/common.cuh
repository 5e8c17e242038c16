// Shared plumbing for libgputils_b200: context layout, error macros, small device helpers.
// Internal header -- the public boundary is include/gputils_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <deque>
#include <vector>

#include "gputils_b200.h"

#define GPUB_SCRATCH_BYTES (64u * 1024u)
#define GPUB_HOSTBUF_BYTES 256u

struct gpub_stream_slot {
    cudaStream_t stream = nullptr;
    bool owned = false;
    void *d_scratch = nullptr;         // GPUB_SCRATCH_BYTES of device scratch (reduction partials)
    unsigned int *d_counter = nullptr; // "last block" ticket, always left at zero
    void *h_result = nullptr;          // pinned + mapped host buffer the final block writes into
    void *d_big = nullptr;             // grow-only device scratch (alias copies of small operands), see gpub_slot_big
    size_t big_bytes = 0;
};

#define GPUB_RING_SLOTS 4
#define GPUB_RING_CHUNK_BYTES (8ull << 20)

// stream slots from this index on are the library's own: side streams for work that can run beside the caller's stream
// (chunks of a batched SVD); gpub_ctx_num_streams does not count them
constexpr int GPUB_INTERNAL_SLOT0 = 4064;

struct gpub_ctx {
    int device = 0;
    int sm_count = 148;
    int max_smem_optin = 0;
    std::deque<gpub_stream_slot> slots;  // deque: growing never moves existing slots
    size_t user_slots = 0;               // slots below GPUB_INTERNAL_SLOT0 in use (the library's own side streams live above it)
    std::mutex mu;
    // stream-ordered allocator behind Session::cudaAllocate (gpub_mem_alloc / gpub_mem_free): a context-owned pool whose
    // release threshold keeps freed blocks cached, so a DTensor constructor / destructor is a pool hit, not a cudaMalloc
    cudaMemPool_t pool = nullptr;
    bool pool_tried = false;
    // host <-> device staging (gpub_upload / gpub_download / gpub_chol_solve_from_host): pinned bounce ring + events,
    // two private non-blocking streams for the upload / download legs of the host pipeline
    void *ring[GPUB_RING_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ring_ev[GPUB_RING_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t aux[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> events;     // cached, timing disabled (host pipeline)
    std::mutex io_mu;                    // one host transfer at a time per context
};

#define GPUB_CUDA(expr)                              \
    do {                                             \
        cudaError_t gpub_e_ = (expr);                \
        if (gpub_e_ != cudaSuccess) return (int) gpub_e_; \
    } while (0)

#define GPUB_LAUNCH_CHECK() GPUB_CUDA(cudaGetLastError())

// Makes ctx->device current for the lifetime of the guard (one-process-per-GPU callers never switch).
struct gpub_device_guard {
    int prev = -1;
    bool switched = false;
    explicit gpub_device_guard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            switched = (cudaSetDevice(dev) == cudaSuccess);
        }
    }
    ~gpub_device_guard() {
        if (switched) cudaSetDevice(prev);
    }
};

// Resolves (ctx, sidx) to a usable slot, creating streams lazily. Returns nullptr on bad arguments.
gpub_stream_slot *gpub_slot(gpub_ctx_t ctx, int sidx, int *err);

// Context-owned, grow-only scratch of the slot (at least `bytes`, 256-byte aligned), or nullptr when bytes exceeds
// GPUB_BIG_MAX_BYTES (callers then fall back to a stream-ordered allocation). Growing synchronises the slot's stream.
#define GPUB_BIG_MAX_BYTES (256ull << 20)
void *gpub_slot_big(gpub_stream_slot *slot, size_t bytes);

// pageable / pinned host <-> device copies of any size on `stream` (ctx.cu): pinned host memory is DMA'd directly, pageable
// memory goes through the context's pinned ring in GPUB_RING_CHUNK_BYTES pieces staged by several host threads.
// H2D returns once the last piece has been QUEUED (the host buffer may be reused); D2H returns once the data is in `dst`.
int gpub_h2d(gpub_ctx_t ctx, cudaStream_t stream, void *dst_dev, const void *src_host, size_t bytes);
int gpub_d2h(gpub_ctx_t ctx, cudaStream_t stream, void *dst_host, const void *src_dev, size_t bytes);
// the two private non-blocking streams of the host pipeline and n cached events
int gpub_ctx_aux(gpub_ctx_t ctx, cudaStream_t *up, cudaStream_t *down, size_t n_events, cudaEvent_t **events);
// a private stream to run an independent kernel beside the caller's stream, and two fresh events (timing disabled) for the fork
// and the join; the caller destroys the events once it has queued the waits (CUDA defers the release until they have completed)
int gpub_ctx_fork(gpub_ctx_t ctx, cudaStream_t *side, cudaEvent_t ev[2]);
int gpub_internal_nullspace_pack_f64(gpub_ctx_t ctx, cudaStream_t stream, size_t n, const double *U, size_t sU, const unsigned int *rank, double *N, size_t sN, size_t batch);
int gpub_internal_nullspace_pack_f32(gpub_ctx_t ctx, cudaStream_t stream, size_t n, const float *U, size_t sU, const unsigned int *rank, float *N, size_t sN, size_t batch);

#define GPUB_ENTER(ctx, sidx)                              \
    if (!(ctx)) return GPUB_EINVAL;                        \
    gpub_device_guard gpub_guard_((ctx)->device);          \
    int gpub_slot_err_ = 0;                                \
    gpub_stream_slot *slot = gpub_slot((ctx), (sidx), &gpub_slot_err_); \
    if (!slot) return gpub_slot_err_;                      \
    cudaStream_t stream = slot->stream;                    \
    (void) stream

template<typename T>
__device__ __forceinline__ T gpub_abs(T x) { return x < T(0) ? -x : x; }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// 64-bit mix (splitmix64 finaliser): the counter-based generator of gpub_fill_uniform_*.
__host__ __device__ __forceinline__ uint64_t gpub_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// uniform in [0,1) with 53 bits, identical on host and device (the oracle mirrors it in numpy)
__host__ __device__ __forceinline__ double gpub_u01(uint64_t seed, uint64_t i) {
    uint64_t h = gpub_mix64(seed ^ gpub_mix64(i));
    return (double) (h >> 11) * (1.0 / 9007199254740992.0);
}

// v[0..P-1] per lane -> v[0] = sum over the 32 lanes of value number lane / (32 / P)   (P a power of two <= 32)
template<typename T, int P, int O>
struct TReduce {
    static __device__ __forceinline__ void run(T *v, int lane) {
        if constexpr (P > 1) {
            constexpr int H = P / 2;
            const bool up = (lane & O) != 0;
#pragma unroll
            for (int k = 0; k < H; k++) {
                const T keep = up ? v[H + k] : v[k];
                const T send = up ? v[k] : v[H + k];
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, O);
            }
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], O);
        }
        if constexpr (O > 1) TReduce<T, (P > 1 ? P / 2 : 1), O / 2>::run(v, lane);
    }
};

static inline size_t gpub_ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }
