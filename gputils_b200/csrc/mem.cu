// Memory and host-transfer plumbing of the stream context:
//   * gpub_mem_alloc / gpub_mem_free -- the stream-ordered pool behind Session::cudaAllocate. The reference does two cudaMalloc
//     and two cudaFree per DTensor (ref: tensor.cuh:1106-1126, 283-294), plus one more pair per tr() (1169), binary operator
//     (634, 640) and Nullspace loop iteration (2076); after the kernels got fast those calls are the run time of the solver loops
//     that use the library. Here every allocation is a cudaMallocFromPoolAsync on the LEGACY default stream out of a pool that
//     keeps freed blocks cached: the legacy stream orders against every blocking stream of the context (they are created with
//     cudaStreamCreate like the reference's, tensor.cuh:161), so an allocation or a free is a device-side ordering point for all
//     of them -- what cudaMalloc / cudaFree gave the reference -- without stopping the host.
//   * gpub_upload / gpub_download -- DTensor::upload / download (ref: tensor.cuh:1128-1154). Pageable host memory is cut into
//     8 MB pieces that several host threads copy into a ring of pinned buffers while the DMA of the previous piece runs.
//   * gpub_chol_solve_from_host_* -- the host pipeline of CholeskyBatchFactoriser (additive): upload, factorise + solve and
//     download of successive chunks of the batch overlap on three streams.
#include "common.cuh"

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <thread>
#include <unordered_map>

namespace {

// ---- owner registry: gpub_mem_free(ptr) needs the context that handed ptr out (a sharded tensor may be destroyed while another
// device is current)
std::mutex g_owner_mu;
std::unordered_map<void *, gpub_ctx_t> g_owner;

int ensure_pool(gpub_ctx_t ctx) {
    if (ctx->pool_tried) return GPUB_OK;
    ctx->pool_tried = true;
    int supported = 0;
    GPUB_CUDA(cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, ctx->device));
    if (!supported) return GPUB_OK;                       // plain cudaMalloc below
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = ctx->device;
    GPUB_CUDA(cudaMemPoolCreate(&ctx->pool, &props));
    uint64_t keep = UINT64_MAX;                           // freed blocks stay in the pool until gpub_mem_trim / release
    GPUB_CUDA(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    return GPUB_OK;
}

// ---- a few persistent host threads for the pageable <-> pinned staging copies
class CopyCrew {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    struct Piece { char *dst; const char *src; size_t n; };
    std::vector<Piece> pieces;
    size_t next = 0, pending = 0;
    bool stop = false;

    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || next < pieces.size(); });
            if (stop) return;
            Piece p = pieces[next++];
            lk.unlock();
            std::memcpy(p.dst, p.src, p.n);
            lk.lock();
            if (--pending == 0) cv_done.notify_all();
        }
    }

public:
    CopyCrew() {
        unsigned hw = std::thread::hardware_concurrency();
        unsigned n = hw >= 16 ? 6 : (hw >= 8 ? 3 : 1);
        for (unsigned i = 0; i < n; i++) workers.emplace_back([this] { loop(); });
    }
    ~CopyCrew() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &t: workers) t.join();
    }
    // the calling thread takes a share too
    void copy(void *dst, const void *src, size_t n) {
        const size_t parts = workers.size() + 1;
        if (n < (1u << 20) || parts == 1) {
            std::memcpy(dst, src, n);
            return;
        }
        const size_t each = ((n / parts) + 4095) & ~(size_t) 4095;
        std::unique_lock<std::mutex> lk(mu);
        pieces.clear();
        next = 0;
        size_t off = each;                                 // piece 0 is the caller's
        while (off < n) {
            pieces.push_back({(char *) dst + off, (const char *) src + off, std::min(each, n - off)});
            off += each;
        }
        pending = pieces.size();
        lk.unlock();
        cv_work.notify_all();
        std::memcpy(dst, src, std::min(each, n));
        lk.lock();
        cv_done.wait(lk, [&] { return pending == 0; });
        pieces.clear();
        next = 0;
    }
};

CopyCrew &crew() {
    static CopyCrew c;
    return c;
}

int ensure_ring(gpub_ctx_t ctx) {
    for (int i = 0; i < GPUB_RING_SLOTS; i++) {
        if (!ctx->ring[i]) GPUB_CUDA(cudaHostAlloc(&ctx->ring[i], GPUB_RING_CHUNK_BYTES, cudaHostAllocDefault));
        if (!ctx->ring_ev[i]) GPUB_CUDA(cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming));
    }
    return GPUB_OK;
}

bool host_is_pinned(const void *p) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

} // namespace

int gpub_h2d(gpub_ctx_t ctx, cudaStream_t stream, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return GPUB_OK;
    if (bytes <= (256u << 10) || host_is_pinned(src)) {
        GPUB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return GPUB_OK;
    }
    std::lock_guard<std::mutex> lock(ctx->io_mu);
    int e = ensure_ring(ctx);
    if (e) return e;
    size_t off = 0;
    for (int i = 0; off < bytes; i++) {
        const int s = i % GPUB_RING_SLOTS;
        const size_t n = std::min((size_t) GPUB_RING_CHUNK_BYTES, bytes - off);
        GPUB_CUDA(cudaEventSynchronize(ctx->ring_ev[s]));                 // the DMA that last read this buffer is done
        crew().copy(ctx->ring[s], (const char *) src + off, n);
        GPUB_CUDA(cudaMemcpyAsync((char *) dst + off, ctx->ring[s], n, cudaMemcpyHostToDevice, stream));
        GPUB_CUDA(cudaEventRecord(ctx->ring_ev[s], stream));
        off += n;
    }
    return GPUB_OK;
}

int gpub_d2h(gpub_ctx_t ctx, cudaStream_t stream, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return GPUB_OK;
    if (bytes <= (256u << 10) || host_is_pinned(dst)) {
        GPUB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
        GPUB_CUDA(cudaStreamSynchronize(stream));
        return GPUB_OK;
    }
    std::lock_guard<std::mutex> lock(ctx->io_mu);
    int e = ensure_ring(ctx);
    if (e) return e;
    const size_t chunk = GPUB_RING_CHUNK_BYTES;
    const size_t count = (bytes + chunk - 1) / chunk;
    // the DMA runs GPUB_RING_SLOTS - 1 pieces ahead of the host threads that empty the ring
    for (size_t i = 0; i < count + GPUB_RING_SLOTS - 1; i++) {
        if (i < count) {
            const int s = (int) (i % GPUB_RING_SLOTS);
            const size_t off = i * chunk, n = std::min(chunk, bytes - off);
            GPUB_CUDA(cudaMemcpyAsync(ctx->ring[s], (const char *) src + off, n, cudaMemcpyDeviceToHost, stream));
            GPUB_CUDA(cudaEventRecord(ctx->ring_ev[s], stream));
        }
        if (i + 1 >= GPUB_RING_SLOTS) {
            const size_t j = i + 1 - GPUB_RING_SLOTS;
            const int s = (int) (j % GPUB_RING_SLOTS);
            const size_t off = j * chunk, n = std::min(chunk, bytes - off);
            GPUB_CUDA(cudaEventSynchronize(ctx->ring_ev[s]));
            crew().copy((char *) dst + off, ctx->ring[s], n);
        }
    }
    return GPUB_OK;
}

int gpub_ctx_fork(gpub_ctx_t ctx, cudaStream_t *side, cudaEvent_t ev[2]) {
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        if (!ctx->aux[0]) GPUB_CUDA(cudaStreamCreateWithFlags(&ctx->aux[0], cudaStreamNonBlocking));
        *side = ctx->aux[0];
    }
    GPUB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    cudaError_t e = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
    if (e != cudaSuccess) { cudaEventDestroy(ev[0]); return (int) e; }
    return GPUB_OK;
}

int gpub_ctx_aux(gpub_ctx_t ctx, cudaStream_t *up, cudaStream_t *down, size_t n_events, cudaEvent_t **events) {
    for (int i = 0; i < 2; i++)
        if (!ctx->aux[i]) GPUB_CUDA(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
    while (ctx->events.size() < n_events) {
        cudaEvent_t ev;
        GPUB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->events.push_back(ev);
    }
    *up = ctx->aux[0];
    *down = ctx->aux[1];
    *events = ctx->events.data();
    return GPUB_OK;
}

// called by gpub_ctx_release (ctx.cu) with the device current and every stream idle
int gpub_mem_release(gpub_ctx_t ctx) {
    int first = GPUB_OK;
    auto keep = [&](cudaError_t e) { if (e != cudaSuccess && first == GPUB_OK) first = (int) e; };
    for (int i = 0; i < GPUB_RING_SLOTS; i++) {
        if (ctx->ring_ev[i]) keep(cudaEventDestroy(ctx->ring_ev[i]));
        if (ctx->ring[i]) keep(cudaFreeHost(ctx->ring[i]));
        ctx->ring_ev[i] = nullptr;
        ctx->ring[i] = nullptr;
    }
    for (auto ev: ctx->events) keep(cudaEventDestroy(ev));
    ctx->events.clear();
    for (int i = 0; i < 2; i++) {
        if (ctx->aux[i]) {
            keep(cudaStreamSynchronize(ctx->aux[i]));
            keep(cudaStreamDestroy(ctx->aux[i]));
        }
        ctx->aux[i] = nullptr;
    }
    if (ctx->pool) {
        // blocks still held by live tensors keep the pool alive inside the runtime; cached blocks go back to the driver
        keep(cudaStreamSynchronize(cudaStreamLegacy));
        keep(cudaMemPoolTrimTo(ctx->pool, 0));
        bool live = false;
        {
            std::lock_guard<std::mutex> lk(g_owner_mu);
            for (auto &kv: g_owner) live = live || kv.second == ctx;
        }
        if (!live) {
            keep(cudaMemPoolDestroy(ctx->pool));
            ctx->pool = nullptr;
            ctx->pool_tried = false;
        }
    }
    return first;
}

// peers of a sharded tensor read each other's pool memory (cudaMemcpyPeerAsync pulls, NCCL): called by
// gpub_multi_enable_peer_access (multi.cu)
int gpub_mem_pool_allow_peer(gpub_ctx_t ctx, int peer_device) {
    gpub_device_guard guard(ctx->device);
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        int e = ensure_pool(ctx);
        if (e) return e;
    }
    if (!ctx->pool) return GPUB_OK;
    cudaMemAccessDesc desc{};
    desc.location.type = cudaMemLocationTypeDevice;
    desc.location.id = peer_device;
    desc.flags = cudaMemAccessFlagsProtReadWrite;
    GPUB_CUDA(cudaMemPoolSetAccess(ctx->pool, &desc, 1));
    return GPUB_OK;
}

extern "C" {

int gpub_mem_alloc(gpub_ctx_t ctx, size_t bytes, void **ptr) {
    if (!ctx || !ptr) return GPUB_EINVAL;
    *ptr = nullptr;
    if (bytes == 0) return GPUB_OK;
    gpub_device_guard guard(ctx->device);
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        int e = ensure_pool(ctx);
        if (e) return e;
    }
    if (ctx->pool) {
        GPUB_CUDA(cudaMallocFromPoolAsync(ptr, bytes, ctx->pool, cudaStreamLegacy));
    } else {
        GPUB_CUDA(cudaMalloc(ptr, bytes));
    }
    std::lock_guard<std::mutex> lk(g_owner_mu);
    g_owner[*ptr] = ctx;
    return GPUB_OK;
}

int gpub_mem_free(void *ptr) {
    if (!ptr) return GPUB_OK;
    gpub_ctx_t ctx = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_owner_mu);
        auto it = g_owner.find(ptr);
        if (it == g_owner.end()) return GPUB_EINVAL;
        ctx = it->second;
        g_owner.erase(it);
    }
    gpub_device_guard guard(ctx->device);
    if (ctx->pool) {
        GPUB_CUDA(cudaFreeAsync(ptr, cudaStreamLegacy));
    } else {
        GPUB_CUDA(cudaFree(ptr));
    }
    return GPUB_OK;
}

int gpub_mem_stats(gpub_ctx_t ctx, size_t *reserved_bytes, size_t *used_bytes) {
    if (!ctx) return GPUB_EINVAL;
    uint64_t r = 0, u = 0;
    if (ctx->pool) {
        gpub_device_guard guard(ctx->device);
        GPUB_CUDA(cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &r));
        GPUB_CUDA(cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrUsedMemCurrent, &u));
    }
    if (reserved_bytes) *reserved_bytes = (size_t) r;
    if (used_bytes) *used_bytes = (size_t) u;
    return GPUB_OK;
}

int gpub_mem_trim(gpub_ctx_t ctx, size_t keep_bytes) {
    if (!ctx) return GPUB_EINVAL;
    if (!ctx->pool) return GPUB_OK;
    gpub_device_guard guard(ctx->device);
    GPUB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    GPUB_CUDA(cudaMemPoolTrimTo(ctx->pool, keep_bytes));
    return GPUB_OK;
}

int gpub_upload(gpub_ctx_t ctx, int sidx, void *dst_dev, const void *src_host, size_t bytes) {
    if (bytes && (!dst_dev || !src_host)) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    int e = gpub_h2d(ctx, stream, dst_dev, src_host, bytes);
    if (e) return e;
    // like the cudaMemcpy it replaces: when the call returns the data is on the device, whichever stream reads it next
    GPUB_CUDA(cudaStreamSynchronize(stream));
    return GPUB_OK;
}

int gpub_download(gpub_ctx_t ctx, int sidx, void *dst_host, const void *src_dev, size_t bytes) {
    if (bytes && (!dst_host || !src_dev)) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    return gpub_d2h(ctx, stream, dst_host, src_dev, bytes);
}

} // extern "C"

// ---- host pipeline of the batched Cholesky solve ---------------------------------------------------------------------------
namespace {

template<typename T> int potrf_l(gpub_ctx_t, int, size_t, T *, size_t, size_t, int *, size_t);
template<> int potrf_l<double>(gpub_ctx_t c, int s, size_t n, double *A, size_t lda, size_t sA, int *info, size_t b) { return gpub_potrf_batched_f64(c, s, n, A, lda, sA, info, b); }
template<> int potrf_l<float>(gpub_ctx_t c, int s, size_t n, float *A, size_t lda, size_t sA, int *info, size_t b) { return gpub_potrf_batched_f32(c, s, n, A, lda, sA, info, b); }
template<typename T> int potrs_l(gpub_ctx_t, int, size_t, const T *, size_t, size_t, T *, size_t, size_t);
template<> int potrs_l<double>(gpub_ctx_t c, int s, size_t n, const double *L, size_t ldl, size_t sL, double *b, size_t sB, size_t k) { return gpub_potrs_batched_f64(c, s, n, L, ldl, sL, b, sB, k); }
template<> int potrs_l<float>(gpub_ctx_t c, int s, size_t n, const float *L, size_t ldl, size_t sL, float *b, size_t sB, size_t k) { return gpub_potrs_batched_f32(c, s, n, L, ldl, sL, b, sB, k); }

template<typename T>
int chol_solve_from_host(gpub_ctx_t ctx, int sidx, size_t n, T *A, T *b, int *info, const T *hA, const T *hb, T *hx, int *hinfo,
                         size_t batch, size_t chunks, bool lower_only) {
    if (!ctx || !A || !hA || (b && !hb) || (hx && !b)) return GPUB_EINVAL;
    if (n == 0 || batch == 0) return GPUB_OK;
    GPUB_ENTER(ctx, sidx);
    if (chunks == 0) chunks = 16;
    chunks = std::min(chunks, batch);
    cudaStream_t up, down;
    cudaEvent_t *ev;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        int e = gpub_ctx_aux(ctx, &up, &down, 2 * chunks + 1, &ev);
        if (e) return e;
    }
    // the device buffers may still be in use by earlier work of the compute stream
    GPUB_CUDA(cudaEventRecord(ev[2 * chunks], stream));
    GPUB_CUDA(cudaStreamWaitEvent(up, ev[2 * chunks], 0));
    GPUB_CUDA(cudaStreamWaitEvent(down, ev[2 * chunks], 0));
    const bool pinned_out = (!hx || host_is_pinned(hx)) && (!hinfo || host_is_pinned(hinfo));
    // pinned input, rows of the half-height strip at least 128 bytes (narrower strided rows cost the copy engine more than they
    // save: measured with scripts/microbench/strip_copy.cu, 32 x 32 fp64: 75 % of the bytes in 0.84 x the time of the dense copy)
    const size_t h = n / 2;
    const bool tri = lower_only && (n % 2 == 0) && h * sizeof(T) >= 128 && host_is_pinned(hA);
    for (size_t c = 0; c < chunks; c++) {
        const size_t lo = c * batch / chunks, hi = (c + 1) * batch / chunks, cnt = hi - lo;
        if (cnt == 0) continue;
        int e = GPUB_OK;
        if (tri) {
            // potrf reads the lower triangle only: the left half of the columns goes as one contiguous block per matrix, of the right
            // half only rows n/2 .. n-1 (a 3-D copy with rows of n/2 elements); the block above the diagonal stays what it was
            GPUB_CUDA(cudaMemcpy2DAsync(A + lo * n * n, n * n * sizeof(T), hA + lo * n * n, n * n * sizeof(T), n * h * sizeof(T), cnt,
                                        cudaMemcpyHostToDevice, up));
            cudaMemcpy3DParms p3 = {};
            const size_t off = h + h * n;                  // element (n/2, n/2) of the first matrix of the chunk
            p3.srcPtr = make_cudaPitchedPtr((void *) (hA + lo * n * n + off), n * sizeof(T), n, n);
            p3.dstPtr = make_cudaPitchedPtr((void *) (A + lo * n * n + off), n * sizeof(T), n, n);
            p3.extent = make_cudaExtent(h * sizeof(T), h, cnt);
            p3.kind = cudaMemcpyHostToDevice;
            GPUB_CUDA(cudaMemcpy3DAsync(&p3, up));
        } else {
            e = gpub_h2d(ctx, up, A + lo * n * n, hA + lo * n * n, cnt * n * n * sizeof(T));
        }
        if (!e && b) e = gpub_h2d(ctx, up, b + lo * n, hb + lo * n, cnt * n * sizeof(T));
        if (e) return e;
        GPUB_CUDA(cudaEventRecord(ev[2 * c], up));
        GPUB_CUDA(cudaStreamWaitEvent(stream, ev[2 * c], 0));
        e = potrf_l<T>(ctx, sidx, n, A + lo * n * n, n, n * n, info ? info + lo : nullptr, cnt);
        if (!e && b) e = potrs_l<T>(ctx, sidx, n, A + lo * n * n, n, n * n, b + lo * n, n, cnt);
        if (e) return e;
        GPUB_CUDA(cudaEventRecord(ev[2 * c + 1], stream));
        GPUB_CUDA(cudaStreamWaitEvent(down, ev[2 * c + 1], 0));
        if (pinned_out) {
            if (hx) GPUB_CUDA(cudaMemcpyAsync(hx + lo * n, b + lo * n, cnt * n * sizeof(T), cudaMemcpyDeviceToHost, down));
            if (hinfo && info) GPUB_CUDA(cudaMemcpyAsync(hinfo + lo, info + lo, cnt * sizeof(int), cudaMemcpyDeviceToHost, down));
        }
    }
    if (!pinned_out) {
        // pageable results: one staged download behind the last chunk (x is 1 / n of the input volume)
        if (hx) { int e = gpub_d2h(ctx, down, hx, b, batch * n * sizeof(T)); if (e) return e; }
        if (hinfo && info) { int e = gpub_d2h(ctx, down, hinfo, info, batch * sizeof(int)); if (e) return e; }
    }
    GPUB_CUDA(cudaStreamSynchronize(down));
    GPUB_CUDA(cudaStreamSynchronize(stream));
    return GPUB_OK;
}

} // namespace

extern "C" {

int gpub_chol_solve_from_host_f64(gpub_ctx_t ctx, int sidx, size_t n, double *A_dev, double *b_dev, int *info_dev, const double *A_host,
                                  const double *b_host, double *x_host, int *info_host, size_t batch, size_t chunks) {
    return chol_solve_from_host<double>(ctx, sidx, n, A_dev, b_dev, info_dev, A_host, b_host, x_host, info_host, batch, chunks & GPUB_CHUNKS_MASK,
                                        (chunks & GPUB_LOWER_ONLY) != 0);
}
int gpub_chol_solve_from_host_f32(gpub_ctx_t ctx, int sidx, size_t n, float *A_dev, float *b_dev, int *info_dev, const float *A_host,
                                  const float *b_host, float *x_host, int *info_host, size_t batch, size_t chunks) {
    return chol_solve_from_host<float>(ctx, sidx, n, A_dev, b_dev, info_dev, A_host, b_host, x_host, info_host, batch, chunks & GPUB_CHUNKS_MASK,
                                       (chunks & GPUB_LOWER_ONLY) != 0);
}

} // extern "C"
