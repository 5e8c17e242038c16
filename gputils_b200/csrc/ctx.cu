// Per-device stream context: the B200-native replacement for the reference's Session singleton of
// cuBLAS/cuSOLVER handles (ref: tensor.cuh:133-247). One context per device, N streams per
// context, each stream with its own reduction scratch and a mapped pinned result slot.
#include "common.cuh"

#include <algorithm>
#include <map>
#include <memory>

int gpub_mem_release(gpub_ctx_t ctx);   // mem.cu: pinned ring, aux streams, cached events, memory pool

namespace {
std::mutex g_registry_mu;
std::map<int, std::unique_ptr<gpub_ctx>> g_registry;

int init_slot(gpub_stream_slot &s, cudaStream_t external, bool internal = false) {
    if (external) {
        s.stream = external;
        s.owned = false;
    } else if (internal) {
        // the library's own side streams: ordered against the caller's stream by events only (a blocking stream would also order
        // against every legacy-stream operation in between, which serialises the chunks when the caller's stream IS the legacy stream)
        GPUB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        s.owned = true;
    } else {
        // blocking stream, like the reference's cudaStreamCreate (tensor.cuh:161): legacy
        // default-stream work and synchronous cudaMemcpy order against it.
        GPUB_CUDA(cudaStreamCreate(&s.stream));
        s.owned = true;
    }
    if (!s.d_scratch) {
        GPUB_CUDA(cudaMalloc(&s.d_scratch, GPUB_SCRATCH_BYTES));
        GPUB_CUDA(cudaMalloc((void **) &s.d_counter, sizeof(unsigned int)));
        GPUB_CUDA(cudaMemset(s.d_counter, 0, sizeof(unsigned int)));
        GPUB_CUDA(cudaHostAlloc(&s.h_result, GPUB_HOSTBUF_BYTES, cudaHostAllocMapped));
    }
    return GPUB_OK;
}
} // namespace

gpub_stream_slot *gpub_slot(gpub_ctx_t ctx, int sidx, int *err) {
    if (sidx < 0 || sidx >= 4096) {
        *err = GPUB_EINVAL;
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(ctx->mu);
    if ((size_t) sidx >= ctx->slots.size()) ctx->slots.resize(sidx + 1);
    if (sidx < GPUB_INTERNAL_SLOT0 && (size_t) sidx + 1 > ctx->user_slots) ctx->user_slots = (size_t) sidx + 1;
    gpub_stream_slot &s = ctx->slots[sidx];
    if (!s.stream) {
        int e = init_slot(s, nullptr, sidx >= GPUB_INTERNAL_SLOT0);
        if (e != GPUB_OK) {
            *err = e;
            return nullptr;
        }
    }
    return &s;
}

void *gpub_slot_big(gpub_stream_slot *slot, size_t bytes) {
    if (!slot || bytes > GPUB_BIG_MAX_BYTES) return nullptr;
    if (bytes <= slot->big_bytes) return slot->d_big;
    if (slot->d_big) {
        if (cudaStreamSynchronize(slot->stream) != cudaSuccess) return nullptr;
        cudaFree(slot->d_big);
        slot->d_big = nullptr;
        slot->big_bytes = 0;
    }
    size_t want = bytes + bytes / 2;
    want = (want + 255) & ~(size_t) 255;
    if (want > GPUB_BIG_MAX_BYTES) want = GPUB_BIG_MAX_BYTES;
    if (cudaMalloc(&slot->d_big, want) != cudaSuccess) {
        slot->d_big = nullptr;
        return nullptr;
    }
    slot->big_bytes = want;
    return slot->d_big;
}

extern "C" {

const char *gpub_version(void) { return "gputils_b200 0.1 (sm_100a)"; }

int gpub_ctx_get(int device, gpub_ctx_t *out) {
    if (!out || device < 0) return GPUB_EINVAL;
    std::lock_guard<std::mutex> lock(g_registry_mu);
    auto it = g_registry.find(device);
    if (it == g_registry.end()) {
        int count = 0;
        GPUB_CUDA(cudaGetDeviceCount(&count));
        if (device >= count) return GPUB_EINVAL;
        auto ctx = std::make_unique<gpub_ctx>();
        ctx->device = device;
        GPUB_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
        GPUB_CUDA(cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        it = g_registry.emplace(device, std::move(ctx)).first;
    }
    *out = it->second.get();
    return GPUB_OK;
}

int gpub_ctx_ensure_streams(gpub_ctx_t ctx, int n) {
    if (!ctx || n < 1) return GPUB_EINVAL;
    gpub_device_guard guard(ctx->device);
    for (int i = 0; i < n; i++) {
        int err = 0;
        if (!gpub_slot(ctx, i, &err)) return err;
    }
    return GPUB_OK;
}

int gpub_ctx_num_streams(gpub_ctx_t ctx) {
    if (!ctx) return 0;
    std::lock_guard<std::mutex> lock(ctx->mu);
    return (int) ctx->user_slots;
}

int gpub_ctx_stream(gpub_ctx_t ctx, int sidx, void **cuda_stream) {
    if (!cuda_stream) return GPUB_EINVAL;
    GPUB_ENTER(ctx, sidx);
    *cuda_stream = (void *) stream;
    return GPUB_OK;
}

int gpub_ctx_bind_stream(gpub_ctx_t ctx, int sidx, void *cuda_stream) {
    if (!ctx || sidx < 0 || sidx >= 4096) return GPUB_EINVAL;
    gpub_device_guard guard(ctx->device);
    std::lock_guard<std::mutex> lock(ctx->mu);
    if ((size_t) sidx >= ctx->slots.size()) ctx->slots.resize(sidx + 1);
    if (sidx < GPUB_INTERNAL_SLOT0 && (size_t) sidx + 1 > ctx->user_slots) ctx->user_slots = (size_t) sidx + 1;
    gpub_stream_slot &s = ctx->slots[sidx];
    if (s.stream && s.owned) {
        GPUB_CUDA(cudaStreamSynchronize(s.stream));
        GPUB_CUDA(cudaStreamDestroy(s.stream));
    }
    s.stream = nullptr;
    // the legacy default stream is a valid binding: represent it with cudaStreamLegacy
    cudaStream_t ext = cuda_stream ? (cudaStream_t) cuda_stream : cudaStreamLegacy;
    return init_slot(s, ext);
}

int gpub_ctx_sync(gpub_ctx_t ctx, int sidx) {
    GPUB_ENTER(ctx, sidx);
    GPUB_CUDA(cudaStreamSynchronize(stream));
    return GPUB_OK;
}

int gpub_ctx_sync_all(gpub_ctx_t ctx) {
    if (!ctx) return GPUB_EINVAL;
    gpub_device_guard guard(ctx->device);
    std::vector<cudaStream_t> streams;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        for (auto &s: ctx->slots)
            if (s.stream) streams.push_back(s.stream);
    }
    for (auto st: streams) GPUB_CUDA(cudaStreamSynchronize(st));
    return GPUB_OK;
}

int gpub_ctx_release(gpub_ctx_t ctx) {
    if (!ctx) return GPUB_EINVAL;
    gpub_device_guard guard(ctx->device);
    std::lock_guard<std::mutex> lock(ctx->mu);
    int first_err = GPUB_OK;
    auto keep = [&](cudaError_t e) { if (e != cudaSuccess && first_err == GPUB_OK) first_err = (int) e; };
    for (auto &s: ctx->slots) {
        if (s.stream) keep(cudaStreamSynchronize(s.stream));
        if (s.d_scratch) keep(cudaFree(s.d_scratch));
        if (s.d_counter) keep(cudaFree(s.d_counter));
        if (s.h_result) keep(cudaFreeHost(s.h_result));
        if (s.d_big) keep(cudaFree(s.d_big));
        if (s.stream && s.owned) keep(cudaStreamDestroy(s.stream));
        s = gpub_stream_slot();   // recreated lazily if the slot is used again
    }
    keep((cudaError_t) std::max(gpub_mem_release(ctx), 0));
    return first_err;
}

int gpub_ctx_release_all(void) {
    std::vector<gpub_ctx_t> all;
    {
        std::lock_guard<std::mutex> lock(g_registry_mu);
        for (auto &kv: g_registry) all.push_back(kv.second.get());
    }
    int first_err = GPUB_OK;
    for (gpub_ctx_t c: all) {
        int e = gpub_ctx_release(c);
        if (e != GPUB_OK && first_err == GPUB_OK) first_err = e;
    }
    return first_err;
}

int gpub_ctx_device(gpub_ctx_t ctx) { return ctx ? ctx->device : -1; }

int gpub_ctx_sm_count(gpub_ctx_t ctx) { return ctx ? ctx->sm_count : 0; }

} // extern "C"
