// Multi-GPU plumbing for the sharded mats axis: one process drives the G GPUs of a box, every device owns a
// contiguous block of matrices of every operand and runs the unchanged single-GPU launchers on it. Nothing here touches the
// data path of a batched op (it shards embarrassingly); these entry points only
//   * make the devices peers over NVLink / NVSwitch (gpub_multi_enable_peer_access),
//   * all-gather result shards to every device: NCCL (ncclAllGather for equal shards, grouped ncclBroadcast for ragged ones),
//     loaded lazily with dlopen so that libgputils_b200.so has no link-time dependency on it, or plain peer copies
//     (cudaMemcpyPeerAsync over NVLink) when the same device appears twice in the list or NCCL is not present.
// The reference has no multi-GPU path at all (tensor.cuh:133-247 binds one device); this is the additive type north_star asks for.
#include "common.cuh"

#include <dlfcn.h>
#include <map>
#include <set>
#include <string>

int gpub_mem_pool_allow_peer(gpub_ctx_t ctx, int peer_device);   // mem.cu

namespace {

// the handful of NCCL entry points used, declared locally so that no NCCL header is needed at build time
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;   // ncclSuccess == 0
enum { GPUB_NCCL_CHAR = 0 };  // ncclInt8 / ncclChar
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    bool ok = false;
};

std::mutex g_multi_mu;
NcclApi g_nccl;
bool g_nccl_tried = false;
std::map<std::string, std::vector<ncclComm_t>> g_comms;   // one communicator clique per device list

NcclApi &nccl_api() {
    if (g_nccl_tried) return g_nccl;
    g_nccl_tried = true;
    // "libnccl.so.2" resolves to an already loaded copy with that soname (e.g. the one bundled with torch) before the system one
    const char *names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    for (const char *nm: names) {
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return g_nccl;
    auto sym = [&](const char *s) { return dlsym(g_nccl.handle, s); };
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll)) sym("ncclCommInitAll");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy)) sym("ncclCommDestroy");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart)) sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd)) sym("ncclGroupEnd");
    g_nccl.AllGather = (decltype(g_nccl.AllGather)) sym("ncclAllGather");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast)) sym("ncclBroadcast");
    g_nccl.GetVersion = (decltype(g_nccl.GetVersion)) sym("ncclGetVersion");
    g_nccl.ok = g_nccl.CommInitAll && g_nccl.CommDestroy && g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.AllGather && g_nccl.Broadcast;
    return g_nccl;
}

bool distinct_devices(const gpub_ctx_t *ctxs, int n) {
    std::set<int> seen;
    for (int i = 0; i < n; i++)
        if (!seen.insert(ctxs[i]->device).second) return false;
    return true;
}

// returns the communicators of the clique (rank g <-> ctxs[g]), creating them on first use; nullptr if NCCL cannot serve the list
std::vector<ncclComm_t> *clique(const gpub_ctx_t *ctxs, int n) {
    NcclApi &api = nccl_api();
    if (!api.ok || !distinct_devices(ctxs, n)) return nullptr;
    std::string key;
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) {
        devs[i] = ctxs[i]->device;
        key += std::to_string(devs[i]) + ",";
    }
    auto it = g_comms.find(key);
    if (it != g_comms.end()) return &it->second;
    std::vector<ncclComm_t> comms(n, nullptr);
    if (api.CommInitAll(comms.data(), n, devs.data()) != 0) return nullptr;
    return &g_comms.emplace(key, std::move(comms)).first->second;
}

int gather_p2p(const gpub_ctx_t *ctxs, int n, int sidx, const void *const *send, const size_t *bytes, void *const *recv) {
    // every destination device pulls the n shards into its receive buffer on its own stream; a source shard must be complete
    // first, so the destination stream waits on an event recorded on the source stream
    struct Events {                                     // destroyed on every exit path (destruction is deferred by the
        std::vector<cudaEvent_t> v;                     // runtime until the event has completed)
        explicit Events(int n) : v(n, nullptr) {}
        ~Events() { for (auto e: v) if (e) cudaEventDestroy(e); }
        cudaEvent_t &operator[](int i) { return v[i]; }
    };
    Events ready(n), done(n);
    std::vector<cudaStream_t> streams(n, nullptr);
    for (int g = 0; g < n; g++) {
        gpub_device_guard guard(ctxs[g]->device);
        int err = 0;
        gpub_stream_slot *slot = gpub_slot(ctxs[g], sidx, &err);
        if (!slot) return err;
        streams[g] = slot->stream;
        GPUB_CUDA(cudaEventCreateWithFlags(&ready[g], cudaEventDisableTiming));
        GPUB_CUDA(cudaEventRecord(ready[g], streams[g]));
    }
    int rc = GPUB_OK;
    for (int g = 0; g < n && rc == GPUB_OK; g++) {
        gpub_device_guard guard(ctxs[g]->device);
        size_t off = 0;
        for (int r = 0; r < n; r++) {
            if (bytes[r]) {
                cudaError_t e = cudaStreamWaitEvent(streams[g], ready[r], 0);
                if (e == cudaSuccess)
                    e = cudaMemcpyPeerAsync((char *) recv[g] + off, ctxs[g]->device, send[r], ctxs[r]->device, bytes[r], streams[g]);
                if (e != cudaSuccess) { rc = (int) e; break; }
            }
            off += bytes[r];
        }
    }
    // a source buffer may be reused by its owner as soon as every destination has read it: the owner's stream waits for all
    // the copies (events recorded after the copies on the destination streams)
    for (int g = 0; g < n && rc == GPUB_OK; g++) {
        gpub_device_guard guard(ctxs[g]->device);
        cudaError_t e = cudaEventCreateWithFlags(&done[g], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(done[g], streams[g]);
        if (e != cudaSuccess) rc = (int) e;
    }
    for (int g = 0; g < n && rc == GPUB_OK; g++) {
        gpub_device_guard guard(ctxs[g]->device);
        for (int r = 0; r < n; r++)
            if (r != g && done[r]) cudaStreamWaitEvent(streams[g], done[r], 0);
    }
    return rc;
}

} // namespace

extern "C" {

int gpub_multi_device_count(int *count) {
    if (!count) return GPUB_EINVAL;
    GPUB_CUDA(cudaGetDeviceCount(count));
    return GPUB_OK;
}

int gpub_multi_enable_peer_access(const int *devices, int n, int *n_pairs_enabled) {
    if (!devices || n < 1) return GPUB_EINVAL;
    int pairs = 0;
    for (int i = 0; i < n; i++) {
        gpub_device_guard guard(devices[i]);
        for (int j = 0; j < n; j++) {
            if (devices[i] == devices[j]) continue;
            int can = 0;
            GPUB_CUDA(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
            if (!can) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                e = cudaSuccess;
            }
            GPUB_CUDA(e);
            // tensors come out of the context's memory pool: pool memory needs its own grant for the peer
            gpub_ctx_t peer_ctx = nullptr;
            int pe = gpub_ctx_get(devices[j], &peer_ctx);
            if (pe) return pe;
            pe = gpub_mem_pool_allow_peer(peer_ctx, devices[i]);
            if (pe) return pe;
            pairs++;
        }
    }
    if (n_pairs_enabled) *n_pairs_enabled = pairs;
    return GPUB_OK;
}

int gpub_multi_nccl_version(int *version) {
    if (!version) return GPUB_EINVAL;
    std::lock_guard<std::mutex> lock(g_multi_mu);
    NcclApi &api = nccl_api();
    *version = 0;
    if (api.ok && api.GetVersion) api.GetVersion(version);
    return GPUB_OK;
}

int gpub_multi_allgather(const gpub_ctx_t *ctxs, int n, int sidx, const void *const *send, const size_t *bytes, void *const *recv,
                         int transport, int *transport_used) {
    if (!ctxs || !send || !bytes || !recv || n < 1 || transport < 0 || transport > 2) return GPUB_EINVAL;
    for (int g = 0; g < n; g++)
        if (!ctxs[g] || !recv[g] || (bytes[g] && !send[g])) return GPUB_EINVAL;
    std::lock_guard<std::mutex> lock(g_multi_mu);
    std::vector<ncclComm_t> *comms = (transport == GPUB_GATHER_P2P) ? nullptr : clique(ctxs, n);
    if (!comms) {
        if (transport == GPUB_GATHER_NCCL) return GPUB_ENOTSUP;
        if (transport_used) *transport_used = GPUB_GATHER_P2P;
        return gather_p2p(ctxs, n, sidx, send, bytes, recv);
    }
    if (transport_used) *transport_used = GPUB_GATHER_NCCL;
    NcclApi &api = nccl_api();
    std::vector<cudaStream_t> streams(n);
    for (int g = 0; g < n; g++) {
        gpub_device_guard guard(ctxs[g]->device);
        int err = 0;
        gpub_stream_slot *slot = gpub_slot(ctxs[g], sidx, &err);
        if (!slot) return err;
        streams[g] = slot->stream;
    }
    bool equal = true;
    for (int g = 1; g < n; g++) equal = equal && bytes[g] == bytes[0];
    int rc = 0;
    if (api.GroupStart() != 0) return GPUB_ENOTSUP;
    if (equal) {
        for (int g = 0; g < n && !rc; g++) rc = api.AllGather(send[g], recv[g], bytes[0], GPUB_NCCL_CHAR, (*comms)[g], streams[g]);
    } else {
        size_t off = 0;
        for (int r = 0; r < n && !rc; r++) {          // ragged shards: one broadcast per owner, all inside one group
            if (bytes[r])
                for (int g = 0; g < n && !rc; g++)
                    rc = api.Broadcast(g == r ? send[r] : nullptr, (char *) recv[g] + off, bytes[r], GPUB_NCCL_CHAR, r, (*comms)[g], streams[g]);
            off += bytes[r];
        }
    }
    int rc_end = api.GroupEnd();
    return (rc || rc_end) ? GPUB_ENOTSUP : GPUB_OK;
}

int gpub_multi_release(void) {
    std::lock_guard<std::mutex> lock(g_multi_mu);
    if (g_comms.empty()) return GPUB_OK;
    NcclApi &api = nccl_api();
    for (auto &kv: g_comms)
        for (ncclComm_t c: kv.second)
            if (c && api.ok) api.CommDestroy(c);
    g_comms.clear();
    return GPUB_OK;
}

} // extern "C"
