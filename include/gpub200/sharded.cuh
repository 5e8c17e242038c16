/*
 * sharded.cuh -- the mats axis of a DTensor sharded over the GPUs of one box (additive API).
 *
 * The reference binds one device for the life of the process (Session, tensor.cuh:133-247) and has no
 * multi-GPU path. Every batched operation of the hot path is independent per matrix, so sharding is
 * embarrassing: device g of G owns the contiguous block [g*ceil(k/G), min(k,(g+1)*ceil(k/G))) of EVERY
 * operand, as an ordinary DTensor allocated on that device, and the unchanged single-GPU launchers run on
 * it through that device's own stream context. One host thread drives all devices: launches are
 * asynchronous, so the G kernels of one call run concurrently. No collective on the data path; NCCL over
 * NVLink / NVSwitch only all-gathers result shards when the caller wants them on every device
 * (allGather), and the flat reductions combine G per-shard scalars on the host.
 */
#ifndef GPUB200_SHARDED_CUH
#define GPUB200_SHARDED_CUH

#include "core.cuh"
#include "dtensor.cuh"
#include "factorisers.cuh"

#include <algorithm>
#include <cmath>
#include <memory>
#include <utility>
#include <vector>

namespace gpub200 {

/** Makes `device` current for the lifetime of the object. */
class DeviceScope {
    int m_prev = 0;
public:
    explicit DeviceScope(int device) {
        gpuErrChk(cudaGetDevice(&m_prev));
        if (m_prev != device) gpuErrChk(cudaSetDevice(device));
    }

    ~DeviceScope() { cudaSetDevice(m_prev); }

    DeviceScope(const DeviceScope &) = delete;

    DeviceScope &operator=(const DeviceScope &) = delete;
};

/** Devices 0..count-1 of this box. */
inline std::vector<int> allDevices() {
    int count = 0;
    gpuErrChk(gpub_multi_device_count(&count));
    std::vector<int> d(count);
    for (int i = 0; i < count; i++) d[i] = i;
    return d;
}

/** Contiguous block partition of k matrices over G shards: [from, to) of shard g (trailing shards may be short or empty). */
inline std::pair<size_t, size_t> shardRange(size_t k, size_t G, size_t g) {
    if (G == 0 || g >= G) throw std::invalid_argument("[shardRange] bad shard index");
    const size_t per = (k + G - 1) / G;
    const size_t from = std::min(k, g * per);
    return {from, std::min(k, from + per)};
}

} // namespace gpub200

TEMPLATE_WITH_TYPE_T
class ShardedDTensor {
private:
    std::vector<int> m_devices;
    std::vector<std::unique_ptr<DTensor<T> > > m_shards;
    size_t m_numRows = 0, m_numCols = 0, m_numMats = 0;

    void allocate(bool zero) {
        if (m_devices.empty()) m_devices = gpub200::allDevices();
        if (m_devices.empty()) throw std::invalid_argument("[ShardedDTensor] no devices");
        gpuErrChk(gpub_multi_enable_peer_access(m_devices.data(), (int) m_devices.size(), nullptr));
        for (size_t g = 0; g < m_devices.size(); g++) {
            gpub200::DeviceScope scope(m_devices[g]);
            auto [from, to] = shardRange(g);
            m_shards.push_back(std::make_unique<DTensor<T> >(m_numRows, m_numCols, to - from, zero));
        }
    }

public:
    ShardedDTensor() = delete;

    /** (m, n, k)-tensor with its k matrices spread over `devices` (all devices of the box when empty). */
    ShardedDTensor(size_t m, size_t n, size_t k, std::vector<int> devices = {}, bool zero = false)
        : m_devices(std::move(devices)), m_numRows(m), m_numCols(n), m_numMats(k) {
        allocate(zero);
    }

    /** Same, initialised from host data (layout as DTensor: column-major matrices, mats axis slowest). */
    ShardedDTensor(const std::vector<T> &data, size_t m, size_t n, size_t k, std::vector<int> devices = {})
        : m_devices(std::move(devices)), m_numRows(m), m_numCols(n), m_numMats(k) {
        allocate(false);
        upload(data);
    }

    size_t numRows() const { return m_numRows; }

    size_t numCols() const { return m_numCols; }

    size_t numMats() const { return m_numMats; }

    size_t numEl() const { return m_numRows * m_numCols * m_numMats; }

    size_t numShards() const { return m_devices.size(); }

    int device(size_t g) const { return m_devices.at(g); }

    const std::vector<int> &devices() const { return m_devices; }

    std::pair<size_t, size_t> shardRange(size_t g) const { return gpub200::shardRange(m_numMats, m_devices.size(), g); }

    /** The block of matrices device(g) owns: an ordinary DTensor living on that device. */
    DTensor<T> &shard(size_t g) { return *m_shards.at(g); }

    const DTensor<T> &shard(size_t g) const { return *m_shards.at(g); }

    /** f(g, shard) with device(g) current; launches made inside are asynchronous, so shards overlap. */
    template<typename F>
    void forEachShard(F &&f) {
        for (size_t g = 0; g < m_devices.size(); g++) {
            if (m_shards[g]->numMats() == 0) continue;
            gpub200::DeviceScope scope(m_devices[g]);
            f(g, *m_shards[g]);
        }
    }

    template<typename F>
    void forEachShard(F &&f) const {
        for (size_t g = 0; g < m_devices.size(); g++) {
            if (m_shards[g]->numMats() == 0) continue;
            gpub200::DeviceScope scope(m_devices[g]);
            f(g, static_cast<const DTensor<T> &>(*m_shards[g]));
        }
    }

    void checkSameSharding(const ShardedDTensor<T> &o, const char *what) const {
        if (o.m_devices != m_devices || o.m_numMats != m_numMats)
            throw std::invalid_argument(std::string("[ShardedDTensor] ") + what + ": operands are sharded differently");
    }

    /** Waits for everything queued on every device of this tensor. */
    void synchronize() const {
        forEachShard([](size_t, const DTensor<T> &) { gpuErrChk(gpub_ctx_sync_all(gpub200::ctx())); });
    }

    /** Host -> devices; the G copies are queued asynchronously on the owners' streams and run concurrently. */
    void upload(const std::vector<T> &vec) {
        if (vec.size() != numEl()) throw std::invalid_argument("[ShardedDTensor::upload] vec has wrong size");
        const size_t per = m_numRows * m_numCols;
        forEachShard([&](size_t g, DTensor<T> &s) {
            const size_t from = shardRange(g).first;
            gpuErrChk(cudaMemcpyAsync(s.raw(), vec.data() + from * per, s.numEl() * sizeof(T), cudaMemcpyHostToDevice,
                                      Session::getInstance().streamOfCurrentDevice(s.streamIdx())));
        });
        synchronize();
    }

    void download(std::vector<T> &vec) const {
        vec.resize(numEl());
        const size_t per = m_numRows * m_numCols;
        forEachShard([&](size_t g, const DTensor<T> &s) {
            const size_t from = shardRange(g).first;
            gpuErrChk(cudaMemcpyAsync(vec.data() + from * per, s.raw(), s.numEl() * sizeof(T), cudaMemcpyDeviceToHost,
                                      Session::getInstance().streamOfCurrentDevice(s.streamIdx())));
        });
        synchronize();
    }

    /* ---- shard-wise operations: the single-GPU methods, one asynchronous launch per device ---- */

    ShardedDTensor &operator*=(T scalar) {
        forEachShard([&](size_t, DTensor<T> &s) { s *= scalar; });
        return *this;
    }

    ShardedDTensor &operator+=(const ShardedDTensor &rhs) {
        checkSameSharding(rhs, "operator+=");
        forEachShard([&](size_t g, DTensor<T> &s) { s += rhs.shard(g); });
        return *this;
    }

    ShardedDTensor &operator-=(const ShardedDTensor &rhs) {
        checkSameSharding(rhs, "operator-=");
        forEachShard([&](size_t g, DTensor<T> &s) { s -= rhs.shard(g); });
        return *this;
    }

    /** C_i <- beta C_i + alpha A_i B_i on every shard (DTensor::addAB, tensor.cuh:1286-1338). */
    void addAB(const ShardedDTensor &A, const ShardedDTensor &B, T alpha = 1, T beta = 0) {
        checkSameSharding(A, "addAB");
        checkSameSharding(B, "addAB");
        forEachShard([&](size_t g, DTensor<T> &s) { s.addAB(A.shard(g), B.shard(g), alpha, beta); });
    }

    /** Batched least squares on every shard (DTensor::leastSquaresBatched, tensor.cuh:1340-1394). */
    void leastSquaresBatched(ShardedDTensor &b) {
        checkSameSharding(b, "leastSquaresBatched");
        if (b.numRows() != m_numRows) throw std::invalid_argument("[Least squares batched] rhs rows does not equal lhs rows");
        if (b.numCols() != 1) throw std::invalid_argument("[Least squares batched] rhs are not vectors");
        if (m_numCols > m_numRows) throw std::invalid_argument("[Least squares batched] supports square or tall matrices only");
        forEachShard([&](size_t g, DTensor<T> &s) { s.leastSquaresBatched(b.shard(g)); });
    }

    /* ---- flat reductions: per-shard partials combined on the host (G scalars) ---- */

    T normF() const {
        double acc = 0;
        forEachShard([&](size_t, const DTensor<T> &s) {
            const double v = (double) s.normF();
            acc += v * v;
        });
        return (T) std::sqrt(acc);
    }

    T sumAbs() const {
        double acc = 0;
        forEachShard([&](size_t, const DTensor<T> &s) { acc += (double) s.sumAbs(); });
        return (T) acc;
    }

    T dotF(const ShardedDTensor &other) {
        checkSameSharding(other, "dotF");
        double acc = 0;
        forEachShard([&](size_t g, DTensor<T> &s) { acc += (double) s.dotF(other.shard(g)); });
        return (T) acc;
    }

    T maxAbs() const {
        T best = 0;
        forEachShard([&](size_t, const DTensor<T> &s) { best = std::max(best, s.maxAbs()); });
        return best;
    }

    T minAbs() const {
        bool first = true;
        T best = 0;
        forEachShard([&](size_t, const DTensor<T> &s) {
            const T v = s.minAbs();
            best = first ? v : std::min(best, v);
            first = false;
        });
        return best;
    }

    /**
     * The whole (m, n, k) tensor on every device: result[g] lives on device(g).
     * NCCL over NVLink / NVSwitch (ncclAllGather for equal shards, grouped ncclBroadcast for ragged ones), or peer
     * copies -- see gpub_multi_allgather. `transportUsed` (optional) receives GPUB_GATHER_NCCL or GPUB_GATHER_P2P.
     */
    std::vector<std::unique_ptr<DTensor<T> > > allGather(int transport = GPUB_GATHER_AUTO, int *transportUsed = nullptr) const {
        const size_t G = m_devices.size();
        std::vector<std::unique_ptr<DTensor<T> > > full(G);
        std::vector<gpub_ctx_t> ctxs(G);
        std::vector<const void *> send(G);
        std::vector<void *> recv(G);
        std::vector<size_t> bytes(G);
        for (size_t g = 0; g < G; g++) {
            gpub200::DeviceScope scope(m_devices[g]);
            full[g] = std::make_unique<DTensor<T> >(m_numRows, m_numCols, m_numMats);
            ctxs[g] = gpub200::ctx();
            send[g] = m_shards[g]->raw();
            recv[g] = full[g]->raw();
            bytes[g] = m_shards[g]->numEl() * sizeof(T);
            /* the gather runs on stream 0 of every device: work queued on a shard's own stream comes first */
            if (m_shards[g]->streamIdx() != 0) {
                cudaEvent_t ev;
                gpuErrChk(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                gpuErrChk(cudaEventRecord(ev, Session::getInstance().streamOfCurrentDevice(m_shards[g]->streamIdx())));
                gpuErrChk(cudaStreamWaitEvent(Session::getInstance().streamOfCurrentDevice(0), ev, 0));
                gpuErrChk(cudaEventDestroy(ev));
            }
        }
        if (numEl() == 0) return full;
        gpuErrChk(gpub_multi_allgather(ctxs.data(), (int) G, 0, send.data(), bytes.data(), recv.data(), transport, transportUsed));
        for (size_t g = 0; g < G; g++) {
            gpub200::DeviceScope scope(m_devices[g]);
            gpuErrChk(gpub_ctx_sync(ctxs[g], 0));
        }
        return full;
    }
};

/**
 * CholeskyBatchFactoriser over a sharded batch (tensor.cuh:2098-2197 per shard): factorise() and solve() queue one
 * launch per device and return; info(g) is the status tensor of shard g.
 */
TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
class ShardedCholeskyBatchFactoriser {
private:
    ShardedDTensor<T> *m_matrix;
    std::vector<std::unique_ptr<CholeskyBatchFactoriser<T> > > m_factorisers;
    bool m_factorisationDone = false;

public:
    ShardedCholeskyBatchFactoriser() = delete;

    ShardedCholeskyBatchFactoriser(ShardedDTensor<T> &A, bool factorised = false) : m_matrix(&A), m_factorisationDone(factorised) {
        if (A.numRows() != A.numCols()) throw std::invalid_argument("[CholeskyBatch] A must be square");
        m_factorisers.resize(A.numShards());
        A.forEachShard([&](size_t g, DTensor<T> &s) { m_factorisers[g] = std::make_unique<CholeskyBatchFactoriser<T> >(s, factorised); });
    }

    void factorise() {
        m_matrix->forEachShard([&](size_t g, DTensor<T> &) { m_factorisers[g]->factorise(); });
        m_factorisationDone = true;
    }

    void solve(ShardedDTensor<T> &b) {
        m_matrix->checkSameSharding(b, "CholeskyBatchSolve");
        m_matrix->forEachShard([&](size_t g, DTensor<T> &) { m_factorisers[g]->solve(b.shard(g)); });
    }

    /**
     * solve() and the all-gather of the solutions in ONE kernel per device: every device solves its shard and stores each
     * solution straight into the gathered (n, 1, k) tensor of every device over NVLink (peer stores from inside the solve
     * kernel, gpub_potrs_allgather_batched), so the gather costs no second pass over x and no NCCL launch. `b` is updated in
     * place as by solve(). Returns the gathered solutions, result[g] living on device(g); blocks until they are complete.
     */
    std::vector<std::unique_ptr<DTensor<T> > > solveAllGather(ShardedDTensor<T> &b) {
        if (!m_factorisationDone) throw std::logic_error("[CholeskyBatchSolve] no factor to solve with");
        m_matrix->checkSameSharding(b, "CholeskyBatchSolve");
        if (b.numCols() != 1 || b.numRows() != m_matrix->numRows()) throw std::invalid_argument("[CholeskyBatchSolve] A and b incompatible");
        const size_t G = m_matrix->numShards(), n = m_matrix->numRows(), k = m_matrix->numMats();
        if (G > 8) throw std::invalid_argument("[solveAllGather] at most 8 devices");
        std::vector<std::unique_ptr<DTensor<T> > > full(G);
        std::vector<T *> peers(G);
        for (size_t g = 0; g < G; g++) {
            gpub200::DeviceScope scope(m_matrix->device(g));
            full[g] = std::make_unique<DTensor<T> >(n, 1, k);
            peers[g] = full[g]->raw();
        }
        /* the result tensors were allocated on the legacy stream of their devices: they exist before any peer writes to them */
        for (size_t g = 0; g < G; g++) {
            gpub200::DeviceScope scope(m_matrix->device(g));
            gpuErrChk(cudaStreamSynchronize(cudaStreamLegacy));
        }
        m_matrix->forEachShard([&](size_t g, DTensor<T> &s) {
            DTensor<T> &bs = b.shard(g);
            gpuErrChk(gpub200::Abi<T>::potrs_allgather(gpub200::ctx(), (int) s.streamIdx(), n, s.raw(), n, n * n, bs.raw(), n, s.numMats(),
                                                       peers.data(), (int) G, m_matrix->shardRange(g).first, n));
        });
        m_matrix->synchronize();
        return full;
    }

    /** Status codes of shard g ((1, 1, shard size)-tensor on device(g)). */
    DTensor<int> &info(size_t g) {
        if (!m_factorisers.at(g)) throw std::invalid_argument("[ShardedCholeskyBatch] shard is empty");
        return m_factorisers[g]->info();
    }

    /** All status codes, in matrix order, on the host. */
    std::vector<int> statuses() {
        std::vector<int> all(m_matrix->numMats(), 0);
        m_matrix->forEachShard([&](size_t g, DTensor<T> &s) {
            std::vector<int> part;
            m_factorisers[g]->info().download(part);
            std::copy(part.begin(), part.end(), all.begin() + m_matrix->shardRange(g).first);
            (void) s;
        });
        return all;
    }
};

#endif /* GPUB200_SHARDED_CUH */
