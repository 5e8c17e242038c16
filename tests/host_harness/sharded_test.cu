// GPU test program for include/gpub200/sharded.cuh: every sharded operation must give exactly what the
// single-GPU DTensor path gives on the same inputs, for ragged and empty shards, over NCCL and over peer copies.
// usage: sharded_test <devices, e.g. 0,1 or 0,0> <transport: auto|nccl|p2p> [gather_mats]
// Built by __graft_entry__.build() into build/tests/sharded_test; driven by tests/test_gpu_sharded.py.
#include <tensor.cuh>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <random>
#include <sstream>
#include <string>

static int g_fail = 0;

static void report(const char *name, bool ok, const std::string &detail = "") {
    std::printf("%s %s %s\n", ok ? "PASS" : "FAIL", name, detail.c_str());
    if (!ok) g_fail++;
}

template<typename T>
static std::vector<T> uniform(size_t n, uint64_t seed) {
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<double> d(-1.0, 1.0);
    std::vector<T> v(n);
    for (auto &x: v) x = (T) d(gen);
    return v;
}

// SPD batch: A_i = G_i G_i^T + n I (column-major)
static std::vector<double> spd(size_t n, size_t k, uint64_t seed) {
    std::vector<double> g = uniform<double>(n * n * k, seed), a(n * n * k);
    for (size_t m = 0; m < k; m++)
        for (size_t i = 0; i < n; i++)
            for (size_t j = 0; j < n; j++) {
                double s = (i == j) ? (double) n : 0.0;
                for (size_t l = 0; l < n; l++) s += g[m * n * n + i + l * n] * g[m * n * n + j + l * n];
                a[m * n * n + i + j * n] = s;
            }
    return a;
}

template<typename T>
static bool same(const std::vector<T> &a, const std::vector<T> &b) {
    return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(T)) == 0;
}

// lower triangles only: potrf leaves the strict upper triangle untouched, so both sides hold the input there anyway
int main(int argc, char **argv) {
    std::vector<int> devices;
    {
        std::stringstream ss(argc > 1 ? argv[1] : "0");
        std::string tok;
        while (std::getline(ss, tok, ',')) devices.push_back(std::stoi(tok));
    }
    int transport = GPUB_GATHER_AUTO;
    if (argc > 2 && !std::strcmp(argv[2], "nccl")) transport = GPUB_GATHER_NCCL;
    if (argc > 2 && !std::strcmp(argv[2], "p2p")) transport = GPUB_GATHER_P2P;
    const size_t gatherMats = argc > 3 ? std::stoull(argv[3]) : 20000;
    int count = 0, ncclVersion = 0;
    gpuErrChk(gpub_multi_device_count(&count));
    gpuErrChk(gpub_multi_nccl_version(&ncclVersion));
    std::printf("INFO devices_on_box=%d shards=%zu nccl_version=%d\n", count, devices.size(), ncclVersion);
    const size_t G = devices.size();

    {   // ---- CholeskyBatchFactoriser, 32 x 32 fp64 (BASELINE config 2 shape), ragged shards, one non-SPD matrix ----
        const size_t n = 32, k = 1003, badIdx = 777;
        std::vector<double> a = spd(n, k, 1), b = uniform<double>(n * k, 2);
        a[badIdx * n * n + 5 + 5 * n] = -1.0;   // pivot 6 of matrix 777 is not positive
        DTensor<double> A1(a, n, n, k), b1(b, n, 1, k);
        CholeskyBatchFactoriser<double> f1(A1);
        f1.factorise();
        f1.solve(b1);
        std::vector<double> L1, x1;
        std::vector<int> info1;
        A1.download(L1);
        b1.download(x1);
        f1.info().download(info1);

        ShardedDTensor<double> AS(a, n, n, k, devices), bS(b, n, 1, k, devices);
        ShardedCholeskyBatchFactoriser<double> fS(AS);
        fS.factorise();
        fS.solve(bS);
        std::vector<double> LS, xS;
        AS.download(LS);
        bS.download(xS);
        std::vector<int> infoS = fS.statuses();
        // the failed matrix holds NaN after its bad column on both sides: compare the bytes
        report("cholesky_factor_bit_identical", same(L1, LS));
        report("cholesky_solution_bit_identical", same(x1, xS));
        report("cholesky_info", same(info1, infoS) && infoS[badIdx] == 6 && infoS[0] == 0, "info[777]=" + std::to_string(infoS[badIdx]));
        size_t covered = 0;
        for (size_t g = 0; g < G; g++) covered += AS.shard(g).numMats();
        report("shards_cover_the_batch", covered == k && AS.shardRange(0).first == 0 && AS.shardRange(G - 1).second == k);

        int used = -1;
        auto full = AS.allGather(transport, &used);
        bool ok = full.size() == G;
        for (size_t g = 0; g < G && ok; g++) {
            gpub200::DeviceScope scope(devices[g]);
            std::vector<double> got;
            full[g]->download(got);
            ok = same(got, L1);
        }
        report("allgather_ragged_every_device_holds_the_batch", ok, std::string("transport=") + (used == GPUB_GATHER_NCCL ? "nccl" : "p2p"));

        // solve + all-gather fused into the solve kernel (peer stores over NVLink): same bits on every device, and in b itself
        if (G <= 8) {
            ShardedDTensor<double> bF(b, n, 1, k, devices);
            auto xs = fS.solveAllGather(bF);
            bool okf = xs.size() == G;
            for (size_t g = 0; g < G && okf; g++) {
                gpub200::DeviceScope scope(devices[g]);
                std::vector<double> got;
                xs[g]->download(got);
                okf = same(got, x1);
            }
            std::vector<double> xF;
            bF.download(xF);
            report("fused_solve_allgather_bit_identical_on_every_device", okf && same(xF, x1));
            bool threw = false;
            try {
                ShardedDTensor<double> A3(a, n, n, k, devices);
                ShardedCholeskyBatchFactoriser<double> f3(A3);
                f3.solveAllGather(bF);
            } catch (const std::logic_error &) { threw = true; }
            report("fused_solve_allgather_needs_a_factor", threw);
        }
    }

    {   // ---- addAB, 8 x 8 fp64, k = 4096 (BASELINE config 1), operators, reductions ----
        const size_t n = 8, k = 4096;
        std::vector<double> a = uniform<double>(n * n * k, 3), b = uniform<double>(n * n * k, 4), c = uniform<double>(n * n * k, 5);
        DTensor<double> A1(a, n, n, k), B1(b, n, n, k), C1(c, n, n, k);
        C1.addAB(A1, B1, 0.5, -2.0);
        C1 += A1;
        C1 *= 3.0;
        C1 -= B1;
        std::vector<double> r1;
        C1.download(r1);
        ShardedDTensor<double> AS(a, n, n, k, devices), BS(b, n, n, k, devices), CS(c, n, n, k, devices);
        CS.addAB(AS, BS, 0.5, -2.0);
        CS += AS;
        CS *= 3.0;
        CS -= BS;
        std::vector<double> rS;
        CS.download(rS);
        report("addAB_and_operators_bit_identical", same(r1, rS));
        auto close = [](double x, double y) { return std::fabs(x - y) <= 1e-12 * std::fabs(y); };
        report("normF", close(CS.normF(), C1.normF()));
        report("sumAbs", close(CS.sumAbs(), C1.sumAbs()));
        report("dotF", close(CS.dotF(AS), C1.dotF(A1)));
        report("maxAbs_minAbs", CS.maxAbs() == C1.maxAbs() && CS.minAbs() == C1.minAbs());

        int used = -1;
        auto full = CS.allGather(transport, &used);   // equal shards when G divides 4096: ncclAllGather
        bool ok = true;
        for (size_t g = 0; g < G && ok; g++) {
            gpub200::DeviceScope scope(devices[g]);
            std::vector<double> got;
            full[g]->download(got);
            ok = same(got, r1);
        }
        report("allgather_equal_shards", ok, std::string("transport=") + (used == GPUB_GATHER_NCCL ? "nccl" : "p2p"));
    }

    {   // ---- leastSquaresBatched, 64 x 16 fp32 (BASELINE config 3 shape) ----
        const size_t m = 64, n = 16, k = 515;
        std::vector<float> a = uniform<float>(m * n * k, 6), b = uniform<float>(m * k, 7);
        DTensor<float> A1(a, m, n, k), b1(b, m, 1, k);
        A1.leastSquaresBatched(b1);
        std::vector<float> x1, q1;
        b1.download(x1);
        A1.download(q1);
        ShardedDTensor<float> AS(a, m, n, k, devices), bS(b, m, 1, k, devices);
        AS.leastSquaresBatched(bS);
        std::vector<float> xS, qS;
        bS.download(xS);
        AS.download(qS);
        report("least_squares_bit_identical", same(x1, xS) && same(q1, qS));
    }

    {   // ---- fewer matrices than shards: trailing shards are empty ----
        const size_t n = 4, k = 1;
        std::vector<double> a = spd(n, k, 8), b = uniform<double>(n * k, 9);
        DTensor<double> A1(a, n, n, k), b1(b, n, 1, k);
        CholeskyBatchFactoriser<double> f1(A1);
        f1.factorise();
        f1.solve(b1);
        std::vector<double> x1, xS;
        b1.download(x1);
        ShardedDTensor<double> AS(a, n, n, k, devices), bS(b, n, 1, k, devices);
        ShardedCholeskyBatchFactoriser<double> fS(AS);
        fS.factorise();
        fS.solve(bS);
        bS.download(xS);
        auto full = bS.allGather(transport);
        std::vector<double> got;
        {
            gpub200::DeviceScope scope(devices[G - 1]);
            full[G - 1]->download(got);
        }
        report("empty_trailing_shards", same(x1, xS) && same(got, x1));
    }

    {   // ---- argument errors are exceptions, as everywhere in the header ----
        bool threw = false;
        try {
            ShardedDTensor<double> X(4, 4, 10, devices), Y(4, 4, 11, devices);
            X += Y;
        } catch (const std::invalid_argument &) { threw = true; }
        report("mismatched_sharding_throws", threw);
    }

    {   // ---- all-gather bandwidth on a DRAM-sized result (reported, not asserted) ----
        const size_t n = 32, k = gatherMats - gatherMats % G;
        ShardedDTensor<double> L(n, n, k, devices, true);
        int used = -1;
        auto warm = L.allGather(transport, &used);
        warm.clear();
        L.synchronize();
        auto t0 = std::chrono::steady_clock::now();
        auto full = L.allGather(transport, &used);
        auto t1 = std::chrono::steady_clock::now();
        const double sec = std::chrono::duration<double>(t1 - t0).count();
        const double bytesIn = (double) (n * n * k * sizeof(double)) * (double) (G - 1) / (double) G;   // received per device
        std::printf("INFO allgather transport=%s shards=%zu bytes_per_device=%.0f seconds=%.6f (includes allocating the %zu result tensors) "
                    "GBps_into_each_device=%.1f\n", used == GPUB_GATHER_NCCL ? "nccl" : "p2p", G, bytesIn, sec, G, bytesIn / sec / 1e9);
    }

    if (G <= 8) {   // ---- fused solve + all-gather against solve followed by all-gather of x (reported, not asserted) ----
        const size_t n = 32, k = gatherMats - gatherMats % G;
        std::vector<double> a1 = spd(n, 1, 7);
        std::vector<double> a(n * n * k), b = uniform<double>(n * k, 8);
        for (size_t i = 0; i < k; i++) std::copy(a1.begin(), a1.end(), a.begin() + i * n * n);
        ShardedDTensor<double> AS(a, n, n, k, devices), bS(b, n, 1, k, devices), bF(b, n, 1, k, devices);
        ShardedCholeskyBatchFactoriser<double> fS(AS);
        fS.factorise();
        { auto w1 = fS.solveAllGather(bF); fS.solve(bS); auto w2 = bS.allGather(transport); }   // warm-up (pool, NCCL clique)
        bS.upload(b); bF.upload(b);
        AS.synchronize();
        auto t0 = std::chrono::steady_clock::now();
        fS.solve(bS);
        auto sep = bS.allGather(transport);
        auto t1 = std::chrono::steady_clock::now();
        auto fused = fS.solveAllGather(bF);
        auto t2 = std::chrono::steady_clock::now();
        const double tsep = std::chrono::duration<double>(t1 - t0).count(), tfus = std::chrono::duration<double>(t2 - t1).count();
        const double bytesIn = (double) (n * k * sizeof(double)) * (double) (G - 1) / (double) G;
        std::printf("INFO solve_allgather systems=%zu shards=%zu separate_ms=%.3f fused_ms=%.3f x_bytes_into_each_device=%.0f\n", k, G, tsep * 1e3,
                    tfus * 1e3, bytesIn);
    }

    std::printf("%s failures=%d\n", g_fail ? "FAILED" : "ALL PASSED", g_fail);
    return g_fail ? 1 : 0;
}
