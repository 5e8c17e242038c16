// GPU test program for the additive batch API of the header: QRBatchFactoriser must give, for every
// matrix of the batch, what the reference-shaped single-matrix QRFactoriser gives on that matrix, and satisfy the same
// properties the reference's own QR tests check (testTensor.cu qrFactorisation / qrLeastSquares: Q R = A, Q'Q = I, the
// least-squares solution solves the normal equations).
// Built by __graft_entry__.build() into build/tests/additive_test; driven by tests/test_gpu_additive.py.
#include <tensor.cuh>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>

static int g_fail = 0;

static std::string sci(double v) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "%.2e", v);
    return buf;
}

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

template<typename T>
static double relDiff(const std::vector<T> &a, const std::vector<T> &b) {
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); i++) {
        num += ((double) a[i] - (double) b[i]) * ((double) a[i] - (double) b[i]);
        den += (double) b[i] * (double) b[i];
    }
    return std::sqrt(num / (den > 0 ? den : 1));
}

template<typename T>
static void qrCase(size_t m, size_t n, size_t k, double tol, const char *tag) {
    std::vector<T> a = uniform<T>(m * n * k, 11 + m), b = uniform<T>(m * k, 12 + n);
    // batch API
    DTensor<T> AB(a, m, n, k), bB(b, m, 1, k);
    QRBatchFactoriser<T> qb(AB);
    qb.factorise();
    qb.leastSquares(bB);
    std::vector<T> qrB, xB, tauB;
    AB.download(qrB);
    bB.download(xB);
    qb.householder().download(tauB);
    // reference-shaped loop: one QRFactoriser per matrix
    std::vector<T> qrL(m * n * k), xL(m * k), tauChk;
    for (size_t i = 0; i < k; i++) {
        std::vector<T> ai(a.begin() + i * m * n, a.begin() + (i + 1) * m * n), bi(b.begin() + i * m, b.begin() + (i + 1) * m);
        DTensor<T> A1(ai, m, n, 1), b1(bi, m, 1, 1);
        QRFactoriser<T> q1(A1);
        q1.factorise();
        q1.leastSquares(b1);
        std::vector<T> t;
        A1.download(t);
        std::copy(t.begin(), t.end(), qrL.begin() + i * m * n);
        b1.download(t);
        std::copy(t.begin(), t.end(), xL.begin() + i * m);
    }
    const double dqr = relDiff(qrB, qrL), dx = relDiff(xB, xL);
    const bool exact = std::memcmp(qrB.data(), qrL.data(), qrB.size() * sizeof(T)) == 0 && std::memcmp(xB.data(), xL.data(), xB.size() * sizeof(T)) == 0;
    report((std::string("batch_equals_loop_") + tag).c_str(), dqr <= tol && dx <= tol,
           "rel_qr=" + sci(dqr) + " rel_x=" + sci(dx) + (exact ? " bit-identical" : ""));

    // properties: Q R = A, Q'Q = I, A'(A x - b) = 0
    DTensor<T> A2(a, m, n, k), Q(m, n, k), R(n, n, k);
    QRBatchFactoriser<T> q2(A2);
    q2.factorise();
    q2.getQR(Q, R);
    DTensor<T> QRp(m, n, k);
    QRp.addAB(Q, R);
    std::vector<T> rec;
    QRp.download(rec);
    report((std::string("QR_reconstructs_A_") + tag).c_str(), relDiff(rec, a) <= tol, "rel=" + sci(relDiff(rec, a)));
    DTensor<T> Qt = Q.tr(), QtQ(n, n, k);
    QtQ.addAB(Qt, Q);
    std::vector<T> g, eye(n * n * k, T(0));
    QtQ.download(g);
    for (size_t i = 0; i < k; i++)
        for (size_t c = 0; c < n; c++) eye[i * n * n + c + c * n] = T(1);
    report((std::string("Q_is_orthonormal_") + tag).c_str(), relDiff(g, eye) <= tol, "rel=" + sci(relDiff(g, eye)));
    double worst = 0;
    for (size_t i = 0; i < k; i++) {
        std::vector<double> r(m);
        double nb = 0, nA = 0;
        for (size_t row = 0; row < m; row++) {
            double s = -(double) b[i * m + row];
            for (size_t c = 0; c < n; c++) s += (double) a[i * m * n + row + c * m] * (double) xB[i * m + c];
            r[row] = s;
            nb += (double) b[i * m + row] * (double) b[i * m + row];
        }
        double gn = 0;
        for (size_t c = 0; c < n; c++) {
            double s = 0;
            for (size_t row = 0; row < m; row++) {
                s += (double) a[i * m * n + row + c * m] * r[row];
                nA += (double) a[i * m * n + row + c * m] * (double) a[i * m * n + row + c * m];
            }
            gn += s * s;
        }
        worst = std::max(worst, std::sqrt(gn) / (std::sqrt(nA / n) * std::sqrt(nb) + 1e-300));
    }
    report((std::string("normal_equations_") + tag).c_str(), worst <= 50 * tol, "worst=" + sci(worst));
}

// SPD batch: A_i = G_i G_i' + n I
template<typename T>
static std::vector<T> spdBatch(size_t n, size_t k, uint64_t seed) {
    std::vector<T> g = uniform<T>(n * n * k, seed), a(n * n * k);
    for (size_t i = 0; i < k; i++)
        for (size_t c = 0; c < n; c++)
            for (size_t r = 0; r < n; r++) {
                double s = (r == c) ? (double) n : 0.0;
                for (size_t l = 0; l < n; l++) s += (double) g[i * n * n + r + l * n] * (double) g[i * n * n + c + l * n];
                a[i * n * n + r + c * n] = (T) s;
            }
    return a;
}

// CholeskyBatchFactoriser::factoriseAndSolveFromHost (three-stream host pipeline) must give, bit for bit, what the reference-shaped
// sequence upload -> factorise -> solve -> download gives; pageable and pinned host memory, more chunks than matrices
template<typename T>
static void hostPipelineCase(size_t n, size_t k, size_t chunks, bool pinned, const char *tag) {
    std::vector<T> a = spdBatch<T>(n, k, 21 + n), b = uniform<T>(n * k, 22 + k);
    a[(k / 2) * n * n] = T(-1);                            // one matrix that is not positive definite: info must say so
    DTensor<T> A1(a, n, n, k), b1(b, n, 1, k);
    CholeskyBatchFactoriser<T> c1(A1);
    c1.factorise();
    c1.solve(b1);
    std::vector<T> L1, x1;
    std::vector<int> i1;
    A1.download(L1);
    b1.download(x1);
    c1.info().download(i1);

    T *hA = a.data(), *hB = b.data();
    std::vector<T> xPage(n * k);
    std::vector<int> iPage(k);
    T *hX = xPage.data();
    int *hI = iPage.data();
    if (pinned) {
        gpuErrChk(cudaHostAlloc((void **) &hA, a.size() * sizeof(T), cudaHostAllocDefault));
        gpuErrChk(cudaHostAlloc((void **) &hB, b.size() * sizeof(T), cudaHostAllocDefault));
        gpuErrChk(cudaHostAlloc((void **) &hX, b.size() * sizeof(T), cudaHostAllocDefault));
        gpuErrChk(cudaHostAlloc((void **) &hI, k * sizeof(int), cudaHostAllocDefault));
        std::memcpy(hA, a.data(), a.size() * sizeof(T));
        std::memcpy(hB, b.data(), b.size() * sizeof(T));
    }
    DTensor<T> A2(n, n, k), b2(n, 1, k);
    CholeskyBatchFactoriser<T> c2(A2);
    c2.factoriseAndSolveFromHost(hA, hB, b2, hX, hI, chunks);
    std::vector<T> L2, x2;
    std::vector<int> i2;
    A2.download(L2);
    b2.download(x2);
    c2.info().download(i2);
    bool lower = true;                                      // only the lower triangle is specified
    for (size_t i = 0; i < k && lower; i++) {
        if (i1[i]) continue;                                // columns after a bad pivot are unspecified
        for (size_t c = 0; c < n; c++)
            for (size_t r = c; r < n; r++) lower = lower && L1[i * n * n + r + c * n] == L2[i * n * n + r + c * n];
    }
    bool xs = true;
    for (size_t i = 0; i < k; i++)
        if (!i1[i]) xs = xs && std::memcmp(&x1[i * n], &x2[i * n], n * sizeof(T)) == 0 && std::memcmp(&x1[i * n], hX + i * n, n * sizeof(T)) == 0;
    const bool infos = i1 == i2 && std::memcmp(i1.data(), hI, k * sizeof(int)) == 0 && i1[k / 2] == 1;
    report((std::string("host_pipeline_equals_whole_tensor_calls_") + tag).c_str(), lower && xs && infos,
           std::string(lower ? "" : "L differs ") + (xs ? "" : "x differs ") + (infos ? "" : "info differs"));
    bool threw = false;
    try { c2.factoriseAndSolveFromHost(hA, hB, b2, hX); } catch (const std::logic_error &) { threw = true; }
    report((std::string("host_pipeline_refuses_a_second_factorisation_") + tag).c_str(), threw);
    if (pinned) {
        cudaFreeHost(hA);
        cudaFreeHost(hB);
        cudaFreeHost(hX);
        cudaFreeHost(hI);
    }
}

// upload / download through the pinned ring (pageable vectors larger than several ring pieces, odd byte counts, row-major)
static void stagedCopies() {
    const size_t count = (40u << 20) / sizeof(double) + 12345;          // > 4 ring pieces of 8 MB, not a multiple of anything
    std::vector<double> h = uniform<double>(count, 31), back;
    DTensor<double> d(count);
    d.upload(h);
    d.download(back);
    report("staged_upload_download_roundtrip_40MB", back == h);
    DTensor<double> e(d);                                               // copy constructor on the pool
    e *= 2.0;
    e.download(back);
    bool ok = true;
    for (size_t i = 0; i < count; i += 997) ok = ok && back[i] == 2.0 * h[i];
    report("copy_then_scale_after_staged_upload", ok);
    std::vector<float> rm = uniform<float>(300 * 200 * 7, 32), cm;
    DTensor<float> r(rm, 300, 200, 7, rowMajor);
    r.download(cm);
    ok = true;
    for (size_t k = 0; k < 7; k++)
        for (size_t i = 0; i < 300; i += 7)
            for (size_t j = 0; j < 200; j += 3) ok = ok && cm[k * 60000 + i + j * 300] == rm[k * 60000 + j + i * 200];
    report("row_major_upload", ok);
}

// the stream-ordered pool: a loop of constructions / destructions reuses cached blocks, and the byte counter balances
static void poolReuse() {
    const size_t before = Session::getInstance().totalAllocatedBytes();
    size_t reserved0 = 0, reserved1 = 0, used = 0;
    { DTensor<double> warm(64, 64, 100); DTensor<double> t = warm.tr(); }
    gpuErrChk(gpub_mem_stats(gpub200::ctx(), &reserved0, &used));
    for (int it = 0; it < 200; it++) {
        DTensor<double> a(64, 64, 100, true);
        DTensor<double> t = a.tr();
        DTensor<double> s = a + t;
        (void) s;
    }
    gpuErrChk(gpub_mem_stats(gpub200::ctx(), &reserved1, &used));
    report("pool_reuses_cached_blocks", reserved1 <= reserved0 + (16u << 20), "reserved " + std::to_string(reserved0) + " -> " + std::to_string(reserved1));
    report("byte_counter_balances", Session::getInstance().totalAllocatedBytes() == before);
    // results stay correct when blocks are recycled while other streams still use them: free on stream-1 work, reallocate, reuse
    std::vector<double> ones(1 << 20, 1.0), got;
    bool ok = true;
    for (int it = 0; it < 20; it++) {
        DTensor<double> *x = new DTensor<double>(ones, 1 << 20);
        x->setStreamIdx(1);
        DTensor<double> y(ones, 1 << 20);
        DTensor<double> xv(*x, 0, 0, (1 << 20) - 1);                     // view on stream 0 ...
        y += xv;
        delete x;                                                        // ... freed while the axpy may still be queued
        DTensor<double> z(1 << 20, 1, 1, true);                          // probably the same block, zeroed
        y += z;
        y.download(got);
        ok = ok && got[0] == 2.0 && got[(1 << 20) - 1] == 2.0 && got[12345] == 2.0;
    }
    report("recycled_blocks_are_ordered_across_streams", ok);
}

// Nullspace built from a tensor that lives on another stream must equal the stream-0 result bit for bit (one stream end to end)
static void nullspaceOnAnotherStream() {
    const size_t m = 16, n = 48, k = 9;
    std::vector<double> a = uniform<double>(m * n * k, 41), b = uniform<double>(n * k, 42);
    DTensor<double> A0(a, m, n, k), b0(b, n, 1, k);
    Nullspace<double> ns0(A0);
    ns0.project(b0);
    DTensor<double> A2(a, m, n, k), b2(b, n, 1, k);
    A2.setStreamIdx(2);
    Nullspace<double> ns2(A2);
    ns2.project(b2);
    std::vector<double> n0, n2, p0, p2;
    ns0.nullspace().download(n0);
    ns2.nullspace().download(n2);
    b0.download(p0);
    b2.download(p2);
    report("nullspace_on_stream_2_equals_stream_0", n0 == n2 && p0 == p2);
    bool threw = false;
    try { DTensor<double> wide(400, 300, 1); Svd<double> svd(wide); } catch (const std::invalid_argument &) { threw = true; }
    report("svd_refuses_unsupported_shapes_with_an_exception", threw);
}

// GivensBatchAnnihilator on a (m, n, k) tensor == GivensAnnihilator looped over the k matrices (what a reference user writes)
template<typename T>
static void givensBatchCase(size_t m, size_t n, size_t k, double tol, const char *tag) {
    std::vector<T> a = uniform<T>(m * n * k, 51 + m);
    DTensor<T> AB(a, m, n, k);
    GivensBatchAnnihilator<T> gb(AB);
    // reduce column 0, then column 1, to upper-triangular form like the reference's test (testTensor.cu givensAnnihilate...)
    for (size_t j = 0; j < std::min<size_t>(n, 3); j++)
        for (size_t r = m - 1; r > j; r--) gb.annihilate(j, r, j);
    std::vector<T> outB;
    AB.download(outB);
    std::vector<T> outL(m * n * k);
    for (size_t i = 0; i < k; i++) {
        std::vector<T> ai(a.begin() + i * m * n, a.begin() + (i + 1) * m * n);
        DTensor<T> A1(ai, m, n, 1);
        GivensAnnihilator<T> g1(A1);
        for (size_t j = 0; j < std::min<size_t>(n, 3); j++)
            for (size_t r = m - 1; r > j; r--) g1.annihilate(j, r, j);
        std::vector<T> t;
        A1.download(t);
        std::copy(t.begin(), t.end(), outL.begin() + i * m * n);
    }
    const double d = relDiff(outB, outL);
    const bool exact = std::memcmp(outB.data(), outL.data(), outB.size() * sizeof(T)) == 0;
    report((std::string("givens_batch_equals_loop_") + tag).c_str(), d <= tol, "rel=" + sci(d) + (exact ? " bit-identical" : ""));
    double below = 0, na = 0, nb = 0;
    for (size_t i = 0; i < k; i++)
        for (size_t j = 0; j < n; j++)
            for (size_t r = 0; r < m; r++) {
                const double v = outB[i * m * n + r + j * m], w = a[i * m * n + r + j * m];
                if (j < 3 && r > j) below = std::max(below, std::fabs(v));
                nb += v * v; na += w * w;
            }
    report((std::string("givens_batch_zeroes_and_preserves_norm_") + tag).c_str(), below <= 50 * tol && std::fabs(std::sqrt(nb / na) - 1.0) <= 50 * tol,
           "below=" + sci(below) + " norm_ratio-1=" + sci(std::sqrt(nb / na) - 1.0));
    // per-matrix (c, s) rotations from device arrays
    std::vector<T> cs(2 * k);
    for (size_t i = 0; i < k; i++) { const double th = 0.1 * (double) (i + 1); cs[i] = (T) std::cos(th); cs[k + i] = (T) std::sin(th); }
    DTensor<T> dcs(cs, 2 * k), X(a, m, n, k);
    GivensBatchAnnihilator<T> gx(X);
    gx.applyLeftGivensRotations(0, m - 1, dcs.raw(), dcs.raw() + k);
    gx.applyRightGivensRotations(0, n - 1, dcs.raw(), dcs.raw() + k);
    std::vector<T> got;
    X.download(got);
    std::vector<double> ref(a.begin(), a.end());
    for (size_t i = 0; i < k; i++) {
        const double c = cs[i], s_ = cs[k + i];
        double *ai = ref.data() + i * m * n;
        for (size_t j = 0; j < n; j++) { const double x = ai[0 + j * m], y = ai[m - 1 + j * m]; ai[0 + j * m] = c * x + s_ * y; ai[m - 1 + j * m] = c * y - s_ * x; }
        for (size_t r = 0; r < m; r++) { const double x = ai[r], y = ai[r + (n - 1) * m]; ai[r] = c * x + s_ * y; ai[r + (n - 1) * m] = c * y - s_ * x; }
    }
    std::vector<T> refT(ref.begin(), ref.end());
    report((std::string("rot_batched_matches_host_") + tag).c_str(), relDiff(got, refT) <= 10 * tol, "rel=" + sci(relDiff(got, refT)));
    bool threw = false;
    try { gb.annihilate(0, m, 0); } catch (const std::invalid_argument &) { threw = true; }
    report((std::string("givens_batch_bad_index_throws_") + tag).c_str(), threw);
}

// PitchedDTensor: matrices on 128-byte boundaries; results equal the dense DTensor path, the padding is never written
template<typename T>
static void pitchedCase(size_t n, size_t k, double tol, const char *tag) {
    std::vector<T> a = spdBatch<T>(n, k, 61 + n), b = uniform<T>(n * k, 62), c = uniform<T>(n * n * k, 63);
    DTensor<T> A1(a, n, n, k), b1(b, n, 1, k), C1(c, n, n, k), P1(n, n, k);
    P1.addAB(A1, C1, T(0.5), T(0));
    CholeskyBatchFactoriser<T> f1(A1);
    f1.factorise();
    f1.solve(b1);
    std::vector<T> L1, x1, p1;
    A1.download(L1); b1.download(x1); P1.download(p1);

    PitchedDTensor<T> A2(a, n, n, k), b2(b, n, 1, k), C2(c, n, n, k), P2(n, n, k, true);
    const bool aligned = (A2.matStride() * sizeof(T)) % 128 == 0 && (b2.matStride() * sizeof(T)) % 128 == 0 &&
                         ((uintptr_t) A2.matrix(k - 1)) % 128 == 0 && A2.matStride() >= n * n;
    report((std::string("pitched_matrices_start_on_128_byte_boundaries_") + tag).c_str(), aligned, "stride=" + std::to_string(A2.matStride()));
    P2.addAB(A2, C2, T(0.5), T(0));
    PitchedCholeskyBatchFactoriser<T> f2(A2);
    f2.factorise();
    f2.solve(b2);
    std::vector<T> L2, x2, p2;
    A2.download(L2); b2.download(x2); P2.download(p2);
    std::vector<int> i1, i2;
    f1.info().download(i1); f2.info().download(i2);
    report((std::string("pitched_cholesky_equals_dense_") + tag).c_str(), relDiff(L2, L1) <= tol && relDiff(x2, x1) <= 10 * tol && i1 == i2,
           "rel_L=" + sci(relDiff(L2, L1)) + " rel_x=" + sci(relDiff(x2, x1)));
    report((std::string("pitched_addAB_equals_dense_") + tag).c_str(), relDiff(p2, p1) <= tol, "rel=" + sci(relDiff(p2, p1)));
    // padding: P2 was zero-initialised, so every byte between the matrices must still be zero; round trip through toDense / fromDense
    std::vector<T> flat(P2.matStride() * k);
    gpuErrChk(cudaMemcpy(flat.data(), P2.raw(), flat.size() * sizeof(T), cudaMemcpyDeviceToHost));
    bool clean = true;
    for (size_t i = 0; i < k; i++)
        for (size_t e = n * n; e < P2.matStride(); e++) clean = clean && flat[i * P2.matStride() + e] == T(0);
    DTensor<T> dense = P2.toDense();
    PitchedDTensor<T> back = PitchedDTensor<T>::fromDense(dense);
    std::vector<T> p3;
    back.download(p3);
    report((std::string("pitched_padding_untouched_and_round_trip_") + tag).c_str(), clean && p3 == p2);
}

int main() {
    Session::setStreams(3);
    pitchedCase<double>(5, 1000, 1e-14, "f64_5x5");
    pitchedCase<float>(10, 333, 1e-6, "f32_10x10");
    pitchedCase<double>(3, 4097, 1e-14, "f64_3x3");
    givensBatchCase<double>(6, 5, 37, 1e-14, "f64_6x5x37");
    givensBatchCase<float>(9, 4, 300, 2e-6, "f32_9x4x300");
    hostPipelineCase<double>(32, 5000, 7, false, "f64_32_pageable");
    hostPipelineCase<double>(32, 5000, 16, true, "f64_32_pinned");
    hostPipelineCase<float>(16, 3, 16, false, "f32_16_more_chunks_than_matrices");
    hostPipelineCase<double>(64, 600, 5, false, "f64_64_pageable");
    stagedCopies();
    poolReuse();
    nullspaceOnAnotherStream();
    qrCase<double>(20, 3, 5, 1e-12, "f64_20x3");        // the reference's qrLeastSquares size, batched
    qrCase<double>(64, 16, 37, 1e-12, "f64_64x16");
    qrCase<double>(512, 64, 6, 1e-12, "f64_512x64");    // tensor-pipe kernel
    qrCase<float>(64, 16, 33, 2e-5, "f32_64x16");
    qrCase<float>(300, 40, 3, 2e-5, "f32_300x40");
    {   // argument errors
        bool t1 = false, t2 = false;
        try { DTensor<double> fat(3, 5, 2); QRBatchFactoriser<double> q(fat); } catch (const std::invalid_argument &) { t1 = true; }
        try {
            DTensor<double> A(8, 3, 4), rhs(8, 1, 3);
            QRBatchFactoriser<double> q(A);
            q.leastSquares(rhs);
        } catch (const std::invalid_argument &) { t2 = true; }
        report("argument_errors_throw", t1 && t2);
    }
    std::printf("%s failures=%d\n", g_fail ? "FAILED" : "ALL PASSED", g_fail);
    return g_fail ? 1 : 0;
}
