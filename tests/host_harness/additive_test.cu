// GPU test program for the additive batch API of the header: QRBatchFactoriser must give, for every
// matrix of the batch, what the reference-shaped single-matrix QRFactoriser gives on that matrix, and satisfy the same
// properties the reference's own QR tests check (testTensor.cu qrFactorisation / qrLeastSquares: Q R = A, Q'Q = I, the
// least-squares solution solves the normal equations).
// Built by __graft_entry__.build() into build/tests/additive_test; driven by tests/test_gpu_additive.py.
#include <tensor.cuh>

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

int main() {
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
