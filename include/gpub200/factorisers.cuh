/*
 * factorisers.cuh -- IStatus, Svd, CholeskyFactoriser, QRFactoriser, Nullspace,
 * CholeskyBatchFactoriser, GivensAnnihilator.
 *
 * Same classes, constructors, methods, exceptions and ownership rules as the reference
 * (ref: tensor.cuh:1453-2312). Differences underneath:
 *   - Svd::factorise is ONE batched launch sequence, not numMats cuSOLVER calls in a host loop
 *     (ref: tensor.cuh:1630-1648); Svd::rank is one launch, not one per matrix (1602-1607).
 *   - Nullspace is assembled on the device (pack + N N^T kernels): no rank download, no per-matrix
 *     slices / tr() allocations / addAB calls (ref: tensor.cuh:2055-2078).
 *   - No cuSOLVER workspace queries: the only scratch is the SVD workspace, sized by the C ABI.
 *   - Every launch goes to the stream of the tensor being factorised (the reference sent all cuSOLVER
 *     work to handle 0).
 */
#ifndef GPUB200_FACTORISERS_CUH
#define GPUB200_FACTORISERS_CUH

#include "dtensor.cuh"

/* ================================================================================================
 *  STATUS INTERFACE
 * ================================================================================================ */
/** Holds the device-side status codes (LAPACK-style info) of a factorisation, one per matrix. */
class IStatus {
protected:
    std::unique_ptr<DTensor<int> > m_info;

    IStatus(size_t n = 1) {
        m_info = std::make_unique<DTensor<int> >(1, 1, n, true);
    }

public:
    /** @return (1, 1, n)-tensor of status codes */
    virtual DTensor<int> &info() {
        return *m_info;
    }
};

/* ================================================================================================
 *  SINGULAR VALUE DECOMPOSITION (SVD)
 * ================================================================================================ */

/**
 * Kept for source compatibility (it is part of the reference's public header): counts the entries of
 * d_array above epsilon. Svd::rank() uses the batched launcher gpub_count_gt_batched_* instead.
 */
TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
__global__ void k_countNonzeroSingularValues(const T *d_array, size_t n, unsigned int *d_count, T epsilon) {
    size_t idx = threadIdx.x + (size_t) blockIdx.x * blockDim.x;
    if (idx < n && d_array[idx] > epsilon) atomicAdd(d_count, 1u);
}

TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
class Svd : public IStatus {
private:
    size_t m_lwork = 0;                              ///< workspace size in bytes
    DTensor<T> *m_tensor = nullptr;                  ///< matrices to factorise (owned iff !m_destroyMatrix)
    std::shared_ptr<DTensor<T> > m_Vtr;              ///< V' (n, n, k)
    std::shared_ptr<DTensor<T> > m_S;                ///< singular values (min(m,n), 1, k)
    std::shared_ptr<DTensor<T> > m_U;                ///< full U (m, m, k) when requested
    std::unique_ptr<DTensor<unsigned char> > m_workspace;
    std::shared_ptr<DTensor<unsigned int> > m_rank;  ///< (1, 1, k)
    bool m_computeU = false;
    bool m_destroyMatrix = true;

    void checkMatrix(DTensor<T> &tensor) const {
        if (tensor.numRows() < tensor.numCols()) {
            throw std::invalid_argument("[svd] your matrix is fat (no offence)");
        }
    };

    void computeWorkspaceSize(size_t m, size_t n) {
        m_lwork = gpub200::Abi<T>::gesvd_worksize(m, n, m_computeU ? 'A' : 'N', m_tensor->numMats());
    }

public:
    /**
     * @param mat tall or square matrices (m, n, k)
     * @param computeU whether to compute the full U
     * @param destroyMatrix whether the factorisation may overwrite `mat`
     */
    Svd(DTensor<T> &mat, bool computeU = false, bool destroyMatrix = true) : IStatus(mat.numMats()) {
        checkMatrix(mat);
        m_destroyMatrix = destroyMatrix;
        m_tensor = (destroyMatrix) ? &mat : new DTensor<T>(mat);
        m_computeU = computeU;
        const size_t m = mat.numRows(), n = mat.numCols(), nMats = mat.numMats();
        computeWorkspaceSize(m, n);
        /* shapes the batched kernels cannot serve are refused here, as an exception, not at factorise() time as an exit */
        if (m_lwork == 0 && mat.numEl() > 0) {
            if (!m_destroyMatrix) delete m_tensor;
            throw std::invalid_argument("[svd] matrix shape not supported by the batched SVD kernels (see INTEGRATION.md)");
        }
        m_workspace = std::make_unique<DTensor<unsigned char> >(m_lwork, 1, 1);
        m_Vtr = std::make_shared<DTensor<T> >(n, n, nMats);
        m_S = std::make_shared<DTensor<T> >(std::min(m, n), 1, nMats);
        m_rank = std::make_unique<DTensor<unsigned int> >(1, 1, nMats, true);
        if (computeU) m_U = std::make_shared<DTensor<T> >(m, m, nMats);
    }

    /** Factorise all matrices (batched). The input is overwritten unless destroyMatrix was false. */
    void factorise() {
        const size_t m = m_tensor->numRows(), n = m_tensor->numCols(), nMats = m_tensor->numMats();
        gpuErrChk(gpub200::Abi<T>::gesvd(gpub200::ctx(), (int) m_tensor->streamIdx(), m_computeU ? 'A' : 'N', m, n,
                                         m_tensor->raw(), m, m * n,
                                         m_S->raw(), std::min(m, n),
                                         m_computeU ? m_U->raw() : nullptr, m, m * m,
                                         m_Vtr->raw(), n, n * n,
                                         m_workspace->raw(), m_lwork, m_info->raw(), nMats));
    }

    DTensor<T> &singularValues() const { return *m_S; }

    DTensor<T> const &rightSingularVectors() const { return *m_Vtr; }

    std::optional<std::shared_ptr<DTensor<T> > > leftSingularVectors() const {
        if (!m_computeU) return std::nullopt;
        return m_U;
    }

    ~Svd() {
        if (!m_destroyMatrix && m_tensor) delete m_tensor;
    }

    /**
     * Numerical rank of every matrix: number of singular values above epsilon.
     * Deliberate fix: the counters are reset first, so calling rank() twice does not double the result
     * (the reference accumulates into a counter it never resets, tensor.cuh:1553, 1600-1609).
     * @return (1, 1, nMats)-tensor
     */
    DTensor<unsigned int> const &rank(T epsilon = 1e-6) const {
        const size_t len = m_S->numCols() * m_S->numRows();
        gpuErrChk(cudaMemset(m_rank->raw(), 0, m_rank->numMats() * sizeof(unsigned int)));
        gpuErrChk(gpub200::Abi<T>::count_gt(gpub200::ctx(), (int) m_tensor->streamIdx(), m_S->raw(), len, len, epsilon,
                                            m_rank->raw(), m_rank->numMats()));
        return *m_rank;
    }
};

/* ================================================================================================
 *  CHOLESKY FACTORISATION (CF)
 * ================================================================================================ */

TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
class CholeskyFactoriser : public IStatus {
private:
    DTensor<T> *m_matrix; ///< matrix to factorise (not owned)

public:
    CholeskyFactoriser(DTensor<T> &A) : IStatus() {
        if (A.numMats() > 1) throw std::invalid_argument("[Cholesky] 3D tensors require `CholeskyBatchFactoriser`");
        if (A.numRows() != A.numCols()) throw std::invalid_argument("[Cholesky] Matrix A must be square");
        m_matrix = &A;
    }

    /** A = L L', lower triangle overwritten with L. */
    void factorise() {
        const size_t n = m_matrix->numRows();
        gpuErrChk(gpub200::Abi<T>::potrf(gpub200::ctx(), (int) m_matrix->streamIdx(), n, m_matrix->raw(), n, n * n,
                                         m_info->raw(), 1));
    }

    /** Solves A x = b in place using the factor (one right-hand side). */
    void solve(DTensor<T> &rhs) {
        const size_t n = m_matrix->numRows();
        gpuErrChk(gpub200::Abi<T>::potrs(gpub200::ctx(), (int) m_matrix->streamIdx(), n, m_matrix->raw(), n, n * n,
                                         rhs.raw(), n, 1));
    }
};

/* ================================================================================================
 *  QR DECOMPOSITION (QR)
 * ================================================================================================ */

TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
class QRFactoriser : public IStatus {
private:
    std::unique_ptr<DTensor<T> > m_householder; ///< tau (n)
    DTensor<T> *m_matrix;                       ///< matrix to factorise (not owned)

public:
    QRFactoriser(DTensor<T> &A) : IStatus() {
        if (A.numMats() > 1) throw std::invalid_argument("[QR] 3D tensors require `leastSquaresBatched`");
        if (A.numRows() < A.numCols()) throw std::invalid_argument("[QR] Matrix A must be tall or square");
        m_matrix = &A;
        m_householder = std::make_unique<DTensor<T> >(m_matrix->numCols());
    }

    /** Householder QR in LAPACK storage: R above, reflectors below the diagonal. */
    void factorise() {
        const size_t m = m_matrix->numRows(), n = m_matrix->numCols();
        gpuErrChk(gpub200::Abi<T>::geqrf(gpub200::ctx(), (int) m_matrix->streamIdx(), m, n, m_matrix->raw(), m, m * n,
                                         m_householder->raw(), n, 1));
    }

    /** Least squares: rhs[0:n] <- argmin ||A x - rhs|| (rhs is overwritten by Q'rhs, then solved). */
    void leastSquares(DTensor<T> &rhs) {
        const size_t m = m_matrix->numRows(), n = m_matrix->numCols();
        const int s = (int) m_matrix->streamIdx();
        gpuErrChk(gpub200::Abi<T>::ormqr(gpub200::ctx(), s, 1, m, 1, n, m_matrix->raw(), m, m * n,
                                         m_householder->raw(), n, rhs.raw(), m, m, 1));
        gpuErrChk(gpub200::Abi<T>::trsv(gpub200::ctx(), s, n, m_matrix->raw(), m, m * n, rhs.raw(), m, 1));
    }

    /**
     * Debug helper: explicit thin Q (m x n) and R (n x n).
     * @throws std::invalid_argument if Q or R have invalid dimensions
     */
    void getQR(DTensor<T> &Q, DTensor<T> &R) {
        const size_t m = m_matrix->numRows(), n = m_matrix->numCols();
        if (Q.numRows() != m || Q.numCols() != n) throw std::invalid_argument("[QR] invalid shape of Q.");
        if (R.numRows() != n || R.numCols() != n) throw std::invalid_argument("[QR] invalid shape of R.");
        /* Q = H_0 ... H_{n-1} applied to the first n columns of the identity */
        std::vector<T> eye(m * n, T(0));
        for (size_t c = 0; c < n; c++) eye[c + c * m] = T(1);
        Q.upload(eye);
        gpuErrChk(gpub200::Abi<T>::ormqr(gpub200::ctx(), (int) m_matrix->streamIdx(), 0, m, n, n, m_matrix->raw(), m, m * n,
                                         m_householder->raw(), n, Q.raw(), m, m * n, 1));
        /* R: one download of the factored matrix instead of n^2/2 single-element copies */
        std::vector<T> qr;
        m_matrix->download(qr);
        std::vector<T> upper(n * n, T(0));
        for (size_t c = 0; c < n; c++)
            for (size_t r = 0; r <= c; r++) upper[r + c * n] = qr[r + c * m];
        R.upload(upper);
    }
};

/**
 * Additive API: Householder QR of every matrix of a (m, n, k)-tensor at once.
 * The reference's QRFactoriser takes one matrix (it throws for k > 1, tensor.cuh:1811-1813) and its users loop over the
 * batch with three library calls per matrix (BASELINE config 4: 256 x QRFactoriser). Here factorise() is ONE launch of the
 * batched geqrf kernel and leastSquares() two (Q'b, then the triangular solves); storage and results per matrix are
 * exactly those of QRFactoriser (LAPACK layout: R above, reflectors below the diagonal, tau per matrix).
 */
TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
class QRBatchFactoriser : public IStatus {
private:
    std::unique_ptr<DTensor<T> > m_householder; ///< tau (n, 1, k)
    DTensor<T> *m_matrix;                       ///< matrices to factorise (not owned)

public:
    QRBatchFactoriser() = delete;

    QRBatchFactoriser(DTensor<T> &A) : IStatus(A.numMats()) {
        if (A.numRows() < A.numCols()) throw std::invalid_argument("[QRBatch] matrices must be tall or square");
        m_matrix = &A;
        m_householder = std::make_unique<DTensor<T> >(A.numCols(), 1, A.numMats(), true);
    }

    /** tau of every matrix, (n, 1, k). */
    DTensor<T> &householder() { return *m_householder; }

    void factorise() {
        const size_t m = m_matrix->numRows(), n = m_matrix->numCols();
        gpuErrChk(gpub200::Abi<T>::geqrf(gpub200::ctx(), (int) m_matrix->streamIdx(), m, n, m_matrix->raw(), m, m * n,
                                         m_householder->raw(), n, m_matrix->numMats()));
    }

    /** C_i <- Q_i' C_i (transpose = true) or Q_i C_i; C is (m, c, k). */
    void applyQ(DTensor<T> &C, bool transpose) {
        const size_t m = m_matrix->numRows(), n = m_matrix->numCols(), k = m_matrix->numMats();
        if (C.numRows() != m || C.numMats() != k) throw std::invalid_argument("[QRBatch] C incompatible with the factorised tensor");
        gpuErrChk(gpub200::Abi<T>::ormqr(gpub200::ctx(), (int) m_matrix->streamIdx(), transpose ? 1 : 0, m, C.numCols(), n,
                                         m_matrix->raw(), m, m * n, m_householder->raw(), n, C.raw(), m, m * C.numCols(), k));
    }

    /** rhs_i[0:n] <- argmin ||A_i x - rhs_i|| for every matrix; rhs is (m, 1, k) and is overwritten (Q_i' rhs_i, then solved). */
    void leastSquares(DTensor<T> &rhs) {
        const size_t m = m_matrix->numRows(), n = m_matrix->numCols(), k = m_matrix->numMats();
        if (rhs.numRows() != m || rhs.numCols() != 1 || rhs.numMats() != k)
            throw std::invalid_argument("[QRBatch] rhs must be (m, 1, k)");
        applyQ(rhs, true);
        gpuErrChk(gpub200::Abi<T>::trsv(gpub200::ctx(), (int) m_matrix->streamIdx(), n, m_matrix->raw(), m, m * n, rhs.raw(), m, k));
    }

    /** Debug helper, as QRFactoriser::getQR: explicit thin Q (m, n, k) and R (n, n, k). */
    void getQR(DTensor<T> &Q, DTensor<T> &R) {
        const size_t m = m_matrix->numRows(), n = m_matrix->numCols(), k = m_matrix->numMats();
        if (Q.numRows() != m || Q.numCols() != n || Q.numMats() != k) throw std::invalid_argument("[QRBatch] invalid shape of Q.");
        if (R.numRows() != n || R.numCols() != n || R.numMats() != k) throw std::invalid_argument("[QRBatch] invalid shape of R.");
        std::vector<T> eye(m * n * k, T(0));
        for (size_t i = 0; i < k; i++)
            for (size_t c = 0; c < n; c++) eye[i * m * n + c + c * m] = T(1);
        Q.upload(eye);
        applyQ(Q, false);
        std::vector<T> qr;
        m_matrix->download(qr);
        std::vector<T> upper(n * n * k, T(0));
        for (size_t i = 0; i < k; i++)
            for (size_t c = 0; c < n; c++)
                for (size_t r = 0; r <= c; r++) upper[i * n * n + r + c * n] = qr[i * m * n + r + c * m];
        R.upload(upper);
    }
};

/* ================================================================================================
 *  Nullspace (N)
 * ================================================================================================ */

TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
class Nullspace {
private:
    std::unique_ptr<DTensor<T> > m_nullspace; ///< N_i, left-packed, zero padded (n, n, k)
    std::unique_ptr<DTensor<T> > m_projOp;    ///< N_i N_i' (n, n, k)

public:
    Nullspace(DTensor<T> &a);

    DTensor<T> const &nullspace() const { return *m_nullspace; }

    /** b_i <- N_i N_i' b_i */
    void project(DTensor<T> &b);
};

template<typename T>
TEMPLATE_CONSTRAINT_REQUIRES_FPX
inline Nullspace<T>::Nullspace(DTensor<T> &a) {
    const size_t m = a.numRows(), n = a.numCols(), nMats = a.numMats();
    if (m > n) throw std::invalid_argument("[nullspace] I was expecting a square or fat matrix");
    m_nullspace = std::make_unique<DTensor<T> >(n, n, nMats);
    m_projOp = std::make_unique<DTensor<T> >(n, n, nMats);
    auto aTranspose = a.tr();   /* carries a's stream: the whole constructor runs on ONE stream */
    Svd<T> svd(aTranspose, true);
    svd.factorise();
    DTensor<unsigned int> const &devRank = svd.rank();
    std::shared_ptr<DTensor<T> > U = svd.leftSingularVectors().value();
    const int s = (int) aTranspose.streamIdx();
    /* N_i = last (n - rank_i) columns of U_i, moved to the front, zero elsewhere; N_i N_i' = I - U1 U1' with the side that has
     * fewer columns multiplied out, per matrix. One call: the packing runs beside the projector on a private stream */
    gpuErrChk(gpub200::Abi<T>::nullspace_build(gpub200::ctx(), s, n, U->raw(), n * n, devRank.raw(), m_nullspace->raw(), n * n,
                                               m_projOp->raw(), n * n, nMats));
    /* U and the rank tensor die with `svd` at scope exit: wait for the two launches that read them */
    Session::getInstance().synchronizeStream(aTranspose.streamIdx());
}

template<typename T>
TEMPLATE_CONSTRAINT_REQUIRES_FPX
inline void Nullspace<T>::project(DTensor<T> &b) {
    b.addAB(*m_projOp, b, 1, 0);
}

/* ================================================================================================
 *  CHOLESKY BATCH FACTORISATION (CBF)
 * ================================================================================================ */

TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
class CholeskyBatchFactoriser : public IStatus {
private:
    DTensor<T> *m_matrix;              ///< matrices to factorise, or their lower Cholesky factors (not owned)
    size_t m_numRows = 0;
    size_t m_numMats = 0;
    bool m_factorisationDone = false;

public:
    CholeskyBatchFactoriser() = delete;

    /**
     * @param A matrices to factorise, or precomputed lower-triangular factors
     * @param factorised true if A already holds the factors
     */
    CholeskyBatchFactoriser(DTensor<T> &A, bool factorised = false) : IStatus(A.numMats()),
                                                                      m_factorisationDone(factorised) {
        if (A.numRows() != A.numCols()) throw std::invalid_argument("[CholeskyBatch] A must be square");
        m_matrix = &A;
        m_numRows = A.numRows();
        m_numMats = A.numMats();
    }

    /** A_i = L_i L_i' for every matrix, in place (lower). */
    void factorise() {
        if (m_factorisationDone) return;
        gpuErrChk(gpub200::Abi<T>::potrf(gpub200::ctx(), (int) m_matrix->streamIdx(), m_numRows, m_matrix->raw(),
                                         m_numRows, m_numRows * m_numRows, m_info->raw(), m_numMats));
        m_factorisationDone = true;
    }

    /** Solves A_i x_i = b_i in place (one right-hand side per matrix). */
    void solve(DTensor<T> &b) {
        if (!m_factorisationDone) throw std::logic_error("[CholeskyBatchSolve] no factor to solve with");
        if (m_numRows != b.numRows() || m_numMats != b.numMats()) {
            throw std::invalid_argument("[CholeskyBatchSolve] A and b incompatible");
        }
        if (b.numCols() != 1) throw std::invalid_argument("[CholeskyBatchSolve] only supports `b` with one column");
        gpuErrChk(gpub200::Abi<T>::potrs(gpub200::ctx(), (int) m_matrix->streamIdx(), m_numRows, m_matrix->raw(),
                                         m_numRows, m_numRows * m_numRows, b.raw(), m_numRows, m_numMats));
    }

    /**
     * Additive API: the whole job from host memory at the speed of the host link. The reference offers whole-tensor calls only
     * (upload, factorise, solve, download: tensor.cuh:1128-1154, 2135-2197), which serialise three transfers and two kernels;
     * here the batch is cut into `chunks` pieces that flow through three streams, so the kernels and the download of one piece
     * hide behind the upload of the next. hostA holds the matrices (numRows^2 values each, column-major, mats slowest), hostB the
     * right-hand sides; pinned host memory is DMA'd directly, pageable memory is staged. On return the factorised tensor holds
     * the factors (as after factorise()), `b` and hostX the solutions, info() -- and hostInfo, if given -- the status codes.
     * lowerTriangleOnly: the factorisation never reads above the diagonal, so the block of each matrix that lies wholly above it is
     * not transferred (pinned hostA, fp64 n >= 32 / fp32 n >= 64, even n): 75 % of the bytes; that block of the tensor keeps its
     * previous content instead of the input's.
     */
    void factoriseAndSolveFromHost(const T *hostA, const T *hostB, DTensor<T> &b, T *hostX, int *hostInfo = nullptr,
                                   size_t chunks = 16, bool lowerTriangleOnly = false) {
        if (m_factorisationDone) throw std::logic_error("[CholeskyBatch] already factorised");
        if (m_numRows != b.numRows() || m_numMats != b.numMats() || b.numCols() != 1)
            throw std::invalid_argument("[CholeskyBatch] A and b incompatible");
        if (!hostA || !hostB) throw std::invalid_argument("[CholeskyBatch] null host buffer");
        gpuErrChk(gpub200::Abi<T>::chol_from_host(gpub200::ctx(), (int) m_matrix->streamIdx(), m_numRows, m_matrix->raw(), b.raw(),
                                                  m_info->raw(), hostA, hostB, hostX, hostInfo, m_numMats,
                                                  chunks | (lowerTriangleOnly ? GPUB_LOWER_ONLY : 0)));
        m_factorisationDone = true;
    }
};

/* ================================================================================================
 *  GIVENS ANNIHILATOR
 * ================================================================================================ */

TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
class GivensAnnihilator {
private:
    DTensor<T> *m_matrix;
    std::unique_ptr<DTensor<T> > m_d_rhyp_cos_sin; ///< device {rhypot, cos, -sin}

    void init() {
        m_d_rhyp_cos_sin = std::make_unique<DTensor<T> >(3);
    }

public:
    GivensAnnihilator() {
        init();
    }

    GivensAnnihilator(DTensor<T> &a) {
        if (a.numMats() > 1) {
            throw std::invalid_argument("[GivensAnnihilator] tensors (numMats > 1) not supported");
        }
        m_matrix = &a;
        init();
    }

    void setMatrix(DTensor<T> &a) {
        if (a.numMats() > 1) {
            throw std::invalid_argument("[GivensAnnihilator] tensors (numMats > 1) not supported");
        }
        m_matrix = &a;
    }

    /** Left Givens rotation G(i, k) that zeroes element (k, j). */
    void annihilate(size_t i, size_t k, size_t j);
};

/**
 * Additive API: Givens rotations on EVERY matrix of a tensor at once. The reference's applyLeft/RightGivensRotation and
 * GivensAnnihilator refuse tensors (numMats > 1: tensor.cuh:1076, 1090, 2229, 2240), so a batch costs numMats x (a one-thread kernel +
 * a cuBLAS rot); here annihilate() is one launch for the whole batch and gives, matrix by matrix, exactly what GivensAnnihilator gives.
 */
TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
class GivensBatchAnnihilator {
private:
    DTensor<T> *m_tensor;

public:
    GivensBatchAnnihilator() = delete;

    GivensBatchAnnihilator(DTensor<T> &a) : m_tensor(&a) {}

    void setTensor(DTensor<T> &a) { m_tensor = &a; }

    /** In every matrix: left Givens rotation G(i, k) that zeroes element (k, j). */
    void annihilate(size_t i, size_t k, size_t j) {
        const size_t nR = m_tensor->numRows(), nC = m_tensor->numCols();
        if (i >= nR or k >= nR or i == k) throw std::invalid_argument("[GivensBatchAnnihilator::annihilate] invalid row index");
        if (j >= nC) throw std::invalid_argument("[GivensBatchAnnihilator::annihilate] invalid column index j");
        gpuErrChk(gpub200::Abi<T>::annihilate_batched(gpub200::ctx(), (int) m_tensor->streamIdx(), m_tensor->raw(), nR, nC, nR * nC, i, k, j,
                                                      m_tensor->numMats()));
    }

    /** In every matrix b: rows i and j rotated with (c[b], minus_s[b]), device arrays of numMats values (cuBLAS rot convention). */
    void applyLeftGivensRotations(size_t i, size_t j, const T *c, const T *minus_s) {
        const size_t nR = m_tensor->numRows(), nC = m_tensor->numCols();
        if (i >= nR or j >= nR) throw std::invalid_argument("[GivensBatchAnnihilator] invalid row index");
        gpuErrChk(gpub200::Abi<T>::rot_batched(gpub200::ctx(), (int) m_tensor->streamIdx(), nC, m_tensor->raw() + i, nR, m_tensor->raw() + j, nR,
                                               nR * nC, c, minus_s, m_tensor->numMats()));
    }

    /** In every matrix b: columns i and j rotated with (c[b], minus_s[b]). */
    void applyRightGivensRotations(size_t i, size_t j, const T *c, const T *minus_s) {
        const size_t nR = m_tensor->numRows(), nC = m_tensor->numCols();
        if (i >= nC or j >= nC) throw std::invalid_argument("[GivensBatchAnnihilator] invalid column index");
        gpuErrChk(gpub200::Abi<T>::rot_batched(gpub200::ctx(), (int) m_tensor->streamIdx(), nR, m_tensor->raw() + i * nR, 1,
                                               m_tensor->raw() + j * nR, 1, nR * nC, c, minus_s, m_tensor->numMats()));
    }
};

/** Kept for source compatibility with the reference header (annihilate() uses gpub_givens_rhypot_*). */
TEMPLATE_WITH_TYPE_T
TEMPLATE_CONSTRAINT_REQUIRES_FPX
__global__ void k_givensAnnihilateRHypot(const T *data, T *res, size_t i, size_t k, size_t j, size_t nRows) {
    T xij = data[i + j * nRows];
    T xkj = data[k + j * nRows];
    res[0] = rhypot(xij, xkj);
    res[1] = xij * (*res);
    res[2] = xkj * (*res);
}

template<typename T>
TEMPLATE_CONSTRAINT_REQUIRES_FPX
inline void GivensAnnihilator<T>::annihilate(size_t i, size_t k, size_t j) {
    const size_t nR = m_matrix->numRows(), nC = m_matrix->numCols();
    if (i >= nR or k >= nR) throw std::invalid_argument("[GivensAnnihilator::annihilate] invalid row index");
    if (j >= nC) throw std::invalid_argument("[GivensAnnihilator::annihilate] invalid column index j");
    T *aux = m_d_rhyp_cos_sin->raw();
    /* cos and -sin stay on the device: the rotation kernel reads them there (no download) */
    gpuErrChk(gpub200::Abi<T>::rhypot(gpub200::ctx(), (int) m_matrix->streamIdx(), m_matrix->raw(), aux, i, k, j, nR));
    m_matrix->applyLeftGivensRotation(i, k, aux + 1, aux + 2);
}

#endif /* GPUB200_FACTORISERS_CUH */
