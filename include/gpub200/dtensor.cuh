/*
 * dtensor.cuh -- DTensor<T>: (rows x cols x mats) device tensor, column-major, mats axis slowest.
 *
 * Public surface identical to the reference (ref: tensor.cuh:271-661). Storage stays logically dense
 * (element (i,j,k) at i + rows*(j + cols*k)) because the reference's tests pin slices that alias the
 * parent, O(1) reshape of views and dense download order (testTensor.cu:287-477). The kernels do not need
 * padding to coalesce: they stage contiguous multi-matrix chunks with 128-bit accesses, and the C ABI
 * takes explicit leading dimensions and batch strides so padded layouts work as well.
 *
 * What changed underneath (every numerical method is one launch into libgputils_b200):
 *   - pointer table filled by a device kernel instead of a host vector + blocking H2D copy;
 *   - tr(): one batched transpose instead of numMats geam calls;
 *   - getRows(): one gather launch instead of one strided copy per row;
 *   - addAB / leastSquaresBatched / reductions / scal / axpy / rot: sm_100a kernels, base + stride
 *     addressing (the pointer table is kept only because ptrMatrices() is public API).
 */
#ifndef GPUB200_DTENSOR_CUH
#define GPUB200_DTENSOR_CUH

#include "core.cuh"

/* ================================================================================================
 *  TENSOR
 * ================================================================================================ */

/** Storage mode of host data handed to upload() / the vector constructor. */
enum StorageMode {
    columnMajor, ///< column major (the device layout)
    rowMajor,    ///< row major (transposed on the host before upload)
    defaultMajor = columnMajor
};

TEMPLATE_WITH_TYPE_T
class DTensor {
private:
    T *m_d_data = nullptr;           ///< device data
    T **m_d_ptrMatrices = nullptr;   ///< device table of pointers to each matrix (public API only)
    size_t m_numRows = 0;
    size_t m_numCols = 0;
    size_t m_numMats = 0;
    bool m_doDestroyData = false;         ///< owns m_d_data
    bool m_doDestroyPtrMatrices = false;  ///< owns m_d_ptrMatrices
    size_t m_idxStream = 0;

    size_t bytes() const { return m_numRows * m_numCols * m_numMats * sizeof(T); }

    void destroy() {
        if (m_doDestroyData) {
            if (m_d_data) gpuErrChk(Session::getInstance().cudaRelease(m_d_data));
            m_d_data = nullptr;
            Session::getInstance().adjustAllocatedBytes(-static_cast<long long>(bytes()));
            m_doDestroyData = false;
        }
        if (m_doDestroyPtrMatrices) {
            if (m_d_ptrMatrices) gpuErrChk(Session::getInstance().cudaRelease(m_d_ptrMatrices));
            m_d_ptrMatrices = nullptr;
            Session::getInstance().adjustAllocatedBytes(-static_cast<long long>(m_numMats * sizeof(T *)));
            m_doDestroyPtrMatrices = false;
        }
    }

    void allocateOnDevice(size_t size, bool zero = false);

    /** Host-side row-major -> column-major reordering, matrix by matrix. */
    void rm2cm(const std::vector<T> &rm, std::vector<T> &cm) const {
        const size_t perMat = m_numRows * m_numCols;
        for (size_t k = 0; k < m_numMats; k++) {
            const T *src = rm.data() + k * perMat;
            T *dst = cm.data() + k * perMat;
            for (size_t c = 0; c < m_numCols; c++)
                for (size_t r = 0; r < m_numRows; r++) dst[r + c * m_numRows] = src[c + r * m_numCols];
        }
    }

    std::ostream &print(std::ostream &out) const;

    void initialisePointersToMatricesData();

public:
    DTensor<T> setStreamIdx(size_t);

    size_t streamIdx() const { return m_idxStream; }

    static DTensor<T> createRandomTensor(size_t numRows, size_t numCols, size_t numMats, T low, T hi);

    static DTensor<T> parseFromFile(std::string path_to_file, StorageMode mode = StorageMode::defaultMajor);

    DTensor() = default;

    ~DTensor() { destroy(); }

    DTensor(size_t m, size_t n = 1, size_t k = 1, bool zero = false);

    DTensor(const std::vector<T> &data, size_t m, size_t n = 1, size_t k = 1,
            StorageMode mode = StorageMode::defaultMajor);

    DTensor(const DTensor &other);

    DTensor(DTensor &&other);

    /** Slice (view): axis 0 = rows of column 0, 1 = columns of matrix 0, 2 = matrices; `to` inclusive. */
    DTensor(const DTensor &other, size_t axis, size_t from, size_t to);

    T *raw() const;

    T **ptrMatrices() const;

    size_t numRows() const;

    size_t numCols() const;

    size_t numMats() const;

    size_t numEl() const;

    bool upload(const std::vector<T> &vec, StorageMode mode = StorageMode::defaultMajor);

    void download(std::vector<T> &vec) const;

    void deviceCopyTo(DTensor<T> &other) const;

    DTensor<T> getRows(size_t rowsFrom, size_t rowsTo, size_t matIdx) const;

    DTensor<T> tr() const;

    T dotF(const DTensor &other);

    T normF() const;

    T sumAbs() const;

    T maxAbs() const;

    T minAbs() const;

    void applyRightGivensRotation(size_t i, size_t j, const T *c, const T *minus_s);

    void applyLeftGivensRotation(size_t i, size_t j, const T *c, const T *minus_s);

    void leastSquaresBatched(DTensor &b);

    void addAB(const DTensor<T> &A, const DTensor<T> &B, T alpha = 1, T beta = 0);

    void reshape(size_t newNumRows, size_t newNumCols, size_t newNumMats = 1);

    void saveToFile(std::string pathToFile);

    /* ------------- OPERATORS ------------- */

    DTensor &operator=(const DTensor &other);

    T operator()(size_t i, size_t j = 0, size_t k = 0) const;

    DTensor &operator*=(T scalar);

    DTensor &operator+=(const DTensor &rhs);

    DTensor &operator-=(const DTensor &rhs);

    /* ------------- FRIENDS ------------- */

    friend DTensor operator+(DTensor &first, const DTensor &second) {
        DTensor result(first);
        result += second;
        return result;
    }

    friend DTensor operator-(DTensor &first, const DTensor &second) {
        DTensor result(first);
        result -= second;
        return result;
    }

    friend DTensor<T> operator*(DTensor &A, DTensor &B) {
        DTensor<T> result(A.m_numRows, B.m_numCols, B.m_numMats);
        result.addAB(A, B);
        return result;
    }

    friend DTensor<T> operator*(T a, DTensor &B) {
        DTensor<T> result(B);
        result *= a;
        return result;
    }

    friend std::ostream &operator<<(std::ostream &out, const DTensor<T> &data) {
        return data.print(out);
    }
}; /* END OF DTENSOR */

/* ------------------------------------------------------------------------------------------------
 *  storage
 * ------------------------------------------------------------------------------------------------ */

template<typename T>
DTensor<T> DTensor<T>::setStreamIdx(size_t idx) {
    if (idx >= s_numStreams) {
        throw std::invalid_argument("Invalid stream index; it exceeds the max allocated streams");
    }
    m_idxStream = idx;
    return *this;
}

template<typename T>
inline void DTensor<T>::allocateOnDevice(size_t size, bool zero) {
    if (size == 0) return;
    destroy();
    const size_t nbytes = size * sizeof(T);
    gpuErrChk(Session::getInstance().cudaAllocate(reinterpret_cast<void **>(&m_d_data), nbytes));
    m_doDestroyData = true;
    /* on the legacy stream, like the allocation: ordered against every blocking stream, asynchronous to the host */
    if (zero) gpuErrChk(cudaMemsetAsync(m_d_data, 0, nbytes, cudaStreamLegacy));
    if (m_numMats > 1) {
        cudaError_t st = Session::getInstance().cudaAllocate(reinterpret_cast<void **>(&m_d_ptrMatrices),
                                                             m_numMats * sizeof(T *));
        if (st != cudaSuccess) {
            gpuErrChk(Session::getInstance().cudaRelease(m_d_data));
            gpuErrChk(st);
        }
        m_doDestroyPtrMatrices = true;
    }
}

template<typename T>
void DTensor<T>::initialisePointersToMatricesData() {
    if (m_numMats <= 1 || !m_d_ptrMatrices || !m_doDestroyPtrMatrices) return;
    /* filled on the device: table[i] = data + i * rows * cols */
    gpuErrChk(gpub_fill_ptr_table(gpub200::ctx(), static_cast<int>(m_idxStream), m_d_data,
                                  m_numRows * m_numCols * sizeof(T), m_numMats,
                                  reinterpret_cast<void **>(m_d_ptrMatrices)));
}

template<typename T>
DTensor<T>::DTensor(size_t m, size_t n, size_t k, bool zero)
    : m_numRows(m), m_numCols(n), m_numMats(k) {
    allocateOnDevice(m * n * k, zero);
    initialisePointersToMatricesData();
}

template<typename T>
DTensor<T>::DTensor(const std::vector<T> &data, size_t m, size_t n, size_t k, StorageMode mode)
    : m_numRows(m), m_numCols(n), m_numMats(k) {
    allocateOnDevice(m * n * k);
    upload(data, mode);
    initialisePointersToMatricesData();
}

template<typename T>
DTensor<T>::DTensor(const DTensor<T> &other)
    : m_numRows(other.m_numRows), m_numCols(other.m_numCols), m_numMats(other.m_numMats),
      m_idxStream(other.m_idxStream) {
    allocateOnDevice(numEl());
    if (numEl() > 0) gpuErrChk(cudaMemcpyAsync(m_d_data, other.raw(), bytes(), cudaMemcpyDeviceToDevice, cudaStreamLegacy));
    initialisePointersToMatricesData();
}

template<typename T>
DTensor<T>::DTensor(const DTensor<T> &other, size_t axis, size_t from, size_t to) {
    if (from > to) throw std::invalid_argument("from > to");
    const size_t len = to - from + 1;
    size_t offset = 0;
    switch (axis) {
        case 2:
            offset = other.m_numRows * other.m_numCols * from;
            m_numRows = other.m_numRows;
            m_numCols = other.m_numCols;
            m_numMats = len;
            m_d_ptrMatrices = other.m_d_ptrMatrices ? other.m_d_ptrMatrices + from : nullptr;
            break;
        case 1:
            offset = other.m_numRows * from;
            m_numRows = other.m_numRows;
            m_numCols = len;
            m_numMats = 1;
            break;
        case 0:
            offset = from;
            m_numRows = len;
            m_numCols = 1;
            m_numMats = 1;
            break;
        default:
            break;
    }
    m_d_data = other.m_d_data + offset;
    m_idxStream = other.m_idxStream;
}

template<typename T>
DTensor<T>::DTensor(DTensor<T> &&other) {
    m_d_data = other.m_d_data;
    m_d_ptrMatrices = other.m_d_ptrMatrices;
    m_numRows = other.m_numRows;
    m_numCols = other.m_numCols;
    m_numMats = other.m_numMats;
    m_doDestroyData = other.m_doDestroyData;
    m_doDestroyPtrMatrices = other.m_doDestroyPtrMatrices;
    m_idxStream = other.m_idxStream;
    other.m_d_data = nullptr;
    other.m_d_ptrMatrices = nullptr;
    other.m_numRows = other.m_numCols = other.m_numMats = 0;
    other.m_doDestroyData = other.m_doDestroyPtrMatrices = false;
}

template<typename T>
void DTensor<T>::reshape(size_t newNumRows, size_t newNumCols, size_t newNumMats) {
    if (m_numRows == newNumRows && m_numCols == newNumCols && m_numMats == newNumMats) return;
    const size_t newNumElements = newNumRows * newNumCols * newNumMats;
    if (numEl() != newNumElements) {
        char msg[256];
        snprintf(msg, sizeof(msg),
                 "DTensor[%zu x %zu x %zu] with %zu elements cannot be reshaped into DTensor[%zu x %zu x %zu] (%zu elements)",
                 numRows(), numCols(), numMats(), numEl(), newNumRows, newNumCols, newNumMats, newNumElements);
        throw std::invalid_argument(msg);
    }
    /* the pointer table is reallocated only when it has to grow */
    if (newNumMats > m_numMats) {
        if (m_d_ptrMatrices && m_doDestroyPtrMatrices) {
            gpuErrChk(Session::getInstance().cudaRelease(m_d_ptrMatrices));
            Session::getInstance().adjustAllocatedBytes(-static_cast<long long>(m_numMats * sizeof(T *)));
        }
        m_d_ptrMatrices = nullptr;
        m_doDestroyPtrMatrices = false;
        if (newNumMats > 1) {
            gpuErrChk(Session::getInstance().cudaAllocate(reinterpret_cast<void **>(&m_d_ptrMatrices),
                                                          newNumMats * sizeof(T *)));
            m_doDestroyPtrMatrices = true;
        }
    }
    m_numRows = newNumRows;
    m_numCols = newNumCols;
    m_numMats = newNumMats;
    initialisePointersToMatricesData();
}

template<typename T>
inline size_t DTensor<T>::numRows() const { return m_numRows; }

template<typename T>
inline size_t DTensor<T>::numCols() const { return m_numCols; }

template<typename T>
inline size_t DTensor<T>::numMats() const { return m_numMats; }

template<typename T>
inline size_t DTensor<T>::numEl() const { return m_numRows * m_numCols * m_numMats; }

template<typename T>
inline T *DTensor<T>::raw() const { return m_d_data; }

template<typename T>
inline T **DTensor<T>::ptrMatrices() const { return m_d_ptrMatrices; }

/* ------------------------------------------------------------------------------------------------
 *  host <-> device
 *  Blocking, as in the reference (synchronous cudaMemcpy, tensor.cuh:1128-1154): "after a method returns, a later
 *  download observes it" holds across streams. gpub_upload / gpub_download queue the copy on the tensor's stream,
 *  stage pageable host memory through the context's pinned ring with several host threads (the DMA of one 8 MB piece
 *  overlaps the staging of the next) and return when the data has arrived. A download first joins the other blocking
 *  streams through the legacy stream, like cudaMemcpy did.
 * ------------------------------------------------------------------------------------------------ */

namespace gpub200 {
/* makes stream `idx` of the current device's context wait for everything queued on the Session's other blocking streams */
inline void joinStreams(size_t idx) {
    if (Session::getInstance().numStreams() <= 1) return;
    cudaEvent_t ev;
    gpuErrChk(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    gpuErrChk(cudaEventRecord(ev, cudaStreamLegacy));   /* the legacy stream waits for all blocking streams */
    gpuErrChk(cudaStreamWaitEvent(Session::getInstance().streamOfCurrentDevice(idx), ev, 0));
    gpuErrChk(cudaEventDestroy(ev));
}
} // namespace gpub200

template<typename T>
inline bool DTensor<T>::upload(const std::vector<T> &vec, StorageMode mode) {
    if (vec.size() != numEl()) throw std::invalid_argument("[upload] vec has wrong size");
    if (vec.empty()) return true;
    gpub200::joinStreams(m_idxStream);   /* earlier readers / writers of this tensor on other streams finish first */
    if (mode == StorageMode::rowMajor) {
        std::vector<T> cm(vec.size());
        rm2cm(vec, cm);
        gpuErrChk(gpub_upload(gpub200::ctx(), (int) m_idxStream, m_d_data, cm.data(), bytes()));
    } else {
        /* no host-side copy of the vector (the reference makes one, tensor.cuh:1134-1139) */
        gpuErrChk(gpub_upload(gpub200::ctx(), (int) m_idxStream, m_d_data, vec.data(), bytes()));
    }
    return true;
}

template<typename T>
inline void DTensor<T>::download(std::vector<T> &vec) const {
    vec.resize(numEl());
    if (vec.empty()) return;
    gpub200::joinStreams(m_idxStream);
    gpuErrChk(gpub_download(gpub200::ctx(), (int) m_idxStream, vec.data(), m_d_data, bytes()));
}

template<typename T>
inline void DTensor<T>::deviceCopyTo(DTensor<T> &elsewhere) const {
    if (elsewhere.numEl() < numEl()) {
        throw std::invalid_argument("[deviceCopyTo] tensor does not fit into destination");
    }
    if (numEl() == 0) return;
    gpuErrChk(cudaMemcpy(elsewhere.raw(), m_d_data, bytes(), cudaMemcpyDeviceToDevice));
}

template<typename T>
inline T DTensor<T>::operator()(size_t i, size_t j, size_t k) const {
    T host;
    const size_t offset = i + m_numRows * (j + m_numCols * k);
    gpuErrChk(cudaMemcpy(&host, m_d_data + offset, sizeof(T), cudaMemcpyDeviceToHost));
    return host;
}

template<typename T>
DTensor<T> &DTensor<T>::operator=(const DTensor<T> &other) {
    /* Shallow alias, like the reference (ref: tensor.cuh:1219-1228): the left-hand side becomes a
     * non-owning view of `other`. Deliberate fixes: buffers this tensor owned are released instead of
     * leaked, and the pointer table is aliased too so ptrMatrices() stays consistent with raw(). */
    if (this == &other) return *this;
    destroy();
    m_numMats = other.m_numMats;
    m_numRows = other.m_numRows;
    m_numCols = other.m_numCols;
    m_d_data = other.m_d_data;
    m_d_ptrMatrices = other.m_d_ptrMatrices;
    m_doDestroyData = false;
    m_doDestroyPtrMatrices = false;
    m_idxStream = other.m_idxStream;
    return *this;
}

/* ------------------------------------------------------------------------------------------------
 *  random / files
 * ------------------------------------------------------------------------------------------------ */

template<typename T>
DTensor<T> DTensor<T>::createRandomTensor(size_t numRows, size_t numCols, size_t numMats, T low, T hi) {
    if constexpr (std::is_floating_point<T>::value) {
        auto randVec = generateRealRandomVector<T>(numRows * numCols * numMats, low, hi);
        return DTensor<T>(randVec, numRows, numCols, numMats);
    } else if constexpr (std::is_same_v<T, int>) {
        auto randVec = generateIntRandomVector(numRows * numCols * numMats, low, hi);
        return DTensor<T>(randVec, numRows, numCols, numMats);
    } else {
        throw std::invalid_argument("[createRandomTensor] unsupported type T");
    }
}

template<typename T>
struct data_t {
    size_t numRows;
    size_t numCols;
    size_t numMats;
    std::vector<T> data;
};

/** Text format: rows, cols, mats on three lines, then one value per line (column-major, mats slowest). */
template<typename T>
data_t<T> vectorFromTextFile(std::string path_to_file) {
    std::ifstream file(path_to_file, std::ios::in);
    if (!file.is_open()) throw std::invalid_argument("[vectorFromTextFile] the file does not exist");
    data_t<T> out;
    std::string line;
    std::getline(file, line);
    out.numRows = std::strtoull(line.c_str(), nullptr, 10);
    std::getline(file, line);
    out.numCols = std::strtoull(line.c_str(), nullptr, 10);
    std::getline(file, line);
    out.numMats = std::strtoull(line.c_str(), nullptr, 10);
    const size_t count = out.numRows * out.numCols * out.numMats;
    out.data.resize(count);
    size_t i = 0;
    while (i < count && std::getline(file, line)) {
        if constexpr (std::is_same_v<T, int>) out.data[i] = std::atoi(line.c_str());
        else if constexpr (std::is_same_v<T, double>) out.data[i] = std::stod(line);
        else if constexpr (std::is_same_v<T, float>) out.data[i] = std::stof(line);
        else if constexpr (std::is_same_v<T, long double>) out.data[i] = std::stold(line);
        else if constexpr (std::is_same_v<T, long>) out.data[i] = std::stol(line);
        else if constexpr (std::is_same_v<T, long long>) out.data[i] = std::stoll(line);
        else if constexpr (std::is_same_v<T, unsigned long>) out.data[i] = std::stoul(line);
        else if constexpr (std::is_same_v<T, unsigned long long>) out.data[i] = std::stoull(line);
        else throw std::invalid_argument("data type not supported");
        i++;
    }
    return out;
}

/** Binary .bt format: three little-endian uint64 (rows, cols, mats) + raw column-major payload. */
template<typename T>
data_t<T> vectorFromBinaryFile(std::string path_to_file) {
    std::ifstream file(path_to_file, std::ios::binary);
    if (!file.is_open()) throw std::invalid_argument("[vectorFromBinaryFile] the file does not exist");
    uint64_t dims[3] = {0, 0, 0};
    file.read(reinterpret_cast<char *>(dims), sizeof(dims));
    data_t<T> out;
    out.numRows = dims[0];
    out.numCols = dims[1];
    out.numMats = dims[2];
    out.data.resize(dims[0] * dims[1] * dims[2]);
    /* one block read instead of one read() per element */
    file.read(reinterpret_cast<char *>(out.data.data()), static_cast<std::streamsize>(out.data.size() * sizeof(T)));
    return out;
}

template<typename T>
DTensor<T> DTensor<T>::parseFromFile(std::string path_to_file, StorageMode mode) {
    const bool binary = path_to_file.size() >= 3 && path_to_file.compare(path_to_file.size() - 3, 3, ".bt") == 0;
    data_t<T> parsed = binary ? vectorFromBinaryFile<T>(path_to_file) : vectorFromTextFile<T>(path_to_file);
    return DTensor<T>(parsed.data, parsed.numRows, parsed.numCols, parsed.numMats, mode);
}

template<typename T>
void DTensor<T>::saveToFile(std::string pathToFile) {
    std::vector<T> host;
    download(host);
    const bool binary = pathToFile.size() >= 3 && pathToFile.compare(pathToFile.size() - 3, 3, ".bt") == 0;
    if (binary) {
        const uint64_t dims[3] = {(uint64_t) numRows(), (uint64_t) numCols(), (uint64_t) numMats()};
        std::ofstream file(pathToFile, std::ios::binary);
        file.write(reinterpret_cast<const char *>(dims), sizeof(dims));
        file.write(reinterpret_cast<const char *>(host.data()), static_cast<std::streamsize>(host.size() * sizeof(T)));
    } else {
        std::ofstream file(pathToFile);
        file << numRows() << '\n' << numCols() << '\n' << numMats() << '\n';
        if constexpr (std::is_floating_point<T>::value) {
            file << std::setprecision(std::numeric_limits<T>::max_digits10);
        }
        for (const T &el: host) file << el << '\n';
    }
}

template<typename T>
std::ostream &DTensor<T>::print(std::ostream &out) const {
    out << "Tensor [" << m_numRows << " x " << m_numCols << " x " << m_numMats << "]:" << std::endl;
    std::vector<T> host;
    download(host);
    for (size_t k = 0; k < m_numMats; k++) {
        out << ">> layer: " << k << std::endl;
        for (size_t i = 0; i < m_numRows; i++) {
            for (size_t j = 0; j < m_numCols; j++) {
                out << std::setw(10) << host[m_numRows * (m_numCols * k + j) + i] << ", ";
            }
            out << std::endl;
        }
    }
    return out;
}

/* ------------------------------------------------------------------------------------------------
 *  numerical methods (float / double): one launch each
 * ------------------------------------------------------------------------------------------------ */

#define GPUB200_FP_ONLY(T) static_assert(std::is_floating_point<T>::value, "this DTensor method needs float or double")

template<typename T>
inline T DTensor<T>::dotF(const DTensor<T> &other) {
    GPUB200_FP_ONLY(T);
    if (m_numRows != other.m_numRows || m_numCols != other.m_numCols || m_numMats != other.m_numMats)
        throw std::invalid_argument("[dotF] incompatible dimensions");
    T result;
    gpuErrChk(gpub200::Abi<T>::dot(gpub200::ctx(), (int) m_idxStream, numEl(), raw(), other.raw(), &result));
    return result;
}

template<typename T>
inline T DTensor<T>::normF() const {
    GPUB200_FP_ONLY(T);
    T result;
    gpuErrChk(gpub200::Abi<T>::nrm2(gpub200::ctx(), (int) m_idxStream, numEl(), m_d_data, &result));
    return result;
}

template<typename T>
inline T DTensor<T>::sumAbs() const {
    GPUB200_FP_ONLY(T);
    T result;
    gpuErrChk(gpub200::Abi<T>::asum(gpub200::ctx(), (int) m_idxStream, numEl(), m_d_data, &result));
    return result;
}

template<typename T>
inline T DTensor<T>::maxAbs() const {
    GPUB200_FP_ONLY(T);
    T result;
    gpuErrChk(gpub200::Abi<T>::amax(gpub200::ctx(), (int) m_idxStream, numEl(), m_d_data, &result, nullptr));
    return result;
}

template<typename T>
inline T DTensor<T>::minAbs() const {
    GPUB200_FP_ONLY(T);
    T result;
    gpuErrChk(gpub200::Abi<T>::amin(gpub200::ctx(), (int) m_idxStream, numEl(), m_d_data, &result, nullptr));
    return result;
}

namespace gpub200 {
/* c / s may live on the host or on the device (GivensAnnihilator passes device pointers) */
inline bool isDevicePointer(const void *p) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}
} // namespace gpub200

template<typename T>
void DTensor<T>::applyRightGivensRotation(size_t i, size_t j, const T *c, const T *minus_s) {
    if (m_numMats > 1) throw std::invalid_argument("[applyRightGivensRotation] tensors (nMat>1) not supported");
    if constexpr (std::is_floating_point<T>::value) {
        T *col_i = m_d_data + i * m_numRows;
        T *col_j = m_d_data + j * m_numRows;
        gpuErrChk(gpub200::Abi<T>::rot(gpub200::ctx(), (int) m_idxStream, m_numRows, col_i, 1, col_j, 1, c, minus_s,
                                       gpub200::isDevicePointer(c) ? 1 : 0));
    } else {
        throw std::invalid_argument("[applyRightGivensRotation] Unsupported type T");
    }
}

template<typename T>
void DTensor<T>::applyLeftGivensRotation(size_t i, size_t j, const T *c, const T *minus_s) {
    if (m_numMats > 1) throw std::invalid_argument("[applyLeftGivensRotation] tensors (nMat>1) not supported");
    if constexpr (std::is_floating_point<T>::value) {
        gpuErrChk(gpub200::Abi<T>::rot(gpub200::ctx(), (int) m_idxStream, m_numCols, m_d_data + i, m_numRows,
                                       m_d_data + j, m_numRows, c, minus_s, gpub200::isDevicePointer(c) ? 1 : 0));
    } else {
        throw std::invalid_argument("[applyLeftGivensRotation] Unsupported type T");
    }
}

template<typename T>
inline DTensor<T> DTensor<T>::tr() const {
    GPUB200_FP_ONLY(T);
    DTensor<T> transposes(m_numCols, m_numRows, m_numMats);
    /* the result carries the stream that fills it, so whatever the caller queues on it next is ordered behind the transpose
     * (the reference hands back a stream-0 tensor filled on stream m_idxStream: Nullspace then mixes two unordered streams) */
    transposes.m_idxStream = m_idxStream;
    const size_t perMat = m_numRows * m_numCols;
    gpuErrChk(gpub200::Abi<T>::transpose(gpub200::ctx(), (int) m_idxStream, m_numRows, m_numCols, raw(), perMat,
                                         transposes.raw(), perMat, m_numMats));
    return transposes;
}

template<typename T>
inline DTensor<T> DTensor<T>::getRows(size_t rowsFrom, size_t rowsTo, size_t matIdx) const {
    GPUB200_FP_ONLY(T);
    const size_t len = rowsTo - rowsFrom + 1;
    DTensor<T> rowsOnly(len, m_numCols, 1);
    rowsOnly.m_idxStream = m_idxStream;
    gpuErrChk(gpub200::Abi<T>::gather_rows(gpub200::ctx(), (int) m_idxStream, raw() + matIdx * m_numRows * m_numCols,
                                           m_numRows, rowsFrom, len, m_numCols, rowsOnly.raw()));
    return rowsOnly;
}

template<typename T>
inline DTensor<T> &DTensor<T>::operator*=(T scalar) {
    GPUB200_FP_ONLY(T);
    gpuErrChk(gpub200::Abi<T>::scal(gpub200::ctx(), (int) m_idxStream, numEl(), scalar, m_d_data));
    return *this;
}

template<typename T>
inline DTensor<T> &DTensor<T>::operator+=(const DTensor<T> &rhs) {
    GPUB200_FP_ONLY(T);
    gpuErrChk(gpub200::Abi<T>::axpy(gpub200::ctx(), (int) m_idxStream, numEl(), T(1), rhs.m_d_data, m_d_data));
    return *this;
}

template<typename T>
inline DTensor<T> &DTensor<T>::operator-=(const DTensor<T> &rhs) {
    GPUB200_FP_ONLY(T);
    gpuErrChk(gpub200::Abi<T>::axpy(gpub200::ctx(), (int) m_idxStream, numEl(), T(-1), rhs.m_d_data, m_d_data));
    return *this;
}

template<typename T>
inline void DTensor<T>::addAB(const DTensor<T> &A, const DTensor<T> &B, T alpha, T beta) {
    GPUB200_FP_ONLY(T);
    /* dimensions are taken from A and B unchecked, exactly as the reference does (tensor.cuh:1288-1291) */
    const size_t nMat = A.numMats(), nRA = A.numRows(), nCA = A.numCols(), nCB = B.numCols();
    gpuErrChk(gpub200::Abi<T>::gemm(gpub200::ctx(), (int) m_idxStream, nRA, nCB, nCA, alpha,
                                    A.raw(), nRA, nRA * nCA,
                                    B.raw(), nCA, nCA * nCB, beta,
                                    raw(), nRA, nRA * nCB, nMat));
}

template<typename T>
inline void DTensor<T>::leastSquaresBatched(DTensor<T> &B) {
    GPUB200_FP_ONLY(T);
    const size_t batchSize = numMats();
    if (B.numRows() != m_numRows)
        throw std::invalid_argument("[Least squares batched] rhs rows does not equal lhs rows");
    if (B.numCols() != 1)
        throw std::invalid_argument("[Least squares batched] rhs are not vectors");
    if (B.numMats() != batchSize)
        throw std::invalid_argument("[Least squares batched] rhs numMats does not equal lhs numMats");
    if (m_numCols > m_numRows)
        throw std::invalid_argument("[Least squares batched] supports square or tall matrices only");
    /* no per-call info tensor: the reference allocated and freed a DTensor<int>(batch) here (tensor.cuh:1353) */
    gpuErrChk(gpub200::Abi<T>::gels(gpub200::ctx(), (int) m_idxStream, m_numRows, m_numCols, raw(), m_numRows,
                                    m_numRows * m_numCols, B.raw(), m_numRows, nullptr, batchSize));
}

#endif /* GPUB200_DTENSOR_CUH */
