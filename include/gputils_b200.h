/*
 * gputils_b200.h -- C ABI of libgputils_b200 (hand-written sm_100a kernels).
 *
 * This is the drop-in boundary for the batched linear-algebra hot path of
 * GPUtils.  Every entry point replaces one cuBLAS / cuSOLVER call site of the
 * reference header (include/tensor.cuh in the reference tree); the call site
 * is cited above each declaration as `ref: tensor.cuh:<line>`.
 *
 * Conventions
 *  - plain C types only: device pointers, sizes, strides (no torch / C++ types);
 *  - all matrices are column-major; `ld*` is the leading dimension in
 *    elements, `stride*` the element distance between consecutive matrices of
 *    a batch (the reference's layout is ld = rows, stride = rows*cols;
 *    padded strides are accepted everywhere);
 *  - every launcher is asynchronous on stream `sidx` of context `ctx`
 *    unless it returns a host scalar (the reductions), which blocks like the
 *    cuBLAS call it replaces;
 *  - launchers allocate nothing: scratch comes from the caller (`work`,
 *    sized by the matching `*_worksize`) or from the context's per-stream
 *    scratch (reductions only); device memory for tensors comes from the
 *    context's stream-ordered pool (gpub_mem_alloc);
 *  - return value: 0 on success, a positive `cudaError_t` code if the CUDA
 *    runtime failed, a negative GPUB_E* code for argument errors.  Launchers
 *    never throw, print or exit; the C++ header maps non-zero to the
 *    reference's print-and-exit convention (ref: tensor.cuh:87-113);
 *  - numerical failure is reported LAPACK-style through device `info` arrays
 *    (potrf: 0 = ok, i>0 = leading minor i not positive definite).
 */
#ifndef GPUTILS_B200_H
#define GPUTILS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPUB_OK 0
#define GPUB_EINVAL (-1)   /* bad argument (negative size, null pointer, ...) */
#define GPUB_ENOTSUP (-2)  /* shape outside what the kernels support         */
#define GPUB_EWORK (-3)    /* workspace too small                            */

typedef struct gpub_ctx *gpub_ctx_t;

/* ---- context: replaces the Session singleton's handle vectors ------------
 * ref: tensor.cuh:133-247 (Session), 154-166 (handle + stream creation)     */
const char *gpub_version(void);
/* per-device context (created on first use, lives until process exit) */
int gpub_ctx_get(int device, gpub_ctx_t *out);
/* make sure streams 0..n-1 exist (blocking streams, like cudaStreamCreate) */
int gpub_ctx_ensure_streams(gpub_ctx_t ctx, int n);
int gpub_ctx_num_streams(gpub_ctx_t ctx);
/* raw cudaStream_t of stream `sidx` (as void*) */
int gpub_ctx_stream(gpub_ctx_t ctx, int sidx, void **cuda_stream);
/* adopt an externally owned cudaStream_t as stream `sidx` (e.g. torch's) */
int gpub_ctx_bind_stream(gpub_ctx_t ctx, int sidx, void *cuda_stream);
/* ref: tensor.cuh:232-246 */
int gpub_ctx_sync(gpub_ctx_t ctx, int sidx);
int gpub_ctx_sync_all(gpub_ctx_t ctx);
/* ref: tensor.cuh:168-173 (~Session destroys its handles): synchronise, then free the streams this context
 * owns and the per-stream scratch. The context stays valid; slots are recreated lazily if used again. */
int gpub_ctx_release(gpub_ctx_t ctx);
/* the same for every context created so far (one per device used) */
int gpub_ctx_release_all(void);
int gpub_ctx_device(gpub_ctx_t ctx);
int gpub_ctx_sm_count(gpub_ctx_t ctx);

/* ---- memory: stream-ordered pool behind Session::cudaAllocate ----------------
 * ref: tensor.cuh:1106-1126 (two cudaMalloc per DTensor), 283-294 (two cudaFree), and the per-call allocations of tr() (1169),
 * the binary operators (634, 640) and the Nullspace loop (2076).
 * gpub_mem_alloc takes `bytes` from the context's pool, ordered on the LEGACY default stream: like cudaMalloc it is an ordering
 * point for every blocking stream of the context (and for synchronous cudaMemcpy), but it does not stop the host, and freed
 * blocks stay cached in the pool, so constructing / destroying tensors in a solver loop costs no driver allocation.
 * gpub_mem_free(ptr) returns a block to the pool it came from (any device may be current). Streams created with
 * cudaStreamNonBlocking (e.g. adopted with gpub_ctx_bind_stream) are NOT ordered by the legacy stream: order them yourself. */
int gpub_mem_alloc(gpub_ctx_t ctx, size_t bytes, void **ptr);
int gpub_mem_free(void *ptr);
/* bytes the pool holds from the driver / bytes currently handed out */
int gpub_mem_stats(gpub_ctx_t ctx, size_t *reserved_bytes, size_t *used_bytes);
/* give cached blocks back to the driver, keeping at most keep_bytes (synchronises the legacy stream) */
int gpub_mem_trim(gpub_ctx_t ctx, size_t keep_bytes);

/* ---- host <-> device copies ---------------------------------------------------
 * ref: tensor.cuh:1128-1145 (upload: host-side copy + one pageable cudaMemcpy), 1147-1154 (download).
 * Blocking like the cudaMemcpy they replace (on return the data has arrived), queued on stream `sidx`. Pinned host memory is
 * DMA'd directly; pageable memory is cut into 8 MB pieces that several host threads stage through a ring of pinned buffers
 * while the DMA of the previous piece is in flight. */
int gpub_upload(gpub_ctx_t ctx, int sidx, void *dst_dev, const void *src_host, size_t bytes);
int gpub_download(gpub_ctx_t ctx, int sidx, void *dst_host, const void *src_dev, size_t bytes);

/* ---- storage helpers ------------------------------------------------------
 * ref: tensor.cuh:672-688 (pointer table built on the host, then H2D)        */
int gpub_fill_ptr_table(gpub_ctx_t ctx, int sidx, void *base, size_t stride_bytes,
                        size_t count, void **table);

/* ---- BLAS-1 style flat ops (whole tensor as one vector) -------------------
 * ref: tensor.cuh:968-1072 (dot, nrm2, asum, iamax, iamin)                   */
int gpub_dot_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *x, const double *y, double *result_host);
int gpub_dot_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *x, const float *y, float *result_host);
/* nrm2: one pass, squares accumulated in fp64; if that sum over- or underflows (fp64 data beyond ~1e154 / below ~1e-154) a second,
 * scaled pass returns what cublasDnrm2's scaled algorithm returns */
int gpub_nrm2_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *x, double *result_host);
int gpub_nrm2_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *x, float *result_host);
int gpub_asum_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *x, double *result_host);
int gpub_asum_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *x, float *result_host);
/* max / min of |x_i|; `index_host` (may be NULL) gets the first 0-based index attaining it */
int gpub_amax_abs_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *x, double *result_host, long long *index_host);
int gpub_amax_abs_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *x, float *result_host, long long *index_host);
int gpub_amin_abs_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *x, double *result_host, long long *index_host);
int gpub_amin_abs_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *x, float *result_host, long long *index_host);
/* ref: tensor.cuh:1210-1276 (scal, axpy) */
int gpub_scal_f64(gpub_ctx_t ctx, int sidx, size_t n, double alpha, double *x);
int gpub_scal_f32(gpub_ctx_t ctx, int sidx, size_t n, float alpha, float *x);
int gpub_axpy_f64(gpub_ctx_t ctx, int sidx, size_t n, double alpha, const double *x, double *y);
int gpub_axpy_f32(gpub_ctx_t ctx, int sidx, size_t n, float alpha, const float *x, float *y);
/* ref: tensor.cuh:1074-1104 (rot).  c / s are DEVICE or HOST pointers:
 * `cs_on_device` != 0 mirrors CUBLAS_POINTER_MODE_DEVICE (tensor.cuh:2297)   */
int gpub_rot_f64(gpub_ctx_t ctx, int sidx, size_t n, double *x, size_t incx, double *y, size_t incy,
                 const double *c, const double *s, int cs_on_device);
int gpub_rot_f32(gpub_ctx_t ctx, int sidx, size_t n, float *x, size_t incx, float *y, size_t incy,
                 const float *c, const float *s, int cs_on_device);
/* ref: tensor.cuh:2272-2281 (k_givensAnnihilateRHypot): res = {rhypot, cos, -sin} */
int gpub_givens_rhypot_f64(gpub_ctx_t ctx, int sidx, const double *data, double *res, size_t i, size_t k, size_t j, size_t nrows);
int gpub_givens_rhypot_f32(gpub_ctx_t ctx, int sidx, const float *data, float *res, size_t i, size_t k, size_t j, size_t nrows);
/* Batched Givens (additive: the reference rotates one matrix per call and refuses tensors, tensor.cuh:1076, 1090, 2229).
 * rot_batched: x_b <- c_b x_b + s_b y_b, y_b <- c_b y_b - s_b x_b for every matrix b (x, y given for matrix 0, `stride` elements
 * between matrices), c and s DEVICE arrays of `batch` values.
 * givens_annihilate_batched: for every matrix the left rotation G(i, k) that zeroes element (k, j), built from elements (i, j) and
 * (k, j) exactly as k_givensAnnihilateRHypot (cos = x_ij rhypot, -sin = x_kj rhypot), applied to rows i and k: one launch.    */
int gpub_rot_batched_f64(gpub_ctx_t ctx, int sidx, size_t n, double *x, size_t incx, double *y, size_t incy, size_t stride,
                         const double *c, const double *s, size_t batch);
int gpub_rot_batched_f32(gpub_ctx_t ctx, int sidx, size_t n, float *x, size_t incx, float *y, size_t incy, size_t stride,
                         const float *c, const float *s, size_t batch);
int gpub_givens_annihilate_batched_f64(gpub_ctx_t ctx, int sidx, double *A, size_t nrows, size_t ncols, size_t strideA,
                                       size_t i, size_t k, size_t j, size_t batch);
int gpub_givens_annihilate_batched_f32(gpub_ctx_t ctx, int sidx, float *A, size_t nrows, size_t ncols, size_t strideA,
                                       size_t i, size_t k, size_t j, size_t batch);
/* ref: tensor.cuh:1396-1424 (getRows: one strided copy per row) -> one launch:
 * dst(r, c) = src(row_from + r, c) for r < nrows_out, c < ncols             */
int gpub_gather_rows_f64(gpub_ctx_t ctx, int sidx, const double *src, size_t ld_src, size_t row_from,
                         size_t nrows_out, size_t ncols, double *dst);
int gpub_gather_rows_f32(gpub_ctx_t ctx, int sidx, const float *src, size_t ld_src, size_t row_from,
                         size_t nrows_out, size_t ncols, float *dst);
/* ref: tensor.cuh:1167-1197 (tr: k geam launches) -> one batched launch      */
int gpub_transpose_batched_f64(gpub_ctx_t ctx, int sidx, size_t m, size_t n, const double *A, size_t strideA,
                               double *At, size_t strideAt, size_t batch);
int gpub_transpose_batched_f32(gpub_ctx_t ctx, int sidx, size_t m, size_t n, const float *A, size_t strideA,
                               float *At, size_t strideAt, size_t batch);

/* ---- batched GEMM ----------------------------------------------------------
 * C_i <- beta*C_i + alpha*A_i*B_i (NN), i < batch.  C may alias B when
 * n == k (Nullspace::project, tensor.cuh:2084).
 * ref: tensor.cuh:1294, 1321 (gemmBatched), 1303, 1330 (gemm, batch == 1)    */
int gpub_gemm_batched_f64(gpub_ctx_t ctx, int sidx, size_t m, size_t n, size_t k, double alpha,
                          const double *A, size_t lda, size_t strideA,
                          const double *B, size_t ldb, size_t strideB, double beta,
                          double *C, size_t ldc, size_t strideC, size_t batch);
int gpub_gemm_batched_f32(gpub_ctx_t ctx, int sidx, size_t m, size_t n, size_t k, float alpha,
                          const float *A, size_t lda, size_t strideA,
                          const float *B, size_t ldb, size_t strideB, float beta,
                          float *C, size_t ldc, size_t strideC, size_t batch);

/* ---- batched Cholesky -------------------------------------------------------
 * potrf: A_i = L_i L_i^T, lower triangle overwritten by L_i, strict upper
 * triangle left untouched; info[i] = 0 or the 1-based index of the first
 * non-positive pivot.
 * ref: tensor.cuh:2138, 2151 (potrfBatched), 1745, 1756 (potrf, batch == 1)  */
int gpub_potrf_batched_f64(gpub_ctx_t ctx, int sidx, size_t n, double *A, size_t lda, size_t strideA,
                           int *info, size_t batch);
int gpub_potrf_batched_f32(gpub_ctx_t ctx, int sidx, size_t n, float *A, size_t lda, size_t strideA,
                           int *info, size_t batch);
/* potrs: solves L_i L_i^T x = b_i in place, one right-hand side.
 * ref: tensor.cuh:2168, 2187 (potrsBatched), 1766, 1777 (potrs)              */
int gpub_potrs_batched_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *L, size_t ldl, size_t strideL,
                           double *b, size_t strideB, size_t batch);
int gpub_potrs_batched_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *L, size_t ldl, size_t strideL,
                           float *b, size_t strideB, size_t batch);

/* Fused solve + all-gather over NVLink peer memory (additive; the gather of SURVEY 8(e) folded into the solve): potrs_batched on
 * this device's shard of `batch` systems, and every solution x_i is ALSO stored, from inside the kernel, at
 * peer_x[p] + (shard_offset + i) * stride_x for p < n_peers (<= 8) -- straight into the gathered (n, 1, k_total) tensor of every
 * device, which must be mapped in this device's address space (gpub_multi_enable_peer_access). No second pass, no second launch.
 * n <= 32 runs the fused kernel; larger n solve first and push the shard with one peer copy per destination on the same stream. */
int gpub_potrs_allgather_batched_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *L, size_t ldl, size_t strideL,
                                     double *b, size_t strideB, size_t batch, double *const *peer_x, int n_peers,
                                     size_t shard_offset, size_t stride_x);
int gpub_potrs_allgather_batched_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *L, size_t ldl, size_t strideL,
                                     float *b, size_t strideB, size_t batch, float *const *peer_x, int n_peers,
                                     size_t shard_offset, size_t stride_x);
/* Host pipeline (additive; the reference has whole-tensor upload -> factorise -> solve -> download, tensor.cuh:1128-1154,
 * 2135-2197): the dense batch A_host (n*n per matrix) / b_host (n per matrix) is cut into `chunks` pieces (0 = default 16);
 * each piece is uploaded into A_dev / b_dev, factorised and solved as soon as it has landed, and x / info are downloaded behind
 * the kernels, on three streams (upload, stream `sidx`, download), so the whole job runs at the host link's speed. On return
 * (blocking) A_dev holds the factors, b_dev and x_host the solutions, info_dev / info_host the status codes. b_* / x_host /
 * info_* may be NULL (factorise only / no download). Host buffers may be pinned (DMA'd directly) or pageable (staged).
 * `chunks` may be OR-ed with GPUB_LOWER_ONLY: the factorisation reads the lower triangle only, so for pinned A_host and rows of the
 * half-height strip of at least 128 bytes (fp64: even n >= 32) the n/2 x n/2 block above the diagonal is not transferred (75 % of
 * the bytes, strided DMA); that block of A_dev then keeps whatever it held. Without the flag every byte of A_host is copied. */
#define GPUB_LOWER_ONLY ((size_t) 1 << 32)
#define GPUB_CHUNKS_MASK (((size_t) 1 << 32) - 1)
int gpub_chol_solve_from_host_f64(gpub_ctx_t ctx, int sidx, size_t n, double *A_dev, double *b_dev, int *info_dev,
                                  const double *A_host, const double *b_host, double *x_host, int *info_host,
                                  size_t batch, size_t chunks);
int gpub_chol_solve_from_host_f32(gpub_ctx_t ctx, int sidx, size_t n, float *A_dev, float *b_dev, int *info_dev,
                                  const float *A_host, const float *b_host, float *x_host, int *info_host,
                                  size_t batch, size_t chunks);

/* ---- batched Householder QR / least squares --------------------------------
 * geqrf: LAPACK storage (R in the upper triangle, reflectors below, tau[n]).
 * ref: tensor.cuh:1870, 1883 (geqrf; single matrix in the reference)         */
int gpub_geqrf_batched_f64(gpub_ctx_t ctx, int sidx, size_t m, size_t n, double *A, size_t lda, size_t strideA,
                           double *tau, size_t strideTau, size_t batch);
int gpub_geqrf_batched_f32(gpub_ctx_t ctx, int sidx, size_t m, size_t n, float *A, size_t lda, size_t strideA,
                           float *tau, size_t strideTau, size_t batch);
/* ormqr, side = left: C_i <- Q_i^T C_i (trans != 0) or Q_i C_i (trans == 0),
 * Q_i = H_1 ... H_k from geqrf; C_i is m x ncols.
 * ref: tensor.cuh:1896, 1915 (Q^T b), 1946, 1980 (Q * I)                      */
int gpub_ormqr_batched_f64(gpub_ctx_t ctx, int sidx, int trans, size_t m, size_t ncols, size_t k,
                           const double *A, size_t lda, size_t strideA, const double *tau, size_t strideTau,
                           double *C, size_t ldc, size_t strideC, size_t batch);
int gpub_ormqr_batched_f32(gpub_ctx_t ctx, int sidx, int trans, size_t m, size_t ncols, size_t k,
                           const float *A, size_t lda, size_t strideA, const float *tau, size_t strideTau,
                           float *C, size_t ldc, size_t strideC, size_t batch);
/* trsv, upper, non-unit, no transpose: solves R_i x = b_i in place (n x n, one rhs).
 * ref: tensor.cuh:1903, 1922 (trsm LEFT UPPER N NONUNIT, nrhs = 1)            */
int gpub_trsv_upper_batched_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *R, size_t ldr, size_t strideR,
                                double *b, size_t strideB, size_t batch);
int gpub_trsv_upper_batched_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *R, size_t ldr, size_t strideR,
                                float *b, size_t strideB, size_t batch);
/* gels (m >= n, one rhs): A_i <- QR factors, b_i[0:n] <- argmin ||A_i x - b_i||,
 * b_i[n:m] <- tail of Q_i^T b_i; info[i] = 0 (or j>0 if R(j,j) == 0).
 * `info` may be NULL.
 * ref: tensor.cuh:1354, 1382 (gelsBatched)                                    */
int gpub_gels_batched_f64(gpub_ctx_t ctx, int sidx, size_t m, size_t n, double *A, size_t lda, size_t strideA,
                          double *b, size_t strideB, int *info, size_t batch);
int gpub_gels_batched_f32(gpub_ctx_t ctx, int sidx, size_t m, size_t n, float *A, size_t lda, size_t strideA,
                          float *b, size_t strideB, int *info, size_t batch);

/* ---- batched SVD -------------------------------------------------------------
 * A_i (m x n, m >= n) = U_i diag(S_i) Vt_i.  S descending, Vt is n x n, U is
 * the FULL m x m factor when jobu == 'A' and not referenced when jobu == 'N'.
 * A is destroyed.  info[i] = 0, or >0 if the iteration did not converge.
 * A matrix whose largest entry is outside LAPACK's gesvd range [sqrt(safmin)/eps, eps/sqrt(safmin)] is scaled by a
 * power of two first and its singular values scaled back (dlascl).  A batch of more matrices than the GPU has SMs
 * runs as four sub-batches on streams of the library's own, forked from stream `sidx` and joined to it again before
 * the call returns: work queued on stream `sidx` behind the call sees every result.
 * Shapes: n <= 32 any m; 32 < n <= 256 while the n x n factor fits shared memory (fp64: n <= 167), else GPUB_ENOTSUP
 * (worksize returns 0).
 * ref: tensor.cuh:1637, 1664 (gesvd, called numMats times in a host loop)     */
size_t gpub_gesvd_batched_worksize_f64(size_t m, size_t n, int jobu, size_t batch);
size_t gpub_gesvd_batched_worksize_f32(size_t m, size_t n, int jobu, size_t batch);
int gpub_gesvd_batched_f64(gpub_ctx_t ctx, int sidx, int jobu, size_t m, size_t n,
                           double *A, size_t lda, size_t strideA, double *S, size_t strideS,
                           double *U, size_t ldu, size_t strideU, double *Vt, size_t ldvt, size_t strideVt,
                           void *work, size_t work_bytes, int *info, size_t batch);
int gpub_gesvd_batched_f32(gpub_ctx_t ctx, int sidx, int jobu, size_t m, size_t n,
                           float *A, size_t lda, size_t strideA, float *S, size_t strideS,
                           float *U, size_t ldu, size_t strideU, float *Vt, size_t ldvt, size_t strideVt,
                           void *work, size_t work_bytes, int *info, size_t batch);
/* rank: count[i] += #{ j < len : S_i[j] > eps }  (accumulates, like the reference)
 * ref: tensor.cuh:1486-1491, 1600-1609 (one launch per matrix)                */
int gpub_count_gt_batched_f64(gpub_ctx_t ctx, int sidx, const double *S, size_t len, size_t strideS,
                              double eps, unsigned int *count, size_t batch);
int gpub_count_gt_batched_f32(gpub_ctx_t ctx, int sidx, const float *S, size_t len, size_t strideS,
                              float eps, unsigned int *count, size_t batch);
/* Nullspace assembly: N_i <- [ U_i(:, rank_i : n-1) | 0 ] (n x n, left-packed,
 * zero padded); rank read from the device.
 * ref: tensor.cuh:2062-2078 (host loop of slices, copies, tr and addAB)       */
int gpub_nullspace_pack_batched_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *U, size_t strideU,
                                    const unsigned int *rank, double *N, size_t strideN, size_t batch);
int gpub_nullspace_pack_batched_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *U, size_t strideU,
                                    const unsigned int *rank, float *N, size_t strideN, size_t batch);
/* P_i <- N_i N_i^T (n x n), the projector of Nullspace (tensor.cuh:2076-2077) */
int gpub_aat_batched_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *N, size_t strideN,
                         double *P, size_t strideP, size_t batch);
int gpub_aat_batched_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *N, size_t strideN,
                         float *P, size_t strideP, size_t batch);

/* The same projector from the orthogonal factor N_i was cut out of: N_i N_i^T = U2 U2^T = I - U1 U1^T (U1 = the first rank_i columns
 * of U_i), so per matrix the side with FEWER columns is multiplied out (rank read from the device). Same result as
 * gpub_aat_batched up to rounding; for a fat 128 x 1024 matrix 7 x fewer flops. */
int gpub_nullspace_projector_batched_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *U, size_t strideU,
                                         const unsigned int *rank, const double *N, size_t strideN,
                                         double *P, size_t strideP, size_t batch);
int gpub_nullspace_projector_batched_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *U, size_t strideU,
                                         const unsigned int *rank, const float *N, size_t strideN,
                                         float *P, size_t strideP, size_t batch);

/* Both of the above in one call (Nullspace's constructor, tensor.cuh:2058-2079): N_i = [ U_i(:, rank_i:n) | 0 ] and P_i = N_i N_i^T.
 * The packing (a copy) runs on a private stream beside the projector (a contraction that reads U, not N); the call's stream waits
 * for both. */
int gpub_nullspace_build_batched_f64(gpub_ctx_t ctx, int sidx, size_t n, const double *U, size_t strideU, const unsigned int *rank,
                                     double *N, size_t strideN, double *P, size_t strideP, size_t batch);
int gpub_nullspace_build_batched_f32(gpub_ctx_t ctx, int sidx, size_t n, const float *U, size_t strideU, const unsigned int *rank,
                                     float *N, size_t strideN, float *P, size_t strideP, size_t batch);

/* ---- synthetic data (bench / tests): counter-based generator, SURVEY 8(d) ---
 * x[i] = lo + (hi - lo) * u(seed, i),  u in [0,1) from a 64-bit hash of i     */
int gpub_fill_uniform_f64(gpub_ctx_t ctx, int sidx, size_t n, double *x, double lo, double hi, uint64_t seed);
int gpub_fill_uniform_f32(gpub_ctx_t ctx, int sidx, size_t n, float *x, float lo, float hi, uint64_t seed);
/* A_i <- G_i G_i^T + shift * I with G_i ~ U[-1,1] (n x n): SPD batches */
int gpub_fill_spd_batched_f64(gpub_ctx_t ctx, int sidx, size_t n, double *A, size_t strideA, double shift,
                              uint64_t seed, size_t batch);
int gpub_fill_spd_batched_f32(gpub_ctx_t ctx, int sidx, size_t n, float *A, size_t strideA, float shift,
                              uint64_t seed, size_t batch);

/* ---- multi-GPU: the mats axis sharded over the GPUs of one box ---
 * The reference binds one device (Session, tensor.cuh:133-247) and has no multi-GPU path; these are additive.
 * Batched ops shard embarrassingly: each device runs the launchers above on its own block of matrices with its own
 * context, so there is NO collective on the data path. The entry points below only make the devices NVLink peers and
 * all-gather result shards (device g contributes bytes[g] bytes from send[g]; every recv[g] receives the concatenation
 * in device order). Asynchronous on stream `sidx` of each context.
 *   transport: GPUB_GATHER_AUTO -> NCCL when it can be loaded and the devices are distinct, else peer copies;
 *              GPUB_GATHER_NCCL -> ncclAllGather (equal shards) / grouped ncclBroadcast (ragged), GPUB_ENOTSUP if unavailable;
 *              GPUB_GATHER_P2P  -> cudaMemcpyPeerAsync pulls over NVLink (also serves a device listed twice). */
#define GPUB_GATHER_AUTO 0
#define GPUB_GATHER_NCCL 1
#define GPUB_GATHER_P2P 2
int gpub_multi_device_count(int *count);
int gpub_multi_enable_peer_access(const int *devices, int n, int *n_pairs_enabled);
/* NCCL version code (e.g. 22809) of the library found with dlopen, 0 if none */
int gpub_multi_nccl_version(int *version);
int gpub_multi_allgather(const gpub_ctx_t *ctxs, int n, int sidx, const void *const *send, const size_t *bytes,
                         void *const *recv, int transport, int *transport_used);
/* destroys the cached NCCL communicators */
int gpub_multi_release(void);

#ifdef __cplusplus
}
#endif
#endif /* GPUTILS_B200_H */
