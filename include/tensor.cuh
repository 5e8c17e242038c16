/*
 * tensor.cuh -- drop-in replacement for GPUtils' include/tensor.cuh on B200 (sm_100a).
 *
 * Same public API as the reference header (DTensor, Session, Svd, CholeskyFactoriser, QRFactoriser,
 * Nullspace, CholeskyBatchFactoriser, GivensAnnihilator, the helper macros and free functions), so
 * the reference's test/testTensor.cu, main.cu and example/main.cu compile against it unchanged.
 * Underneath there is no cuBLAS / cuSOLVER: every numerical method forwards to the extern "C"
 * launchers of libgputils_b200 (include/gputils_b200.h), hand-written sm_100a kernels.
 *
 * Link with -lgputils_b200. The parts live in include/gpub200/.
 */
#ifndef TENSOR_CUH
#define TENSOR_CUH

#include "gpub200/core.cuh"
#include "gpub200/dtensor.cuh"
#include "gpub200/factorisers.cuh"
#include "gpub200/sharded.cuh" /* additive: mats axis sharded over the GPUs of one box */
#include "gpub200/pitched.cuh" /* additive: per-matrix 128-byte-aligned (padded) storage for shapes that are not multiples of 128 B */

#endif /* TENSOR_CUH */
