// Development/test harness only: compiles the __host__ __device__ small-SVD core with the host compiler so
// that tests/test_svd_core_host.py can compare it with LAPACK (scipy) on the CPU box. Not linked into the
// shipped library and never used by the product path (which runs the same code inside CUDA kernels).
#include "../../gputils_b200/csrc/svd_small.cuh"

extern "C" {
int harness_gesvd_small_f64(int m, int n, double *A, double *S, double *U, double *Vt, int want_u, double *scratch) {
    return gpub_svd::gesvd_small<double>(m, n, A, m, S, U, m, Vt, n, want_u != 0, scratch);
}
int harness_gesvd_small_f32(int m, int n, float *A, float *S, float *U, float *Vt, int want_u, float *scratch) {
    return gpub_svd::gesvd_small<float>(m, n, A, m, S, U, m, Vt, n, want_u != 0, scratch);
}
}
