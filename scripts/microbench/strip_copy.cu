// Microbenchmark: host -> device transfer of the LOWER TRIANGLES of a batch of 32 x 32 fp64 matrices as strided 3-D copies
// (column strips of W columns, rows from the strip's first column down) against the dense copy.  nvcc -O2 strip_copy.cu -o strip_copy
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
int main() {
    const size_t n = 32, k = 250000, bytes = n * n * k * sizeof(double);
    double *h, *d;
    CK(cudaHostAlloc((void **) &h, bytes, cudaHostAllocDefault));
    CK(cudaMalloc((void **) &d, bytes));
    for (size_t i = 0; i < n * n * k; i += 4096) h[i] = 1.0;
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(a, s));
        CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(b, s)); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b);
    }
    printf("dense        : %.3f ms  %.1f GB/s\n", ms, bytes / ms / 1e6);
    for (int W: {16, 8, 4, 2}) {
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(a, s));
            size_t moved = 0;
            for (size_t c0 = 0; c0 < n; c0 += W) {
                cudaMemcpy3DParms p = {};
                const size_t rows = n - c0;                      // rows c0 .. n-1 of columns c0 .. c0+W-1
                p.srcPtr = make_cudaPitchedPtr((char *) h + (c0 + c0 * n) * sizeof(double), n * sizeof(double), n, W);
                p.dstPtr = make_cudaPitchedPtr((char *) d + (c0 + c0 * n) * sizeof(double), n * sizeof(double), n, W);
                p.extent = make_cudaExtent(rows * sizeof(double), W, k);
                // slice pitch = pitch * height must equal one matrix: height = n columns, so describe the matrix as n rows of the 3-D volume
                p.srcPtr.ysize = n; p.dstPtr.ysize = n;
                p.kind = cudaMemcpyHostToDevice;
                CK(cudaMemcpy3DAsync(&p, s));
                moved += rows * W * k * sizeof(double);
            }
            CK(cudaEventRecord(b, s)); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b);
            if (rep) printf("strips of %2d : %.3f ms  %.1f GB/s on the bytes moved (%.1f %% of dense), %.2fx the dense copy's time\n", W, ms,
                            moved / ms / 1e6, 100.0 * moved / bytes, ms / (bytes / 56.0e6));
        }
    }
    return 0;
}
