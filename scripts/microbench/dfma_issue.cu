// Does an FP64 instruction cost the warp scheduler one issue slot or two? 16 warps per SM run DFMA chains with K independent
// FP32 / integer instructions interleaved per DFMA; if FP64 held the issue port for both of its pipe cycles, the DFMA rate would drop
// as soon as K > 0. build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_issue dfma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;

template<int K>
__global__ void k_mix(double *out, double a, double b, float fa, float fb) {
    double x[8];
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 0.5f + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x[i] = fma(x[i], a, b);
#pragma unroll
            for (int k = 0; k < K; k++) y[(i + k) & 7] = fmaf(y[(i + k) & 7], fa, fb);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int K>
static void run(int sms, double *out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<K><<<sms, 512>>>(out, 1.0000001, 1e-9, 1.0000001f, 1e-9f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_mix<K><<<sms, 512>>>(out, 1.0000001, 1e-9, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double) sms * 512 * ITERS * 8;
    printf("FFMA per DFMA %d: %.1f DFMA lanes per clock per SM, %.1f FFMA lanes per clock per SM, %.2f cycles per DFMA warp instruction per scheduler\n", K,
           dfma / (ms * 1e-3) / 1.965e9 / sms, dfma * K / (ms * 1e-3) / 1.965e9 / sms, (ms * 1e-3) * 1.965e9 / (4.0 * ITERS * 8));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 512);
    run<0>(p.multiProcessorCount, out);
    run<1>(p.multiProcessorCount, out);
    run<2>(p.multiProcessorCount, out);
    run<3>(p.multiProcessorCount, out);
    run<4>(p.multiProcessorCount, out);
    return 0;
}
