// Register-resident microbenchmarks for the roofline denominators MEASURED_PEAKS.json does not have
// (SURVEY.md 8d): FP64 DFMA, FP64 DMMA (mma.sync m8n8k4 -> DMMA.8x8x4), FP32 FFMA, and DFMA+DMMA issued together.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_dfma(double *out, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(float *out, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double *out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mixed(double *out, double a, double b) {
    double c[4][2], x[8];
#pragma unroll
    for (int i = 0; i < 4; i++) c[i][0] = c[i][1] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x - i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent-chain latencies (one warp): cycles per op
__global__ void k_lat(double *out, long long *cyc, double a, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < 1024; it++) x = fma(x, a, b);
    long long t1 = clock64();
    double y = x;
    for (int it = 0; it < 1024; it++) y = __shfl_sync(0xffffffffu, y, (threadIdx.x + 1) & 31);
    long long t2 = clock64();
    double z = y + 2.0;
    for (int it = 0; it < 1024; it++) z = rsqrt(z) + 1.5;
    long long t3 = clock64();
    double c0 = z, c1 = z;
    for (int it = 0; it < 1024; it++) dmma(c0, c1, a, b);
    long long t4 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
    out[threadIdx.x] = c0 + c1;
}

template<typename F> float run(F launch, int reps) {
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(s); launch(); cudaEventRecord(e); cudaEventSynchronize(e);
        float ms; cudaEventElapsedTime(&ms, s, e); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256;
    double *d; cudaMalloc(&d, sizeof(double) * blocks * threads);
    long long *cyc; cudaMalloc(&cyc, 64);
    double n = (double) blocks * threads;
    float t;
    t = run([&] { k_dfma<<<blocks, threads>>>(d, 1.0000001, 1e-9); }, 5);
    printf("{\"dfma_tflops\": %.2f, ", n * 16 * ITERS * 2 / t / 1e9);
    t = run([&] { k_dmma<<<blocks, threads>>>(d, 1.0000001, 1e-9); }, 5);
    printf("\"dmma_tflops\": %.2f, ", (double) blocks * (threads / 32) * 8.0 * ITERS * 512 / t / 1e9);
    t = run([&] { k_mixed<<<blocks, threads>>>(d, 1.0000001, 1e-9); }, 5);
    printf("\"mixed_dmma_plus_dfma_tflops\": %.2f, ", ((double) blocks * (threads / 32) * 4.0 * ITERS * 512 + n * 8 * ITERS * 2) / t / 1e9);
    t = run([&] { k_ffma<<<blocks, threads>>>((float *) d, 1.0000001f, 1e-9f); }, 5);
    printf("\"ffma_tflops\": %.2f, ", n * 16 * ITERS * 2 / t / 1e9);
    k_lat<<<1, 32>>>(d, cyc, 1.0000001, 1e-9); cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("\"lat_cycles\": {\"dfma\": %.1f, \"shfl64\": %.1f, \"rsqrt_f64_plus_add\": %.1f, \"dmma\": %.1f}, \"sms\": %d}\n",
           h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, sms);
    return 0;
}
