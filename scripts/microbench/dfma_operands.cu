// FP64 issue rate against the number of distinct register operands (the roofline denominator of the rotation kernels):
// peaks.cu measures fma(x, a, b) with a, b shared by every instruction; a plane rotation reads three different register
// pairs per DFMA. build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands dfma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;

__global__ void k_one(double *out, double a, double b) {            // one varying operand
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

__global__ void k_two(double *out, double a, double b) {            // two varying operands
    double x[16], y[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fma(x[i], y[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_three(double *out, double a, double b) {          // three varying operands
    double x[12], y[12], z[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-3 * i + a; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 12; i++) x[i] = fma(x[i], y[i], z[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the rotation as the Jacobi kernels apply it: 2 DMUL + 2 DFMA per element of a column pair, (cs, sn) per pair
__global__ void k_rot(double *out, double cs, double sn) {
    double x[16], y[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const double a = x[i], b = y[i];
            y[i] = fma(sn, a, cs * b);
            x[i] = fma(cs, a, -(sn * b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same rotation as tangent form: x' = x - t y, y' = y + t x (two DFMA, the scaling by cs deferred)
__global__ void k_rot_fast(double *out, double t) {
    double x[16], y[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const double a = x[i], b = y[i];
            x[i] = fma(-t, b, a);
            y[i] = fma(t, a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<typename K, typename... A>
static float run(K k, int threads, int ctas_per_sm, int sms, double *out, A... a) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<sms * ctas_per_sm, threads>>>(out, a...);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<sms * ctas_per_sm, threads>>>(out, a...);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    for (int warps : {4, 8, 16, 32}) {
        const int threads = 32 * (warps < 32 ? warps : 32), ctas = 1;
        const double inst1 = (double) sms * ctas * threads * ITERS;
        float m1 = run(k_one, threads, ctas, sms, out, 1.0000001, 1e-9);
        float m2 = run(k_two, threads, ctas, sms, out, 1.0000001, 1e-9);
        float m3 = run(k_three, threads, ctas, sms, out, 1.0000001, 1e-9);
        float mr = run(k_rot, threads, ctas, sms, out, 0.8, 0.6);
        float mf = run(k_rot_fast, threads, ctas, sms, out, 1e-9);
        printf("warps/SM %2d: DP instructions per clock per SM (lanes): fma(x,a,b) %.1f  fma(x,y,b) %.1f  fma(x,y,z) %.1f  rotation (2 DMUL + 2 DFMA) %.1f  tangent rotation (2 DFMA) %.1f\n",
               warps, inst1 * 16 / (m1 * 1e-3) / 1.965e9 / sms, inst1 * 16 / (m2 * 1e-3) / 1.965e9 / sms, inst1 * 12 / (m3 * 1e-3) / 1.965e9 / sms,
               inst1 * 64 / (mr * 1e-3) / 1.965e9 / sms, inst1 * 32 / (mf * 1e-3) / 1.965e9 / sms);
    }
    return 0;
}
