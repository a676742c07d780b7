// Microbenchmark: peak FP64 throughput of DMMA (mma.sync m8n8k4 / m16n8k8 f64) vs DFMA on this GPU.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dmma884(double* out, int iters) {
    double c[16][2];
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1688(double* out, int iters) {
    double c[8][4];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 + threadIdx.x * 1e-4, b1 = b0 + 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(double* out, int iters) {
    double c[32];
    for (int i = 0; i < 32; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 32; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs\n", p.name, sms);
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32;
        int blocks = sms;
        float ms = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
        double flops = 2.0 * 8 * 8 * 4 * 16.0 * iters * warps * blocks;
        printf("DMMA m8n8k4   %2d warps/SM: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
        ms = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, iters); });
        flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * blocks;
        printf("DMMA m16n8k8  %2d warps/SM: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
        ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
        flops = 2.0 * 32 * 32.0 * iters * warps * blocks;
        printf("DFMA          %2d warps/SM: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
