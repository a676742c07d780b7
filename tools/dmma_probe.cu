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

// 4 x 4 outer-product operand pattern of a GEMM inner loop (distinct A and B registers), optionally
// with the 6 DADDs per 24 DMMAs of the 3M kernel
template <int WITH_DADD>
__global__ void k_dmma_ab(double* out, int iters) {
    double c[24][2];
    for (int i = 0; i < 24; ++i) c[i][0] = c[i][1] = 0.0;
    double a[2][2], b[4][2];
    for (int i = 0; i < 2; ++i) { a[i][0] = threadIdx.x * 1e-3 + i; a[i][1] = a[i][0] * 0.5; }
    for (int j = 0; j < 4; ++j) { b[j][0] = 1.0 + threadIdx.x * 1e-4 + j; b[j][1] = b[j][0] * 0.25; }
    for (int it = 0; it < iters; ++it) {
        double a2[2], b2[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) a2[i] = WITH_DADD ? a[i][0] + a[i][1] : a[i][0];
#pragma unroll
        for (int j = 0; j < 4; ++j) b2[j] = WITH_DADD ? b[j][1] - b[j][0] : b[j][1];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i * 4 + j][0]), "+d"(c[i * 4 + j][1]) : "d"(a[i][0]), "d"(b[j][0]));
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[8 + i * 4 + j][0]), "+d"(c[8 + i * 4 + j][1]) : "d"(a[i][1]), "d"(b[j][1]));
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[16 + i * 4 + j][0]), "+d"(c[16 + i * 4 + j][1]) : "d"(a2[i]), "d"(b2[j]));
        // make the operands loop-carried so that the DADDs are not hoisted
        if (WITH_DADD) {
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i][0] = a2[i] * 0.999;
        }
    }
    double s = 0;
    for (int i = 0; i < 24; ++i) s += c[i][0] + c[i][1];
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
    for (int warps : {4, 8, 16}) {
        int threads = warps * 32;
        int blocks = sms;
        float ms = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
        double flops = 2.0 * 8 * 8 * 4 * 16.0 * iters * warps * blocks;
        printf("DMMA m8n8k4   %2d warps/SM: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
        ms = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, iters); });
        flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * blocks;
        printf("DMMA m16n8k8  %2d warps/SM: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
        ms = timeit([&] { k_dmma_ab<0><<<blocks, threads>>>(out, iters); });
        flops = 2.0 * 8 * 8 * 4 * 24.0 * iters * warps * blocks;
        printf("DMMA 2x4 tile %2d warps/SM: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
        ms = timeit([&] { k_dmma_ab<1><<<blocks, threads>>>(out, iters); });
        printf("DMMA 2x4+DADD %2d warps/SM: %8.3f ms  %7.2f TFLOP/s (DMMA flops only)\n", warps, ms, flops / ms / 1e9);
        ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
        flops = 2.0 * 32 * 32.0 * iters * warps * blocks;
        printf("DFMA          %2d warps/SM: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
