// Microbenchmark: legacy mma.sync.m16n8k16 (f16 in, f32 acc) throughput on sm_100a.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>
__global__ void __launch_bounds__(256) mma_kernel(float* out, int iters) {
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) ffma_kernel(float* out, int iters) {
    float c[16]; for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 0.001f + i;
    float a = 1.0001f, b = 0.0001f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bps = 1; bps <= 4; bps *= 2) {
        int iters = 20000;
        mma_kernel<<<148 * bps, 256>>>(d, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); mma_kernel<<<148 * bps, 256>>>(d, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double macs = (double)148 * bps * 8 /*warps*/ * iters * 8 * 2048.0;
        printf("mma.sync m16n8k16: %d CTA/SM x 8 warps: %.3f ms  %.1f TFLOP/s  (%.0f MAC/clk/SM at 1.9GHz)\n", bps, ms, 2 * macs / ms / 1e9, macs / (ms * 1e-3) / 148 / 1.9e9);
    }
    {
        int iters = 20000;
        ffma_kernel<<<148 * 4, 256>>>(d, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); ffma_kernel<<<148 * 4, 256>>>(d, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fmas = (double)148 * 4 * 256 * iters * 16.0;
        printf("ffma: %.3f ms %.1f TFLOP/s (%.0f FMA/clk/SM at 1.9GHz)\n", ms, 2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.9e9);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
