// Microbenchmark: tcgen05.mma kind::f16 M=128, K=16 issue/throughput vs N, A from shared memory (SS) or TMEM (TS).
// Answers: (1) is a pixel-major (M = pixels, N = couts <= 128) conv tile faster than the M-padded one, (2) could a
// banded-Toeplitz FIR run on tcgen05 with small N.
#include "../maua_b200/csrc/common.cuh"
using namespace mb;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int R>
__global__ void __launch_bounds__(128, 1) bench(int N, int ts, int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_f16(128, N, 0, 0);
        const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
        uint64_t da[4], db[4];
        for (int j = 0; j < 4; ++j) {
            da[j] = make_smem_desc(sa + j * 32, 16, 1024, 2);
            db[j] = make_smem_desc(sb + j * 32, 16, 1024, 2);
        }
        long long t0 = clock64();
        for (int it = 0; it < iters; it += 4 * R) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    // accumulator r at columns [r*N, (r+1)*N); its A rows 4 KB further on (a different image row block)
                    if (leader) {
                        if (ts) umma_f16_ts(tm + r * N, tm + 480 + j * 8, db[j], idesc, 1);
                        else umma_f16(tm + r * N, da[j] + static_cast<uint64_t>((r & 7) * (4096 >> 4)), db[j], idesc, 1);
                    }
                }
            }
        }
        if (leader) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (cycles && threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int R>
void run(int N, int ts, long long* d) {
    if (N * R > (ts ? 448 : 512)) return;
    const int iters = 4096;
    cudaFuncSetAttribute(bench<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    bench<R><<<148, 128, 100 * 1024>>>(N, ts, 64, nullptr);
    bench<R><<<148, 128, 100 * 1024>>>(N, ts, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d ts=%d: %s\n", N, ts, cudaGetErrorString(e)); exit(1); }
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%s M=128 N=%3d R=%2d K=16: %6.1f cycles/MMA (floor N/2 = %3d), %5.0f MAC/clk/SM\n", ts ? "TS(A in TMEM)" : "SS(A in smem)", N, R,
           avg / iters, N / 2, 128.0 * N * 16 / (avg / iters));
}

int main() {
    long long* d; cudaMalloc(&d, 148 * 8);
    const int Ns[] = {8, 16, 32, 48, 64, 96, 128, 256};
    for (int ts = 0; ts < 2; ++ts)
        for (int N : Ns) {
            run<1>(N, ts, d); run<2>(N, ts, d); run<4>(N, ts, d); run<8>(N, ts, d); run<16>(N, ts, d);
        }
    return 0;
}
