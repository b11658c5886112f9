// Standalone probe: which 4-D tensor maps does the TMA unit accept, and what smem layout results?
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../maua_b200/csrc/common.cuh"
namespace mb { void set_error(const char*, ...) {} }
using namespace mb;

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, __half* out, int nbytes, int c0, int c1, int c2, int c3) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, nbytes);
        tma_load_4d(smem, &tm, &bar, c0, c1, c2, c3);
    }
    mbar_wait(&bar, 0);
    __syncthreads();
    for (int i = threadIdx.x; i < nbytes / 2; i += blockDim.x) out[i] = reinterpret_cast<__half*>(smem)[i];
}

int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int W = 20, C = 64, H = 20, B = 1, Wp = 24;
    int bw = 64, bc = 64, bh = 6;
    int c0 = -2, c1 = 0, c2 = -2, c3 = 0;
    bool monotonic = false;
    CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B;
    if (variant == 1) { c0 = 0; c2 = 0; }
    if (variant == 2) { c0 = -2; c2 = 0; }
    if (variant == 3) { c0 = 0; c2 = -2; }
    if (variant == 4) { c0 = -8; c2 = 0; }
    if (variant == 5) { c0 = 3; c2 = 0; }
    if (variant == 6) { c0 = 8; c2 = 1; }
    if (variant == 7) { bw = 32; swz = CU_TENSOR_MAP_SWIZZLE_64B; bh = 10; c0 = 0; c2 = -2; }
    if (variant == 8) { c0 = 0; c2 = -2; monotonic = true; }
    if (variant == 9) { c0 = -8; c2 = -2; }
    size_t plane = (size_t)H * Wp;
    std::vector<__half> hx(plane * C * B);
    for (int c = 0; c < C; ++c) for (int h = 0; h < H; ++h) for (int w = 0; w < Wp; ++w)
        { unsigned short v = (unsigned short)((c << 10) | (h << 5) | w); hx[(size_t)c * plane + h * Wp + w] = *reinterpret_cast<__half*>(&v); }
    __half* dx; cudaMalloc(&dx, hx.size() * 2); cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    PFN_encodeTiled enc = nullptr; { void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); enc = (PFN_encodeTiled)p; }
    CUtensorMap tm;
    CUresult r;
    int box_elems;
    if (!monotonic) {
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {plane * 2, (cuuint64_t)Wp * 2, plane * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bc, (cuuint32_t)bh, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        box_elems = bw * bc * bh;
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Wp * 2, plane * 2, plane * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        box_elems = bw * bc * bh;
        int t = c1; c1 = c2; c2 = t;
    }
    printf("variant %d encode result %d box bytes %d\n", variant, (int)r, box_elems * 2);
    if (r != CUDA_SUCCESS) return 1;
    __half* dout; cudaMalloc(&dout, box_elems * 2);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, box_elems * 2 + 2048);
    probe_kernel<<<1, 128, box_elems * 2 + 2048>>>(tm, dout, box_elems * 2, c0, c1, c2, c3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d kernel: %s\n", variant, cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<__half> ho(box_elems); cudaMemcpy(ho.data(), dout, box_elems * 2, cudaMemcpyDeviceToHost);
    if (!monotonic) {
        // check assumed layout [h][c][w] with 16B-chunk XOR swizzle
        int rowb = bw * 2; int chunks = rowb / 16; int bad = 0, checked = 0;
        for (int h = 0; h < bh; ++h) for (int c = 0; c < bc; ++c) for (int w = 0; w < bw; ++w) {
            int gh = c2 + h, gw = c0 + w, gc = c1 + c;
            unsigned short want = (gh >= 0 && gh < H && gw >= 0 && gw < W && gc < C) ? (unsigned short)((gc << 10) | (gh << 5) | gw) : 0;
            size_t row = (size_t)h * bc + c;  // 'rowb'-byte rows
            size_t byte = row * rowb;
            int chunk = w / 8;
            int sw = (swz == CU_TENSOR_MAP_SWIZZLE_128B) ? (chunk ^ (int)((byte / 128) % 8)) : (chunk ^ (int)((byte / 128) % 4));
            size_t addr = byte + sw * 16 + (w % 8) * 2;
            unsigned short got = *reinterpret_cast<unsigned short*>(&ho[addr / 2]);
            ++checked; if (got != want) { if (bad < 5) printf("  mismatch h%d c%d w%d got %04x want %04x\n", h, c, w, got, want); ++bad; }
        }
        printf("variant %d layout check: %d / %d mismatches\n", variant, bad, checked);
    }
    return 0;
}
