// filtered_lrelu for sm_100a: bias -> zero-insert upsample -> FIR -> leaky-ReLU*gain -> clamp ->
// FIR -> decimate, one shared-memory resident tile per CTA; the (up*in)^2 intermediate never
// touches HBM.  Replaces upstream torch_utils/ops/filtered_lrelu.{py,cpp,cu} (un-vendored
// submodule maua/GAN/nv; call site upstream networks_stylegan3.py SynthesisLayer.forward).
//
// Index convention (upstream _filtered_lrelu_ref / _upfirdn2d_ref), per axis:
//   xu[m]  = x[m/up] if m % up == 0 else 0                       (zero insertion)
//   t[J]   = up * sum_k fu[UT-1-k] * xu[J + k - p0]              (pad p0 may be negative = crop)
//   a[J]   = clamp(lrelu(t[J]) * gain)
//   out[o] = sum_k fd[DT-1-k] * a[o*down + k]
// Two kernels:
//   flrelu_generic_kernel  any up/down/taps, separable or 2-D (radial) down filter; simple loops.
//   flrelu_sep_kernel<UP>  the two shapes StyleGAN3 layers use at down=2 / 12 taps
//                          (up=2/12 taps, up=4/24 taps): polyphase, register-blocked.
#include "common.cuh"
#include "kernels.h"

namespace mb {
namespace {

__device__ __forceinline__ int floor_div(int a, int b) {
    int q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
}
__device__ __forceinline__ int ceil_div_s(int a, int b) { return -floor_div(-a, b); }

struct TileGeo {
    int o0, J0, TT, i_lo, IN_T;
};
__device__ __forceinline__ TileGeo tile_geo(int o0, int OT, int up, int down, int UT, int DT, int p0) {
    TileGeo g;
    g.o0 = o0;
    g.J0 = o0 * down;
    g.TT = (OT - 1) * down + DT;
    g.i_lo = ceil_div_s(g.J0 - p0, up);
    const int i_hi = floor_div(g.J0 + g.TT + UT - 2 - p0, up);
    g.IN_T = i_hi - g.i_lo + 1;
    return g;
}

constexpr int kGenOT = 32;

// ---------------------------------------------------------------------------------------
// generic kernel
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flrelu_generic_kernel(FlreluArgs a, int tiles_x, int max_in, int max_tt) {
    extern __shared__ float sm[];
    const int UT = a.up_taps, DT = a.down_taps, up = a.up, down = a.down;
    float* s_fu = sm;                         // [UT]
    float* s_fd = s_fu + 32;                  // [DT*DT] or [DT]
    float* s_in = s_fd + 160;                 // [max_in][max_in]
    float* s_uh = s_in + max_in * max_in;     // [max_in][max_tt]
    float* s_t = s_uh + max_in * max_tt;      // [max_tt][max_tt]
    float* s_dh = s_t + max_tt * max_tt;      // [max_tt][kGenOT]

    const int tid = threadIdx.x;
    const int c = blockIdx.y, b = blockIdx.z;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;

    for (int i = tid; i < UT; i += 256) s_fu[i] = a.fu[i] * up;
    for (int i = tid; i < (a.fd_2d ? DT * DT : DT); i += 256) s_fd[i] = a.fd[i];

    const TileGeo gx = tile_geo(tx * kGenOT, kGenOT, up, down, UT, DT, a.px0);
    const TileGeo gy = tile_geo(ty * kGenOT, kGenOT, up, down, UT, DT, a.py0);
    const float bias = a.bias ? a.bias[c] : 0.0f;
    const __half* xp = a.x + (static_cast<long long>(b) * a.C + c) * a.Hin * a.Wp_in;

    // load input tile (+bias), zeros outside the image
    for (int idx = tid; idx < gy.IN_T * gx.IN_T; idx += 256) {
        const int ly = idx / gx.IN_T, lx = idx % gx.IN_T;
        const int iy = gy.i_lo + ly, ix = gx.i_lo + lx;
        float v = 0.0f;
        if (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) v = __half2float(xp[static_cast<long long>(iy) * a.Wp_in + ix]) + bias;
        s_in[ly * max_in + lx] = v;
    }
    __syncthreads();
    // pass A: horizontal upsampling FIR on the real input rows
    for (int idx = tid; idx < gy.IN_T * gx.TT; idx += 256) {
        const int ly = idx / gx.TT, lj = idx % gx.TT;
        const int J = gx.J0 + lj;
        int k0 = (a.px0 - J) % up;
        if (k0 < 0) k0 += up;
        float acc = 0.0f;
        for (int k = k0; k < UT; k += up) {
            const int i = (J + k - a.px0) / up - gx.i_lo;
            acc = fmaf(s_fu[UT - 1 - k], s_in[ly * max_in + i], acc);
        }
        s_uh[ly * max_tt + lj] = acc;
    }
    __syncthreads();
    // pass B: vertical upsampling FIR + activation
    for (int idx = tid; idx < gy.TT * gx.TT; idx += 256) {
        const int lJ = idx / gx.TT, lj = idx % gx.TT;
        const int J = gy.J0 + lJ;
        int k0 = (a.py0 - J) % up;
        if (k0 < 0) k0 += up;
        float acc = 0.0f;
        for (int k = k0; k < UT; k += up) {
            const int i = (J + k - a.py0) / up - gy.i_lo;
            acc = fmaf(s_fu[UT - 1 - k], s_uh[i * max_tt + lj], acc);
        }
        acc = (acc < 0.0f ? acc * a.slope : acc) * a.gain;
        if (a.clamp >= 0.0f) acc = fminf(fmaxf(acc, -a.clamp), a.clamp);
        s_t[lJ * max_tt + lj] = acc;
    }
    __syncthreads();
    __half* yp = a.y + (static_cast<long long>(b) * a.C + c) * a.Hout * a.Wp_out;
    const float oscale = a.scale ? a.scale[b * a.C + c] : 1.0f;
    if (a.fd_2d) {
        for (int idx = tid; idx < kGenOT * kGenOT; idx += 256) {
            const int ly = idx / kGenOT, lx = idx % kGenOT;
            const int oy = gy.o0 + ly, ox = gx.o0 + lx;
            if (oy >= a.Hout || ox >= a.Wout) continue;
            float acc = 0.0f;
            for (int ky = 0; ky < DT; ++ky)
                for (int kx = 0; kx < DT; ++kx)
                    acc = fmaf(s_fd[(DT - 1 - ky) * DT + (DT - 1 - kx)], s_t[(ly * down + ky) * max_tt + lx * down + kx], acc);
            yp[static_cast<long long>(oy) * a.Wp_out + ox] = __float2half_rn(acc * oscale);
        }
        return;
    }
    // pass C: horizontal decimating FIR
    for (int idx = tid; idx < gy.TT * kGenOT; idx += 256) {
        const int lJ = idx / kGenOT, lx = idx % kGenOT;
        float acc = 0.0f;
        for (int k = 0; k < DT; ++k) acc = fmaf(s_fd[DT - 1 - k], s_t[lJ * max_tt + lx * down + k], acc);
        s_dh[lJ * kGenOT + lx] = acc;
    }
    __syncthreads();
    // pass D: vertical decimating FIR, scale by the next layer's style, store fp16
    for (int idx = tid; idx < kGenOT * kGenOT; idx += 256) {
        const int ly = idx / kGenOT, lx = idx % kGenOT;
        const int oy = gy.o0 + ly, ox = gx.o0 + lx;
        if (oy >= a.Hout || ox >= a.Wout) continue;
        float acc = 0.0f;
        for (int k = 0; k < DT; ++k) acc = fmaf(s_fd[DT - 1 - k], s_dh[(ly * down + k) * kGenOT + lx], acc);
        yp[static_cast<long long>(oy) * a.Wp_out + ox] = __float2half_rn(acc * oscale);
    }
}

}  // namespace

int flrelu_sep_launch(const FlreluArgs& a, cudaStream_t stream);  // flrelu_sep.cu (CUDA cores)
bool flrelu_sep_supported(const FlreluArgs& a);
int flrelu_mma_launch(const FlreluArgs& a, cudaStream_t stream);  // flrelu_mma.cu (tensor cores)
bool flrelu_mma_supported(const FlreluArgs& a);

static int flrelu_generic_launch(const FlreluArgs& a, cudaStream_t stream) {
    MB_REQUIRE(a.up >= 1 && a.down >= 1 && a.up_taps >= 1 && a.up_taps <= 32 && a.down_taps >= 1 &&
                   a.down_taps <= 12,
               "filtered_lrelu: unsupported filter configuration up=%d/%d taps down=%d/%d taps", a.up, a.up_taps,
               a.down, a.down_taps);
    MB_REQUIRE(a.up_taps % a.up == 0 || a.up == 1, "filtered_lrelu: up_taps must be a multiple of up");
    const int TT = (kGenOT - 1) * a.down + a.down_taps;
    const int max_in = (TT + a.up_taps - 2) / a.up + 3;
    const int tiles_x = ceil_div(a.Wout, kGenOT), tiles_y = ceil_div(a.Hout, kGenOT);
    const size_t smem = sizeof(float) * (32 + 160 + static_cast<size_t>(max_in) * max_in +
                                         static_cast<size_t>(max_in) * TT + static_cast<size_t>(TT) * TT +
                                         static_cast<size_t>(TT) * kGenOT);
    MB_REQUIRE(smem <= 200 * 1024, "filtered_lrelu: tile does not fit shared memory (%zu bytes)", smem);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        MB_CUDA(cudaFuncSetAttribute(flrelu_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
        smem_set = smem;
    }
    dim3 grid(tiles_x * tiles_y, a.C, a.B);
    flrelu_generic_kernel<<<grid, 256, smem, stream>>>(a, tiles_x, max_in, TT);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

// impl: 0 = best available (tensor-core chain, else generic), 1 = generic loops, 2 = CUDA-core polyphase kernel
int flrelu_launch_impl(const FlreluArgs& a, int impl, cudaStream_t stream) {
    if (a.B == 0 || a.C == 0) return MB_OK;
    if (impl == 0 && flrelu_mma_supported(a)) return flrelu_mma_launch(a, stream);
    MB_REQUIRE(a.y_nhwc == nullptr, "filtered_lrelu: channels-last output needs the tensor-core kernel");
    if (impl == 2 && flrelu_sep_supported(a)) return flrelu_sep_launch(a, stream);
    return flrelu_generic_launch(a, stream);
}

int flrelu_launch(const FlreluArgs& a, cudaStream_t stream) { return flrelu_launch_impl(a, 0, stream); }

}  // namespace mb
