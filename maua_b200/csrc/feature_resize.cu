// Output-size hooks of the StyleGAN3 wrapper on the device.
//
// maua/GAN/wrappers/stylegan3.py:62-117 registers a forward hook on one synthesis module that either
//   "stretch":  interpolate(output, size, mode="bicubic", align_corners=False)       (:101-104)
//   "pad-zero": pad(output, (pad_w, pad_w, pad_h, pad_h), mode="constant", value=0)  (:108-115)
// and lets every later layer run on the resized (possibly non-square) feature map.  Here the hook is a kernel
// between two layers of mb_net_forward, working on the fp16 activation in the layout the next kernel consumes:
// channels-last [B][H][W][Cp] in front of a conv, planar [B][C][H][Wp] in front of the ToRGB kernel.  The stored
// activation is already multiplied by the next layer's per-channel style; both strategies are linear per channel,
// so they commute with that scale.  Bicubic = torch's upsample_bicubic2d (A = -0.75, taps clamped to the image,
// fp32 arithmetic); negative padding crops, as torch.nn.functional.pad does.
#include "common.cuh"
#include "kernels.h"
#include "resize_taps.cuh"

namespace mb {
namespace {

// one thread = 8 channels (16 bytes) of one output pixel
__global__ void __launch_bounds__(256) resize_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int h,
                                                          int w, int Cp, int oh, int ow, int mode, int pad_t, int pad_l) {
    const int groups = Cp / 8;
    const long long total = static_cast<long long>(B) * oh * ow * groups;
    const float sh = static_cast<float>(h) / static_cast<float>(oh);
    const float sw = static_cast<float>(w) / static_cast<float>(ow);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(idx % groups);
        long long r = idx / groups;
        const int ox = static_cast<int>(r % ow); r /= ow;
        const int oy = static_cast<int>(r % oh);
        const int b = static_cast<int>(r / oh);
        const __half* src = x + static_cast<long long>(b) * h * w * Cp + g * 8;
        uint4 outv = make_uint4(0u, 0u, 0u, 0u);
        if (mode == MB_RESIZE_PAD_ZERO) {
            const int sy = oy - pad_t, sx = ox - pad_l;
            if (sy >= 0 && sy < h && sx >= 0 && sx < w)
                outv = *reinterpret_cast<const uint4*>(src + (static_cast<long long>(sy) * w + sx) * Cp);
        } else {
            const Taps t = make_taps(oy, ox, h, w, sh, sw);
            float acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float row[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) row[c] = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 v = *reinterpret_cast<const uint4*>(src + (static_cast<long long>(t.iy[j]) * w + t.ix[i]) * Cp);
                    const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float2 f = __half22float2(hv[c]);
                        row[2 * c] += t.wx[i] * f.x;
                        row[2 * c + 1] += t.wx[i] * f.y;
                    }
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] += t.wy[j] * row[c];
            }
            __half2* ov = reinterpret_cast<__half2*>(&outv);
#pragma unroll
            for (int c = 0; c < 4; ++c) ov[c] = __floats2half2_rn(acc[2 * c], acc[2 * c + 1]);
        }
        *reinterpret_cast<uint4*>(y + ((static_cast<long long>(b) * oh + oy) * ow + ox) * Cp + g * 8) = outv;
    }
}

// one thread = one output element of a planar map (row pitch Wp / Wpo; the pitch padding is written as zero)
__global__ void __launch_bounds__(256) resize_planar_kernel(const __half* __restrict__ x, __half* __restrict__ y, int planes,
                                                            int h, int w, int Wp, int oh, int ow, int Wpo, int mode,
                                                            int pad_t, int pad_l) {
    const long long total = static_cast<long long>(planes) * oh * Wpo;
    const float sh = static_cast<float>(h) / static_cast<float>(oh);
    const float sw = static_cast<float>(w) / static_cast<float>(ow);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(idx % Wpo);
        const int oy = static_cast<int>((idx / Wpo) % oh);
        const long long pl = idx / (static_cast<long long>(Wpo) * oh);
        const __half* src = x + pl * h * Wp;
        float acc = 0.0f;
        if (ox < ow) {
            if (mode == MB_RESIZE_PAD_ZERO) {
                const int sy = oy - pad_t, sx = ox - pad_l;
                if (sy >= 0 && sy < h && sx >= 0 && sx < w) acc = __half2float(src[sy * Wp + sx]);
            } else {
                const Taps t = make_taps(oy, ox, h, w, sh, sw);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float row = 0.0f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) row += t.wx[i] * __half2float(src[t.iy[j] * Wp + t.ix[i]]);
                    acc += t.wy[j] * row;
                }
            }
        }
        y[idx] = __float2half_rn(acc);
    }
}

inline int grid_for(long long total, int num_sms) {
    const long long want = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms > 0 ? num_sms : 148) * 8;
    return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

int resize_nhwc_launch(const __half* x, __half* y, int B, int h, int w, int Cp, int oh, int ow, int mode, int pad_t, int pad_l,
                       int num_sms, cudaStream_t stream) {
    MB_REQUIRE(Cp % 8 == 0, "resize_nhwc: channel pitch %d is not a multiple of 8", Cp);
    MB_REQUIRE(h > 0 && w > 0 && oh > 0 && ow > 0, "resize_nhwc: empty feature map (%dx%d -> %dx%d)", h, w, oh, ow);
    const long long total = static_cast<long long>(B) * oh * ow * (Cp / 8);
    resize_nhwc_kernel<<<grid_for(total, num_sms), 256, 0, stream>>>(x, y, B, h, w, Cp, oh, ow, mode, pad_t, pad_l);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int resize_planar_launch(const __half* x, __half* y, int planes, int h, int w, int Wp, int oh, int ow, int Wpo, int mode,
                         int pad_t, int pad_l, int num_sms, cudaStream_t stream) {
    MB_REQUIRE(h > 0 && w > 0 && oh > 0 && ow > 0, "resize_planar: empty feature map (%dx%d -> %dx%d)", h, w, oh, ow);
    const long long total = static_cast<long long>(planes) * oh * Wpo;
    resize_planar_kernel<<<grid_for(total, num_sms), 256, 0, stream>>>(x, y, planes, h, w, Wp, oh, ow, Wpo, mode, pad_t, pad_l);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

}  // namespace mb
