// Image-space helpers of the render path (rows a12 / a22 of SURVEY §8a):
//   maua/ops/image.py:214-240 resample: separable Lanczos-2 prefilter (reflect padding) when shrinking, then bicubic
//     interpolation with align_corners=True (MauaPatch.force_output_size);
//   maua/GAN/wrappers/stylegan2.py:196-213 make_noise_pyramid: bicubic resize (align_corners=False) of a noise map to
//     each layer's noise size, divided by its per-sample standard deviation;
//   maua/ops/noise.py:27-88 perlin_noise: 3-D gradient noise from host-drawn gradient angles.
// The bicubic kernel follows torch's upsample_bicubic2d (A = -0.75, source index NOT clamped, taps clamped to the
// image), so F.interpolate(mode="bicubic") -- what the reference calls -- is the parity target.  Planar float32.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace mb {
namespace {

__device__ __forceinline__ void cubic_w(float t, float (&w)[4]) {
    const float A = -0.75f;
    const float x0 = t + 1.0f, x3 = 2.0f - t, x2 = 1.0f - t;
    w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
    w[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
    w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
    w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

__global__ void bicubic_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int h, int w, int oh, int ow,
                               int align_corners) {
    const long long total = static_cast<long long>(planes) * oh * ow;
    const float sh = align_corners ? (oh > 1 ? static_cast<float>(h - 1) / static_cast<float>(oh - 1) : 0.0f)
                                   : static_cast<float>(h) / static_cast<float>(oh);
    const float sw = align_corners ? (ow > 1 ? static_cast<float>(w - 1) / static_cast<float>(ow - 1) : 0.0f)
                                   : static_cast<float>(w) / static_cast<float>(ow);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(idx % ow);
        const int oy = static_cast<int>((idx / ow) % oh);
        const long long pl = idx / (static_cast<long long>(ow) * oh);
        const float ry = align_corners ? sh * static_cast<float>(oy) : sh * (static_cast<float>(oy) + 0.5f) - 0.5f;
        const float rx = align_corners ? sw * static_cast<float>(ox) : sw * (static_cast<float>(ox) + 0.5f) - 0.5f;
        const float fy = floorf(ry), fx = floorf(rx);
        const int iy = static_cast<int>(fy), ix = static_cast<int>(fx);
        float wy[4], wx[4];
        cubic_w(ry - fy, wy);
        cubic_w(rx - fx, wx);
        const float* src = x + pl * h * w;
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int yy = min(max(iy - 1 + j, 0), h - 1);
            float row = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int xx = min(max(ix - 1 + i, 0), w - 1);
                row += wx[i] * src[yy * w + xx];
            }
            acc += wy[j] * row;
        }
        y[idx] = acc;
    }
}

// 1-D FIR along H (axis = 0) or W (axis = 1) with torch 'reflect' padding of (taps - 1) / 2 on both sides
__global__ void fir_reflect_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int h, int w, const float* __restrict__ k,
                                   int taps, int axis) {
    const long long total = static_cast<long long>(planes) * h * w;
    const int pad = (taps - 1) / 2;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int xx = static_cast<int>(idx % w);
        const int yy = static_cast<int>((idx / w) % h);
        const float* src = x + (idx / (static_cast<long long>(w) * h)) * h * w;
        const int n = axis == 0 ? h : w;
        float acc = 0.0f;
        for (int t = 0; t < taps; ++t) {
            int p = (axis == 0 ? yy : xx) + t - pad;
            if (p < 0) p = -p;
            if (p >= n) p = 2 * (n - 1) - p;
            acc = fmaf(k[t], axis == 0 ? src[p * w + xx] : src[yy * w + p], acc);
        }
        y[idx] = acc;
    }
}

// x[s, :] /= std(x[s, :]) (unbiased, as torch.std); one CTA per sample, two-pass mean / variance
__global__ void __launch_bounds__(256) std_normalize_kernel(float* __restrict__ x, long long per) {
    __shared__ float red[8];
    __shared__ float s_val;
    float* p = x + blockIdx.x * per;
    auto bsum = [&](float v) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.0f;
            for (int i = 0; i < 8; ++i) s += red[i];
            s_val = s;
        }
        __syncthreads();
        return s_val;
    };
    float s = 0.0f;
    for (long long i = threadIdx.x; i < per; i += blockDim.x) s += p[i];
    const float mean = bsum(s) / static_cast<float>(per);
    float v = 0.0f;
    for (long long i = threadIdx.x; i < per; i += blockDim.x) {
        const float d = p[i] - mean;
        v += d * d;
    }
    const float sd = sqrtf(bsum(v) / static_cast<float>(per - 1));
    for (long long i = threadIdx.x; i < per; i += blockDim.x) p[i] = p[i] / sd;
}

// perlin_noise (ops/noise.py:27-88): grad [r0+1][r1+1][r2+1][3] unit gradients (tileable wrap already applied by the host)
__global__ void perlin_kernel(const float* __restrict__ grad, int s0, int s1, int s2, int r0, int r1, int r2, float* __restrict__ out) {
    const long long total = static_cast<long long>(s0) * s1 * s2;
    const int d0 = s0 / r0, d1 = s1 / r1, d2 = s2 / r2;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(idx % s2), j = static_cast<int>((idx / s2) % s1), i = static_cast<int>(idx / (static_cast<long long>(s2) * s1));
        // np.mgrid[0:res:delta] % 1 in float32: position inside the lattice cell
        float g[3];
        const int ii[3] = {i, j, k};
        const int ss[3] = {s0, s1, s2}, rr[3] = {r0, r1, r2};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float pos = static_cast<float>(static_cast<double>(ii[a]) * (static_cast<double>(rr[a]) / ss[a]));
            g[a] = pos - floorf(pos);
        }
        const int c0 = i / d0, c1 = j / d1, c2 = k / d2;
        auto dotg = [&](int a, int b, int c) {
            const float* gv = grad + ((static_cast<long long>(c0 + a) * (r1 + 1) + (c1 + b)) * (r2 + 1) + (c2 + c)) * 3;
            return (g[0] - a) * gv[0] + (g[1] - b) * gv[1] + (g[2] - c) * gv[2];
        };
        float t[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) t[a] = g[a] * g[a] * g[a] * (g[a] * (g[a] * 6.0f - 15.0f) + 10.0f);
        const float n00 = dotg(0, 0, 0) * (1.0f - t[0]) + t[0] * dotg(1, 0, 0);
        const float n10 = dotg(0, 1, 0) * (1.0f - t[0]) + t[0] * dotg(1, 1, 0);
        const float n01 = dotg(0, 0, 1) * (1.0f - t[0]) + t[0] * dotg(1, 0, 1);
        const float n11 = dotg(0, 1, 1) * (1.0f - t[0]) + t[0] * dotg(1, 1, 1);
        const float n0 = (1.0f - t[1]) * n00 + t[1] * n10;
        const float n1 = (1.0f - t[1]) * n01 + t[1] * n11;
        out[idx] = ((1.0f - t[2]) * n0 + t[2] * n1) * 2.0f - 1.0f;
    }
}

int grid_of(long long total) {
    long long g = (total + 255) / 256;
    if (g > 148 * 32) g = 148 * 32;
    return g < 1 ? 1 : static_cast<int>(g);
}


// tensor2bytes (maua/ops/io.py:47-70) on the device: float32 NCHW in [lo, hi] -> uint8 NHWC.  One thread = 4 adjacent pixels of
// one row, all channels: coalesced 16-byte reads per channel plane, 4 * C contiguous output bytes.
__global__ void frames_to_rgb24_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int B, int C, int H, int W, float lo,
                                       float inv) {
    const long long quads = (static_cast<long long>(W) + 3) / 4;
    const long long total = static_cast<long long>(B) * H * quads;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int qx = static_cast<int>(idx % quads);
        const long long r = idx / quads;
        const int h = static_cast<int>(r % H), b = static_cast<int>(r / H);
        const int w0 = qx * 4;
        const int n = W - w0 < 4 ? W - w0 : 4;
        uint8_t* o = out + ((static_cast<long long>(b) * H + h) * W + w0) * C;
        if (C == 3 && n == 4 && (W & 3) == 0) {
            // rgb24: 4 pixels = 12 bytes at a 12-byte-multiple offset -> three 32-bit stores instead of twelve byte stores
            uint8_t px[12];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 q = *reinterpret_cast<const float4*>(x + ((static_cast<long long>(b) * 3 + c) * H + h) * W + w0);
                const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) px[i * 3 + c] = static_cast<uint8_t>(rintf(fminf(fmaxf((v[i] - lo) * inv, 0.0f), 1.0f) * 255.0f));
            }
            uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                o32[k] = static_cast<uint32_t>(px[4 * k]) | (static_cast<uint32_t>(px[4 * k + 1]) << 8) |
                         (static_cast<uint32_t>(px[4 * k + 2]) << 16) | (static_cast<uint32_t>(px[4 * k + 3]) << 24);
            continue;
        }
        for (int c = 0; c < C; ++c) {
            const float* p = x + ((static_cast<long long>(b) * C + c) * H + h) * W + w0;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (n == 4 && (W & 3) == 0) {
                const float4 q = *reinterpret_cast<const float4*>(p);
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
                for (int i = 0; i < n; ++i) v[i] = p[i];
            }
            for (int i = 0; i < n; ++i)
                o[i * C + c] = static_cast<uint8_t>(rintf(fminf(fmaxf((v[i] - lo) * inv, 0.0f), 1.0f) * 255.0f));
        }
    }
}

// ---- op-level upfirdn2d / bias_act of the in-tree inference network (maua/GAN/wrappers/inference/ops.py:65-114) ----------
// These are the standalone forms of what sg2_act_kernel fuses between two convolutions; exported for callers of the ops
// themselves and for op-level parity tests.  upfirdn2d: zero insertion x `up`, padding (negative = crop),
// correlation with the 2-D filter f * gain (NOT flipped, ops.py:106-111), decimation by `down`.  Polyphase: a thread only
// visits the taps that land on a real sample.
__global__ void upfirdn2d_kernel(const float* __restrict__ x, const float* __restrict__ f, float* __restrict__ y, int planes, int H,
                                 int W, int fh, int fw, int up, int down, int px0, int py0, int Ho, int Wo, float gain) {
    const long long total = static_cast<long long>(planes) * Ho * Wo;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(idx % Wo);
        const long long r = idx / Wo;
        const int oy = static_cast<int>(r % Ho);
        const long long pl = r / Ho;
        const float* xp = x + pl * H * W;
        // out[oy][ox] = sum_{ky,kx} f[ky][kx] * xu[oy*down + ky - py0][ox*down + kx - px0],  xu[a][b] = x[a/up][b/up] on the grid
        const int by = oy * down - py0, bx = ox * down - px0;
        float acc = 0.0f;
        for (int ky = ((-by) % up + up) % up; ky < fh; ky += up) {
            const int iy = (by + ky) / up;
            if (by + ky < 0 || iy >= H) continue;
            for (int kx = ((-bx) % up + up) % up; kx < fw; kx += up) {
                const int ix = (bx + kx) / up;
                if (bx + kx < 0 || ix >= W) continue;
                acc = fmaf(f[ky * fw + kx], xp[static_cast<long long>(iy) * W + ix], acc);
            }
        }
        y[idx] = acc * gain;
    }
}
// act: 0 linear, 1 lrelu(alpha)
__global__ void bias_act_kernel(const float* __restrict__ x, const float* __restrict__ b, float* __restrict__ y, long long n, int C,
                                long long plane, int act, float alpha, float gain, float clamp) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float v = x[i];
        if (b) v += b[(i / plane) % C];
        if (act == 1) v = v < 0.0f ? v * alpha : v;
        v *= gain;
        if (clamp >= 0.0f) v = fminf(fmaxf(v, -clamp), clamp);
        y[i] = v;
    }
}
}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_resize_bicubic(const float* x, float* y, int planes, int h, int w, int out_h, int out_w, int align_corners, mb_stream stream) {
    MB_REQUIRE(x && y && planes > 0 && h > 0 && w > 0 && out_h > 0 && out_w > 0, "mb_resize_bicubic: bad argument");
    bicubic_kernel<<<grid_of(static_cast<long long>(planes) * out_h * out_w), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, planes, h, w, out_h, out_w,
                                                                                                                       align_corners);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_fir_reflect(const float* x, float* y, int planes, int h, int w, const float* taps, int n_taps, int axis, mb_stream stream) {
    MB_REQUIRE(x && y && taps && x != y && planes > 0 && h > 0 && w > 0 && n_taps > 0 && (axis == 0 || axis == 1), "mb_fir_reflect: bad argument");
    MB_REQUIRE((n_taps - 1) / 2 < (axis == 0 ? h : w), "mb_fir_reflect: reflect padding %d does not fit the axis", (n_taps - 1) / 2);
    fir_reflect_kernel<<<grid_of(static_cast<long long>(planes) * h * w), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, planes, h, w, taps, n_taps, axis);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_std_normalize(float* x, int samples, int64_t per_sample, mb_stream stream) {
    MB_REQUIRE(x && samples > 0 && per_sample > 1, "mb_std_normalize: bad argument");
    std_normalize_kernel<<<samples, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, per_sample);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_perlin_noise(const float* gradients, int s0, int s1, int s2, int r0, int r1, int r2, float* out, mb_stream stream) {
    MB_REQUIRE(gradients && out && r0 > 0 && r1 > 0 && r2 > 0 && s0 % r0 == 0 && s1 % r1 == 0 && s2 % r2 == 0,
               "mb_perlin_noise: shape must be a multiple of res");
    perlin_kernel<<<grid_of(static_cast<long long>(s0) * s1 * s2), 256, 0, static_cast<cudaStream_t>(stream)>>>(gradients, s0, s1, s2, r0, r1, r2, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_frames_to_rgb24(const float* x, uint8_t* out, int B, int C, int H, int W, float lo, float hi, mb_stream stream) {
    MB_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0 && hi > lo, "mb_frames_to_rgb24: bad argument");
    const long long total = static_cast<long long>(B) * H * ((W + 3) / 4);
    const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
    frames_to_rgb24_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, B, C, H, W, lo, 1.0f / (hi - lo));
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_upfirdn2d(const float* x, const float* f, float* y, int B, int C, int H, int W, int fh, int fw, int up, int down,
                            int px0, int px1, int py0, int py1, float gain, mb_stream stream) {
    MB_REQUIRE(x && f && y && B > 0 && C > 0 && H > 0 && W > 0 && fh > 0 && fw > 0 && up >= 1 && down >= 1, "mb_upfirdn2d: bad argument");
    const int Ho = (H * up + py0 + py1 - fh) / down + 1, Wo = (W * up + px0 + px1 - fw) / down + 1;
    MB_REQUIRE(H * up + py0 + py1 >= fh && W * up + px0 + px1 >= fw, "mb_upfirdn2d: the padded signal is smaller than the filter");
    const long long total = static_cast<long long>(B) * C * Ho * Wo;
    const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
    // ops.py:106: f * gain^(ndim / 2) with a 2-D filter = f * gain
    upfirdn2d_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, f, y, B * C, H, W, fh, fw, up, down, px0, py0, Ho, Wo, gain);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_bias_act(const float* x, const float* b, float* y, int B, int C, int H, int W, int act, float alpha, float gain, float clamp,
                           mb_stream stream) {
    MB_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0 && (act == 0 || act == 1), "mb_bias_act: bad argument (act: 0 linear, 1 lrelu)");
    const long long n = static_cast<long long>(B) * C * H * W;
    const int grid = static_cast<int>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
    bias_act_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, b, y, n, C, static_cast<long long>(H) * W, act, alpha, gain, clamp);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}
