// Envelope post-ops and latent sequencers on the device (rows a7/a10/a11 of SURVEY §8a):
// maua/audiovisual/audioreactive/signal.py (resample :5, normalize :27, percentile_clip :55, gaussian_filter
// :108) and latent.py (single_weighted :12, multi_weighted :21).  All tensors are [T, C] row-major float32
// with T = frames on the first axis, as in the reference.  These run once per render over a few MB; they
// are written for coalesced access along C, not tuned further.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace mb {
namespace {

// depthwise temporal Gaussian with circular padding; kernel taps built in shared memory (signal.py:125-131)
// pad_mode 1 = torch 'reflect' padding (selfsupervised/features/processing.py:11-50 with mode="reflect", used by
// salience_weighted, selfsupervised/mir.py:16-17): index -i -> i, T-1+i -> T-1-i
__global__ void gaussian_filter_kernel(const float* __restrict__ x, float* __restrict__ y, int T, int C, float sigma,
                                       int radius, int causal_mode, float causal, int pad_mode) {
    extern __shared__ float k[];  // [2*radius+1]
    __shared__ float ksum;
    for (int i = threadIdx.x; i < 2 * radius + 1; i += blockDim.x) {
        const float d = static_cast<float>(i - radius);
        float v = expf(-0.5f / (sigma * sigma) * d * d);
        if (causal_mode && i > radius) v *= causal;
        k[i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int i = 0; i < 2 * radius + 1; ++i) s += k[i];
        ksum = s;
    }
    __syncthreads();
    const float inv = 1.0f / ksum;
    const long long total = static_cast<long long>(T) * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(idx / C), c = static_cast<int>(idx - static_cast<long long>(t) * C);
        float acc = 0.0f;
        for (int i = 0; i < 2 * radius + 1; ++i) {
            int tt = t + i - radius;
            if (pad_mode == 1) {
                if (tt < 0) tt = -tt;
                if (tt > T - 1) tt = 2 * (T - 1) - tt;
            } else {
                tt %= T;
                if (tt < 0) tt += T;
            }
            acc = fmaf(k[i] * inv, x[static_cast<long long>(tt) * C + c], acc);
        }
        y[idx] = acc;
    }
}

// global min / max -> (x - min) / (max - min + eps)
__global__ void minmax_kernel(const float* __restrict__ x, long long n, float* __restrict__ mm) {
    __shared__ float smin[32], smax[32];
    float lo = 3.0e38f, hi = -3.0e38f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        lo = fminf(lo, x[i]);
        hi = fmaxf(hi, x[i]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < static_cast<int>(blockDim.x >> 5); ++i) { lo = fminf(lo, smin[i]); hi = fmaxf(hi, smax[i]); }
        mm[0] = lo;
        mm[1] = hi;
    }
}
__global__ void normalize_apply_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                       const float* __restrict__ mm, float eps) {
    const float lo = mm[0], den = (mm[1] - mm[0]) + eps;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        y[i] = (x[i] - lo) / den;
}

// salience_weighted (selfsupervised/mir.py:13-21): (short / long)^2 * envelope
__global__ void salience_kernel(const float* __restrict__ s, const float* __restrict__ l, const float* __restrict__ e,
                                float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float r = s[i] / l[i];
        out[i] = r * r * e[i];
    }
}

// merge step of latent_patch (selfsupervised/latent.py:69-78) on layers [lay0, lay1) of latents [T, L, D], in place:
// mode 0 average: (lat + seq) / 2; 1 modulate: lat * (1 - m[t]) + m[t] * seq; 2 overwrite: seq
__global__ void latent_merge_kernel(float* __restrict__ lat, const float* __restrict__ seq, const float* __restrict__ mod, int mode,
                                    int lay0, int lay1, int T, int L, int D) {
    const int span = (lay1 - lay0) * D;
    const long long total = static_cast<long long>(T) * span;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(idx / span);
        const long long o = static_cast<long long>(t) * L * D + static_cast<long long>(lay0) * D + (idx - static_cast<long long>(t) * span);
        const float a = lat[o], b = seq[o];
        float r;
        if (mode == 0) {
            r = (a + b) / 2.0f;
        } else if (mode == 1) {
            const float m = mod[t];
            r = a * (1.0f - m);     // the reference multiplies in place, then adds (two roundings)
            r = r + m * b;
        } else {
            r = b;
        }
        lat[o] = r;
    }
}

// F.interpolate(mode="linear", align_corners=False) along time: src = (dst + 0.5) * T / S - 0.5, clamped at 0
__global__ void resample_linear_kernel(const float* __restrict__ x, float* __restrict__ y, int T, int S, int C) {
    const long long total = static_cast<long long>(S) * C;
    const float scale = static_cast<float>(T) / static_cast<float>(S);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int s = static_cast<int>(idx / C), c = static_cast<int>(idx - static_cast<long long>(s) * C);
        float src = (static_cast<float>(s) + 0.5f) * scale - 0.5f;
        if (src < 0.0f) src = 0.0f;
        const int i0 = static_cast<int>(src);
        const int i1 = i0 < T - 1 ? i0 + 1 : i0;
        const float l1 = src - static_cast<float>(i0), l0 = 1.0f - l1;
        y[idx] = l0 * x[static_cast<long long>(i0) * C + c] + l1 * x[static_cast<long long>(i1) * C + c];
    }
}

// out[t, :] = sum_a (env[t,a] / sum_a' env[t,a']) * latents[a % K, :]       (latent.py:21-31)
__global__ void multi_weighted_kernel(const float* __restrict__ lat, const float* __restrict__ env, float* __restrict__ out,
                                      int T, int A, int K, int D) {
    const int t = blockIdx.y;
    float wsum = 0.0f;
    for (int a = 0; a < A; ++a) wsum += env[static_cast<long long>(t) * A + a];
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
        float acc = 0.0f;
        for (int a = 0; a < A; ++a) acc += (env[static_cast<long long>(t) * A + a] / wsum) * lat[static_cast<long long>(a % K) * D + d];
        out[static_cast<long long>(t) * D + d] = acc;
    }
}

// out[t, :] = low * (1 - e[t]) + high * e[t]                               (latent.py:12-17)
__global__ void single_weighted_kernel(const float* __restrict__ low, const float* __restrict__ high,
                                       const float* __restrict__ env, float* __restrict__ out, int T, int D) {
    const long long total = static_cast<long long>(T) * D;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(idx / D), d = static_cast<int>(idx - static_cast<long long>(t) * D);
        const float e = env[t];
        out[idx] = low[d] * (1.0f - e) + high[d] * e;
    }
}

int grid1d(long long total) {
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_gaussian_filter_ex(const float* x, float* y, int T, int C, float sigma, int causal_mode, float causal,
                                     int pad_mode, mb_stream stream) {
    MB_REQUIRE(x && y && x != y && T > 0 && C > 0 && sigma > 0.0f, "mb_gaussian_filter: bad argument");
    MB_REQUIRE(pad_mode == 0 || pad_mode == 1, "mb_gaussian_filter: pad_mode %d unknown (0 circular, 1 reflect)", pad_mode);
    int radius = static_cast<int>(sigma * 4.0f);
    if (radius > 3 * T) radius = 3 * T;
    MB_REQUIRE(radius <= T - pad_mode, "mb_gaussian_filter: radius %d too large for %d frames (the reference's short-sequence branch is not built)", radius, T);
    gaussian_filter_kernel<<<grid1d(static_cast<long long>(T) * C), 256, sizeof(float) * (2 * radius + 1),
                             static_cast<cudaStream_t>(stream)>>>(x, y, T, C, sigma, radius, causal_mode, causal, pad_mode);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_gaussian_filter(const float* x, float* y, int T, int C, float sigma, int causal_mode, float causal,
                                  mb_stream stream) {
    return mb_gaussian_filter_ex(x, y, T, C, sigma, causal_mode, causal, 0, stream);
}

extern "C" int mb_salience(const float* short_env, const float* long_env, const float* envelope, float* out, int64_t n, mb_stream stream) {
    MB_REQUIRE(short_env && long_env && envelope && out && n > 0, "mb_salience: bad argument");
    salience_kernel<<<grid1d(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(short_env, long_env, envelope, out, n);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_latent_merge(float* latents, const float* sequence, const float* modulation, int mode, int lay0, int lay1, int T,
                               int L, int D, mb_stream stream) {
    MB_REQUIRE(latents && sequence && T > 0 && L > 0 && D > 0, "mb_latent_merge: bad argument");
    MB_REQUIRE(mode >= 0 && mode <= 2 && (mode != 1 || modulation), "mb_latent_merge: mode %d / modulation mismatch", mode);
    if (lay1 > L) lay1 = L;   // python slicing clips (the reference's slices reach 18 layers whatever num_ws is)
    if (lay0 >= lay1) return MB_OK;
    latent_merge_kernel<<<grid1d(static_cast<long long>(T) * (lay1 - lay0) * D), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        latents, sequence, modulation, mode, lay0, lay1, T, L, D);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_normalize(const float* x, float* y, int64_t n, float eps, float* scratch2, mb_stream stream) {
    MB_REQUIRE(x && y && scratch2 && n > 0, "mb_normalize: bad argument");
    minmax_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(x, n, scratch2);
    normalize_apply_kernel<<<grid1d(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, scratch2, eps);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_resample_linear(const float* x, float* y, int T, int S, int C, mb_stream stream) {
    MB_REQUIRE(x && y && T > 0 && S > 0 && C > 0, "mb_resample_linear: bad argument");
    resample_linear_kernel<<<grid1d(static_cast<long long>(S) * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, T, S, C);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_multi_weighted(const float* latents, const float* envelopes, float* out, int T, int A, int K, int D,
                                 mb_stream stream) {
    MB_REQUIRE(latents && envelopes && out && T > 0 && A > 0 && K > 0 && D > 0, "mb_multi_weighted: bad argument");
    dim3 grid((D + 255) / 256 < 64 ? (D + 255) / 256 : 64, T);
    multi_weighted_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(latents, envelopes, out, T, A, K, D);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_single_weighted(const float* low, const float* high, const float* envelope, float* out, int T, int D,
                                  mb_stream stream) {
    MB_REQUIRE(low && high && envelope && out && T > 0 && D > 0, "mb_single_weighted: bad argument");
    single_weighted_kernel<<<grid1d(static_cast<long long>(T) * D), 256, 0, static_cast<cudaStream_t>(stream)>>>(low, high, envelope, out, T, D);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}
