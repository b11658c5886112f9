// Latent and noise sequencers on the device (rows a11 / a12 of SURVEY §8a).
//   maua/audiovisual/audioreactive/latent.py: select_modulo :34-45, slerp :54-65, slerp_loops :68-80,
//     spline_loops :83-92 (natural cubic spline through the looped key latents; the reference calls the third-party
//     torchcubicspline, absent from the checkout: the published natural-spline algorithm is restated, see oracle/signal.py)
//   maua/audiovisual/audioreactive/selfsupervised/noise.py: Blend :11-25, Multiply :28-40, Loop :43-54 and the
//     Average / Modulate / ScaleBias combinators :57-86 (one fused a*x*mx[b] + b*y*my[b] + c kernel).
// All of this runs once per render (latents) or once per batch on small maps (noise): written for coalesced access and
// one pass over the data, not tuned further.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace mb {
namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.0f;
    for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) s += red[i];
    return s;
}

// One CTA per (interpolated source row r = s * nseg + m, layer): the slerp of latent.py:54-65 between looped key m and
// m+1 at t_s, written to rows[r][layer][:].
__global__ void __launch_bounds__(128) slerp_rows_kernel(const float* __restrict__ keys, int K, int L, int D, int nseg, int S,
                                                         float* __restrict__ rows) {
    __shared__ float red[4];
    const int r = blockIdx.x, l = blockIdx.y;
    const int s = r / nseg, m = r - s * nseg;
    const float t = S > 1 ? static_cast<float>(s) / static_cast<float>(S - 1) : 0.0f;   // torch.linspace(0, 1, S)
    const float* a = keys + (static_cast<long long>(m % K) * L + l) * D;
    const float* b = keys + (static_cast<long long>((m + 1) % K) * L + l) * D;
    float aa = 0.0f, bb = 0.0f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) { aa += a[i] * a[i]; bb += b[i] * b[i]; }
    const float na = sqrtf(block_sum(aa, red)), nb = sqrtf(block_sum(bb, red));
    float dd = 0.0f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) dd += (a[i] / na) * (b[i] / nb);
    const float d = block_sum(dd, red);
    const float p = t * acosf(d);
    float cc = 0.0f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const float c = b[i] / nb - d * (a[i] / na);
        cc += c * c;
    }
    const float nc = sqrtf(block_sum(cc, red));
    const float cp = cosf(p), sp = sinf(p);
    float oo = 0.0f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const float c = (b[i] / nb - d * (a[i] / na)) / nc;
        const float o = (a[i] / na) * cp + c * sp;
        oo += o * o;
    }
    const float no = sqrtf(block_sum(oo, red));
    float* out = rows + (static_cast<long long>(r) * L + l) * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const float c = (b[i] / nb - d * (a[i] / na)) / nc;
        out[i] = ((a[i] / na) * cp + c * sp) / no;
    }
}

// natural cubic spline, uniform knots: second derivatives z (z_0 = z_{M-1} = 0) by the Thomas algorithm, one thread per
// channel; y_m = keys[m % K] (torch.cat([y] * n_loops + [y[[0]]]), latent.py:86)
__global__ void spline_solve_kernel(const float* __restrict__ keys, int K, int C, int M, float* __restrict__ z) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float h = 1.0f / static_cast<float>(M - 1);
    const float k6 = 6.0f / (h * h);
    auto y = [&](int m) { return keys[static_cast<long long>(m % K) * C + c]; };
    z[c] = 0.0f;
    z[static_cast<long long>(M - 1) * C + c] = 0.0f;
    if (M < 3) return;
    // forward sweep: c'_i = 1 / (4 - c'_{i-1}), d'_i = (rhs_i - d'_{i-1}) * c'_i   (unit off-diagonals)
    float cp = 0.0f, dp = 0.0f;
    for (int i = 1; i <= M - 2; ++i) {
        const float rhs = k6 * (y(i - 1) - 2.0f * y(i) + y(i + 1));
        const float den = 4.0f - cp;
        cp = 1.0f / den;
        dp = (rhs - dp) / den;
        z[static_cast<long long>(i) * C + c] = dp;        // d'_i (c'_i is recomputed in the back substitution)
    }
    // back substitution: z_i = d'_i - c'_i z_{i+1}; c'_i recomputed from its closed recurrence in reverse is awkward,
    // so a second forward pass stores nothing: recompute c' up to i on the fly (M is a few dozen)
    float znext = 0.0f;
    for (int i = M - 2; i >= 1; --i) {
        float ci = 0.0f;
        for (int j = 1; j <= i; ++j) ci = 1.0f / (4.0f - ci);
        const float zi = z[static_cast<long long>(i) * C + c] - ci * znext;
        z[static_cast<long long>(i) * C + c] = zi;
        znext = zi;
    }
}

// t_end = 1, wrap = 0: evaluate at linspace(0, 1, size) (spline_loops).  wrap = 1: at linspace(0, t_end, size) % 1
// (selfsupervised/latent.py:7-13 spline_loop_latents; linspace as ATen computes it: from the start up to the middle,
// from the end beyond it).
__global__ void spline_eval_kernel(const float* __restrict__ keys, const float* __restrict__ z, int K, int C, int M, int size,
                                   float* __restrict__ out, float t_end, int wrap) {
    const long long total = static_cast<long long>(size) * C;
    const float h = 1.0f / static_cast<float>(M - 1);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(idx / C), c = static_cast<int>(idx - static_cast<long long>(j) * C);
        float t = size > 1 ? static_cast<float>(j) / static_cast<float>(size - 1) : 0.0f;
        if (wrap) {
            const float step = size > 1 ? t_end / static_cast<float>(size - 1) : 0.0f;
            t = j < size / 2 ? step * static_cast<float>(j) : t_end - step * static_cast<float>(size - 1 - j);
            t = fmodf(t, 1.0f);
        }
        const float u = t * static_cast<float>(M - 1);
        int i = static_cast<int>(u);
        if (i > M - 2) i = M - 2;
        const float f = u - static_cast<float>(i), g = 1.0f - f;
        const float y0 = keys[static_cast<long long>(i % K) * C + c], y1 = keys[static_cast<long long>((i + 1) % K) * C + c];
        const float z0 = z[static_cast<long long>(i) * C + c], z1 = z[static_cast<long long>(i + 1) * C + c];
        out[idx] = g * y0 + f * y1 + (h * h / 6.0f) * ((g * g * g - g) * z0 + (f * f * f - f) * z1);
    }
}

// select_modulo (latent.py:34-45) up to the final Gaussian: quantile clamp -> min-max normalise -> round to a key index
// -> gather.  `sorted` is the ascending-sorted envelope (index bookkeeping done by the caller).
__global__ void select_modulo_kernel(const float* __restrict__ env, const float* __restrict__ sorted, int T, const float* __restrict__ keys,
                                     int K, int C, float* __restrict__ out) {
    auto quant = [&](float q) {   // torch.quantile, linear interpolation
        const float pos = q * static_cast<float>(T - 1);
        const int lo = static_cast<int>(floorf(pos));
        const int hi = lo + 1 < T ? lo + 1 : lo;
        const float w = pos - static_cast<float>(lo);
        return sorted[lo] + w * (sorted[hi] - sorted[lo]);
    };
    const float low = quant(0.25f), high = quant(0.75f);
    const long long total = static_cast<long long>(T) * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(idx / C), c = static_cast<int>(idx - static_cast<long long>(t) * C);
        const float e = fminf(fmaxf(env[t], low), high);
        // signal.normalize: y = x - min; y / max(y); the clamped envelope attains both bounds whenever T >= 4
        float v = (e - low) / (high - low);
        v *= static_cast<float>(K - 1);
        int k = static_cast<int>(rintf(v));
        k = k < 0 ? 0 : (k > K - 1 ? K - 1 : k);
        out[idx] = keys[static_cast<long long>(k) * C + c];
    }
}

// Blend (two_sided) / Multiply: out[b,p] = sum_m n0[m,p] mod[b,m] (+ sum_m n1[m,p] (1 - mod[b,m]))
__global__ void noise_mix_kernel(const float* __restrict__ noise, const float* __restrict__ mod, int B, int M, int P, int two_sided,
                                 float* __restrict__ out) {
    const long long total = static_cast<long long>(B) * P;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(idx / P), p = static_cast<int>(idx - static_cast<long long>(b) * P);
        float left = 0.0f, right = 0.0f;
        for (int m = 0; m < M; ++m) {
            const float w = mod[b * M + m];
            left = fmaf(noise[static_cast<long long>(m) * P + p], w, left);
            if (two_sided) right = fmaf(noise[(static_cast<long long>(M) + m) * P + p], 1.0f - w, right);
        }
        out[idx] = left + right;
    }
}

// Loop: out[b] = sin(cos(idx[b] + n0) / (sigma / 50) + n1) * n2, divided by its per-frame RMS + eps; one CTA per frame
__global__ void __launch_bounds__(256) noise_loop_kernel(const float* __restrict__ idx, const float* __restrict__ noise, int P, float sigma,
                                                         float* __restrict__ out) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    const float ph = idx[b], div = sigma / 50.0f;
    float ss = 0.0f;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        const float f = cosf(ph + noise[p]) / div;
        const float o = sinf(f + noise[P + p]) * noise[2 * P + p];
        out[static_cast<long long>(b) * P + p] = o;
        ss += o * o;
    }
    const float rms = sqrtf(block_sum(ss, red) / static_cast<float>(P)) + 1.1920928955078125e-07f;
    for (int p = threadIdx.x; p < P; p += blockDim.x) out[static_cast<long long>(b) * P + p] /= rms;
}

// out[b,p] = a * x[b,p] * (mx ? mx[b] : 1) + c * y[b,p] * (my ? (one_minus ? 1 - my[b] : my[b]) : 1) + bias
__global__ void noise_combine_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ mx,
                                     const float* __restrict__ my, int one_minus, float a, float c, float bias, int B, int P,
                                     float* __restrict__ out) {
    const long long total = static_cast<long long>(B) * P;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(idx / P);
        float v = a * x[idx] * (mx ? mx[b] : 1.0f);
        if (y) v += c * y[idx] * (my ? (one_minus ? 1.0f - my[b] : my[b]) : 1.0f);
        out[idx] = v + bias;
    }
}

int grid_for(long long total) {
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_slerp_rows(const float* keys, int K, int L, int D, int n_loops, int steps, float* rows, mb_stream stream) {
    MB_REQUIRE(keys && rows && K > 0 && L > 0 && D > 0 && n_loops > 0 && steps > 0, "mb_slerp_rows: bad argument");
    const int nseg = K * n_loops;  // len(cat([y] * n_loops + [y[[0]]])) - 1
    dim3 grid(steps * nseg, L);
    slerp_rows_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(keys, K, L, D, nseg, steps, rows);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_spline_loops(const float* keys, int K, int C, int n_loops, int size, float* out, float* workspace, mb_stream stream) {
    MB_REQUIRE(keys && out && workspace && K > 0 && C > 0 && n_loops > 0 && size > 0, "mb_spline_loops: bad argument");
    const int M = K * n_loops + 1;
    MB_REQUIRE(M >= 2 && M <= 4096, "mb_spline_loops: %d knots unsupported", M);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    spline_solve_kernel<<<(C + 127) / 128, 128, 0, s>>>(keys, K, C, M, workspace);
    spline_eval_kernel<<<grid_for(static_cast<long long>(size) * C), 256, 0, s>>>(keys, workspace, K, C, M, size, out, 1.0f, 0);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_spline_loop_latents(const float* keys, int K, int C, float n_loops, int size, float* out, float* workspace,
                                      mb_stream stream) {
    MB_REQUIRE(keys && out && workspace && K > 0 && C > 0 && n_loops >= 0.0f && size > 0, "mb_spline_loop_latents: bad argument");
    const int M = K + 1;  // torch.cat((y, y[[0]])): ONE loop of knots, walked n_loops times by the wrapped positions
    MB_REQUIRE(M >= 2 && M <= 4096, "mb_spline_loop_latents: %d knots unsupported", M);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    spline_solve_kernel<<<(C + 127) / 128, 128, 0, s>>>(keys, K, C, M, workspace);
    spline_eval_kernel<<<grid_for(static_cast<long long>(size) * C), 256, 0, s>>>(keys, workspace, K, C, M, size, out, n_loops, 1);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_select_modulo(const float* envelope, const float* sorted_envelope, int T, const float* keys, int K, int C,
                                float* out, mb_stream stream) {
    MB_REQUIRE(envelope && sorted_envelope && keys && out && T >= 4 && K > 0 && C > 0, "mb_select_modulo: bad argument");
    select_modulo_kernel<<<grid_for(static_cast<long long>(T) * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        envelope, sorted_envelope, T, keys, K, C, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_noise_mix(const float* noise, const float* modulator, int B, int M, int P, int two_sided, float* out, mb_stream stream) {
    MB_REQUIRE(noise && modulator && out && B > 0 && M > 0 && P > 0, "mb_noise_mix: bad argument");
    noise_mix_kernel<<<grid_for(static_cast<long long>(B) * P), 256, 0, static_cast<cudaStream_t>(stream)>>>(noise, modulator, B, M, P,
                                                                                                           two_sided, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_noise_loop(const float* idx, const float* noise, int B, int P, float sigma, float* out, mb_stream stream) {
    MB_REQUIRE(idx && noise && out && B > 0 && P > 0 && sigma > 0.0f, "mb_noise_loop: bad argument");
    noise_loop_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(idx, noise, P, sigma, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_noise_combine(const float* x, const float* y, const float* mx, const float* my, int one_minus, float a, float c,
                                float bias, int B, int P, float* out, mb_stream stream) {
    MB_REQUIRE(x && out && B > 0 && P > 0, "mb_noise_combine: bad argument");
    noise_combine_kernel<<<grid_for(static_cast<long long>(B) * P), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, y, mx, my, one_minus, a, c, bias, B, P, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}
