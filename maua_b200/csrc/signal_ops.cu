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


// ---- midpoint quantile (the reference's compiled efficient_quantile behind features/efficient_quantile/__init__.py:6-7) ----
// quantile(t, q) = method 3 of efficient_quantile.cpp:86-208 called with the quantile as a FLOAT32 tensor and NaNs dropped:
//   qs = double(float(q));  qm = qs * (size - 1);  ql = trunc(qm);  qu = ceil(qm);
//   out = float( ql == qu ? v[ql] : v[qu] - (v[qu] - v[ql]) * 0.5 )      (torch::lerp in double, weight 0.5)
// with v the ascending order statistics.  The reference partially sorts on the host (std::nth_element after a device ->
// host copy); here the two order statistics come from a most-significant-byte-first radix SELECT over order-preserving
// 32-bit keys: four 256-bin histogram passes per rank, no sort, the data never leaves the device.  One CTA: the arrays
// of this path are envelopes of a few thousand frames.
__device__ __forceinline__ uint32_t order_key(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// key of the element of rank r (0-based, NaNs excluded); all threads of the CTA call it and receive the result
__device__ uint32_t radix_select(const float* __restrict__ x, long long n, long long r, unsigned int* hist, unsigned int* sh) {
    uint32_t prefix = 0, mask = 0;
    for (int pass = 3; pass >= 0; --pass) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const int shift = 8 * pass;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const float v = x[i];
            if (v != v) continue;
            const uint32_t k = order_key(v);
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long acc = 0;
            int b = 0;
            for (; b < 255; ++b) {
                if (acc + hist[b] > r) break;
                acc += hist[b];
            }
            sh[0] = static_cast<unsigned int>(b);
            sh[1] = static_cast<unsigned int>(acc);          // elements below the chosen bin (fits: <= n of one CTA pass)
            sh[2] = static_cast<unsigned int>(acc >> 32);
        }
        __syncthreads();
        prefix |= sh[0] << shift;
        mask |= 255u << shift;
        r -= static_cast<long long>(sh[1]) | (static_cast<long long>(sh[2]) << 32);
        __syncthreads();
    }
    return prefix;
}
__global__ void __launch_bounds__(1024) quantile_mid_kernel(const float* __restrict__ x, long long n, float q, float* __restrict__ out) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int sh[4];
    __shared__ unsigned long long cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    unsigned long long mine = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) mine += (x[i] == x[i]) ? 1u : 0u;
    for (int s = 16; s > 0; s >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, s);
    if ((threadIdx.x & 31) == 0) atomicAdd(&cnt, mine);
    __syncthreads();
    const long long size = static_cast<long long>(cnt);
    if (size == 0) {
        if (threadIdx.x == 0) out[0] = __int_as_float(0x7fc00000);
        return;
    }
    const double qm = static_cast<double>(q) * static_cast<double>(size - 1);
    const long long ql = static_cast<long long>(qm), qu = static_cast<long long>(ceil(qm));
    const double lo = static_cast<double>(key_value(radix_select(x, n, ql, hist, sh)));
    double res = lo;
    if (qu > ql) {
        const double hi = static_cast<double>(key_value(radix_select(x, n, qu, hist, sh)));
        res = hi - (hi - lo) * 0.5;
    }
    if (threadIdx.x == 0) out[0] = static_cast<float>(res);
}

// ---- cascaded second-order sections (the reference's low_pass / high_pass / band_pass: scipy.signal.sosfilt of a
// Butterworth design, maua/audiovisual/audioreactive/audio.py:96-110) -------------------------------------------------
// One section in transposed direct form II (what scipy's _sosfilt runs):  y = b0 x + z0;  z0' = b1 x - a1 y + z1;
// z1' = b2 x - a2 y.  The state is linear in (z, x):  z' = A z + B x  with  A = [[-a1, 1], [-a2, 0]],
// B = [b1 - a1 b0, b2 - a2 b0].  A recurrence over 1.4 M samples is serial on the host; here the signal is cut into
// chunks: (1) every chunk runs from a zero state and keeps its end state, (2) one thread chains the true start states
// through A^L (2x2, from the host), (3) every chunk runs again from its true start state and writes y.  Double
// precision throughout, like scipy on a float64 coefficient array.
constexpr int kSosChunk = 256;
struct SosCoef {
    double b0, a1, a2, B0, B1;   // B = (b1 - a1 b0, b2 - a2 b0)
    double AL[4];                // A^L row-major
};
__global__ void sos_zero_state_kernel(const double* __restrict__ x, long long n, SosCoef c, double* __restrict__ zend) {
    const long long chunk = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long i0 = chunk * kSosChunk;
    if (i0 >= n) return;
    const long long i1 = i0 + kSosChunk < n ? i0 + kSosChunk : n;
    double z0 = 0.0, z1 = 0.0;
    for (long long i = i0; i < i1; ++i) {
        const double xv = x[i];
        const double nz0 = -c.a1 * z0 + z1 + c.B0 * xv;
        z1 = -c.a2 * z0 + c.B1 * xv;
        z0 = nz0;
    }
    zend[2 * chunk] = z0;
    zend[2 * chunk + 1] = z1;
}
__global__ void sos_chain_kernel(const double* __restrict__ zend, double* __restrict__ zstart, long long chunks, SosCoef c) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double z0 = 0.0, z1 = 0.0;
    for (long long k = 0; k < chunks; ++k) {
        zstart[2 * k] = z0;
        zstart[2 * k + 1] = z1;
        const double n0 = c.AL[0] * z0 + c.AL[1] * z1 + zend[2 * k];
        const double n1 = c.AL[2] * z0 + c.AL[3] * z1 + zend[2 * k + 1];
        z0 = n0;
        z1 = n1;
    }
}
__global__ void sos_apply_kernel(const double* __restrict__ x, double* __restrict__ y, long long n, SosCoef c,
                                 const double* __restrict__ zstart) {
    const long long chunk = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long i0 = chunk * kSosChunk;
    if (i0 >= n) return;
    const long long i1 = i0 + kSosChunk < n ? i0 + kSosChunk : n;
    double z0 = zstart[2 * chunk], z1 = zstart[2 * chunk + 1];
    for (long long i = i0; i < i1; ++i) {
        const double xv = x[i];
        y[i] = c.b0 * xv + z0;
        const double nz0 = -c.a1 * z0 + z1 + c.B0 * xv;
        z1 = -c.a2 * z0 + c.B1 * xv;
        z0 = nz0;
    }
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

extern "C" int mb_quantile_mid(const float* x, int64_t n, float q, float* out, mb_stream stream) {
    MB_REQUIRE(x && out && n >= 0, "mb_quantile_mid: bad argument");
    MB_REQUIRE(q >= 0.0f && q <= 1.0f, "mb_quantile_mid: the quantile must be in [0, 1] (got %g)", static_cast<double>(q));
    quantile_mid_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(x, n, q, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_sosfilt(const double* x, double* y, int64_t n, const double* sos_host, int n_sections, double* scratch,
                          mb_stream stream_) {
    MB_REQUIRE(x && y && sos_host && scratch && n > 0 && n_sections > 0, "mb_sosfilt: bad argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long chunks = (n + kSosChunk - 1) / kSosChunk;
    double* zend = scratch;
    double* zstart = scratch + 2 * chunks;
    const double* src = x;
    for (int s = 0; s < n_sections; ++s) {
        const double* q = sos_host + 6 * s;
        MB_REQUIRE(q[3] != 0.0, "mb_sosfilt: a0 of section %d is zero", s);
        SosCoef c;
        const double b0 = q[0] / q[3], b1 = q[1] / q[3], b2 = q[2] / q[3];
        c.b0 = b0; c.a1 = q[4] / q[3]; c.a2 = q[5] / q[3];
        c.B0 = b1 - c.a1 * b0; c.B1 = b2 - c.a2 * b0;
        // A^L by repeated squaring (L = 256 = 2^8)
        double m[4] = {-c.a1, 1.0, -c.a2, 0.0};
        for (int k = 0; k < 8; ++k) {
            const double r[4] = {m[0] * m[0] + m[1] * m[2], m[0] * m[1] + m[1] * m[3], m[2] * m[0] + m[3] * m[2], m[2] * m[1] + m[3] * m[3]};
            for (int j = 0; j < 4; ++j) m[j] = r[j];
        }
        for (int j = 0; j < 4; ++j) c.AL[j] = m[j];
        const int grid = static_cast<int>((chunks + 127) / 128);
        sos_zero_state_kernel<<<grid, 128, 0, stream>>>(src, n, c, zend);
        sos_chain_kernel<<<1, 32, 0, stream>>>(zend, zstart, chunks, c);
        sos_apply_kernel<<<grid, 128, 0, stream>>>(src, y, n, c, zstart);
        MB_CUDA(cudaGetLastError());
        src = y;   // the next section filters in place (each chunk reads x[i] before it writes y[i])
    }
    return MB_OK;
}
