// Small kernels around the two hot ops of the StyleGAN3 synthesis path: weight packing,
// style / demodulation coefficients, the Fourier-feature input layer, the fused ToRGB+output
// layer, layout conversion helpers and a plain CUDA-core convolution used to bisect the
// tensor-core kernel in the parity tests.
// Reference semantics: upstream networks_stylegan3.py (SynthesisInput.forward,
// SynthesisLayer.forward, modulated_conv2d, SynthesisNetwork.forward); call sites
// maua/GAN/wrappers/stylegan3.py:33,51-60.
#include "common.cuh"
#include "kernels.h"

namespace mb {

size_t packed_weight_elems(int Cout, int Cin, int ksz) {
    const size_t nCC = (Cin + 63) / 64;
    return static_cast<size_t>(round_up(Cout, 128)) * ksz * ksz * nCC * 64;
}

namespace {

__device__ __forceinline__ float block_reduce_sum(float v, float* s_red) {
    // s_red: >= 32 floats of shared memory
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? s_red[threadIdx.x] : 0.0f;
    if (warp == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) s_red[0] = v;
    __syncthreads();
    v = s_red[0];
    __syncthreads();
    return v;
}

// grid = round_up(Cout,128) rows; row >= Cout is zero padding.
__global__ void pack_weights_kernel(const float* __restrict__ w, __half* __restrict__ wpk, float* __restrict__ wsqT,
                                    int Cout, int Cin, int ksz, int prenorm, int flip) {
    __shared__ float s_red[32];
    const int o = blockIdx.x;
    const int nCC = (Cin + 63) / 64;
    const int kk = ksz * ksz;
    const long long Ktot = static_cast<long long>(kk) * nCC * 64;
    __half* row = wpk + o * Ktot;
    if (o >= Cout) {
        for (long long i = threadIdx.x; i < Ktot; i += blockDim.x) row[i] = __float2half_rn(0.0f);
        return;
    }
    const float* wo = w + static_cast<long long>(o) * Cin * kk;
    float scale = 1.0f;
    if (prenorm) {
        float ss = 0.0f;
        for (int i = threadIdx.x; i < Cin * kk; i += blockDim.x) ss += wo[i] * wo[i];
        ss = block_reduce_sum(ss, s_red);
        scale = rsqrtf(ss / static_cast<float>(Cin * kk));
    }
    for (long long i = threadIdx.x; i < Ktot; i += blockDim.x) {
        const int cl = static_cast<int>(i % 64);
        long long r = i / 64;
        const int kh = static_cast<int>(r % ksz); r /= ksz;
        const int cc = static_cast<int>(r % nCC); r /= nCC;
        const int kw = static_cast<int>(r);
        const int ci = cc * 64 + cl;
        float v = 0.0f;
        if (ci < Cin) v = wo[(ci * ksz + (flip ? ksz - 1 - kh : kh)) * ksz + (flip ? ksz - 1 - kw : kw)] * scale;
        row[i] = __float2half_rn(v);
    }
    if (wsqT) {
        for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
            float ss = 0.0f;
            for (int k = 0; k < kk; ++k) {
                const float v = wo[ci * kk + k] * scale;
                ss += v * v;
            }
            wsqT[static_cast<long long>(ci) * Cout + o] = ss;
        }
    }
}

__global__ void conv_simt_kernel(ConvTcArgs p, int Hout, int Wout) {
    const long long total = static_cast<long long>(p.B) * p.Cout * Hout * Wout;
    const int nCC = (p.Cin + 63) / 64;
    const int pad = p.pad;
    const long long Ktot = static_cast<long long>(p.ksz) * p.ksz * nCC * 64;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int wo = static_cast<int>(idx % Wout);
        long long r = idx / Wout;
        const int ho = static_cast<int>(r % Hout); r /= Hout;
        const int co = static_cast<int>(r % p.Cout);
        const int b = static_cast<int>(r / p.Cout);
        float acc = 0.0f;
        for (int kw = 0; kw < p.ksz; ++kw) {
            const int wi = wo + kw - pad;
            if (wi < 0 || wi >= p.Win) continue;
            for (int kh = 0; kh < p.ksz; ++kh) {
                const int hi = ho + kh - pad;
                if (hi < 0 || hi >= p.Hin) continue;
                for (int ci = 0; ci < p.Cin; ++ci) {
                    const int cc = ci >> 6, cl = ci & 63;
                    const float wv = __half2float(p.wpk[co * Ktot + (((kw * nCC + cc) * p.ksz + kh) << 6) + cl]);
                    const float xv =
                        __half2float(p.x[((static_cast<long long>(b) * p.Hin + hi) * p.Win + wi) * p.Cp_in + ci]);
                    acc = fmaf(wv, xv, acc);
                }
            }
        }
        if (p.d) acc *= p.d[b * p.Cout + co];
        if (p.bias) acc += p.bias[co];
        p.y[((static_cast<long long>(b) * p.Cout + co) * Hout + ho) * p.Wp_out + wo] = __float2half_rn(acc);
    }
}

// grid (num_layers, B), 256 threads.  styles = affine(w) [*style_scale]; demodulate: s *= rsqrt(mean s^2),
// d[o] = rsqrt(sum_i s_i^2 wsq[o,i] + 1e-8); stored style additionally carries input_gain.
__global__ void __launch_bounds__(256) styles_kernel(const __grid_constant__ StylesArgs a) {
    extern __shared__ float sm[];
    float* s_w = sm;                 // [w_dim]
    float* s_s = sm + a.w_dim;       // [Cin]
    __shared__ float s_red[32];
    const StyleLayerDesc& L = a.L[blockIdx.x];
    const int b = blockIdx.y;
    const float* w = a.ws + (static_cast<long long>(b) * a.num_ws + L.ws_index) * a.w_dim;
    for (int i = threadIdx.x; i < a.w_dim; i += blockDim.x) s_w[i] = w[i];
    __syncthreads();
    const float wgain = rsqrtf(static_cast<float>(a.w_dim));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int i = warp; i < L.Cin; i += nwarps) {
        const float* aw = L.affine_w + static_cast<long long>(i) * a.w_dim;
        float acc = 0.0f;
        for (int k = lane; k < a.w_dim; k += 32) acc = fmaf(s_w[k], aw[k] * wgain, acc);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_s[i] = (acc + L.affine_b[i]) * L.style_scale;
    }
    __syncthreads();
    if (L.demodulate && L.normalize_style) {
        float ss = 0.0f;
        for (int i = threadIdx.x; i < L.Cin; i += blockDim.x) ss += s_s[i] * s_s[i];
        ss = block_reduce_sum(ss, s_red);
        const float nrm = rsqrtf(ss / static_cast<float>(L.Cin));
        for (int i = threadIdx.x; i < L.Cin; i += blockDim.x) s_s[i] *= nrm;
        __syncthreads();
    }
    const float input_gain = L.magnitude_ema ? rsqrtf(*L.magnitude_ema) : 1.0f;
    for (int i = threadIdx.x; i < L.Cin; i += blockDim.x) L.s_out[static_cast<long long>(b) * L.Cin + i] = s_s[i] * input_gain;
    if (L.d_out) {
        for (int o = threadIdx.x; o < L.Cout; o += blockDim.x) {
            float acc = 0.0f;
            if (L.demodulate) {
                for (int i = 0; i < L.Cin; ++i) acc = fmaf(s_s[i] * s_s[i], L.wsqT[static_cast<long long>(i) * L.Cout + o], acc);
                acc = rsqrtf(acc + 1e-8f);
            } else {
                acc = 1.0f;
            }
            L.d_out[static_cast<long long>(b) * L.Cout + o] = acc;
        }
    }
}

// One block per sample: affine -> rotation/translation -> transformed frequencies, phases, amplitudes.
__global__ void __launch_bounds__(128) input_prep_kernel(InputArgs a) {
    __shared__ float s_t[4];
    __shared__ float s_m[6];
    const int b = blockIdx.x;
    const float* w = a.ws + static_cast<long long>(b) * a.num_ws * a.w_dim;  // ws[b, 0]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float wgain = rsqrtf(static_cast<float>(a.w_dim));
    {
        float acc = 0.0f;
        for (int k = lane; k < a.w_dim; k += 32) acc = fmaf(w[k], a.affine_w[warp * a.w_dim + k] * wgain, acc);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_t[warp] = acc + a.affine_b[warp];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t0 = s_t[0], t1 = s_t[1], t2 = s_t[2], t3 = s_t[3];
        const float n = sqrtf(t0 * t0 + t1 * t1);
        t0 /= n; t1 /= n; t2 /= n; t3 /= n;
        // m_r = [[t0,-t1,0],[t1,t0,0],[0,0,1]], m_t = [[1,0,-t2],[0,1,-t3],[0,0,1]]
        // m_rt = m_r @ m_t
        float mrt[9] = {t0, -t1, t0 * (-t2) + (-t1) * (-t3), t1, t0, t1 * (-t2) + t0 * (-t3), 0.f, 0.f, 1.f};
        float u[9];
        const float* tf = a.transform ? a.transform + static_cast<long long>(b) * a.transform_stride : nullptr;
        for (int i = 0; i < 9; ++i) u[i] = tf ? tf[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
        // M = mrt @ u ; keep rows 0,1
        for (int r = 0; r < 2; ++r)
            for (int cidx = 0; cidx < 3; ++cidx)
                s_m[r * 3 + cidx] = mrt[r * 3 + 0] * u[0 * 3 + cidx] + mrt[r * 3 + 1] * u[1 * 3 + cidx] + mrt[r * 3 + 2] * u[2 * 3 + cidx];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < a.C; j += blockDim.x) {
        const float f0 = a.freqs[j * 2 + 0], f1 = a.freqs[j * 2 + 1];
        // phases += freqs @ M[:2, 2];  freqs = freqs @ M[:2,:2]
        const float ph = a.phases[j] + (f0 * s_m[2] + f1 * s_m[5]);
        const float g0 = f0 * s_m[0] + f1 * s_m[3];
        const float g1 = f0 * s_m[1] + f1 * s_m[4];
        const float nrm = sqrtf(g0 * g0 + g1 * g1);
        float amp = 1.0f - (nrm - a.bandwidth) / (a.sampling_rate / 2.0f - a.bandwidth);
        amp = fminf(fmaxf(amp, 0.0f), 1.0f);
        float* o = a.scratch + (static_cast<long long>(b) * a.C + j) * 4;
        o[0] = g0; o[1] = g1; o[2] = ph; o[3] = amp;
    }
}

constexpr int kInPix = 16;   // pixels per CTA.  32 halves the L2 re-reads of the [C][C] mixing matrix but needs 177 registers (one CTA per SM): measured 0.41 -> 0.61 ms at B=8, so 16 stays
// grid (ceil(size*size/kInPix), B), 256 threads: Fourier features for kInPix pixels, then the channel mix (fp32 FFMA: exact
// parity with the oracle matters more here than tensor-core speed; the whole layer is 0.7 GFLOP/frame).
// Features are staged [channel j][kInPix pixels] so one LDS.128 feeds four FMAs, and a thread owns output channels
// tid and tid + 256 so every staged value is used twice: 4 LDS + 2 LDG per 32 FMA (was 16 LDS + 1 LDG per 16 FMA).
__global__ void __launch_bounds__(256) input_feat_kernel(InputArgs a) {
    extern __shared__ __align__(16) float sm[];  // [C][kInPix]
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * kInPix;
    const int npix = a.size * a.size;
    const float theta = 0.5f * static_cast<float>(a.size) / a.sampling_rate;
    const float* sc = a.scratch + static_cast<long long>(b) * a.C * 4;
    for (int idx = threadIdx.x; idx < kInPix * a.C; idx += blockDim.x) {
        const int j = idx / kInPix, pl = idx - j * kInPix;
        const int pix = p0 + pl;
        float v = 0.0f;
        if (pix < npix) {
            const int h = pix / a.size, w = pix - h * a.size;
            // F.affine_grid(align_corners=False): base coordinate (2i+1)/size - 1, times theta
            const float gx = ((2.0f * w + 1.0f) / static_cast<float>(a.size) - 1.0f) * theta;
            const float gy = ((2.0f * h + 1.0f) / static_cast<float>(a.size) - 1.0f) * theta;
            const float4 f = *reinterpret_cast<const float4*>(sc + j * 4);
            float arg = gx * f.x + gy * f.y;
            arg = arg + f.z;
            v = sinf(arg * 6.283185307179586f) * f.w;
        }
        sm[j * kInPix + pl] = v;
    }
    __syncthreads();
    const float wscale = rsqrtf(static_cast<float>(a.C));
    for (int c0 = threadIdx.x; c0 < a.C; c0 += 2 * blockDim.x) {
        const int c1 = c0 + blockDim.x;
        const bool two = c1 < a.C;
        float acc0[kInPix], acc1[kInPix];
#pragma unroll
        for (int i = 0; i < kInPix; ++i) acc0[i] = acc1[i] = 0.0f;
        for (int j = 0; j < a.C; ++j) {
            const float* wrow = a.weightT + static_cast<long long>(j) * a.C;
            const float w0 = wrow[c0] * wscale;
            const float w1 = two ? wrow[c1] * wscale : 0.0f;
#pragma unroll
            for (int q = 0; q < kInPix / 4; ++q) {
                const float4 x = *reinterpret_cast<const float4*>(sm + j * kInPix + 4 * q);
                acc0[4 * q + 0] = fmaf(x.x, w0, acc0[4 * q + 0]); acc1[4 * q + 0] = fmaf(x.x, w1, acc1[4 * q + 0]);
                acc0[4 * q + 1] = fmaf(x.y, w0, acc0[4 * q + 1]); acc1[4 * q + 1] = fmaf(x.y, w1, acc1[4 * q + 1]);
                acc0[4 * q + 2] = fmaf(x.z, w0, acc0[4 * q + 2]); acc1[4 * q + 2] = fmaf(x.z, w1, acc1[4 * q + 2]);
                acc0[4 * q + 3] = fmaf(x.w, w0, acc0[4 * q + 3]); acc1[4 * q + 3] = fmaf(x.w, w1, acc1[4 * q + 3]);
            }
        }
        const float st0 = a.style ? a.style[static_cast<long long>(b) * a.C + c0] : 1.0f;
        const float st1 = (a.style && two) ? a.style[static_cast<long long>(b) * a.C + c1] : 1.0f;
        __half* o = a.out + static_cast<long long>(b) * npix * a.Cp;
#pragma unroll
        for (int i = 0; i < kInPix; ++i) {
            const int pix = p0 + i;
            if (pix < npix) {
                o[static_cast<long long>(pix) * a.Cp + c0] = __float2half_rn(acc0[i] * st0);
                if (two) o[static_cast<long long>(pix) * a.Cp + c1] = __float2half_rn(acc1[i] * st1);
            }
        }
    }
    // zero the channel padding [C, Cp)
    for (int idx = threadIdx.x; idx < kInPix * (a.Cp - a.C); idx += blockDim.x) {
        const int i = idx / (a.Cp - a.C), cpad = a.C + idx % (a.Cp - a.C);
        const int pix = p0 + i;
        if (pix < npix) a.out[(static_cast<long long>(b) * npix + pix) * a.Cp + cpad] = __float2half_rn(0.0f);
    }
}

// Fused ToRGB layer + network output: 1x1 modulated conv without demodulation (style already folded
// into x), + bias, clamp, * output_scale; then either f32 NCHW or the uint8 NHWC wire format.
// One thread = 8 consecutive pixels of a row (16-byte loads per channel plane).
// COUT = a.Cout as a compile-time constant (3 for RGB: every accumulator index is static, no local-memory array)
template <int COUT>
__global__ void __launch_bounds__(256) torgb_out_kernel(ToRgbArgs a) {
    __shared__ float s_w[4 * 64];
    __shared__ float s_b[4];
    for (int i = threadIdx.x; i < COUT * a.Cin; i += blockDim.x) s_w[i] = a.w[i];
    if (threadIdx.x < COUT) s_b[threadIdx.x] = a.bias ? a.bias[threadIdx.x] : 0.0f;
    __syncthreads();
    const int groups_per_row = a.Wp / 8;
    const long long total = static_cast<long long>(a.B) * a.H * groups_per_row;
    const long long plane = static_cast<long long>(a.H) * a.Wp;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(idx % groups_per_row);
        long long r = idx / groups_per_row;
        const int h = static_cast<int>(r % a.H);
        const int b = static_cast<int>(r / a.H);
        const int w0 = g * 8;
        if (w0 >= a.W) continue;
        float acc[COUT][8];
#pragma unroll
        for (int o = 0; o < COUT; ++o)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[o][i] = 0.0f;
        const __half* xp = a.x + static_cast<long long>(b) * a.Cin * plane + static_cast<long long>(h) * a.Wp + w0;
#pragma unroll 8
        for (int ci = 0; ci < a.Cin; ++ci) {
            const uint4 raw = *reinterpret_cast<const uint4*>(xp + ci * plane);
            const __half2* hp = reinterpret_cast<const __half2*>(&raw);
            float xv[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(hp[i]);
                xv[2 * i] = f.x; xv[2 * i + 1] = f.y;
            }
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                const float wv = s_w[o * a.Cin + ci];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[o][i] = fmaf(wv, xv[i], acc[o][i]);
            }
        }
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = acc[o][i] + s_b[o];
                if (a.clamp >= 0.0f) v = fminf(fmaxf(v, -a.clamp), a.clamp);
                acc[o][i] = v * a.output_scale;
            }
        }
        if (a.out_fmt == MB_OUT_F32_NCHW || a.out_fmt == MB_OUT_F32_NCHW_01 || a.out_fmt == MB_OUT_F32_NCHW_UNIT) {
            const bool unit = a.out_fmt == MB_OUT_F32_NCHW_01;   // (x + 1) / 2 clamped to [0, 1]
            const bool unit_raw = a.out_fmt == MB_OUT_F32_NCHW_UNIT;   // (x + 1) / 2, not clamped
            float* out = static_cast<float*>(a.out);
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                float* op = out + ((static_cast<long long>(b) * COUT + o) * a.H + h) * a.W + w0;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (w0 + i < a.W)
                        op[i] = unit ? fminf(fmaxf((acc[o][i] + 1.0f) * 0.5f, 0.0f), 1.0f) : (unit_raw ? (acc[o][i] + 1.0f) * 0.5f : acc[o][i]);
            }
        } else {
            uint8_t* out = static_cast<uint8_t*>(a.out);
            uint8_t* op = out + ((static_cast<long long>(b) * a.H + h) * a.W + w0) * COUT;
            if (COUT == 3 && w0 + 8 <= a.W && (a.W & 7) == 0) {
                // 8 RGB pixels = 24 bytes at a 24-byte-multiple offset: three 8-byte stores instead of 24 single bytes
                uint32_t pk[6] = {0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int o = 0; o < 3; ++o) {
                        float v = (acc[o][i] + 1.0f) * 0.5f;
                        v = fminf(fmaxf(v, 0.0f), 1.0f);
                        const uint32_t q = static_cast<uint32_t>(rintf(v * 255.0f));
                        const int byte = i * 3 + o;
                        pk[byte >> 2] |= q << (8 * (byte & 3));
                    }
                }
                uint2* o8 = reinterpret_cast<uint2*>(op);
                o8[0] = make_uint2(pk[0], pk[1]);
                o8[1] = make_uint2(pk[2], pk[3]);
                o8[2] = make_uint2(pk[4], pk[5]);
                continue;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (w0 + i < a.W) {
#pragma unroll
                    for (int o = 0; o < COUT; ++o) {
                        float v = (acc[o][i] + 1.0f) * 0.5f;
                        v = fminf(fmaxf(v, 0.0f), 1.0f);
                        op[i * COUT + o] = static_cast<uint8_t>(rintf(v * 255.0f));
                    }
                }
            }
        }
    }
}

__global__ void modulate_to_half_kernel(const float* __restrict__ x, const float* __restrict__ s, float gain,
                                        __half* __restrict__ out, int B, int C, int H, int W, int Wp,
                                        const float* __restrict__ bias) {
    const long long total = static_cast<long long>(B) * C * H * Wp;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int w = static_cast<int>(idx % Wp);
        const long long r = idx / Wp;  // (b*C + c)*H + h
        const long long bc = r / H;
        float v = 0.0f;
        if (w < W) {
            v = x[r * W + w] * gain;
            if (s) v *= s[bc];
            if (bias) v += bias[bc % C];
        }
        out[idx] = __float2half_rn(v);
    }
}

__global__ void half_to_float_kernel(const __half* __restrict__ x, float* __restrict__ out, int B, int C, int H, int W,
                                     int Wp) {
    const long long total = static_cast<long long>(B) * C * H * W;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int w = static_cast<int>(idx % W);
        const long long r = idx / W;
        out[idx] = __half2float(x[r * Wp + w]);
    }
}

__global__ void modulate_to_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ s, float gain,
                                        __half* __restrict__ out, int B, int C, int H, int W, int Cp) {
    const long long total = static_cast<long long>(B) * H * W * Cp;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % Cp);
        long long r = idx / Cp;
        const int w = static_cast<int>(r % W); r /= W;
        const int h = static_cast<int>(r % H);
        const int b = static_cast<int>(r / H);
        float v = 0.0f;
        if (c < C) {
            v = x[((static_cast<long long>(b) * C + c) * H + h) * W + w] * gain;
            if (s) v *= s[b * C + c];
        }
        out[idx] = __float2half_rn(v);
    }
}

__global__ void nhwc_to_float_kernel(const __half* __restrict__ x, float* __restrict__ out, int B, int C, int H, int W,
                                     int Cp) {
    const long long total = static_cast<long long>(B) * C * H * W;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int w = static_cast<int>(idx % W);
        long long r = idx / W;
        const int h = static_cast<int>(r % H); r /= H;
        const int c = static_cast<int>(r % C);
        const int b = static_cast<int>(r / C);
        out[idx] = __half2float(x[((static_cast<long long>(b) * H + h) * W + w) * Cp + c]);
    }
}

// planar -> channels-last through a shared-memory tile: one CTA = (b, h, 32 pixels) x all channels.
// Reads are 64-byte row segments per channel, writes are one contiguous run of 32*Cp halves.
constexpr int kTrP = 32;
__global__ void __launch_bounds__(256) planar_to_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ out,
                                                             int C, int H, int W, int Wp, int Cp) {
    extern __shared__ __half tile[];  // [Cp][kTrP + 2]
    const int w0 = blockIdx.x * kTrP, h = blockIdx.y, b = blockIdx.z;
    const __half2 zero2 = __float2half2_rn(0.0f);
    // load: thread -> (channel, pixel pair)
    for (int idx = threadIdx.x; idx < Cp * (kTrP / 2); idx += blockDim.x) {
        const int c = idx / (kTrP / 2), pp = idx % (kTrP / 2);
        const int w = w0 + pp * 2;
        __half2 v = zero2;
        if (c < C && w < Wp) v = *reinterpret_cast<const __half2*>(x + ((static_cast<long long>(b) * C + c) * H + h) * Wp + w);
        *reinterpret_cast<__half2*>(tile + c * (kTrP + 2) + pp * 2) = v;
    }
    __syncthreads();
    // store: thread -> (pixel, channel pair)
    const int npx = (W - w0) < kTrP ? (W - w0) : kTrP;
    __half* o = out + ((static_cast<long long>(b) * H + h) * W + w0) * Cp;
    for (int idx = threadIdx.x; idx < npx * (Cp / 2); idx += blockDim.x) {
        const int px = idx / (Cp / 2), cp2 = idx % (Cp / 2);
        __half2 v;
        v.x = tile[(cp2 * 2) * (kTrP + 2) + px];
        v.y = tile[(cp2 * 2 + 1) * (kTrP + 2) + px];
        *reinterpret_cast<__half2*>(o + static_cast<long long>(px) * Cp + cp2 * 2) = v;
    }
}

// op-level helper, grid B.
__global__ void style_demod_kernel(const float* __restrict__ s, const float* __restrict__ wsqT, float* s_out,
                                   float* d_out, int Cin, int Cout, int demodulate, float input_gain) {
    extern __shared__ float sm[];
    __shared__ float s_red[32];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < Cin; i += blockDim.x) sm[i] = s[static_cast<long long>(b) * Cin + i];
    __syncthreads();
    if (demodulate) {
        float ss = 0.0f;
        for (int i = threadIdx.x; i < Cin; i += blockDim.x) ss += sm[i] * sm[i];
        ss = block_reduce_sum(ss, s_red);
        const float nrm = rsqrtf(ss / static_cast<float>(Cin));
        for (int i = threadIdx.x; i < Cin; i += blockDim.x) sm[i] *= nrm;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < Cin; i += blockDim.x) s_out[static_cast<long long>(b) * Cin + i] = sm[i] * input_gain;
    for (int o = threadIdx.x; o < Cout; o += blockDim.x) {
        float acc = 1.0f;
        if (demodulate) {
            acc = 0.0f;
            for (int i = 0; i < Cin; ++i) acc = fmaf(sm[i] * sm[i], wsqT[static_cast<long long>(i) * Cout + o], acc);
            acc = rsqrtf(acc + 1e-8f);
        }
        d_out[static_cast<long long>(b) * Cout + o] = acc;
    }
}

static int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    if (g > 148LL * 32) g = 148LL * 32;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

}  // namespace

int pack_weights_launch(const float* w, __half* wpk, float* wsqT, int Cout, int Cin, int ksz, int prenorm,
                        cudaStream_t stream, int flip) {
    pack_weights_kernel<<<round_up(Cout, 128), 256, 0, stream>>>(w, wpk, wsqT, Cout, Cin, ksz, prenorm, flip);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int conv_simt_launch(const ConvTcArgs& p, cudaStream_t stream) {
    const int Hout = p.Hin + 2 * p.pad - (p.ksz - 1), Wout = p.Win + 2 * p.pad - (p.ksz - 1);
    const long long total = static_cast<long long>(p.B) * p.Cout * Hout * Wout;
    conv_simt_kernel<<<grid_for(total, 256), 256, 0, stream>>>(p, Hout, Wout);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int styles_launch(const StylesArgs& a, cudaStream_t stream) {
    int max_cin = 0;
    for (int i = 0; i < a.num_layers; ++i) max_cin = a.L[i].Cin > max_cin ? a.L[i].Cin : max_cin;
    const size_t smem = sizeof(float) * (a.w_dim + max_cin);
    dim3 grid(a.num_layers, a.B);
    styles_kernel<<<grid, 256, smem, stream>>>(a);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

// ---- tensor-core variant of the input layer's channel mix --------------------------------------------------------
// x[b, p, c] = sum_j feat[b, p, j] * (weight[c, j] / sqrt(C)) is a dense [pixels x C] x [C x C] contraction per frame: it
// runs on the tcgen05 1x1 conv (conv_tc.cu) with BOTH operands split into fp16 hi + lo parts,
//     feat * w ~= fh * wh + fl * wh + fh * wl      (the dropped fl * wl term is ~2^-22 relative),
// i.e. K = 3 C: features stored as [fh | fl | fh], weights as [wh | wh | wl].  Products of fp16 values are exact in the
// fp32 accumulator, so the result keeps fp32-class accuracy (the CUDA-core kernel above ran at 23 % of the fp32 FMA
// peak: 0.65 ms per 16 frames).  The layer-0 style rides in the conv epilogue's per-(frame, cout) scale, so the
// activation is rounded to fp16 once, exactly like the fp32 kernel.
__global__ void __launch_bounds__(256) input_feat_split_kernel(InputArgs a, __half* __restrict__ out) {
    const int npix = a.size * a.size;
    const long long total = static_cast<long long>(a.B) * npix * a.C;
    const float theta = 0.5f * static_cast<float>(a.size) / a.sampling_rate;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(idx % a.C);
        const long long q = idx / a.C;
        const int pix = static_cast<int>(q % npix);
        const int b = static_cast<int>(q / npix);
        const int h = pix / a.size, w = pix - h * a.size;
        const float gx = ((2.0f * w + 1.0f) / static_cast<float>(a.size) - 1.0f) * theta;
        const float gy = ((2.0f * h + 1.0f) / static_cast<float>(a.size) - 1.0f) * theta;
        const float4 f = *reinterpret_cast<const float4*>(a.scratch + (static_cast<long long>(b) * a.C + j) * 4);
        float arg = gx * f.x + gy * f.y;
        arg = arg + f.z;
        const float v = sinf(arg * 6.283185307179586f) * f.w;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        __half* o = out + q * (3LL * a.C);
        o[j] = hi;
        o[a.C + j] = lo;
        o[2 * a.C + j] = hi;
    }
}

// weight [C][C] f32 -> w3 [C][3C] f32 = [wh | wh | wl] of weight / sqrt(C) (every entry exactly representable in fp16)
__global__ void input_w3_kernel(const float* __restrict__ w, float* __restrict__ w3, int C) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C * C) return;
    const int c = idx / C, j = idx - c * C;
    const float v = w[idx] * rsqrtf(static_cast<float>(C));
    const float hi = __half2float(__float2half_rn(v));
    const float lo = __half2float(__float2half_rn(v - hi));
    float* o = w3 + static_cast<long long>(c) * 3 * C;
    o[j] = hi;
    o[C + j] = hi;
    o[2 * C + j] = lo;
}

int sg3_input_split_weights_launch(const float* weight, float* w3, int C, cudaStream_t stream) {
    input_w3_kernel<<<ceil_div(C * C, 256), 256, 0, stream>>>(weight, w3, C);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

// input_prep + split features into `feat` (channels-last [B][size][size][3C] fp16); the caller runs the 1x1 conv
int sg3_input_features_launch(const InputArgs& a, __half* feat, cudaStream_t stream) {
    MB_REQUIRE(a.C % 16 == 0, "input layer: %d channels is not a multiple of 16 (tensor-core path)", a.C);
    input_prep_kernel<<<a.B, 128, 0, stream>>>(a);
    MB_CUDA(cudaGetLastError());
    const long long total = static_cast<long long>(a.B) * a.size * a.size * a.C;
    input_feat_split_kernel<<<grid_for(total, 256), 256, 0, stream>>>(a, feat);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int sg3_input_launch(const InputArgs& a, cudaStream_t stream) {
    input_prep_kernel<<<a.B, 128, 0, stream>>>(a);
    MB_CUDA(cudaGetLastError());
    const size_t smem = sizeof(float) * kInPix * a.C;
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        MB_CUDA(cudaFuncSetAttribute(input_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        smem_set = smem;
    }
    dim3 grid(ceil_div(a.size * a.size, kInPix), a.B);
    input_feat_kernel<<<grid, 256, smem, stream>>>(a);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int torgb_out_launch(const ToRgbArgs& a, cudaStream_t stream) {
    MB_REQUIRE(a.Cout <= 4 && a.Cin <= 64, "torgb: Cout<=4, Cin<=64 supported (got %d, %d)", a.Cout, a.Cin);
    const long long total = static_cast<long long>(a.B) * a.H * (a.Wp / 8);
    switch (a.Cout) {
        case 1: torgb_out_kernel<1><<<grid_for(total, 256), 256, 0, stream>>>(a); break;
        case 2: torgb_out_kernel<2><<<grid_for(total, 256), 256, 0, stream>>>(a); break;
        case 3: torgb_out_kernel<3><<<grid_for(total, 256), 256, 0, stream>>>(a); break;
        default: torgb_out_kernel<4><<<grid_for(total, 256), 256, 0, stream>>>(a); break;
    }
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int modulate_to_half_launch(const float* x, const float* s, float gain, __half* out, int B, int C, int H, int W, int Wp,
                            cudaStream_t stream, const float* bias) {
    const long long total = static_cast<long long>(B) * C * H * Wp;
    modulate_to_half_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, s, gain, out, B, C, H, W, Wp, bias);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int modulate_to_nhwc_launch(const float* x, const float* s, float gain, __half* out, int B, int C, int H, int W, int Cp,
                            cudaStream_t stream) {
    const long long total = static_cast<long long>(B) * H * W * Cp;
    modulate_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, s, gain, out, B, C, H, W, Cp);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int nhwc_to_float_launch(const __half* x, float* out, int B, int C, int H, int W, int Cp, cudaStream_t stream) {
    const long long total = static_cast<long long>(B) * C * H * W;
    nhwc_to_float_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, out, B, C, H, W, Cp);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int planar_to_nhwc_launch(const __half* x, __half* out, int B, int C, int H, int W, int Wp, int Cp, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(Cp) * (kTrP + 2) * sizeof(__half);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        MB_CUDA(cudaFuncSetAttribute(planar_to_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        smem_set = smem;
    }
    dim3 grid(ceil_div(W, kTrP), H, B);
    planar_to_nhwc_kernel<<<grid, 256, smem, stream>>>(x, out, C, H, W, Wp, Cp);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int half_to_float_launch(const __half* x, float* out, int B, int C, int H, int W, int Wp, cudaStream_t stream) {
    const long long total = static_cast<long long>(B) * C * H * W;
    half_to_float_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, out, B, C, H, W, Wp);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int style_demod_launch(const float* s, const float* wsqT, float* s_out, float* d_out, int B, int Cin, int Cout,
                       int demodulate, float input_gain, cudaStream_t stream) {
    style_demod_kernel<<<B, 256, sizeof(float) * Cin, stream>>>(s, wsqT, s_out, d_out, Cin, Cout, demodulate, input_gain);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

}  // namespace mb
