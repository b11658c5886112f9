// StyleGAN2 synthesis network on the sm_100a kernels: replaces the reference's in-tree inference network
// maua/GAN/wrappers/inference/stylegan2.py:195-436 (SynthesisLayer / ToRGBLayer / SynthesisBlock /
// SynthesisNetwork, architecture 'skip') and its ops (inference/ops.py: modulated_conv2d :146, conv2d_resample
// :189, upfirdn2d :87, upsample2d :117, bias_act :65).
//
// Per block (resolution r):  conv0 = stride-2 transposed 3x3 modulated conv + 4x4 FIR  ->  conv1 = 3x3 'same'
// modulated conv  ->  ToRGB 1x1 (no demodulation) added to the FIR-upsampled image of the previous block.
//   * both 3x3 convs run on the tcgen05 implicit-GEMM kernel (conv_tc.cu): conv1 with padding 1; conv0 = the stride-2
//     transposed conv (ops.py:213-230) in POLYPHASE form: the four output parities (row, column even / odd) are four
//     2x2 kernels over the un-inserted input -- taps {0,2}x{0,2}, {0,2}x{1}, {1}x{0,2}, {1}x{1} of the flipped 3x3
//     kernel -- run as ONE 2x2 'full' convolution with 4*Cout output channels (phase-major); no zero-inserted copy of
//     the input exists and no MAC lands on an inserted zero (16 instead of 36 multiply-adds per input pixel, channel
//     pair and output channel);
//   * precision (`precise`, default on): activations and weights go to the tensor core as fp16 hi + lo pairs --
//     x*w = xh*wh + xl*wh + xh*wl as one contraction over 3*Cp channels [xh | xl | xh] x [wh | wh | wl], products exact
//     in the fp32 accumulator -- and the conv output leaves as an fp16 hi / lo plane pair, so the whole path carries
//     fp32-class values.  A random-init StyleGAN2 image swings over +-30: the reference's 1e-3 pixel tolerance asks
//     for ~3e-5 of that range, which fp16 storage (5e-4) cannot hold.  `precise` off = plain fp16 operands (3x fewer MACs);
//   * everything between two convs is ONE fused kernel (sg2_act_kernel): [4x4 FIR, gain 4] + noise + bias +
//     leaky-ReLU*sqrt2 + clamp (bias_act), the next conv's style and the planar -> channels-last store, plus the
//     ToRGB reduction over channels and the skip-image accumulation.
// Styles are folded into the activations and the demodulation coefficient into the conv epilogue, exactly as for
// StyleGAN3 (net.cu); the feature map after conv1 has two consumers (ToRGB and the next conv0) with different
// styles, so the fused kernel applies each while the activation tile is in shared memory.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "sg2.h"

namespace mb {
namespace {

inline int cpad16(int c) { return (c + 15) / 16 * 16; }

// Effective fp32 weights of one 3x3 layer for the conv kernel: out [Cout'][ns * Cp][k'][k'].
//   phase = 0: k' = 3, Cout' = Cout, out = w.
//   phase = 1: k' = 2, Cout' = 4 * Cout: output channel (pa * 2 + pb) * Cout + co holds the 2x2 kernel of output parity
//              (pa, pb) of the stride-2 transposed conv: tap (a', b') = flipped tap (row(pa, a'), col(pb, b')) with
//              row(0, .) = {0, 2}, row(1, .) = {none, 1} (a' = 0 reads input row u - 1, a' = 1 row u).
//   ns = 3 (precise): channel blocks [w | w | w - fp16(w)]; the fp16 packing that follows turns them into [wh | wh | wl].
__global__ void sg2_prep_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int Cp, int phase, int ns) {
    const int kk = phase ? 2 : 3, Co = phase ? 4 * Cout : Cout, Ci = ns * Cp;
    const long long total = static_cast<long long>(Co) * Ci * kk * kk;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(idx % kk);
        long long r = idx / kk;
        const int a = static_cast<int>(r % kk); r /= kk;
        const int cq = static_cast<int>(r % Ci);
        const int cop = static_cast<int>(r / Ci);
        const int part = cq / Cp, ci = cq - part * Cp;
        float v = 0.0f;
        if (ci < Cin) {
            if (!phase) {
                v = w[((static_cast<long long>(cop) * Cin + ci) * 3 + a) * 3 + b];
            } else {
                const int ph = cop / Cout, co = cop - ph * Cout, pa = ph >> 1, pb = ph & 1;
                const int fr = pa ? (a == 1 ? 1 : -1) : (a == 0 ? 0 : 2);   // row of the flipped kernel wf[r][c] = w[2 - r][2 - c]
                const int fc = pb ? (b == 1 ? 1 : -1) : (b == 0 ? 0 : 2);
                if (fr >= 0 && fc >= 0) v = w[((static_cast<long long>(co) * Cin + ci) * 3 + (2 - fr)) * 3 + (2 - fc)];
            }
            if (part == 2) v -= __half2float(__float2half_rn(v));
        }
        out[idx] = v;
    }
}
// wsqT [Cin][Cout] -> [Cin][4 * Cout] (the four parities of a transposed conv share the demodulation coefficient)
__global__ void sg2_tile4_kernel(const float* __restrict__ in, float* __restrict__ out, int Cin, int Cout) {
    const int total = Cin * 4 * Cout;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int o = idx % (4 * Cout), i = idx / (4 * Cout);
        out[idx] = in[i * Cout + o % Cout];
    }
}

// one activation value -> its channels-last slots: fp16(v) at c (and at 2 Cp + c), fp16(v - fp16(v)) at Cp + c when ns == 3
__device__ __forceinline__ void put_split(__half* px, int c, int Cp, int ns, float v) {
    const __half h = __float2half_rn(v);
    px[c] = h;
    if (ns == 3) {
        px[Cp + c] = __float2half_rn(v - __half2float(h));
        px[2 * Cp + c] = h;
    }
}

// const [C][r][r] f32 * style[b][c] -> channels-last fp16 [B][r][r][ns * Cp]
__global__ void const_input_kernel(const float* __restrict__ cst, const float* __restrict__ style, __half* __restrict__ out,
                                   int B, int C, int r, int Cp, int ns) {
    const long long total = static_cast<long long>(B) * r * r * Cp;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % Cp);
        long long q = idx / Cp;
        const int px = static_cast<int>(q % (r * r));
        const int b = static_cast<int>(q / (r * r));
        put_split(out + q * (static_cast<long long>(ns) * Cp), c, Cp, ns, c < C ? cst[c * r * r + px] * style[b * C + c] : 0.0f);
    }
}

struct ActArgs {
    const __half* y;        // planar conv output [B][C][Hy][Wpy]; Hy = R (+1 with FIR).  phase: the four parity planes of the
                            // polyphase transposed conv, [B][4][C][R/2 + 1][Wpy]: y(yy, xx) = plane (yy & 1) * 2 + (xx & 1) at (yy >> 1, xx >> 1)
    long long y_lo;         // > 0: y + y_lo holds the low halves of a split output (value = hi + lo)
    int phase, ns;          // ns: channel blocks of x_next (3 = [xh | xl | xh], 1 = plain fp16)
    const __half* pre;      // if set: the finished activation, channels-last [B][R][R][Cp] (a warped feature map); y, noise, bias unused
    const float* noise;     // [R][R] (noise_bstride 0) or per-frame [B][R][R] (noise_bstride R*R) or nullptr
    const float* bias;      // [C]
    const float* style_next;  // [B][C] or nullptr (last block)
    __half* x_next;         // channels-last [B][R][R][Cp] or nullptr
    const float* rgb_w;     // [3][C] ToRGB weight or nullptr (no ToRGB after conv0)
    const float* rgb_style; // [B][C] (already * 1/sqrt(C))
    const float* rgb_bias;  // [3]
    const float* img_prev;  // [B][3][R][R] upsampled skip image or nullptr
    float* img;             // [B][3][R][R]
    int B, C, R, Hy, Wpy, Cp, fir, nimg;
    long long noise_bstride;
    float clamp;
};
constexpr int kActP = 32;  // pixels per CTA

// one CTA = (b, row h, 32 pixels) x all channels
__global__ void __launch_bounds__(256) sg2_act_kernel(const ActArgs a) {
    extern __shared__ float xs[];  // [C][kActP + 1]
    const int w0 = blockIdx.x * kActP, h = blockIdx.y, b = blockIdx.z;
    // separable [1,3,3,1]/8 * 2 per axis = the 4x4 FIR of setup_filter([1,3,3,1]) with gain up^2 = 4 (ops.py:225,236)
    const float f4[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    for (int idx = threadIdx.x; idx < a.C * kActP; idx += blockDim.x) {
        const int c = idx / kActP, px = idx - c * kActP;
        const int w = w0 + px;
        float v = 0.0f;
        if (a.pre) {
            if (w < a.R) v = __half2float(a.pre[((static_cast<long long>(b) * a.R + h) * a.R + w) * a.Cp + c]);
        } else if (w < a.R) {
            const int hp = a.phase ? a.R / 2 + 1 : a.Hy;               // rows of a stored plane
            const long long plane = static_cast<long long>(hp) * a.Wpy;
            const __half* yp = a.y + (static_cast<long long>(b) * (a.phase ? 4 : 1) * a.C + c) * plane;
            auto at = [&](int yy, int xx) -> float {
                const __half* q = a.phase ? yp + static_cast<long long>(((yy & 1) * 2 + (xx & 1)) * a.C) * plane + static_cast<long long>(yy >> 1) * a.Wpy + (xx >> 1)
                                          : yp + static_cast<long long>(yy) * a.Wpy + xx;
                float v = __half2float(*q);
                if (a.y_lo > 0) v += __half2float(q[a.y_lo]);
                return v;
            };
            if (a.fir) {
                // upfirdn2d(pad 1): out[h][w] = sum_{ky,kx} f[ky] f[kx] y[h - 1 + ky][w - 1 + kx], zero outside
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
                    const int yy = h - 1 + ky;
                    if (yy < 0 || yy >= a.Hy) continue;
                    float row = 0.0f;
#pragma unroll
                    for (int kx = 0; kx < 4; ++kx) {
                        const int xx = w - 1 + kx;
                        if (xx >= 0 && xx < a.Hy) row = fmaf(f4[kx], at(yy, xx), row);
                    }
                    v = fmaf(f4[ky], row, v);
                }
            } else {
                v = at(h, w);
            }
            if (a.noise) v += a.noise[b * a.noise_bstride + h * a.R + w];
            v += a.bias[c];
            v = (v < 0.0f ? v * 0.2f : v) * 1.41421356237309515f;
            v = fminf(fmaxf(v, -a.clamp), a.clamp);
        }
        xs[c * (kActP + 1) + px] = v;
    }
    __syncthreads();
    const int npx = min(kActP, a.R - w0);
    if (a.x_next) {
        const int cpx = a.ns * a.Cp;   // halfs per pixel of x_next
        __half* o = a.x_next + ((static_cast<long long>(b) * a.R + h) * a.R + w0) * cpx;
        for (int idx = threadIdx.x; idx < npx * (a.Cp / 2); idx += blockDim.x) {
            const int px = idx / (a.Cp / 2), c = (idx - px * (a.Cp / 2)) * 2;
            const float s0 = c < a.C ? (a.style_next ? a.style_next[b * a.C + c] : 1.0f) : 0.0f;
            const float s1 = c + 1 < a.C ? (a.style_next ? a.style_next[b * a.C + c + 1] : 1.0f) : 0.0f;
            const float v0 = c < a.C ? xs[c * (kActP + 1) + px] * s0 : 0.0f;
            const float v1 = c + 1 < a.C ? xs[(c + 1) * (kActP + 1) + px] * s1 : 0.0f;
            __half* op = o + static_cast<long long>(px) * cpx;
            const __half2 hi = __floats2half2_rn(v0, v1);
            *reinterpret_cast<__half2*>(op + c) = hi;
            if (a.ns == 3) {
                const float2 hf = __half22float2(hi);
                *reinterpret_cast<__half2*>(op + a.Cp + c) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                *reinterpret_cast<__half2*>(op + 2 * a.Cp + c) = hi;
            }
        }
    }
    if (a.rgb_w) {
        // ToRGB: 1x1 modulated conv without demodulation + bias + clamp, added to the upsampled skip image
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int item = warp; item < npx * a.nimg; item += nw) {
            const int px = item / a.nimg, o = item - px * a.nimg;
            float acc = 0.0f;
            for (int c = lane; c < a.C; c += 32)
                acc = fmaf(xs[c * (kActP + 1) + px], a.rgb_w[o * a.C + c] * a.rgb_style[b * a.C + c], acc);
            for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
            if (lane == 0) {
                float v = fminf(fmaxf(acc + a.rgb_bias[o], -a.clamp), a.clamp);
                const long long oi = ((static_cast<long long>(b) * a.nimg + o) * a.R + h) * a.R + w0 + px;
                if (a.img_prev) v += a.img_prev[oi];
                a.img[oi] = v;
            }
        }
    }
}

// Tiled form of the fused kernel for activations that come straight from a conv (no warped feature map): one CTA = (frame,
// TH image rows, 32 pixels) x all channels.  sg2_act_kernel above reads its 16 FIR taps per output value from global memory
// (32 loads per value with the hi / lo planes: the r2 bench line showed it at 7 % of the HBM roof, 75 % of a StyleGAN2 step).
// Here a warp stages the (TH + 3) x 35 input patch of one channel in shared memory (hi + lo summed to fp32, the four parity
// planes read as two interleaved coalesced streams), runs the separable [1,3,3,1] FIR from there (7 shared loads per output at
// TH = 4), and the finished TH x 32 x C block leaves pixel-major like before.  ToRGB: one thread per (pixel, colour) walks
// the channels of the staged block against a per-frame weight * style table.
template <int TH>
__global__ void __launch_bounds__(256) sg2_act_tiled_kernel(const ActArgs a) {
    extern __shared__ float sm[];
    constexpr int PXT = TH * kActP;             // pixels of the tile
    constexpr int XP = PXT + 1;                 // pitch of a channel row in xs
    constexpr int PR = TH + 3, PC = 36;         // per-warp input patch (rows x padded columns)
    float* xs = sm;                             // [C][XP]
    float* patch = xs + a.C * XP;               // [8 warps][PR][PC]
    float* nz = patch + 8 * PR * PC;            // [PXT] noise of the tile
    float* wrgb = nz + PXT;                     // [nimg][C] ToRGB weight * style of this frame
    const int w0 = blockIdx.x * kActP, h0 = blockIdx.y * TH, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float f4[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    for (int i = threadIdx.x; i < PXT; i += blockDim.x) {
        const int hh = h0 + i / kActP, ww = w0 + i % kActP;
        nz[i] = (a.noise && hh < a.R && ww < a.R) ? a.noise[b * a.noise_bstride + hh * a.R + ww] : 0.0f;
    }
    if (a.rgb_w)
        for (int i = threadIdx.x; i < a.nimg * a.C; i += blockDim.x) wrgb[i] = a.rgb_w[i] * a.rgb_style[b * a.C + i % a.C];
    __syncthreads();
    const int hp = a.phase ? a.R / 2 + 1 : a.Hy;
    const long long plane = static_cast<long long>(hp) * a.Wpy;
    float* pw = patch + warp * PR * PC;
    for (int c = warp; c < a.C; c += 8) {
        const __half* yp = a.y + (static_cast<long long>(b) * (a.phase ? 4 : 1) * a.C + c) * plane;
        auto at = [&](int yy, int xx) -> float {
            const __half* q = a.phase ? yp + static_cast<long long>(((yy & 1) * 2 + (xx & 1)) * a.C) * plane + static_cast<long long>(yy >> 1) * a.Wpy + (xx >> 1)
                                      : yp + static_cast<long long>(yy) * a.Wpy + xx;
            float v = __half2float(*q);
            if (a.y_lo > 0) v += __half2float(q[a.y_lo]);
            return v;
        };
        const float bias = a.bias[c];
        float outv[TH];
        if (a.fir) {
            // patch rows h0 - 1 .. h0 + TH + 1, columns w0 - 1 .. w0 + 33 of the (R + 1)^2 conv output, zero outside
            for (int r = 0; r < PR; ++r) {
                const int yy = h0 - 1 + r;
                for (int cc = lane; cc < 35; cc += 32) {
                    const int xx = w0 - 1 + cc;
                    pw[r * PC + cc] = (yy >= 0 && yy < a.Hy && xx >= 0 && xx < a.Hy) ? at(yy, xx) : 0.0f;
                }
            }
            __syncwarp();
            float hrow[PR];
#pragma unroll
            for (int r = 0; r < PR; ++r) {
                float t = 0.0f;
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) t = fmaf(f4[kx], pw[r * PC + lane + kx], t);
                hrow[r] = t;
            }
#pragma unroll
            for (int t = 0; t < TH; ++t) {
                float v = 0.0f;
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) v = fmaf(f4[ky], hrow[t + ky], v);
                outv[t] = v;
            }
            __syncwarp();   // the patch is free for the next channel
        } else {
#pragma unroll
            for (int t = 0; t < TH; ++t) outv[t] = (h0 + t < a.R && w0 + lane < a.R) ? at(h0 + t, w0 + lane) : 0.0f;
        }
#pragma unroll
        for (int t = 0; t < TH; ++t) {
            float v = outv[t] + nz[t * kActP + lane] + bias;
            v = (v < 0.0f ? v * 0.2f : v) * 1.41421356237309515f;
            v = fminf(fmaxf(v, -a.clamp), a.clamp);
            xs[c * XP + t * kActP + lane] = (h0 + t < a.R && w0 + lane < a.R) ? v : 0.0f;
        }
    }
    __syncthreads();
    const int npx = min(kActP, a.R - w0);
    if (a.x_next) {
        const int cpx = a.ns * a.Cp, half_cp = a.Cp / 2;
        for (int idx = threadIdx.x; idx < PXT * half_cp; idx += blockDim.x) {
            const int pix = idx / half_cp, c = (idx - pix * half_cp) * 2;
            const int t = pix / kActP, px = pix - t * kActP;
            if (px >= npx || h0 + t >= a.R) continue;
            const float s0 = c < a.C ? (a.style_next ? a.style_next[b * a.C + c] : 1.0f) : 0.0f;
            const float s1 = c + 1 < a.C ? (a.style_next ? a.style_next[b * a.C + c + 1] : 1.0f) : 0.0f;
            const float v0 = c < a.C ? xs[c * XP + pix] * s0 : 0.0f;
            const float v1 = c + 1 < a.C ? xs[(c + 1) * XP + pix] * s1 : 0.0f;
            __half* op = a.x_next + ((static_cast<long long>(b) * a.R + h0 + t) * a.R + w0 + px) * cpx;
            const __half2 hi = __floats2half2_rn(v0, v1);
            *reinterpret_cast<__half2*>(op + c) = hi;
            if (a.ns == 3) {
                const float2 hf = __half22float2(hi);
                *reinterpret_cast<__half2*>(op + a.Cp + c) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                *reinterpret_cast<__half2*>(op + 2 * a.Cp + c) = hi;
            }
        }
    }
    if (a.rgb_w) {
        for (int item = threadIdx.x; item < PXT * a.nimg; item += blockDim.x) {
            const int o = item / PXT, pix = item - o * PXT;
            const int t = pix / kActP, px = pix - t * kActP;
            if (px >= npx || h0 + t >= a.R) continue;
            float acc = 0.0f;
            for (int c = 0; c < a.C; ++c) acc = fmaf(xs[c * XP + pix], wrgb[o * a.C + c], acc);
            float v = fminf(fmaxf(acc + a.rgb_bias[o], -a.clamp), a.clamp);
            const long long oi = ((static_cast<long long>(b) * a.nimg + o) * a.R + h0 + t) * a.R + w0 + px;
            if (a.img_prev) v += a.img_prev[oi];
            a.img[oi] = v;
        }
    }
}

// Feature-map warp of the network-bending hooks (maua/GAN/wrappers/stylegan2.py:153-194: kornia translate / rotate / scale
// with padding_mode="reflection"; kornia is an un-pinned, absent dependency: its published warp_affine = affine_grid +
// grid_sample(bilinear, align_corners=True) is restated).  out[b, y, x, :] = bilinear sample of src[b] at
// (sx, sy) = M_b (x, y, 1) in pixel coordinates, coordinates reflected about the centres of the border pixels.
// One thread = 8 channels of one output pixel; channels-last fp16 [B][R][R][Cp].
__device__ __forceinline__ float reflect_coord(float x, int size) {
    // grid_sample reflect_coordinates(x, 0, 2 * (size - 1)) followed by clip_coordinates (align_corners=True)
    if (size <= 1) return 0.0f;
    const float span = static_cast<float>(size - 1);
    x = fabsf(x);
    const float extra = fmodf(x, span);
    const int flips = static_cast<int>(floorf(x / span));
    x = (flips & 1) ? span - extra : extra;
    return fminf(fmaxf(x, 0.0f), span);
}

__global__ void __launch_bounds__(256) sg2_warp_kernel(const __half* __restrict__ src, __half* __restrict__ dst,
                                                       const float* __restrict__ mats /*[B][2][3]*/, int B, int R, int Cp) {
    const int groups = Cp / 8;
    const long long total = static_cast<long long>(B) * R * R * groups;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(idx % groups);
        long long q = idx / groups;
        const int x = static_cast<int>(q % R); q /= R;
        const int y = static_cast<int>(q % R);
        const int b = static_cast<int>(q / R);
        const float* m = mats + b * 6;
        const float sx = reflect_coord(m[0] * x + m[1] * y + m[2], R);
        const float sy = reflect_coord(m[3] * x + m[4] * y + m[5], R);
        const float fx = floorf(sx), fy = floorf(sy);
        const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
        const float tx = sx - fx, ty = sy - fy;
        const float wgt[4] = {(1.0f - tx) * (1.0f - ty), tx * (1.0f - ty), (1.0f - tx) * ty, tx * ty};
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.0f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
            if (xx < 0 || xx >= R || yy < 0 || yy >= R) continue;  // grid_sample: out-of-range taps count as zero
            const uint4 v = *reinterpret_cast<const uint4*>(src + ((static_cast<long long>(b) * R + yy) * R + xx) * Cp + g * 8);
            const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float2 f = __half22float2(hv[c]);
                acc[2 * c] = fmaf(wgt[t], f.x, acc[2 * c]);
                acc[2 * c + 1] = fmaf(wgt[t], f.y, acc[2 * c + 1]);
            }
        }
        uint4 o;
        __half2* ov = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int c = 0; c < 4; ++c) ov[c] = __floats2half2_rn(acc[2 * c], acc[2 * c + 1]);
        *reinterpret_cast<uint4*>(dst + ((static_cast<long long>(b) * R + y) * R + x) * Cp + g * 8) = o;
    }
}

// upsample2d(img, [1,3,3,1]): zero-insert x2, pad (2,1), 4x4 FIR with gain 4 (ops.py:117-133) on [N][r][r] planes
__global__ void upsample_rgb_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int r) {
    const int R = 2 * r;
    const float f4[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    const long long total = static_cast<long long>(N) * R * R;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int X = static_cast<int>(idx % R);
        long long q = idx / R;
        const int Y = static_cast<int>(q % R);
        const int n = static_cast<int>(q / R);
        const float* xp = x + static_cast<long long>(n) * r * r;
        float acc = 0.0f;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const int my = Y + ky - 2;  // index into the zero-inserted signal
            if (my < 0 || (my & 1) || (my >> 1) >= r) continue;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const int mx = X + kx - 2;
                if (mx < 0 || (mx & 1) || (mx >> 1) >= r) continue;
                acc = fmaf(f4[ky] * f4[kx], xp[(my >> 1) * r + (mx >> 1)], acc);
            }
        }
        y[idx] = acc;
    }
}

// (x + 1) / 2 clamped to [0, 1]: MB_OUT_F32_NCHW_01
__global__ void img_to_unit_kernel(const float* __restrict__ img, float* __restrict__ out, long long n, int clamp01) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float v = (img[i] + 1.0f) * 0.5f;
        out[i] = clamp01 ? fminf(fmaxf(v, 0.0f), 1.0f) : v;
    }
}

__global__ void img_to_u8_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, int B, int C, int R) {
    const long long total = static_cast<long long>(B) * R * R * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        long long q = idx / C;
        const int px = static_cast<int>(q % (static_cast<long long>(R) * R));
        const int b = static_cast<int>(q / (static_cast<long long>(R) * R));
        float v = (img[(static_cast<long long>(b) * C + c) * R * R + px] + 1.0f) * 0.5f;
        v = fminf(fmaxf(v, 0.0f), 1.0f);
        out[idx] = static_cast<uint8_t>(rintf(v * 255.0f));
    }
}

int grid1d(long long total) {
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace

struct Sg2Param {
    float* dev = nullptr;
    std::vector<int64_t> shape;
    size_t numel = 0;      // elements currently held
    size_t capacity = 0;   // elements allocated
    size_t plane = 0;      // noise maps only: r*r; numel may then be any multiple of it (per-frame noise [B,1,r,r])
    bool set = false;
};
struct Sg2Layer {   // one modulated conv (conv0 / conv1) or ToRGB
    int cin = 0, cout = 0, res = 0, up = 1, ksz = 3;
    Sg2Param affine_w, affine_b, weight, bias, noise;
    __half* wpk = nullptr;   // packed effective weights (sg2_prep_weights_kernel -> pack_weights_launch)
    size_t wpk_elems = 0;
    float* wsqT = nullptr;   // [cin][cout]: sum over the 3x3 taps of w^2 (demodulation)
    float* wsq4 = nullptr;   // conv0: wsqT tiled to [cin][4 * cout] for the four parity planes
};
struct Sg2Block {
    int res = 0, cin = 0, cout = 0;
    Sg2Param cst;
    Sg2Layer conv0, conv1, torgb;
    bool has_conv0 = false;
};
struct Sg2Net {
    int w_dim = 512, res = 0, img_channels = 3, num_ws = 0;
    std::vector<Sg2Block> blocks;
    std::map<std::string, Sg2Param*> by_name;
    bool finalized = false;
    int conv_impl = 0;
    int precise = 1;         // fp16 hi + lo operands and conv outputs (see the header); 0 = plain fp16
    int act_tiled = 1;       // 1: sg2_act_tiled_kernel between the convs; 0: the untiled sg2_act_kernel (A/B, MB_SG2_ACT_TILED=0)
    int packed_precise = -1; // the mode the packed weights were built for
    int last_launches = 0;
    // feature-map warps applied by the next forwards (sg2_set_warps): layer = index into the wrapper's layer_names
    std::vector<int> warp_layer;
    const float* warp_mats = nullptr;  // device [n_warps][warp_batch][2][3], caller-owned
    int warp_batch = 0;
};

static int sg2_alloc(Sg2Param& p, std::initializer_list<int64_t> shape) {
    p.shape.assign(shape.begin(), shape.end());
    p.numel = 1;
    for (int64_t s : p.shape) p.numel *= static_cast<size_t>(s);
    MB_CUDA(cudaMalloc(&p.dev, sizeof(float) * p.numel));
    p.capacity = p.numel;
    return MB_OK;
}

void sg2_destroy(Sg2Net* n) {
    if (!n) return;
    for (auto& kv : n->by_name)
        if (kv.second->dev) cudaFree(kv.second->dev);
    for (auto& b : n->blocks)
        for (Sg2Layer* L : {&b.conv0, &b.conv1, &b.torgb}) {
            if (L->wpk) cudaFree(L->wpk);
            if (L->wsqT) cudaFree(L->wsqT);
            if (L->wsq4) cudaFree(L->wsq4);
        }
    delete n;
}

int sg2_create(int w_dim, int img_resolution, int img_channels, int channel_base, int channel_max, Sg2Net** out) {
    MB_REQUIRE(img_resolution >= 8 && (img_resolution & (img_resolution - 1)) == 0 && img_resolution <= 2048,
               "mb_sg2_create: img_resolution must be a power of two in [8, 2048]");
    MB_REQUIRE(img_channels >= 1 && img_channels <= 4, "mb_sg2_create: img_channels must be 1..4");
    Sg2Net* n = new Sg2Net();
    if (const char* ev = getenv("MB_SG2_ACT_TILED")) n->act_tiled = atoi(ev);
    n->w_dim = w_dim; n->res = img_resolution; n->img_channels = img_channels;
    auto ch = [&](int r) { int c = channel_base / r; return c < channel_max ? c : channel_max; };
    int bi = 0;
    for (int r = 4; r <= img_resolution; r *= 2, ++bi) {
        n->blocks.emplace_back();
    }
    bi = 0;
    for (int r = 4; r <= img_resolution; r *= 2, ++bi) {
        Sg2Block& b = n->blocks[bi];
        b.res = r; b.cout = ch(r); b.cin = r > 4 ? ch(r / 2) : 0; b.has_conv0 = r > 4;
        const std::string pre = "bs." + std::to_string(bi) + ".";
        auto add_layer = [&](Sg2Layer& L, const std::string& name, int cin, int cout, int ksz, int up, bool noise) -> int {
            L.cin = cin; L.cout = cout; L.res = r; L.up = up; L.ksz = ksz;
            int rc;
            if ((rc = sg2_alloc(L.affine_w, {cin, w_dim})) != MB_OK) return rc;
            if ((rc = sg2_alloc(L.affine_b, {cin})) != MB_OK) return rc;
            if ((rc = sg2_alloc(L.weight, {cout, cin, ksz, ksz})) != MB_OK) return rc;
            if ((rc = sg2_alloc(L.bias, {cout})) != MB_OK) return rc;
            n->by_name[pre + name + ".affine.weight"] = &L.affine_w;
            n->by_name[pre + name + ".affine.bias"] = &L.affine_b;
            n->by_name[pre + name + ".weight"] = &L.weight;
            n->by_name[pre + name + ".bias"] = &L.bias;
            if (noise) {
                if ((rc = sg2_alloc(L.noise, {r, r})) != MB_OK) return rc;
                L.noise.plane = static_cast<size_t>(r) * r;
                n->by_name[pre + name + ".noise_const"] = &L.noise;
            }
            if (ksz == 3) {
                if (cudaMalloc(&L.wsqT, sizeof(float) * cin * cout) != cudaSuccess ||
                    (up == 2 && cudaMalloc(&L.wsq4, sizeof(float) * cin * 4 * cout) != cudaSuccess)) {
                    set_error("mb_sg2_create: cudaMalloc failed");
                    return MB_ECUDA;
                }
            }
            return MB_OK;
        };
        int rc = MB_OK;
        if (!b.has_conv0) {
            rc = sg2_alloc(b.cst, {b.cout, r, r});
            n->by_name[pre + "const"] = &b.cst;
        } else {
            rc = add_layer(b.conv0, "conv0", b.cin, b.cout, 3, 2, true);
        }
        if (rc == MB_OK) rc = add_layer(b.conv1, "conv1", b.cout, b.cout, 3, 1, true);
        if (rc == MB_OK) rc = add_layer(b.torgb, "torgb", b.cout, img_channels, 1, 1, false);
        if (rc != MB_OK) {
            sg2_destroy(n);
            return rc;
        }
        n->num_ws += b.has_conv0 ? 2 : 1;
    }
    n->num_ws += 1;  // the last block's ToRGB
    *out = n;
    return MB_OK;
}

int sg2_set_param(Sg2Net* n, const char* name, const float* data, const int64_t* shape, int ndim, cudaStream_t stream) {
    auto it = n->by_name.find(name);
    if (it == n->by_name.end()) {
        // buffers of the reference state dict that carry no information for this implementation
        const std::string s = name;
        if (s.size() > 15 && s.compare(s.size() - 15, 15, "resample_filter") == 0) return MB_OK;
        set_error("mb_net_set_param: unknown StyleGAN2 parameter '%s'", name);
        return MB_EINVAL;
    }
    Sg2Param& p = *it->second;
    size_t numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= static_cast<size_t>(shape[i]);
    if (p.plane) {
        // the reference swaps noise_const for a per-frame [B,1,r,r] tensor on every call (wrappers/stylegan2.py:81-96)
        MB_REQUIRE(numel >= p.plane && numel % p.plane == 0, "mb_net_set_param: '%s' has %zu elements, expected a multiple of %zu",
                   name, numel, p.plane);
        if (numel > p.capacity) {
            MB_CUDA(cudaStreamSynchronize(stream));
            cudaFree(p.dev);
            p.dev = nullptr; p.capacity = 0;
            MB_CUDA(cudaMalloc(&p.dev, sizeof(float) * numel));
            p.capacity = numel;
        }
        p.numel = numel;
        MB_CUDA(cudaMemcpyAsync(p.dev, data, sizeof(float) * numel, cudaMemcpyDeviceToDevice, stream));
        p.set = true;
        return MB_OK;  // noise maps feed no derived operand: the network stays finalized
    }
    MB_REQUIRE(numel == p.numel, "mb_net_set_param: '%s' has %zu elements, expected %zu", name, numel, p.numel);
    MB_CUDA(cudaMemcpyAsync(p.dev, data, sizeof(float) * p.numel, cudaMemcpyDeviceToDevice, stream));
    p.set = true;
    n->finalized = false;
    return MB_OK;
}

static int sg2_pack_layer(Sg2Net* n, Sg2Layer& L, int phase, cudaStream_t stream) {
    const int ns = n->precise ? 3 : 1, Cp = cpad16(L.cin);
    const int kk = phase ? 2 : 3, Co = phase ? 4 * L.cout : L.cout, Ci = ns * Cp;
    // demodulation sums from the original 3x3 weights (the packed copy this call also writes is scratch)
    __half* scratch = nullptr;
    MB_CUDA(cudaMalloc(&scratch, packed_weight_elems(L.cout, L.cin, 3) * sizeof(__half)));
    int rc = pack_weights_launch(L.weight.dev, scratch, L.wsqT, L.cout, L.cin, 3, 0, stream, 0);
    if (rc == MB_OK && phase) {
        sg2_tile4_kernel<<<grid1d(static_cast<long long>(L.cin) * 4 * L.cout), 256, 0, stream>>>(L.wsqT, L.wsq4, L.cin, L.cout);
        if (cudaGetLastError() != cudaSuccess) rc = MB_ECUDA;
    }
    float* eff = nullptr;
    const size_t eff_elems = static_cast<size_t>(Co) * Ci * kk * kk;
    if (rc == MB_OK && cudaMalloc(&eff, eff_elems * sizeof(float)) != cudaSuccess) rc = MB_ECUDA;
    const size_t need = packed_weight_elems(Co, Ci, kk);
    if (rc == MB_OK && need != L.wpk_elems) {
        if (L.wpk) cudaFree(L.wpk);
        L.wpk = nullptr; L.wpk_elems = 0;
        if (cudaMalloc(&L.wpk, need * sizeof(__half)) != cudaSuccess) rc = MB_ECUDA;
        else L.wpk_elems = need;
    }
    if (rc == MB_OK) {
        sg2_prep_weights_kernel<<<grid1d(static_cast<long long>(eff_elems)), 256, 0, stream>>>(L.weight.dev, eff, L.cout, L.cin, Cp, phase, ns);
        if (cudaGetLastError() != cudaSuccess) rc = MB_ECUDA;
    }
    if (rc == MB_OK) rc = pack_weights_launch(eff, L.wpk, nullptr, Co, Ci, kk, 0, stream, 0);
    cudaStreamSynchronize(stream);
    cudaFree(scratch);
    if (eff) cudaFree(eff);
    if (rc == MB_ECUDA) set_error("mb_net_finalize: packing the StyleGAN2 weights failed (%s)", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

int sg2_finalize(Sg2Net* n, cudaStream_t stream) {
    for (auto& kv : n->by_name)
        if (!kv.second->set) {
            set_error("mb_net_finalize: StyleGAN2 parameter '%s' was never set", kv.first.c_str());
            return MB_ESTATE;
        }
    for (auto& b : n->blocks) {
        int rc;
        if (b.has_conv0 && (rc = sg2_pack_layer(n, b.conv0, 1, stream)) != MB_OK) return rc;
        if ((rc = sg2_pack_layer(n, b.conv1, 0, stream)) != MB_OK) return rc;
    }
    MB_CUDA(cudaStreamSynchronize(stream));
    n->packed_precise = n->precise;
    n->finalized = true;
    return MB_OK;
}

namespace {
struct Sg2Ws {
    size_t styles, d, x, y, img0, img1, t0, t1, total;
    std::vector<size_t> style_l, d_l;  // per layer (conv0, conv1, torgb per block) float offsets
};
Sg2Ws sg2_ws(const Sg2Net* n, int B) {
    Sg2Ws w;
    size_t ns = 0, nd = 0, mx = 0, my = 0, mt = 0;
    const size_t nsx = n->precise ? 3 : 1, nsy = n->precise ? 2 : 1;   // channel blocks of X, planes (hi, lo) of Y
    for (const auto& b : n->blocks) {
        for (const Sg2Layer* L : {&b.conv0, &b.conv1, &b.torgb}) {
            w.style_l.push_back(ns);
            w.d_l.push_back(nd);
            ns += static_cast<size_t>(B) * L->cin;
            nd += static_cast<size_t>(B) * L->cout * (L == &b.conv0 ? 4 : 1);   // conv0: one coefficient per parity plane
        }
        const size_t r = b.res;
        mx = std::max(mx, static_cast<size_t>(B) * r * r * cpad16(b.cout) * nsx);
        mt = std::max(mt, static_cast<size_t>(B) * r * r * cpad16(b.cout));
        if (b.has_conv0) {
            mx = std::max(mx, static_cast<size_t>(B) * (r / 2) * (r / 2) * cpad16(b.cin) * nsx);
            my = std::max(my, static_cast<size_t>(B) * 4 * b.cout * (r / 2 + 1) * pitch8(static_cast<int>(r / 2 + 1)) * nsy);
        }
        my = std::max(my, static_cast<size_t>(B) * b.cout * r * pitch8(static_cast<int>(r)) * nsy);
    }
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = round_up_sz(off + bytes, 1024); return o; };
    w.styles = take(ns * 4); w.d = take(nd * 4);
    w.x = take(mx * 2); w.y = take(my * 2);
    const size_t img = static_cast<size_t>(B) * n->img_channels * n->res * n->res * 4;
    w.img0 = take(img); w.img1 = take(img);
    w.t0 = w.t1 = off;
    if (!n->warp_layer.empty()) { w.t0 = take(mt * 2); w.t1 = take(mt * 2); }  // ping-pong buffers of the warped feature map (plain fp16)
    w.total = off;
    return w;
}
}  // namespace

size_t sg2_workspace_bytes(const Sg2Net* n, int B) { return sg2_ws(n, B).total; }
int sg2_num_ws(const Sg2Net* n) { return n->num_ws; }
int sg2_resolution(const Sg2Net* n) { return n->res; }
int sg2_last_launches(const Sg2Net* n) { return n->last_launches; }
void sg2_set_conv_impl(Sg2Net* n, int impl) { n->conv_impl = impl; }
void sg2_set_precise(Sg2Net* n, int on) { n->precise = on ? 1 : 0; }

int sg2_set_warps(Sg2Net* n, int n_warps, const int32_t* layers, const float* inv_mats, int batch) {
    MB_REQUIRE(n_warps >= 0 && n_warps <= 16, "mb_sg2_set_warps: between 0 and 16 warps, got %d", n_warps);
    if (n_warps == 0) {
        n->warp_layer.clear(); n->warp_mats = nullptr; n->warp_batch = 0;
        return MB_OK;
    }
    MB_REQUIRE(layers && inv_mats && batch > 0, "mb_sg2_set_warps: null argument");
    const int n_names = 2 * static_cast<int>(n->blocks.size());
    for (int i = 0; i < n_warps; ++i)
        MB_REQUIRE(layers[i] >= 0 && layers[i] < n_names, "mb_sg2_set_warps: layer %d out of range [0, %d)", layers[i], n_names);
    n->warp_layer.assign(layers, layers + n_warps);
    n->warp_mats = inv_mats;
    n->warp_batch = batch;
    return MB_OK;
}

int sg2_forward(Sg2Net* n, const float* ws, int B, void* out, int out_fmt, void* workspace, size_t workspace_bytes,
                int num_sms, cudaStream_t stream, const std::function<void(int, int)>& mark_fn) {
    auto mark = [&](int kind, int layer) { if (mark_fn) mark_fn(kind, layer); };
    if (!n->finalized) {
        set_error("mb_net_forward: call mb_net_finalize after setting parameters");
        return MB_ESTATE;
    }
    if (n->packed_precise != n->precise) {   // the precision mode was switched after the weights were packed
        int rr = sg2_finalize(n, stream);
        if (rr != MB_OK) return rr;
    }
    const int nsx = n->precise ? 3 : 1;
    const Sg2Ws wl = sg2_ws(n, B);
    if (workspace_bytes < wl.total) {
        set_error("mb_net_forward: workspace too small (%zu < %zu bytes)", workspace_bytes, wl.total);
        return MB_ENOMEM;
    }
    uint8_t* base = static_cast<uint8_t*>(workspace);
    float* styles = reinterpret_cast<float*>(base + wl.styles);
    float* dco = reinterpret_cast<float*>(base + wl.d);
    __half* X = reinterpret_cast<__half*>(base + wl.x);
    __half* Y = reinterpret_cast<__half*>(base + wl.y);
    float* img[2] = {reinterpret_cast<float*>(base + wl.img0), reinterpret_cast<float*>(base + wl.img1)};
    __half* T[2] = {reinterpret_cast<__half*>(base + wl.t0), reinterpret_cast<__half*>(base + wl.t1)};
    int launches = 0, rc;
    MB_REQUIRE(n->warp_layer.empty() || n->warp_batch == B, "mb_net_forward: feature-map warps were set for a batch of %d, forward got %d",
               n->warp_batch, B);

    // styles (+ demodulation coefficients) of every layer; ws index: block i uses ws[w_idx + {0,1,2}]
    {
        int w_idx = 0, li = 0;
        StylesArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.B = B; sa.num_ws = n->num_ws; sa.w_dim = n->w_dim; sa.ws = ws;
        auto flush = [&]() -> int {
            if (sa.num_layers == 0) return MB_OK;
            int r = styles_launch(sa, stream);
            launches += 1;
            sa.num_layers = 0;
            return r;
        };
        for (const auto& b : n->blocks) {
            int j = 0;
            for (const Sg2Layer* L : {&b.conv0, &b.conv1, &b.torgb}) {
                const bool present = !(L == &b.conv0 && !b.has_conv0);
                if (present) {
                    StyleLayerDesc& d = sa.L[sa.num_layers++];
                    d.affine_w = L->affine_w.dev; d.affine_b = L->affine_b.dev;
                    const bool parity4 = L == &b.conv0;   // polyphase transposed conv: 4 * cout output planes, same coefficient each
                    d.wsqT = parity4 ? L->wsq4 : L->wsqT; d.magnitude_ema = nullptr;
                    d.s_out = styles + wl.style_l[li];
                    d.d_out = L->ksz == 3 ? dco + wl.d_l[li] : nullptr;
                    d.Cin = L->cin; d.Cout = parity4 ? 4 * L->cout : L->cout; d.ws_index = w_idx + j;
                    d.demodulate = L->ksz == 3; d.normalize_style = 0;
                    d.style_scale = L->ksz == 1 ? 1.0f / sqrtf(static_cast<float>(L->cin)) : 1.0f;
                    ++j;
                    if (sa.num_layers == kMaxLayers && (rc = flush()) != MB_OK) return rc;
                }
                ++li;
            }
            w_idx += b.has_conv0 ? 2 : 1;
        }
        if ((rc = flush()) != MB_OK) return rc;
        mark(0, -1);
    }

    int cur = 0;  // which img buffer holds the previous block's image
    bool have_img = false;
    for (size_t bi = 0; bi < n->blocks.size(); ++bi) {
        const Sg2Block& b = n->blocks[bi];
        const int r = b.res;
        const bool last = bi + 1 == n->blocks.size();
        const float* s_conv0 = styles + wl.style_l[bi * 3 + 0];
        const float* s_conv1 = styles + wl.style_l[bi * 3 + 1];
        const float* s_rgb = styles + wl.style_l[bi * 3 + 2];
        // phase = 1: the polyphase transposed conv (2x2 'full' kernel, 4 * cout parity planes of (hin + 1)^2); else 3x3 'same'
        auto conv = [&](const Sg2Layer& L, const __half* xin, int hin, int phase, const float* d) -> int {
            ConvTcArgs ca;
            ca.x = xin; ca.wpk = L.wpk; ca.d = d; ca.bias = nullptr; ca.y = Y;
            ca.B = B; ca.Cin = nsx * cpad16(L.cin); ca.Cout = phase ? 4 * L.cout : L.cout; ca.Hin = hin; ca.Win = hin; ca.Cp_in = ca.Cin;
            ca.ksz = phase ? 2 : 3; ca.pad = 1;
            const int hout = hin + 2 * ca.pad - (ca.ksz - 1);
            ca.Wp_out = pitch8(hout); ca.tile_w = 32; ca.num_sms = num_sms;
            if (n->precise && n->conv_impl == 0) {
                ca.pm_max_cout = 0; ca.cm_shift = 0;   // the split epilogue lives in the plain cout-major tile
                ca.split_lo_off = static_cast<long long>(B) * ca.Cout * hout * ca.Wp_out;
            }
            launches += 1;
            const int rr = n->conv_impl == 0 ? conv_tc_launch(ca, stream) : conv_simt_launch(ca, stream);
            mark(2, static_cast<int>(2 * bi) + (phase ? 0 : 1));
            return rr;
        };
        auto act_raw = [&](const Sg2Layer& L, int fir, const float* style_next, __half* x_next, bool rgb, const float* prev,
                           const __half* pre, int ns_out) -> int {
            ActArgs a;
            a.y = Y; a.pre = pre; a.noise = L.noise.dev; a.bias = L.bias.dev;
            const size_t nb = L.noise.numel / L.noise.plane;
            MB_REQUIRE(nb == 1 || nb == static_cast<size_t>(B), "mb_net_forward: noise of a %dx%d layer holds %zu maps, batch is %d",
                       r, r, nb, B);
            a.noise_bstride = nb == 1 ? 0 : static_cast<long long>(L.noise.plane);
            a.style_next = style_next; a.x_next = x_next;
            a.rgb_w = rgb ? b.torgb.weight.dev : nullptr; a.rgb_style = s_rgb; a.rgb_bias = b.torgb.bias.dev;
            a.img_prev = prev; a.img = img[cur ^ 1];
            a.B = B; a.C = L.cout; a.R = r; a.Hy = fir ? r + 1 : r; a.Cp = cpad16(L.cout);
            a.phase = fir; a.ns = ns_out;
            a.Wpy = fir ? pitch8(r / 2 + 1) : pitch8(r);
            a.y_lo = (n->precise && n->conv_impl == 0) ? static_cast<long long>(B) * (fir ? 4 : 1) * L.cout * (fir ? r / 2 + 1 : r) * a.Wpy : 0;
            a.fir = fir; a.nimg = n->img_channels; a.clamp = 256.0f;
            if (pre == nullptr && n->act_tiled) {
                // tiled kernel: rows per CTA by what the staged block costs in shared memory
                const int th = (L.cout <= 128 && r >= 4) ? 4 : ((L.cout <= 256 && r >= 2) ? 2 : 1);
                const size_t smem = sizeof(float) * (static_cast<size_t>(L.cout) * (th * kActP + 1) + 8 * (th + 3) * 36 + th * kActP + n->img_channels * L.cout);
                auto run = [&](auto kern) -> int {
                    MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                    dim3 grid(ceil_div(r, kActP), ceil_div(r, th), B);
                    kern<<<grid, 256, smem, stream>>>(a);
                    return MB_OK;
                };
                int rr = th == 4 ? run(sg2_act_tiled_kernel<4>) : (th == 2 ? run(sg2_act_tiled_kernel<2>) : run(sg2_act_tiled_kernel<1>));
                if (rr != MB_OK) return rr;
            } else {
                const size_t smem = sizeof(float) * L.cout * (kActP + 1);
                static size_t smem_set = 0;
                if (smem > 48 * 1024 && smem > smem_set) {
                    MB_CUDA(cudaFuncSetAttribute(sg2_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                    smem_set = smem;
                }
                dim3 grid(ceil_div(r, kActP), r, B);
                sg2_act_kernel<<<grid, 256, smem, stream>>>(a);
            }
            MB_CUDA(cudaGetLastError());
            launches += 1;
            mark(3, static_cast<int>(2 * bi) + (fir ? 0 : 1));
            return MB_OK;
        };
        // name_idx = position of this layer in the wrapper's layer_names (stylegan2.py:48-51): block 0 owns entries 0 and 1
        // (both "bs.0.conv1"), block i entries 2i (conv0) and 2i + 1 (conv1).  With warps on the layer, the activation
        // is first written unstyled, warped in hook order (ping-pong), and only then styled / sent through ToRGB.
        auto act = [&](const Sg2Layer& L, int fir, const float* style_next, bool rgb, const float* prev, int name_idx) -> int {
            bool warped = false;
            for (int wl_ : n->warp_layer)
                if (wl_ == name_idx || (bi == 0 && wl_ <= 1)) warped = true;
            if (!warped) return act_raw(L, fir, style_next, style_next ? X : nullptr, rgb, prev, nullptr, nsx);
            int rr;
            if ((rr = act_raw(L, fir, nullptr, T[0], false, nullptr, nullptr, 1)) != MB_OK) return rr;   // the warps run on plain fp16
            int cur_t = 0;
            const int cp = cpad16(L.cout);
            for (size_t wi = 0; wi < n->warp_layer.size(); ++wi) {
                const int wl_ = n->warp_layer[wi];
                if (!(wl_ == name_idx || (bi == 0 && wl_ <= 1))) continue;
                const long long tot = static_cast<long long>(B) * r * r * (cp / 8);
                sg2_warp_kernel<<<grid1d(tot), 256, 0, stream>>>(T[cur_t], T[cur_t ^ 1], n->warp_mats + wi * static_cast<size_t>(B) * 6, B, r, cp);
                MB_CUDA(cudaGetLastError());
                launches += 1;
                mark(4, name_idx);
                cur_t ^= 1;
            }
            return act_raw(L, 0, style_next, style_next ? X : nullptr, rgb, prev, T[cur_t], nsx);
        };
        if (!b.has_conv0) {
            const long long tot = static_cast<long long>(B) * r * r * cpad16(b.cout);
            const_input_kernel<<<grid1d(tot), 256, 0, stream>>>(b.cst.dev, s_conv1, X, B, b.cout, r, cpad16(b.cout), nsx);
            MB_CUDA(cudaGetLastError());
            launches += 1;
            mark(1, -1);
        } else {
            // conv0: X holds x * style(conv0) at r/2 -> polyphase transposed conv (four parity planes) -> FIR + act
            if ((rc = conv(b.conv0, X, r / 2, 1, dco + wl.d_l[bi * 3 + 0])) != MB_OK) return rc;
            if ((rc = act(b.conv0, 1, s_conv1, false, nullptr, static_cast<int>(2 * bi))) != MB_OK) return rc;
        }
        (void)s_conv0;
        if ((rc = conv(b.conv1, X, r, 0, dco + wl.d_l[bi * 3 + 1])) != MB_OK) return rc;
        const float* prev = nullptr;
        if (have_img) {
            // img[cur] (r/2) -> upsampled into img[cur^1]?  keep three-step: upsample into the other buffer, accumulate in place
            const long long tot = static_cast<long long>(B) * n->img_channels * r * r;
            upsample_rgb_kernel<<<grid1d(tot), 256, 0, stream>>>(img[cur], img[cur ^ 1], B * n->img_channels, r / 2);
            MB_CUDA(cudaGetLastError());
            launches += 1;
            mark(5, static_cast<int>(bi));
            prev = img[cur ^ 1];
        }
        const float* s_next = last ? nullptr : styles + wl.style_l[(bi + 1) * 3 + 0];
        if ((rc = act(b.conv1, 0, s_next, true, prev, static_cast<int>(2 * bi + 1))) != MB_OK) return rc;
        cur ^= 1;
        have_img = true;
    }
    const long long nout = static_cast<long long>(B) * n->img_channels * n->res * n->res;
    if (out_fmt == MB_OUT_F32_NCHW) {
        MB_CUDA(cudaMemcpyAsync(out, img[cur], nout * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    } else if (out_fmt == MB_OUT_F32_NCHW_01 || out_fmt == MB_OUT_F32_NCHW_UNIT) {
        img_to_unit_kernel<<<grid1d(nout), 256, 0, stream>>>(img[cur], static_cast<float*>(out), nout, out_fmt == MB_OUT_F32_NCHW_01);
        MB_CUDA(cudaGetLastError());
        launches += 1;
    } else {
        img_to_u8_kernel<<<grid1d(nout), 256, 0, stream>>>(img[cur], static_cast<uint8_t*>(out), B, n->img_channels, n->res);
        MB_CUDA(cudaGetLastError());
        launches += 1;
    }
    mark(5, static_cast<int>(n->blocks.size()));
    n->last_launches = launches;
    return MB_OK;
}

}  // namespace mb
