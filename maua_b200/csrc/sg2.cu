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
#include "resize_taps.cuh"

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

// const [C][npix] f32 * style[b][c] -> channels-last fp16 [B][npix][ns * Cp]
__global__ void const_input_kernel(const float* __restrict__ cst, const float* __restrict__ style, __half* __restrict__ out,
                                   int B, int C, int npix, int Cp, int ns) {
    const long long total = static_cast<long long>(B) * npix * Cp;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % Cp);
        long long q = idx / Cp;
        const int px = static_cast<int>(q % npix);
        const int b = static_cast<int>(q / npix);
        put_split(out + q * (static_cast<long long>(ns) * Cp), c, Cp, ns, c < C ? cst[static_cast<long long>(c) * npix + px] * style[b * C + c] : 0.0f);
    }
}

struct ActArgs {
    const __half* y;        // planar conv output [B][C][Hy][Wpy]; (Hy, Wy) = (RH, RW) (+1 with FIR).  phase: the four parity planes of the
                            // polyphase transposed conv, [B][4][C][RH/2 + 1][Wpy]: y(yy, xx) = plane (yy & 1) * 2 + (xx & 1) at (yy >> 1, xx >> 1)
    long long y_lo;         // > 0: y + y_lo holds the low halves of a split output (value = hi + lo)
    int phase, ns;          // ns: channel blocks of x_next (3 = [xh | xl | xh], 1 = plain fp16)
    const float* pre32;     // like pre, as fp32 (behind an output-size hook: the resized map keeps fp32 precision)
    float* x_next32;        // if set (pass 1 in front of an output-size hook): the unstyled activation as fp32 [B][RH][RW][Cp]
    const __half* pre;      // if set: the finished activation, channels-last [B][RH][RW][Cp] (a warped / resized feature map); y, noise, bias unused
    const float* noise;     // [RH][RW] (noise_bstride 0) or per-frame [B][RH][RW] (noise_bstride RH*RW) or nullptr
    const float* bias;      // [C]
    const float* style_next;  // [B][C] or nullptr (last block)
    __half* x_next;         // channels-last [B][RH][RW][Cp] or nullptr
    const float* rgb_w;     // [3][C] ToRGB weight or nullptr (no ToRGB after conv0)
    const float* rgb_style; // [B][C] (already * 1/sqrt(C))
    const float* rgb_bias;  // [3]
    const float* img_prev;  // [B][3][RH][RW] upsampled skip image or nullptr
    float* img;             // [B][3][RH][RW]
    int B, C, RH, RW, Hy, Wy, Wpy, Cp, fir, nimg;   // the map is RH x RW (non-square behind an output-size hook)
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
        if (a.pre32) {
            if (w < a.RW) v = a.pre32[((static_cast<long long>(b) * a.RH + h) * a.RW + w) * a.Cp + c];
        } else if (a.pre) {
            if (w < a.RW) v = __half2float(a.pre[((static_cast<long long>(b) * a.RH + h) * a.RW + w) * a.Cp + c]);
        } else if (w < a.RW) {
            const int hp = a.phase ? a.RH / 2 + 1 : a.Hy;               // rows of a stored plane
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
                        if (xx >= 0 && xx < a.Wy) row = fmaf(f4[kx], at(yy, xx), row);
                    }
                    v = fmaf(f4[ky], row, v);
                }
            } else {
                v = at(h, w);
            }
            if (a.noise) v += a.noise[b * a.noise_bstride + h * a.RW + w];
            v += a.bias[c];
            v = (v < 0.0f ? v * 0.2f : v) * 1.41421356237309515f;
            v = fminf(fmaxf(v, -a.clamp), a.clamp);
        }
        xs[c * (kActP + 1) + px] = v;
    }
    __syncthreads();
    const int npx = min(kActP, a.RW - w0);
    if (a.x_next32) {
        float* o = a.x_next32 + ((static_cast<long long>(b) * a.RH + h) * a.RW + w0) * a.Cp;
        for (int idx = threadIdx.x; idx < npx * a.Cp; idx += blockDim.x) {
            const int px = idx / a.Cp, c = idx - px * a.Cp;
            o[idx] = c < a.C ? xs[c * (kActP + 1) + px] : 0.0f;
        }
    }
    if (a.x_next) {
        const int cpx = a.ns * a.Cp;   // halfs per pixel of x_next
        __half* o = a.x_next + ((static_cast<long long>(b) * a.RH + h) * a.RW + w0) * cpx;
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
                const long long oi = ((static_cast<long long>(b) * a.nimg + o) * a.RH + h) * a.RW + w0 + px;
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
        nz[i] = (a.noise && hh < a.RH && ww < a.RW) ? a.noise[b * a.noise_bstride + hh * a.RW + ww] : 0.0f;
    }
    if (a.rgb_w)
        for (int i = threadIdx.x; i < a.nimg * a.C; i += blockDim.x) wrgb[i] = a.rgb_w[i] * a.rgb_style[b * a.C + i % a.C];
    __syncthreads();
    const int hp = a.phase ? a.RH / 2 + 1 : a.Hy;
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
            // patch rows h0 - 1 .. h0 + TH + 1, columns w0 - 1 .. w0 + 33 of the (RH + 1) x (RW + 1) conv output, zero outside
            for (int r = 0; r < PR; ++r) {
                const int yy = h0 - 1 + r;
                for (int cc = lane; cc < 35; cc += 32) {
                    const int xx = w0 - 1 + cc;
                    pw[r * PC + cc] = (yy >= 0 && yy < a.Hy && xx >= 0 && xx < a.Wy) ? at(yy, xx) : 0.0f;
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
            for (int t = 0; t < TH; ++t) outv[t] = (h0 + t < a.RH && w0 + lane < a.RW) ? at(h0 + t, w0 + lane) : 0.0f;
        }
#pragma unroll
        for (int t = 0; t < TH; ++t) {
            float v = outv[t] + nz[t * kActP + lane] + bias;
            v = (v < 0.0f ? v * 0.2f : v) * 1.41421356237309515f;
            v = fminf(fmaxf(v, -a.clamp), a.clamp);
            xs[c * XP + t * kActP + lane] = (h0 + t < a.RH && w0 + lane < a.RW) ? v : 0.0f;
        }
    }
    __syncthreads();
    const int npx = min(kActP, a.RW - w0);
    if (a.x_next32) {
        for (int idx = threadIdx.x; idx < PXT * a.Cp; idx += blockDim.x) {
            const int pix = idx / a.Cp, c = idx - pix * a.Cp;
            const int t = pix / kActP, px = pix - t * kActP;
            if (px >= npx || h0 + t >= a.RH) continue;
            a.x_next32[((static_cast<long long>(b) * a.RH + h0 + t) * a.RW + w0 + px) * a.Cp + c] = c < a.C ? xs[c * XP + pix] : 0.0f;
        }
    }
    if (a.x_next) {
        const int cpx = a.ns * a.Cp, half_cp = a.Cp / 2;
        for (int idx = threadIdx.x; idx < PXT * half_cp; idx += blockDim.x) {
            const int pix = idx / half_cp, c = (idx - pix * half_cp) * 2;
            const int t = pix / kActP, px = pix - t * kActP;
            if (px >= npx || h0 + t >= a.RH) continue;
            const float s0 = c < a.C ? (a.style_next ? a.style_next[b * a.C + c] : 1.0f) : 0.0f;
            const float s1 = c + 1 < a.C ? (a.style_next ? a.style_next[b * a.C + c + 1] : 1.0f) : 0.0f;
            const float v0 = c < a.C ? xs[c * XP + pix] * s0 : 0.0f;
            const float v1 = c + 1 < a.C ? xs[(c + 1) * XP + pix] * s1 : 0.0f;
            __half* op = a.x_next + ((static_cast<long long>(b) * a.RH + h0 + t) * a.RW + w0 + px) * cpx;
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
            if (px >= npx || h0 + t >= a.RH) continue;
            float acc = 0.0f;
            for (int c = 0; c < a.C; ++c) acc = fmaf(xs[c * XP + pix], wrgb[o * a.C + c], acc);
            float v = fminf(fmaxf(acc + a.rgb_bias[o], -a.clamp), a.clamp);
            const long long oi = ((static_cast<long long>(b) * a.nimg + o) * a.RH + h0 + t) * a.RW + w0 + px;
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
                                                       const float* __restrict__ mats /*[B][2][3]*/, int B, int RH, int RW, int Cp) {
    const int groups = Cp / 8;
    const long long total = static_cast<long long>(B) * RH * RW * groups;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(idx % groups);
        long long q = idx / groups;
        const int x = static_cast<int>(q % RW); q /= RW;
        const int y = static_cast<int>(q % RH);
        const int b = static_cast<int>(q / RH);
        const float* m = mats + b * 6;
        const float sx = reflect_coord(m[0] * x + m[1] * y + m[2], RW);
        const float sy = reflect_coord(m[3] * x + m[4] * y + m[5], RH);
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
            if (xx < 0 || xx >= RW || yy < 0 || yy >= RH) continue;  // grid_sample: out-of-range taps count as zero
            const uint4 v = *reinterpret_cast<const uint4*>(src + ((static_cast<long long>(b) * RH + yy) * RW + xx) * Cp + g * 8);
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
        *reinterpret_cast<uint4*>(dst + ((static_cast<long long>(b) * RH + y) * RW + x) * Cp + g * 8) = o;
    }
}

// upsample2d(img, [1,3,3,1]): zero-insert x2, pad (2,1), 4x4 FIR with gain 4 (ops.py:117-133) on [N][rh][rw] planes
__global__ void upsample_rgb_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int rh, int rw) {
    const int RH = 2 * rh, RW = 2 * rw;
    const float f4[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    const long long total = static_cast<long long>(N) * RH * RW;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int X = static_cast<int>(idx % RW);
        long long q = idx / RW;
        const int Y = static_cast<int>(q % RH);
        const int n = static_cast<int>(q / RH);
        const float* xp = x + static_cast<long long>(n) * rh * rw;
        float acc = 0.0f;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const int my = Y + ky - 2;  // index into the zero-inserted signal
            if (my < 0 || (my & 1) || (my >> 1) >= rh) continue;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const int mx = X + kx - 2;
                if (mx < 0 || (mx & 1) || (mx >> 1) >= rw) continue;
                acc = fmaf(f4[ky] * f4[kx], xp[(my >> 1) * rw + (mx >> 1)], acc);
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

__global__ void img_to_u8_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, int B, int C, long long npix) {
    const long long total = static_cast<long long>(B) * npix * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        long long q = idx / C;
        const long long px = q % npix;
        const int b = static_cast<int>(q / npix);
        float v = (img[(static_cast<long long>(b) * C + c) * npix + px] + 1.0f) * 0.5f;
        v = fminf(fmaxf(v, 0.0f), 1.0f);
        out[idx] = static_cast<uint8_t>(rintf(v * 255.0f));
    }
}

// ---- output-size hook (maua/GAN/wrappers/stylegan2.py:104-151, get_hook :216-340) ---------------------------------------------
// The reference registers a forward hook on one SynthesisLayer that resizes its output ("stretch": bicubic interpolate;
// "pad-<how>-<where>": torch.nn.functional.pad with a constant / reflect / replicate / circular border) and adds a fixed noise
// map; the block's image is resized the same way and the hooked block's ToRGB output is mapped back first (inverse: bicubic to
// the layer size, or the crop that undoes the padding).  Every later layer then runs on the resized, possibly non-square map.
enum { SG2_RS_STRETCH = 0, SG2_RS_CONST = 1, SG2_RS_REFLECT = 2, SG2_RS_REPLICATE = 3, SG2_RS_CIRCULAR = 4 };

// source index of padded coordinate o (already shifted by the leading pad) in a row of n samples; -1 = the constant
__device__ __forceinline__ int pad_index(int o, int n, int mode) {
    if (o >= 0 && o < n) return o;
    if (mode == SG2_RS_CONST) return -1;
    if (mode == SG2_RS_REPLICATE) return o < 0 ? 0 : n - 1;
    if (mode == SG2_RS_CIRCULAR) { const int m = o % n; return m < 0 ? m + n : m; }
    if (n == 1) return 0;
    const int p = 2 * (n - 1);   // reflect without repeating the edge sample
    int m = o % p;
    if (m < 0) m += p;
    return m < n ? m : p - m;
}

// channels-last fp32 [B][h][w][Cp] -> [B][oh][ow][Cp] (+ noise[c][oh][ow]) as fp32 (y32) or fp16 (y16: warps follow);
// one thread = 4 channels of one output pixel
__global__ void __launch_bounds__(256) sg2_resize_kernel(const float* __restrict__ x, float* __restrict__ y32, __half* __restrict__ y16,
                                                         const float* __restrict__ noise, int B, int C, int h, int w, int Cp, int oh, int ow, int mode,
                                                         int pad_t, int pad_l, float value) {
    const int groups = Cp / 4;
    const long long total = static_cast<long long>(B) * oh * ow * groups;
    const float sh = static_cast<float>(h) / static_cast<float>(oh), sw = static_cast<float>(w) / static_cast<float>(ow);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(idx % groups);
        long long r = idx / groups;
        const int ox = static_cast<int>(r % ow); r /= ow;
        const int oy = static_cast<int>(r % oh);
        const int b = static_cast<int>(r / oh);
        const float* src = x + static_cast<long long>(b) * h * w * Cp + g * 4;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (mode == SG2_RS_STRETCH) {
            const Taps t = make_taps(oy, ox, h, w, sh, sw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float row[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(src + (static_cast<long long>(t.iy[j]) * w + t.ix[i]) * Cp);
                    row[0] += t.wx[i] * v.x; row[1] += t.wx[i] * v.y; row[2] += t.wx[i] * v.z; row[3] += t.wx[i] * v.w;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] += t.wy[j] * row[c];
            }
        } else {
            const int sy = pad_index(oy - pad_t, h, mode), sx = pad_index(ox - pad_l, w, mode);
            if (sy < 0 || sx < 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] = value;
            } else {
                const float4 v = *reinterpret_cast<const float4*>(src + (static_cast<long long>(sy) * w + sx) * Cp);
                acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[3] = v.w;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int ch = g * 4 + c;
            if (ch >= C) acc[c] = 0.0f;   // the pad channels stay zero (conv operand)
            else if (noise) acc[c] += noise[(static_cast<long long>(ch) * oh + oy) * ow + ox];
        }
        const long long o = ((static_cast<long long>(b) * oh + oy) * ow + ox) * Cp + g * 4;
        if (y32) *reinterpret_cast<float4*>(y32 + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else {
            *reinterpret_cast<__half2*>(y16 + o) = __floats2half2_rn(acc[0], acc[1]);
            *reinterpret_cast<__half2*>(y16 + o + 2) = __floats2half2_rn(acc[2], acc[3]);
        }
    }
}

// per-channel mean and unbiased standard deviation over [B][H][W] of a channels-last map -> stats[c], stats[C + c]
// (what the reference's hook measures on the resized features to draw its noise map, stylegan2.py:236-247); one CTA per channel
__global__ void __launch_bounds__(256) sg2_channel_stats_kernel(const float* __restrict__ x, float* __restrict__ stats, long long npix, int C, int Cp) {
    const int c = blockIdx.x;
    double s = 0.0, ss = 0.0;
    for (long long i = threadIdx.x; i < npix; i += blockDim.x) {
        const double v = static_cast<double>(x[i * Cp + c]);
        s += v; ss += v * v;
    }
    __shared__ double sh_s[256], sh_ss[256];
    sh_s[threadIdx.x] = s; sh_ss[threadIdx.x] = ss;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) { sh_s[threadIdx.x] += sh_s[threadIdx.x + k]; sh_ss[threadIdx.x] += sh_ss[threadIdx.x + k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double n = static_cast<double>(npix), mean = sh_s[0] / n;
        const double var = npix > 1 ? fmax((sh_ss[0] - n * mean * mean) / (n - 1.0), 0.0) : 0.0;
        stats[c] = static_cast<float>(mean);
        stats[C + c] = static_cast<float>(sqrt(var));
    }
}

// fp32 planar images [N][h][w] -> [N][oh][ow] (+ add[N][oh][ow]): the image hooks of the hooked block.  Negative pads crop.
__global__ void __launch_bounds__(256) sg2_img_resize_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ add,
                                                             int N, int h, int w, int oh, int ow, int mode, int pad_t, int pad_l, float value) {
    const long long total = static_cast<long long>(N) * oh * ow;
    const float sh = static_cast<float>(h) / static_cast<float>(oh), sw = static_cast<float>(w) / static_cast<float>(ow);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(idx % ow);
        const int oy = static_cast<int>((idx / ow) % oh);
        const float* src = x + (idx / (static_cast<long long>(ow) * oh)) * h * w;
        float acc = 0.0f;
        if (mode == SG2_RS_STRETCH) {
            const Taps t = make_taps(oy, ox, h, w, sh, sw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float row = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) row += t.wx[i] * src[t.iy[j] * w + t.ix[i]];
                acc += t.wy[j] * row;
            }
        } else {
            const int sy = pad_index(oy - pad_t, h, mode), sx = pad_index(ox - pad_l, w, mode);
            acc = (sy < 0 || sx < 0) ? value : src[sy * w + sx];
        }
        y[idx] = acc + (add ? add[idx] : 0.0f);
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
    // output-size hook (sg2_set_resize): layer = index into the wrapper's layer_names, -1 = none
    int rs_layer = -1, rs_mode = 0, rs_th = 0, rs_tw = 0, rs_pad_t = 0, rs_pad_l = 0;
    float rs_value = 0.0f;
    const float* rs_noise = nullptr;   // device [C][th][tw] added to the resized features (layer 0: the resized constant input
                                       // itself, noise included), caller-owned; nullptr = none
    float* rs_stats = nullptr;         // device [2][C]: when set, the next forwards write mean / std of the resized features here
};

namespace {
struct Sg2Dims { int h0, w0, h1, w1, ho, wo; };   // maps of conv0's output, conv1's output and what ToRGB / the next block see
// Layer geometry behind an output-size hook on layer_names[L] (block L / 2): L = 0 resizes the constant input (a forward
// pre-hook in the reference, no image hooks), an even L > 0 the output of conv0 (conv1 and ToRGB run resized), an odd L the
// output of conv1 (ToRGB runs resized); every later block doubles what it receives.
std::vector<Sg2Dims> sg2_dims(const Sg2Net* n) {
    std::vector<Sg2Dims> d(n->blocks.size());
    const int bk = n->rs_layer < 0 ? 1 << 30 : n->rs_layer / 2;
    for (size_t j = 0; j < n->blocks.size(); ++j) {
        const int s = n->blocks[j].res;
        if (static_cast<int>(j) < bk) d[j] = {s, s, s, s, s, s};
        else if (static_cast<int>(j) == bk) {
            if (n->rs_layer == 0) d[j] = {n->rs_th, n->rs_tw, n->rs_th, n->rs_tw, n->rs_th, n->rs_tw};
            else if (n->rs_layer % 2 == 0) d[j] = {s, s, n->rs_th, n->rs_tw, n->rs_th, n->rs_tw};
            else d[j] = {s, s, s, s, n->rs_th, n->rs_tw};
        } else {
            const int h = 2 * d[j - 1].ho, w = 2 * d[j - 1].wo;
            d[j] = {h, w, h, w, h, w};
        }
    }
    return d;
}
}  // namespace

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
        // the reference swaps noise_const for a per-frame [B,1,h,w] tensor on every call (wrappers/stylegan2.py:81-96), and
        // behind an output-size hook for maps of the resized layer's size (:137-147): the last two dimensions are the map
        MB_REQUIRE(ndim >= 2 && shape[ndim - 1] > 0 && shape[ndim - 2] > 0, "mb_net_set_param: '%s' needs a [..., h, w] shape", name);
        p.plane = static_cast<size_t>(shape[ndim - 1]) * static_cast<size_t>(shape[ndim - 2]);
        p.shape.assign(shape + ndim - 2, shape + ndim);
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
    size_t styles, d, x, y, img0, img1, img2, t0, t1, total;
    std::vector<size_t> style_l, d_l;  // per layer (conv0, conv1, torgb per block) float offsets
};
Sg2Ws sg2_ws(const Sg2Net* n, int B) {
    Sg2Ws w;
    size_t ns = 0, nd = 0, mx = 0, my = 0, mt = 0;
    const size_t nsx = n->precise ? 3 : 1, nsy = n->precise ? 2 : 1;   // channel blocks of X, planes (hi, lo) of Y
    const std::vector<Sg2Dims> dims = sg2_dims(n);
    size_t bi = 0, max_img = 0;
    for (const auto& b : n->blocks) {
        for (const Sg2Layer* L : {&b.conv0, &b.conv1, &b.torgb}) {
            w.style_l.push_back(ns);
            w.d_l.push_back(nd);
            ns += static_cast<size_t>(B) * L->cin;
            nd += static_cast<size_t>(B) * L->cout * (L == &b.conv0 ? 4 : 1);   // conv0: one coefficient per parity plane
        }
        const Sg2Dims& g = dims[bi];
        const size_t px_max = std::max({static_cast<size_t>(g.h0) * g.w0, static_cast<size_t>(g.h1) * g.w1, static_cast<size_t>(g.ho) * g.wo});
        mx = std::max(mx, static_cast<size_t>(B) * px_max * cpad16(b.cout) * nsx);
        mt = std::max(mt, static_cast<size_t>(B) * px_max * cpad16(b.cout));
        if (b.has_conv0) {
            const size_t hi = g.h0 / 2, wi = g.w0 / 2;
            mx = std::max(mx, static_cast<size_t>(B) * hi * wi * cpad16(b.cin) * nsx);
            my = std::max(my, static_cast<size_t>(B) * 4 * b.cout * (hi + 1) * pitch8(static_cast<int>(wi + 1)) * nsy);
        }
        my = std::max(my, static_cast<size_t>(B) * b.cout * g.h1 * pitch8(g.w1) * nsy);
        max_img = std::max({max_img, static_cast<size_t>(g.ho) * g.wo, static_cast<size_t>(b.res) * b.res});
        ++bi;
    }
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = round_up_sz(off + bytes, 1024); return o; };
    w.styles = take(ns * 4); w.d = take(nd * 4);
    w.x = take(mx * 2); w.y = take(my * 2);
    const size_t img = static_cast<size_t>(B) * n->img_channels * max_img * 4;
    w.img0 = take(img); w.img1 = take(img);
    w.img2 = n->rs_layer > 0 ? take(img) : off;   // the hooked block's ToRGB output before it is mapped back
    w.t0 = w.t1 = off;
    if (!n->warp_layer.empty() || n->rs_layer > 0) { const size_t eb = n->rs_layer > 0 ? 4 : 2; w.t0 = take(mt * eb); w.t1 = take(mt * eb); }  // ping-pong buffers of the warped / resized feature map (plain fp16)
    w.total = off;
    return w;
}
}  // namespace

size_t sg2_workspace_bytes(const Sg2Net* n, int B) { return sg2_ws(n, B).total; }
int sg2_num_ws(const Sg2Net* n) { return n->num_ws; }
int sg2_resolution(const Sg2Net* n) { return n->res; }
void sg2_output_hw(const Sg2Net* n, int* h, int* w) {
    const Sg2Dims g = sg2_dims(n).back();
    *h = g.ho; *w = g.wo;
}

int sg2_set_resize(Sg2Net* n, int layer, int mode, int th, int tw, int pad_t, int pad_l, float value, const float* noise, float* stats) {
    if (layer < 0) {
        n->rs_layer = -1; n->rs_noise = nullptr; n->rs_stats = nullptr;
        return MB_OK;
    }
    MB_REQUIRE(layer < 2 * static_cast<int>(n->blocks.size()), "mb_sg2_set_resize: layer %d out of range [0, %d)", layer, 2 * static_cast<int>(n->blocks.size()));
    MB_REQUIRE(mode >= SG2_RS_STRETCH && mode <= SG2_RS_CIRCULAR, "mb_sg2_set_resize: unknown mode %d", mode);
    MB_REQUIRE(th >= 1 && tw >= 1 && th <= 8192 && tw <= 8192, "mb_sg2_set_resize: target size %dx%d out of range", th, tw);
    const int s = n->blocks[layer / 2].res;
    if (mode != SG2_RS_STRETCH) {
        MB_REQUIRE(th >= s && tw >= s, "mb_sg2_set_resize: padding cannot shrink a %dx%d layer to %dx%d (negative padding is a TODO of the reference too)", s, s, th, tw);
        MB_REQUIRE(pad_t >= 0 && pad_l >= 0 && pad_t <= th - s && pad_l <= tw - s, "mb_sg2_set_resize: leading pads (%d, %d) do not fit", pad_t, pad_l);
        if (mode == SG2_RS_REFLECT) MB_REQUIRE(th - s < s && tw - s < s, "mb_sg2_set_resize: reflect padding needs pads smaller than the layer (%d)", s);
    }
    MB_REQUIRE(layer != 0 || noise != nullptr, "mb_sg2_set_resize: layer 0 takes the resized constant input as its map");
    n->rs_layer = layer; n->rs_mode = mode; n->rs_th = th; n->rs_tw = tw; n->rs_pad_t = pad_t; n->rs_pad_l = pad_l; n->rs_value = value;
    n->rs_noise = noise; n->rs_stats = stats;
    return MB_OK;
}
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

    const std::vector<Sg2Dims> dims = sg2_dims(n);
    float* img_y = reinterpret_cast<float*>(base + wl.img2);
    int cur = 0;  // which img buffer holds the previous block's image
    bool have_img = false;
    for (size_t bi = 0; bi < n->blocks.size(); ++bi) {
        const Sg2Block& b = n->blocks[bi];
        const Sg2Dims& g = dims[bi];
        const bool last = bi + 1 == n->blocks.size();
        const float* s_conv0 = styles + wl.style_l[bi * 3 + 0];
        const float* s_conv1 = styles + wl.style_l[bi * 3 + 1];
        const float* s_rgb = styles + wl.style_l[bi * 3 + 2];
        // phase = 1: the polyphase transposed conv (2x2 'full' kernel, 4 * cout parity planes of (hin + 1) x (win + 1)); else 3x3 'same'
        auto conv = [&](const Sg2Layer& L, const __half* xin, int hin, int win, int phase, const float* d) -> int {
            ConvTcArgs ca;
            ca.x = xin; ca.wpk = L.wpk; ca.d = d; ca.bias = nullptr; ca.y = Y;
            ca.B = B; ca.Cin = nsx * cpad16(L.cin); ca.Cout = phase ? 4 * L.cout : L.cout; ca.Hin = hin; ca.Win = win; ca.Cp_in = ca.Cin;
            ca.ksz = phase ? 2 : 3; ca.pad = 1;
            const int hout = hin + 2 * ca.pad - (ca.ksz - 1), wout = win + 2 * ca.pad - (ca.ksz - 1);
            ca.Wp_out = pitch8(wout); ca.tile_w = 32; ca.num_sms = num_sms;
            if (n->precise && n->conv_impl == 0) {
                ca.pm_max_cout = 0; ca.cm_shift = 0;   // the split epilogue lives in the plain cout-major tile
                ca.split_lo_off = static_cast<long long>(B) * ca.Cout * hout * ca.Wp_out;
            }
            launches += 1;
            const int rr = n->conv_impl == 0 ? conv_tc_launch(ca, stream) : conv_simt_launch(ca, stream);
            mark(2, static_cast<int>(2 * bi) + (phase ? 0 : 1));
            return rr;
        };
        // one pass of the fused kernel over an rh x rw map; out_img: where the ToRGB result (+ prev) goes
        auto act_raw = [&](const Sg2Layer& L, int fir, const float* style_next, __half* x_next, bool rgb, const float* prev,
                           const __half* pre, int ns_out, int rh, int rw, float* out_img, const float* pre32, float* x_next32) -> int {
            ActArgs a;
            a.y = Y; a.pre = pre; a.pre32 = pre32; a.x_next32 = x_next32; a.noise = L.noise.dev; a.bias = L.bias.dev;
            const bool from_conv = pre == nullptr && pre32 == nullptr;
            if (from_conv) {
                const size_t nb = L.noise.plane ? L.noise.numel / L.noise.plane : 0;
                MB_REQUIRE(L.noise.shape.size() == 2 && L.noise.shape[0] == rh && L.noise.shape[1] == rw,
                           "mb_net_forward: the noise map of this %dx%d layer is %lldx%lld (behind an output-size hook every later layer needs noise of its new size)",
                           rh, rw, L.noise.shape.size() == 2 ? static_cast<long long>(L.noise.shape[0]) : 0LL,
                           L.noise.shape.size() == 2 ? static_cast<long long>(L.noise.shape[1]) : 0LL);
                MB_REQUIRE(nb == 1 || nb == static_cast<size_t>(B), "mb_net_forward: noise of a %dx%d layer holds %zu maps, batch is %d",
                           rh, rw, nb, B);
                a.noise_bstride = nb == 1 ? 0 : static_cast<long long>(L.noise.plane);
            } else {
                a.noise_bstride = 0;
            }
            a.style_next = style_next; a.x_next = x_next;
            a.rgb_w = rgb ? b.torgb.weight.dev : nullptr; a.rgb_style = s_rgb; a.rgb_bias = b.torgb.bias.dev;
            a.img_prev = prev; a.img = out_img;
            a.B = B; a.C = L.cout; a.RH = rh; a.RW = rw; a.Hy = fir ? rh + 1 : rh; a.Wy = fir ? rw + 1 : rw; a.Cp = cpad16(L.cout);
            a.phase = fir; a.ns = ns_out;
            a.Wpy = fir ? pitch8(rw / 2 + 1) : pitch8(rw);
            a.y_lo = (n->precise && n->conv_impl == 0) ? static_cast<long long>(B) * (fir ? 4 : 1) * L.cout * (fir ? rh / 2 + 1 : rh) * a.Wpy : 0;
            a.fir = fir; a.nimg = n->img_channels; a.clamp = 256.0f;
            if (from_conv && n->act_tiled) {
                // tiled kernel: rows per CTA by what the staged block costs in shared memory
                const int th = (L.cout <= 128 && rh >= 4) ? 4 : ((L.cout <= 256 && rh >= 2) ? 2 : 1);
                const size_t smem = sizeof(float) * (static_cast<size_t>(L.cout) * (th * kActP + 1) + 8 * (th + 3) * 36 + th * kActP + n->img_channels * L.cout);
                auto run = [&](auto kern) -> int {
                    MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                    dim3 grid(ceil_div(rw, kActP), ceil_div(rh, th), B);
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
                dim3 grid(ceil_div(rw, kActP), rh, B);
                sg2_act_kernel<<<grid, 256, smem, stream>>>(a);
            }
            MB_CUDA(cudaGetLastError());
            launches += 1;
            mark(3, static_cast<int>(2 * bi) + (fir ? 0 : 1));
            return MB_OK;
        };
        // name_idx = position of this layer in the wrapper's layer_names (stylegan2.py:48-51): block 0 owns entries 0 and 1
        // (both "bs.0.conv1"), block i entries 2i (conv0) and 2i + 1 (conv1).  With an output-size hook or warps on the layer,
        // the activation is first written unstyled (rh x rw), resized (-> oh x ow), warped in hook order (ping-pong), and only
        // then styled / sent through ToRGB.  The reference registers the resize hook at construction, the warps later: torch runs
        // forward hooks in registration order.
        auto act = [&](const Sg2Layer& L, int fir, const float* style_next, bool rgb, const float* prev, int name_idx, int rh, int rw,
                       int oh, int ow, float* out_img) -> int {
            bool warped = false;
            for (int wl_ : n->warp_layer)
                if (wl_ == name_idx || (bi == 0 && wl_ <= 1)) warped = true;
            const bool resized = n->rs_layer > 0 && n->rs_layer == name_idx;
            if (!warped && !resized) return act_raw(L, fir, style_next, style_next ? X : nullptr, rgb, prev, nullptr, nsx, rh, rw, out_img, nullptr, nullptr);
            int rr;
            const int cp = cpad16(L.cout);
            int cur_t = 0;
            const float* final32 = nullptr;
            if (resized) {
                // the hooked map keeps fp32 precision: unstyled activation -> fp32, resize (+ noise) in fp32, second pass reads fp32
                float* T32[2] = {reinterpret_cast<float*>(T[0]), reinterpret_cast<float*>(T[1])};
                if ((rr = act_raw(L, fir, nullptr, nullptr, false, nullptr, nullptr, 1, rh, rw, nullptr, nullptr, T32[0])) != MB_OK) return rr;
                const long long tot = static_cast<long long>(B) * oh * ow * (cp / 4);
                sg2_resize_kernel<<<grid1d(tot), 256, 0, stream>>>(T32[0], T32[1], nullptr, n->rs_noise, B, L.cout, rh, rw, cp, oh, ow, n->rs_mode,
                                                                   n->rs_pad_t, n->rs_pad_l, n->rs_value);
                MB_CUDA(cudaGetLastError());
                launches += 1;
                if (n->rs_stats) {
                    // statistics of the resized features (the wrapper's probe forward runs while the noise map does not exist yet)
                    sg2_channel_stats_kernel<<<L.cout, 256, 0, stream>>>(T32[1], n->rs_stats, static_cast<long long>(B) * oh * ow, L.cout, cp);
                    MB_CUDA(cudaGetLastError());
                    launches += 1;
                }
                final32 = T32[1];
                if (warped) {
                    // the warps run on plain fp16: identity "pad" of the resized map into the other buffer as halves
                    sg2_resize_kernel<<<grid1d(tot), 256, 0, stream>>>(T32[1], nullptr, T[0], nullptr, B, L.cout, oh, ow, cp, oh, ow, SG2_RS_CONST, 0, 0, 0.0f);
                    MB_CUDA(cudaGetLastError());
                    launches += 1;
                    final32 = nullptr;
                    cur_t = 0;
                }
                mark(4, name_idx);
            } else {
                if ((rr = act_raw(L, fir, nullptr, T[0], false, nullptr, nullptr, 1, rh, rw, nullptr, nullptr, nullptr)) != MB_OK) return rr;   // warps run on plain fp16
            }
            for (size_t wi = 0; wi < n->warp_layer.size(); ++wi) {
                const int wl_ = n->warp_layer[wi];
                if (!(wl_ == name_idx || (bi == 0 && wl_ <= 1))) continue;
                const long long tot = static_cast<long long>(B) * oh * ow * (cp / 8);
                sg2_warp_kernel<<<grid1d(tot), 256, 0, stream>>>(T[cur_t], T[cur_t ^ 1], n->warp_mats + wi * static_cast<size_t>(B) * 6, B, oh, ow, cp);
                MB_CUDA(cudaGetLastError());
                launches += 1;
                mark(4, name_idx);
                cur_t ^= 1;
            }
            return act_raw(L, 0, style_next, style_next ? X : nullptr, rgb, prev, final32 ? nullptr : T[cur_t], nsx, oh, ow, out_img, final32, nullptr);
        };
        const bool hooked_block = n->rs_layer > 0 && n->rs_layer / 2 == static_cast<int>(bi);   // image hooks of this block
        const int hin = bi > 0 ? dims[bi - 1].ho : 0, win = bi > 0 ? dims[bi - 1].wo : 0;
        if (!b.has_conv0) {
            const bool override_cst = n->rs_layer == 0;   // pre-hook on bs.0.conv1: the wrapper hands over the resized constant (+ noise)
            const long long tot = static_cast<long long>(B) * g.h1 * g.w1 * cpad16(b.cout);
            const_input_kernel<<<grid1d(tot), 256, 0, stream>>>(override_cst ? n->rs_noise : b.cst.dev, s_conv1, X, B, b.cout, g.h1 * g.w1,
                                                                cpad16(b.cout), nsx);
            MB_CUDA(cudaGetLastError());
            launches += 1;
            mark(1, -1);
        } else {
            // conv0: X holds x * style(conv0) at (hin, win) -> polyphase transposed conv (four parity planes) -> FIR + act
            if ((rc = conv(b.conv0, X, hin, win, 1, dco + wl.d_l[bi * 3 + 0])) != MB_OK) return rc;
            if ((rc = act(b.conv0, 1, s_conv1, false, nullptr, static_cast<int>(2 * bi), g.h0, g.w0, g.h1, g.w1, nullptr)) != MB_OK) return rc;
        }
        (void)s_conv0;
        if ((rc = conv(b.conv1, X, g.h1, g.w1, 0, dco + wl.d_l[bi * 3 + 1])) != MB_OK) return rc;
        // skip image: the previous image upsampled to this block's ORIGINAL size (twice the previous block's output)
        const int ih = have_img ? 2 * hin : b.res, iw = have_img ? 2 * win : b.res;
        const float* prev = nullptr;
        if (have_img) {
            const long long tot = static_cast<long long>(B) * n->img_channels * ih * iw;
            upsample_rgb_kernel<<<grid1d(tot), 256, 0, stream>>>(img[cur], img[cur ^ 1], B * n->img_channels, hin, win);
            MB_CUDA(cudaGetLastError());
            launches += 1;
            mark(5, static_cast<int>(bi));
            prev = img[cur ^ 1];
        }
        const float* s_next = last ? nullptr : styles + wl.style_l[(bi + 1) * 3 + 0];
        // a hook on block 0's second name (layer 1) is the same conv1 as name 1
        const int name1 = static_cast<int>(2 * bi + 1);
        if (!hooked_block) {
            if ((rc = act(b.conv1, 0, s_next, true, prev, name1, g.h1, g.w1, g.ho, g.wo, img[cur ^ 1])) != MB_OK) return rc;
        } else {
            // image hooks (stylegan2.py:129-135, get_hook :330-338): ToRGB sees the resized features; its output is mapped back to the
            // layer size (rgb_hook = inverse), added to the upsampled previous image, and the block's image is resized (img_hook)
            if ((rc = act(b.conv1, 0, s_next, true, nullptr, name1, g.h1, g.w1, g.ho, g.wo, img_y)) != MB_OK) return rc;
            const int N3 = B * n->img_channels;
            const bool stretch = n->rs_mode == SG2_RS_STRETCH;
            float* img_s = img[cur];   // the previous image has been consumed by the upsample above
            const long long tot_s = static_cast<long long>(N3) * ih * iw;
            sg2_img_resize_kernel<<<grid1d(tot_s), 256, 0, stream>>>(img_y, img_s, prev, N3, g.ho, g.wo, ih, iw, stretch ? SG2_RS_STRETCH : SG2_RS_CONST,
                                                                     stretch ? 0 : -n->rs_pad_t, stretch ? 0 : -n->rs_pad_l, 0.0f);
            MB_CUDA(cudaGetLastError());
            const long long tot_o = static_cast<long long>(N3) * g.ho * g.wo;
            sg2_img_resize_kernel<<<grid1d(tot_o), 256, 0, stream>>>(img_s, img[cur ^ 1], nullptr, N3, ih, iw, g.ho, g.wo, n->rs_mode, n->rs_pad_t,
                                                                     n->rs_pad_l, n->rs_value);
            MB_CUDA(cudaGetLastError());
            launches += 2;
            mark(5, static_cast<int>(bi));
        }
        cur ^= 1;
        have_img = true;
    }
    const Sg2Dims& gl = dims.back();
    const long long npix = static_cast<long long>(gl.ho) * gl.wo;
    const long long nout = static_cast<long long>(B) * n->img_channels * npix;
    if (out_fmt == MB_OUT_F32_NCHW) {
        MB_CUDA(cudaMemcpyAsync(out, img[cur], nout * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    } else if (out_fmt == MB_OUT_F32_NCHW_01 || out_fmt == MB_OUT_F32_NCHW_UNIT) {
        img_to_unit_kernel<<<grid1d(nout), 256, 0, stream>>>(img[cur], static_cast<float*>(out), nout, out_fmt == MB_OUT_F32_NCHW_01);
        MB_CUDA(cudaGetLastError());
        launches += 1;
    } else {
        img_to_u8_kernel<<<grid1d(nout), 256, 0, stream>>>(img[cur], static_cast<uint8_t*>(out), B, n->img_channels, npix);
        MB_CUDA(cudaGetLastError());
        launches += 1;
    }
    mark(5, static_cast<int>(n->blocks.size()));
    n->last_launches = launches;
    return MB_OK;
}

}  // namespace mb
