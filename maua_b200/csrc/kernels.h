// Internal launch interfaces between the translation units of libmaua_b200.
//
// Device activation formats (fp16):
//   "planar":        [B][C][H][Wp], Wp = W rounded up to 8 elements (16-byte aligned rows); only
//                    [0,W) of a row is meaningful.  Produced by the conv epilogue and by
//                    filtered_lrelu, consumed by filtered_lrelu and the ToRGB/output kernel.
//   "channels-last": [B][H][W][Cp], Cp = C rounded up to 8 (TMA global strides are multiples of
//                    16 bytes).  The conv's activation operand: the TMA box start must be 16-byte
//                    aligned, so the 1-pixel kw shifts have to live on an outer dimension.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace mb {

static inline int pitch8(int w) { return (w + 7) / 8 * 8; }
static inline int pitch16(int w) { return (w + 15) / 16 * 16; }

// ---- conv_tc.cu ------------------------------------------------------------------------
struct ConvTcArgs {
    const __half* x;    // channels-last [B][Hin][Win][Cp_in], already multiplied by the style
    const __half* wpk;  // packed weights [round_up(Cout,128)][k*k*nCC*64], see pack_weights
    const float* d;     // [B][Cout] demodulation coefficients or nullptr
    const float* bias;  // [Cout] added after demodulation (the layer bias filtered_lrelu would add) or nullptr
    __half* y;          // planar [B][Cout][Hin+k-1][Wp_out]
    int B, Cin, Cout, Hin, Win, Cp_in, Wp_out, ksz;
    int pad;            // zero padding per side (StyleGAN3: ksz-1 'full', StyleGAN2: ksz/2 'same'); output = in + 2*pad - (ksz-1)
    int tile_w;         // pixel-tile width 32 (x8 rows) or 16 (x16 rows)
    int pm_max_cout = 64;  // layers with ceil16(Cout) <= this (and Cin > 32) run the pixel-major tile (0 = never)
    int cm_stack = 0;      // 3x3 layers with <= 32 couts: cout-major tile with the kw taps stacked along M (conv_cms_kernel)
    int pm_stack = 1;      // 3x3 layers with 3 * ceil16(Cout) <= 256: pixel-major tile with the kw taps stacked along N (conv_pms_kernel)
    int pm_shift = 1;      // pixel-major tile: 1 / 2 = one patch load per chunk, kw shifts through the A descriptor start
    int cm_shift = 0;      // cout-major tile with resident weights: one patch load per chunk, kw shifts through the B descriptor start
                           // (correct, measured slower than per-kw loads: see KArgs::shift in conv_tc.cu)
    int row_interleaved = 0;  // 1: output layout [B][H][Cout][Wp_out] instead of planar [B][Cout][H][Wp_out]
    int epi_groups = 0;    // cout-major tile: epilogue warp groups (0 = chosen per layer, 1 = one warp per TMEM quadrant, 3)
    int narrow_a = 1;      // Cout <= 128: load only ceil8(Cout) weight rows per tile and keep them resident when they fit
    int num_sms;
    long long split_lo_off = 0;      // > 0: also store fp16(v - fp16(v)) at y + split_lo_off (needs the cout-major tile: pm_max_cout = 0, cm_shift = 0)
    unsigned int* absmax = nullptr;  // device word (zeroed by the caller): atomicMax of the bits of max |y| over the stored outputs
    // Channels-last output of the stacked pixel-major tile (the RRDB convs of csrc/rrdb.cu): when nhwc.y is set, the epilogue
    // computes v = conv * d + bias, v = v < 0 ? slope * v : v, out = alpha * v + beta * r1 + gamma * r2 and stores fp16 at
    // nhwc.y[((b * Hout + h) * Wout + w) * cp + c_off + cout]; r1 / r2 are channels-last fp16 maps of the same H x W (pixel pitch
    // r*_cp, first channel r*_off) or nullptr.  `y` is ignored.
    struct NhwcOut {
        __half* y = nullptr;
        int cp = 0, c_off = 0;
        float slope = 1.0f, alpha = 1.0f, beta = 0.0f, gamma = 0.0f;
        const __half* r1 = nullptr;
        int r1_cp = 0, r1_off = 0;
        const __half* r2 = nullptr;
        int r2_cp = 0, r2_off = 0;
    } nhwc;
};
int conv_tc_launch(const ConvTcArgs& p, cudaStream_t stream);
int conv_tc_smem_bytes(int tw);

// ---- sg3_misc.cu -----------------------------------------------------------------------
// K index of the packed weight matrix: (((kw*nCC + cc)*k + kh)*64 + c_local)
size_t packed_weight_elems(int Cout, int Cin, int ksz);
// W f32 [Cout][Cin][k][k] -> optional per-cout pre-normalisation (demodulate) -> fp16 packed,
// wsqT f32 [Cin][Cout] = sum_k Wn^2 (for the demodulation coefficients).
int pack_weights_launch(const float* w, __half* wpk, float* wsqT, int Cout, int Cin, int ksz, int prenorm,
                        cudaStream_t stream, int flip = 0);  // flip: spatially mirrored taps (transposed conv as conv)
// Plain CUDA-core direct convolution on the same operands (bisecting aid, see mb_net_set_conv_impl).
int conv_simt_launch(const ConvTcArgs& p, cudaStream_t stream);

struct StyleLayerDesc {
    const float* affine_w;  // [Cin][w_dim]
    const float* affine_b;  // [Cin]
    const float* wsqT;      // [Cin][Cout] or nullptr (no demodulation)
    const float* magnitude_ema;  // scalar on device
    float* s_out;           // [B][Cin]  normalised style * input_gain
    float* d_out;           // [B][Cout] or nullptr
    int Cin, Cout, ws_index, demodulate;
    int normalize_style;    // StyleGAN3 rescales the style to unit RMS before demodulation, StyleGAN2 does not
    float style_scale;      // torgb: 1/sqrt(Cin*k*k), else 1
};
constexpr int kMaxLayers = 24;
struct StylesArgs {
    StyleLayerDesc L[kMaxLayers];
    int num_layers, B, num_ws, w_dim;
    const float* ws;  // [B][num_ws][w_dim]
};
int styles_launch(const StylesArgs& a, cudaStream_t stream);

struct InputArgs {
    const float* ws;          // [B][num_ws][w_dim], uses ws[:,0]
    const float* affine_w;    // [4][w_dim]
    const float* affine_b;    // [4]
    const float* transform;   // [3][3] (transform_stride 0) or one matrix per sample [B][3][3] (transform_stride 9)
    int transform_stride;
    const float* freqs;       // [C][2]
    const float* phases;      // [C]
    const float* weightT;     // [C(j)][C(c)] = weight[c][j] transposed
    const float* style;       // [B][C] style of layer 0 (normalised * input_gain)
    float* scratch;           // [B][C][4]: fx, fy, phase, amplitude
    __half* out;              // channels-last [B][size][size][Cp]
    int B, num_ws, w_dim, C, size, Cp;
    float sampling_rate, bandwidth;
};
int sg3_input_launch(const InputArgs& a, cudaStream_t stream);
// tensor-core input layer: features split into fp16 hi/lo parts [fh | fl | fh] (K = 3C) for a 1x1 tcgen05 conv against
// weights [wh | wh | wl] (sg3_input_split_weights_launch + pack_weights_launch)
int sg3_input_features_launch(const InputArgs& a, __half* feat, cudaStream_t stream);
int sg3_input_split_weights_launch(const float* weight, float* w3, int C, cudaStream_t stream);

struct ToRgbArgs {
    const __half* x;     // [B][Cin][H][Wp] (already * style of the torgb layer)
    const float* w;      // [Cout<=4][Cin] raw weights (no demodulation)
    const float* bias;   // [Cout]
    void* out;
    int B, Cin, Cout, H, W, Wp, out_fmt;
    float clamp, output_scale;
};
int torgb_out_launch(const ToRgbArgs& a, cudaStream_t stream);

// x f32 [B][C][H][W] * s[b][c] * gain (+ bias[c]) -> fp16 planar (s, bias may be nullptr)
int modulate_to_half_launch(const float* x, const float* s, float gain, __half* out, int B, int C, int H, int W, int Wp,
                            cudaStream_t stream, const float* bias = nullptr);
// x f32 [B][C][H][W] * s[b][c] * gain -> fp16 channels-last [B][H][W][Cp]
int modulate_to_nhwc_launch(const float* x, const float* s, float gain, __half* out, int B, int C, int H, int W, int Cp,
                            cudaStream_t stream);
// fp16 planar -> f32 [B][C][H][W]
int half_to_float_launch(const __half* x, float* out, int B, int C, int H, int W, int Wp, cudaStream_t stream);
// fp16 channels-last -> f32 [B][C][H][W]
int nhwc_to_float_launch(const __half* x, float* out, int B, int C, int H, int W, int Cp, cudaStream_t stream);
// fp16 planar [B][C][H][Wp] -> fp16 channels-last [B][H][W][Cp] (pad channels written as zero)
int planar_to_nhwc_launch(const __half* x, __half* out, int B, int C, int H, int W, int Wp, int Cp, cudaStream_t stream);
// op-level helper: s -> normalised s (if demodulate), d[b][o]
int style_demod_launch(const float* s, const float* wsqT, float* s_out, float* d_out, int B, int Cin, int Cout,
                       int demodulate, float input_gain, cudaStream_t stream);

// ---- feature_resize.cu (output-size hooks; mode = MB_RESIZE_STRETCH | MB_RESIZE_PAD_ZERO) -------
int resize_nhwc_launch(const __half* x, __half* y, int B, int h, int w, int Cp, int oh, int ow, int mode, int pad_t, int pad_l,
                       int num_sms, cudaStream_t stream);
int resize_planar_launch(const __half* x, __half* y, int planes, int h, int w, int Wp, int oh, int ow, int Wpo, int mode,
                         int pad_t, int pad_l, int num_sms, cudaStream_t stream);

// ---- flrelu.cu -------------------------------------------------------------------------
struct FlreluArgs {
    const __half* x;     // [B][C][Hin][Wp_in]
    const float* bias;   // [C] or nullptr
    const float* scale;  // [B][C] multiplier applied to the result (next layer's style) or nullptr
    __half* y;           // planar [B][C][Hout][Wp_out]
    __half* y_nhwc;      // if set: write channels-last [B][Hout][Wout][Cp_out] instead (tensor-core kernel only)
    int Cp_out;
    float fu[32];        // HOST copy of the up filter taps (without the up^2 gain); [1]={1} when up == 1
    float fd[144];       // HOST copy: [down_taps] or [down_taps^2] (row-major) when fd_2d
    int B, C, Hin, Win, Wp_in, Hout, Wout, Wp_out;
    int up, down, up_taps, down_taps, fd_2d;
    int px0, py0;        // leading padding of the zero-inserted signal (may be negative = crop)
    float gain, slope, clamp;
    int num_sms;
    int in_row_interleaved = 0;  // 1: x is [B][Hin][C][Wp_in] (ConvTcArgs::row_interleaved) instead of planar; streaming kernels only
    const unsigned int* in_absmax = nullptr;  // device word: bits of max |x| over the input (ConvTcArgs::absmax), or nullptr
};
int flrelu_launch(const FlreluArgs& a, cudaStream_t stream);
// impl: 0 = best available (tensor-core chain), 1 = generic loops, 2 = CUDA-core polyphase kernel
int flrelu_launch_impl(const FlreluArgs& a, int impl, cudaStream_t stream);
bool flrelu_mma_supported(const FlreluArgs& a);

}  // namespace mb
