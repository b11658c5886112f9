/*
 * maua_b200.h -- C ABI of libmaua_b200.so, the Blackwell (sm_100a) replacement for the
 * audio-reactive StyleGAN render path of maua-maua-maua/maua.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes and a raw CUDA stream
 * handle, returns 0 on success or a negative MB_E* code (message via mb_last_error()).
 * No C++ exception crosses the boundary, nothing here allocates inside a forward call:
 * the caller (PyTorch in the reference) owns every tensor including outputs and the
 * workspace, the library owns only the opaque `mb_net` handle (re-packed weights).
 *
 * Each declaration cites the reference interface it replaces (paths relative to the
 * reference checkout, maua @ d968fd9).  `maua/GAN/nv` is an un-vendored submodule
 * (maua-maua-maua/nvGAN @ 7809c05, fork of NVlabs/stylegan3); for those the upstream
 * file is named.
 */
#ifndef MAUA_B200_H
#define MAUA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_OK 0
#define MB_EINVAL (-1)   /* bad argument / shape / name */
#define MB_ECUDA (-2)    /* CUDA runtime or driver error */
#define MB_ESTATE (-3)   /* call order (forward before finalize, ...) */
#define MB_ENOMEM (-4)   /* workspace too small */
#define MB_ENODEV (-5)   /* no sm_100 device */

typedef void* mb_stream; /* cudaStream_t */

/* ---- library ---------------------------------------------------------------------- */

/* Select the device and check it is compute capability 10.x.  No CPU fallback exists. */
int mb_init(int device);
/* Thread-local message of the last failing call on this thread. */
const char* mb_last_error(void);
/* ABI version (bumped on any signature change). */
int mb_abi_version(void);

/* ---- StyleGAN3 synthesis network ---------------------------------------------------
 * Replaces `stylegan3.SynthesisNetwork(w_dim=512, img_resolution=1024, img_channels=3)`
 * as constructed at maua/GAN/wrappers/stylegan3.py:33 and called at :60
 * (`self.G_synth.forward(latents)`); upstream training/networks_stylegan3.py
 * SynthesisNetwork / SynthesisInput / SynthesisLayer.
 */
typedef struct mb_sg3_cfg {
    int32_t w_dim;            /* 512 */
    int32_t img_resolution;   /* 1024 */
    int32_t img_channels;     /* 3 */
    int32_t channel_base;     /* 32768 (T) / 65536 (R) */
    int32_t channel_max;      /* 512 (T) / 1024 (R) */
    int32_t num_layers;       /* 14 */
    int32_t num_critical;     /* 2 */
    int32_t conv_kernel;      /* 3 (T) / 1 (R) */
    int32_t filter_size;      /* 6 */
    int32_t lrelu_upsampling; /* 2 */
    int32_t use_radial_filters; /* 0 (T) / 1 (R) */
    int32_t margin_size;      /* 10 */
    double first_cutoff;      /* 2 */
    double first_stopband;    /* 2**2.1 */
    double last_stopband_rel; /* 2**0.3 */
    double output_scale;      /* 0.25 */
    double conv_clamp;        /* 256 */
} mb_sg3_cfg;

/* Per-layer geometry as derived by upstream SynthesisNetwork.__init__ (index 0..num_layers,
 * the last entry is the ToRGB layer).  Pure host arithmetic: usable without a GPU. */
typedef struct mb_sg3_layer {
    int32_t idx, is_torgb, is_critically_sampled, use_fp16;
    int32_t in_channels, out_channels;
    int32_t in_size, out_size;
    int32_t in_sampling_rate, out_sampling_rate, tmp_sampling_rate;
    int32_t conv_kernel, up, down, up_taps, down_taps, down_radial;
    int32_t pad_lo, pad_hi;
    double in_cutoff, out_cutoff, in_half_width, out_half_width;
    char name[32];            /* "L{idx}_{out_size}_{out_channels}" */
} mb_sg3_layer;

typedef struct mb_net mb_net;

void mb_sg3_default_cfg(mb_sg3_cfg* cfg, int config_r /*0 = StyleGAN3-T, 1 = StyleGAN3-R*/);
/* Host-only geometry query (no device needed). layers must hold cfg->num_layers+1 entries. */
int mb_sg3_geometry(const mb_sg3_cfg* cfg, mb_sg3_layer* layers, int32_t* input_channels,
                    int32_t* input_size, double* input_sampling_rate, double* input_bandwidth);

int mb_sg3_create(const mb_sg3_cfg* cfg, mb_net** out);
void mb_net_destroy(mb_net* net);

/* ---- StyleGAN2 synthesis network --------------------------------------------------------------------
 * Replaces the reference's in-tree inference network `SynthesisNetwork(w_dim=512, img_resolution=R, img_channels=3)`
 * (maua/GAN/wrappers/inference/stylegan2.py:385-436; constructed at maua/GAN/wrappers/stylegan2.py:34-36) including
 * its ops modulated_conv2d / conv2d_resample / upfirdn2d / upsample2d / bias_act (inference/ops.py:65-233).
 * The returned handle is driven through the same mb_net_* entry points; state-dict keys are the reference's
 * ("bs.0.const", "bs.3.conv0.affine.weight", "bs.2.conv1.noise_const", "bs.1.torgb.bias", ...); noise_mode "const". */
int mb_sg2_create(int w_dim, int img_resolution, int img_channels, int channel_base, int channel_max, mb_net** out);
/* num_ws of the network behind the handle (StyleGAN3: num_layers + 2; StyleGAN2: convs + final ToRGB). */
int mb_net_num_ws(const mb_net* net);

/* Upload one tensor of the generator state dict.  `name` is the upstream state-dict key
 * ("input.freqs", "input.phases", "input.weight", "input.affine.weight", "input.affine.bias",
 * "input.transform", "L3_52_512.weight", ".bias", ".affine.weight", ".affine.bias",
 * ".magnitude_ema", ".up_filter", ".down_filter").  `data` is a DEVICE pointer to contiguous
 * float32; the library copies / re-packs it (fp16 K-major tiles for the tensor-core path) on
 * `stream`.  Replaces `load_state_dict` / parameter access of the torch module
 * (maua/GAN/wrappers/stylegan3.py:35,56-59). */
int mb_net_set_param(mb_net* net, const char* name, const float* data, const int64_t* shape,
                     int ndim, mb_stream stream);
/* Derive packed operands (pre-normalised fp16 weights, squared-weight tables). Call after the
 * last set_param and again after any later set_param. */
int mb_net_finalize(mb_net* net, mb_stream stream);

size_t mb_net_workspace_bytes(const mb_net* net, int batch);

#define MB_OUT_F32_NCHW 0 /* float32 [B,3,H,W], the synthesizer's raw output (~[-1,1]) */
#define MB_OUT_F32_NCHW_01 1 /* float32 [B,3,H,W] = clamp((x+1)/2, 0, 1): what MauaGenerator.render yields (wrappers/__init__.py:93) */
#define MB_OUT_U8_NHWC 2  /* uint8 [B,H,W,3] = round(clamp((x+1)/2,0,1)*255): the
                             tensor2bytes() wire format of maua/ops/io.py:47-70 */
#define MB_OUT_F32_NCHW_UNIT 3 /* float32 [B,3,H,W] = (x+1)/2, NOT clamped: what the FFMPEG / MemMap renderers hand to the
                                  patch's postprocess before tensor2bytes clamps (render/ffmpeg.py:72-73, memmap.py:30) */

/* One synthesis forward for `batch` frames.
 *   ws        device float32 [batch, num_ws, w_dim]   (W+ latents, stylegan3.py:51 `latents`)
 *   transform device float32 [3,3] or NULL            (G_synth.input.transform, stylegan3.py:59); see also
 *             mb_net_forward_xf for one matrix per frame
 *   out       device buffer of out_fmt
 *   workspace device buffer of >= mb_net_workspace_bytes(net, batch) bytes, 1024-aligned
 * Enqueues on `stream` and returns; no allocation, no host sync. */
int mb_net_forward(mb_net* net, const float* ws, const float* transform, int batch, void* out,
                   int out_fmt, void* workspace, size_t workspace_bytes, mb_stream stream);

/* mb_net_forward with ONE input transform per frame: transforms = device float32 [batch,3,3].  The reference's
 * make_transform_mat (stylegan3.py:82-93) squeezes its arguments and so only supports one translation / rotation per
 * call; this entry point lifts that limit for batched audio-reactive translation / rotation
 * tracks.  StyleGAN3 handles only. */
int mb_net_forward_xf(mb_net* net, const float* ws, const float* transforms, int batch, void* out,
                      int out_fmt, void* workspace, size_t workspace_bytes, mb_stream stream);

/* Output-size hook of the StyleGAN3 wrapper (maua/GAN/wrappers/stylegan3.py:62-117 change_output_resolution / get_hook):
 * the output of one synthesis module is resized and every later layer runs on the resized, possibly non-square map.
 *   module    0 = SynthesisInput, i = the layer layer_names[i-1] (the reference's `layer` argument); must feed another
 *             layer of the network (0 <= module <= num_layers; resizing the finished image is the caller's business)
 *   strategy  MB_RESIZE_NONE removes the hook;
 *             MB_RESIZE_STRETCH: a, b = target height, width; bicubic, align_corners=False (stylegan3.py:101-104);
 *             MB_RESIZE_PAD_ZERO: a, b = rows / columns of zeros added on EACH side, negative crops (stylegan3.py:108-115)
 * Changes mb_net_workspace_bytes and the output shape (mb_net_output_shape).  StyleGAN3 handles only. */
#define MB_RESIZE_NONE 0
#define MB_RESIZE_STRETCH 1
#define MB_RESIZE_PAD_ZERO 2
int mb_net_set_resize(mb_net* net, int module, int strategy, int a, int b);
/* Host-only (no device needed): the image size a network of this configuration produces under such a hook. */
int mb_sg3_resized_output(const mb_sg3_cfg* cfg, int module, int strategy, int a, int b, int32_t* height, int32_t* width);
/* Height and width of the image mb_net_forward writes (img_resolution unless a resize hook is set). */
int mb_net_output_shape(const mb_net* net, int32_t* height, int32_t* width);

/* Network-bending feature-map warps of StyleGAN2Synthesizer (maua/GAN/wrappers/stylegan2.py:65-80,153-194: forward hooks
 * running kornia translate / scale / rotate with padding_mode="reflection" on the output of layer_names[layer]).
 * Every forward that follows applies warp i to the activation of `layers[i]` (hook order = array order) until
 * n_warps = 0 clears them.  kornia is an absent, un-pinned dependency (setup.py:59): its published warp_affine
 * (affine_grid + bilinear grid_sample, align_corners=True) is what the kernel restates.
 *   layers    host int32 [n_warps], index into the wrapper's layer_names (0..2*blocks-1; 0 and 1 both name bs.0.conv1)
 *   inv_mats  device float32 [n_warps, batch, 2, 3]: destination pixel (x, y, 1) -> source pixel, i.e. the inverse of
 *             the 2x3 matrix kornia builds; caller-owned, must stay valid while the warps are set
 * Changes mb_net_workspace_bytes.  StyleGAN2 handles only. */
int mb_sg2_set_warps(mb_net* net, int n_warps, const int32_t* layers, const float* inv_mats, int batch);

/* Output-size hook of StyleGAN2Synthesizer.change_output_resolution (maua/GAN/wrappers/stylegan2.py:104-151, get_hook :216-340):
 * the output of layer_names[layer] is resized to target_h x target_w and every later layer runs on the resized, possibly
 * non-square map; the hooked block's ToRGB output is mapped back to the layer size before it joins the skip image and the
 * block's image is resized like the features (the reference's rgb_hook / img_hook).
 *   layer   index into the wrapper's layer_names; 0 = the reference's forward PRE-hook on bs.0.conv1: `noise` then IS the
 *           resized constant input [C, target_h, target_w] (resize + noise done by the caller), no image hooks; -1 clears
 *   mode    0 "stretch" (bicubic, align_corners=False), 1 constant pad with `value`, 2 reflect, 3 replicate, 4 circular
 *           (torch.nn.functional.pad semantics; pad_top / pad_left = leading pads, the rest trails)
 *   noise   device float32 [C, target_h, target_w] added to the resized features (the reference draws it once per channel from
 *           N(mean_c, std_c) of the resized features, :236-249) or NULL; caller-owned, must stay valid while the hook is set
 *   stats   device float32 [2, C] or NULL: when set, forwards write the per-channel mean / unbiased std of the resized features
 *           there (read them after a probe forward with noise = NULL to draw the noise map)
 * Every layer behind the hook needs a noise_const of its new size (mb_net_set_param accepts any [..., h, w] noise map; the
 * reference swaps in fresh randn maps, :137-147).  Changes mb_net_workspace_bytes and mb_net_output_shape.  StyleGAN2 only. */
int mb_sg2_set_resize(mb_net* net, int layer, int mode, int target_h, int target_w, int pad_top, int pad_left, float value,
                      const float* noise, float* stats);

/* Debug / parity aid: copy the activation a layer produced during the LAST forward into
 * `out` as float32 [B,C,H,W] (undoing the style pre-multiplication is the caller's business:
 * what is stored is x * style_{next}).  idx = -1 is the SynthesisInput output. */
int mb_net_read_activation(mb_net* net, int idx, int batch, float* out, mb_stream stream);
/* Number of kernels the last forward launched (for bench.py's gpu_launches). */
int mb_net_last_launch_count(const mb_net* net);
/* 0 = tcgen05 tensor-core conv (product path), 1 = plain CUDA-core conv (bisecting aid for the
 * parity tests; never used by the host facade). */
int mb_net_set_conv_impl(mb_net* net, int impl);
/* Test / tuning knobs: "sg2_precise" (StyleGAN2 handles; 1 default: fp16 hi + lo operands and conv outputs = fp32-class
 * pixels, 0: plain fp16 operands, 3x fewer MACs), "conv_impl" (0|1), "conv_tile_w" (32|16), "flrelu_impl" (0 auto | 1 generic),
 * "debug_stop" (stop the forward after layer N; -1 = after the input layer; default: run all),
 * "profile" (0 off | 1 record per-launch CUDA events of the last forward | 2 accumulate over forwards),
 * "profile_reset" (drop accumulated records). */
int mb_net_set_option(mb_net* net, const char* key, int value);
/* Per-launch device times (CUDA events on the forward's stream) of the last forward run with option
 * "profile"=1; call after synchronising the stream.  kind: 0 styles, 1 input, 2 modulated conv,
 * 3 filtered_lrelu, 4 layout transpose, 5 torgb/output.  Returns the number of records (<= cap). */
int mb_net_profile_read(mb_net* net, float* ms, int32_t* kind, int32_t* layer, int cap);
/* Shape [C,H,W] of the activation mb_net_read_activation would return. */
int mb_net_activation_shape(const mb_net* net, int32_t* c, int32_t* h, int32_t* w);

/* ---- op-level entry points (parity tests call the same kernels the network uses) ---- */

/* modulated_conv2d, upstream networks_stylegan3.py modulated_conv2d (grouped-conv
 * formulation; in-tree twin maua/GAN/wrappers/inference/ops.py:146-186).
 *   x [B,Cin,H,W] f32, w [Cout,Cin,k,k] f32, s [B,Cin] f32 -> y [B,Cout,H+k-1,W+k-1] f32
 * padding = k-1, demodulate as given, input_gain scalar.  All device pointers.
 * impl: 0 product dispatch (tcgen05; cout-major tile, pixel-major tile for <= 64 couts with > 32 cins; narrow layers
 * keep their weight tiles resident in shared memory), 1 CUDA-core direct convolution (bisecting aid), 2 cout-major
 * with 16-wide pixel tiles, 3 cout-major with full 128-row weight tiles streamed per stage (first path),
 * 4 pixel-major tile for everything up to 128 couts, 7 cout-major tile everywhere with ONE activation patch load per
 * 64-channel chunk wherever the weights are resident (kw shifts through the B-descriptor start). */
int mb_modulated_conv2d(const float* x, const float* w, const float* s, float* y, int B, int Cin,
                        int Cout, int H, int W, int k, int demodulate, float input_gain,
                        int impl, mb_stream stream);

/* filtered_lrelu, upstream torch_utils/ops/filtered_lrelu.py (reference semantics
 * _filtered_lrelu_ref): bias -> zero-insert x`up` -> FIR fu (gain up^2) with padding
 * [px0,px1,py0,py1] -> leaky-relu(slope)*gain -> clamp -> FIR fd -> decimate x`down`.
 *   x [B,C,H,W] f32, fu [up_taps] (separable) or NULL, fd [down_taps] separable or
 *   [down_taps,down_taps] when fd_2d, b [C] or NULL -> y [B,C,Ho,Wo] f32 */
int mb_filtered_lrelu(const float* x, const float* fu, const float* fd, const float* b, float* y,
                      int B, int C, int H, int W, int up, int down, int up_taps, int down_taps,
                      int fd_2d, int px0, int px1, int py0, int py1, float gain, float slope,
                      float clamp, mb_stream stream);

/* ---- audio features ----------------------------------------------------------------------------
 * Onset envelope + RMS of the torch-native feature path, one fused pass on the device:
 *   onsets(audio, sr) = normalize(onset_strength(percussive(audio), sr))     features/audio.py:20-28
 *   rms(audio, sr)                                                            features/audio.py:31-37
 * (maua/audiovisual/audioreactive/selfsupervised/features/; n_fft 2048, hop 1024, HPSS 31-tap medians,
 * margin 8, 128 Slaney mel bands up to 11025 Hz, top_db 80) and the strict-local-maximum peak rule of
 * audioreactive/signal.py:69-76.
 *   audio          device float32 [n], n a multiple of 1024 (the path resamples to sr = 1024*fps,
 *                  selfsupervised/sample.py:29-30, so one hop = one video frame), n >= 16384
 *   mel_filterbank device float32 [128,1025] (host-designed: depends on sr)
 *   onsets, rms    device float32 [T], T = n / 1024;  peak_idx device int32 [T]; n_peaks device int32 [1]
 *   percussive_out device float32 [n] or NULL (the HPSS percussive signal, for parity tests)
 */
size_t mb_audio_workspace_bytes(int64_t n_samples);
int mb_audio_onsets_rms(const float* audio, int64_t n, const float* mel_filterbank, float margin,
                        float* onsets, float* rms, int32_t* peak_idx, int32_t* n_peaks,
                        float* percussive_out, void* workspace, size_t workspace_bytes, mb_stream stream);

/* harmonic(audio, margin) (which = 0) / percussive(audio, margin) (which = 1), features/audio.py:13-24:
 * STFT -> HPSS soft mask -> iSTFT, device float32 [n] -> [n].  Workspace: mb_audio_workspace_bytes(n). */
int mb_audio_hpss_component(const float* audio, int64_t n, float margin, int which, float* out,
                            void* workspace, size_t workspace_bytes, mb_stream stream);

/* Spectral descriptors of the torch-native feature list (selfsupervised/features/audio.py:59-133, the AFEATFNS of
 * selfsupervised/mir.py:9): the device magnitude spectrogram [T][1025] (spectrogram(y), power 1, last STFT column dropped)
 * and mel power spectrogram [T][128] (melspectrogram(y, sr)), frame-major, and on top of them
 *   mb_spectral_flatness  exp(mean(log(max(amin, S^power)))) / mean(max(amin, S^power)) per frame            (:123-133)
 *   mb_spectral_contrast  per octave band, power_to_db(mean of the top quantile) - power_to_db(mean of the bottom
 *                         quantile); bands as bin ranges [lo, hi) with `cnt` bins per quantile, designed on the host
 *                         exactly as the reference's loop (:88-113); scratch: device float32 [2 * n_bands * T]       (:69-120)
 *   mb_mfcc               power_to_db (global top_db 80) of the mel power spectrogram (in place), orthonormal DCT-II over
 *                         the mel bands, first n_mfcc coefficients -> [T][n_mfcc]                                    (:59-64)
 * Workspace of mb_audio_spectrogram: mb_audio_workspace_bytes(n). */
int mb_audio_spectrogram(const float* audio, int64_t n, const float* mel_filterbank, float* mag_out, float* mel_out,
                         void* workspace, size_t workspace_bytes, mb_stream stream);
int mb_spectral_flatness(const float* mag, int T, float amin, float power, float* out, mb_stream stream);
int mb_spectral_contrast(const float* mag, int T, int n_bands, const int32_t* lo, const int32_t* hi, const int32_t* cnt,
                         int linear, float* scratch, float* out, mb_stream stream);
int mb_mfcc(float* mel, int T, int n_mfcc, float* out, mb_stream stream);

/* Constant-Q chroma, rosa/spectral.py:286-325 chroma_cqt over rosa/constantq.py:13-116 cqt (recursive octaves:
 * kaiser-sinc decimation by 2, rectangular STFT, sparse FFT-domain filter bank) and rosa/convert.py:69-117.
 *   audio        device float32 [n], n a multiple of hop; hop a multiple of 2^(n_octaves-1)
 *   decim_kernel device float32 [decim_taps]: torchaudio sinc_interp_kaiser 2:1 kernel (host-designed), left pad
 *                decim_width
 *   basis_*      CSR of the per-octave one-sided FFT basis [bins_per_octave, n_fft/2+1] (complex64 values as float
 *                pairs), the octave-0 matrix; octave i uses it scaled by sqrt(2^i) (constantq.py:98)
 *   inv_sqrt_len device float32 [n_bins] = 1/sqrt(filter length) (constantq.py:105-113)
 *   fold         device float32 [n_chroma, n_bins] (cq_to_chroma)
 *   cqt_mag      device float32 [n_bins, T] or NULL (|CQT|, for parity tests); chroma device float32 [n_chroma, T]
 */
size_t mb_chroma_workspace_bytes(int64_t n_samples, int n_bins, int hop);
int mb_chroma_cqt(const float* audio, int64_t n, int hop, int n_fft, int n_octaves, int bins_per_octave,
                  const float* decim_kernel, int decim_taps, int decim_width, const int32_t* basis_rowptr,
                  const int32_t* basis_col, const float* basis_val, const float* inv_sqrt_len, const float* fold,
                  int n_chroma, float threshold, int normalize, float* cqt_mag, float* chroma, void* workspace,
                  size_t workspace_bytes, mb_stream stream);

/* estimate_tuning(y, sr, bins_per_octave=...) of rosa/pitch.py:9-120 (piptrack with the reference defaults: n_fft 2048,
 * hop 512, 150..4000 Hz, threshold 0.1; median magnitude gate; histogram peak of the pitch residuals) -> tuning in
 * fractions of a bin, device float32 [1].  audio: device float32 [n], n a multiple of 512. */
size_t mb_tuning_workspace_bytes(int64_t n_samples);
int mb_estimate_tuning(const float* audio, int64_t n, float sr, int bins_per_octave, int bins /* ceil(1 / resolution) */, float* tuning,
                       void* workspace, size_t workspace_bytes, mb_stream stream);

/* chroma_cens of rosa/spectral.py:239-280 after chroma_cqt(norm=False): L1 normalisation per frame, the smooth 4-step
 * quantiser (natural cubic spline through host-designed knots, coef = [a | b | c | d] per interval, then the smooth
 * step function), win_len-tap temporal smoothing (zero-padded 'same'), L2 normalisation per frame.
 * chroma_raw / scratch / out: device float32 [n_chroma, T]. */
int mb_chroma_cens_post(const float* chroma_raw, int n_chroma, int T, const float* knots_x, const float* coef, int n_knots,
                        const float* smooth_win, int win_len, float* scratch, float* out, mb_stream stream);

/* ---- envelope post-ops and latent sequencers (device float32, [T, C] row-major, T = frames) ----------
 * maua/audiovisual/audioreactive/signal.py: gaussian_filter :108-157 (circular padding, optional causal
 * half-kernel factor), normalize :27-38 (eps 0) / processing.py:53-56 (eps 1e-8), resample :5-24 (linear,
 * align_corners False); latent.py: single_weighted :12-17, multi_weighted :21-31. */
int mb_gaussian_filter(const float* x, float* y, int T, int C, float sigma, int causal_mode, float causal, mb_stream stream);
int mb_normalize(const float* x, float* y, int64_t n, float eps, float* scratch2 /* device float[2] */, mb_stream stream);
int mb_resample_linear(const float* x, float* y, int T, int S, int C, mb_stream stream);
/* Op-level entry points of the in-tree inference network's resampling / activation ops (maua/GAN/wrappers/inference/ops.py):
 * upfirdn2d :87-114 -- zero insertion x up, padding [px0, px1, py0, py1] (negative = crop), correlation with the 2-D filter
 * f [fh, fw] scaled by gain (not flipped), decimation by down: float32 [B,C,H,W] -> [B,C,Ho,Wo],
 * Ho = (H*up + py0 + py1 - fh) / down + 1; bias_act :65-84 -- x + b[c], act (0 linear, 1 leaky ReLU with slope alpha),
 * * gain, clamp to [-clamp, clamp] when clamp >= 0.  (Inside mb_net_forward both are fused into one kernel between the convs.) */
int mb_upfirdn2d(const float* x, const float* f, float* y, int B, int C, int H, int W, int fh, int fw, int up, int down, int px0,
                 int px1, int py0, int py1, float gain, mb_stream stream);
int mb_bias_act(const float* x, const float* b /* [C] or NULL */, float* y, int B, int C, int H, int W, int act, float alpha,
                float gain, float clamp, mb_stream stream);
/* tensor2bytes (maua/ops/io.py:47-70) without the host copy: float32 [B,C,H,W] in [lo, hi] ->
 * uint8 [B,H,W,C] = round(clamp((x - lo) / (hi - lo), 0, 1) * 255), one pass (the reference chains
 * permute / clamp / sub / div / mul / round / byte: seven elementwise kernels). */
int mb_frames_to_rgb24(const float* x, uint8_t* out, int B, int C, int H, int W, float lo, float hi, mb_stream stream);
/* scipy.signal.sosfilt(sos, x) as the reference's low_pass / high_pass / band_pass call it (audioreactive/audio.py:96-110):
 * cascade of second-order sections [b0 b1 b2 a0 a1 a2] (HOST array, n_sections x 6) over a device float64 signal, zero
 * initial state, double precision.  Chunk-parallel (zero-state pass, state chain through A^256, apply pass); y may alias x.
 * scratch: device double[4 * ceil(n / 256)]. */
int mb_sosfilt(const double* x, double* y, int64_t n, const double* sos_host, int n_sections, double* scratch, mb_stream stream);
/* quantile(tensor, q) of selfsupervised/features/efficient_quantile/__init__.py:6-7 (the reference's compiled
 * efficient_quantile.cpp:86-208, method 3 "mid point", NaNs ignored, q taken as float32): device float32 [n] -> out[0].
 * Radix select on the device; n == 0 or all-NaN gives NaN. */
int mb_quantile_mid(const float* x, int64_t n, float q, float* out /* device float[1] */, mb_stream stream);
int mb_multi_weighted(const float* latents /*[K,D]*/, const float* envelopes /*[T,A]*/, float* out /*[T,D]*/, int T, int A,
                      int K, int D, mb_stream stream);
int mb_single_weighted(const float* low /*[D]*/, const float* high /*[D]*/, const float* envelope /*[T]*/, float* out /*[T,D]*/,
                       int T, int D, mb_stream stream);

/* Twins of the torch-native feature path (maua/audiovisual/audioreactive/selfsupervised/):
 * features/processing.py:11-50 gaussian_filter with mode="reflect" (pad_mode 1; 0 = circular = mb_gaussian_filter);
 * mir.py:13-21 salience_weighted's final (short / long)^2 * envelope; latent.py:69-78 the merge step of latent_patch on
 * layers [lay0, lay1) of latents [T, L, D] in place (mode 0 average, 1 modulate by modulation[T], 2 overwrite);
 * latent.py:7-13 spline_loop_latents: natural cubic spline through cat(keys, keys[0]) evaluated at
 * linspace(0, n_loops, size) % 1 (fractional n_loops allowed; workspace: device float32 [(K+1) * C]). */
int mb_gaussian_filter_ex(const float* x, float* y, int T, int C, float sigma, int causal_mode, float causal, int pad_mode, mb_stream stream);
int mb_salience(const float* short_env, const float* long_env, const float* envelope, float* out, int64_t n, mb_stream stream);
int mb_latent_merge(float* latents, const float* sequence, const float* modulation, int mode, int lay0, int lay1, int T, int L,
                    int D, mb_stream stream);
int mb_spline_loop_latents(const float* keys, int K, int C, float n_loops, int size, float* out, float* workspace, mb_stream stream);

/* latent.py: slerp_loops :68-80 = mb_slerp_rows (the [steps * K*n_loops, L, D] table of spherical interpolants between
 * consecutive looped keys, row = step * nseg + segment as the reference's reshape orders them) followed by
 * mb_resample_linear to `size`; spline_loops :83-92 = natural cubic spline through cat([keys] * n_loops + [keys[0]]) at
 * uniform knots, evaluated at linspace(0, 1, size) (workspace: device float32 [(K*n_loops+1) * C]); select_modulo :34-45
 * up to its final gaussian_filter (sorted_envelope = ascending sort of envelope). keys [K, C] row-major. */
int mb_slerp_rows(const float* keys, int K, int L, int D, int n_loops, int steps, float* rows, mb_stream stream);
int mb_spline_loops(const float* keys, int K, int C, int n_loops, int size, float* out, float* workspace, mb_stream stream);
int mb_select_modulo(const float* envelope, const float* sorted_envelope, int T, const float* keys, int K, int C,
                     float* out, mb_stream stream);

/* selfsupervised/noise.py: Blend :11-25 (two_sided = 1, noise [2, M, P]) / Multiply :28-40 (two_sided = 0, noise [M, P]) with
 * modulator rows [B, M] -> out [B, P], P = H * W; Loop :43-54 (idx [B] = phases of the batch, noise [3, P]); the
 * Average / Modulate / ScaleBias combinators :57-86 as out = a x mx[b] + c y (my[b] | 1 - my[b]) + bias (y, mx, my may be NULL). */
int mb_noise_mix(const float* noise, const float* modulator, int B, int M, int P, int two_sided, float* out, mb_stream stream);
int mb_noise_loop(const float* idx, const float* noise, int B, int P, float sigma, float* out, mb_stream stream);
int mb_noise_combine(const float* x, const float* y, const float* mx, const float* my, int one_minus, float a, float c,
                     float bias, int B, int P, float* out, mb_stream stream);

/* ---- image-space helpers (planar float32 [planes, h, w]) -----------------------------------------------------
 * mb_resize_bicubic: F.interpolate(mode="bicubic", align_corners=...) as maua/ops/image.py:240 (resample, True) and
 *   maua/GAN/wrappers/stylegan2.py:205 (make_noise_pyramid, False) call it;
 * mb_fir_reflect: the separable Lanczos prefilter of resample (image.py:228-238): 1-D FIR along H (axis 0) or W (axis 1),
 *   reflect padding (n_taps - 1) / 2;
 * mb_std_normalize: x[s] /= std(x[s]) (unbiased), stylegan2.py:212;
 * mb_perlin_noise: maua/ops/noise.py:27-88 for host-drawn unit gradients [r0+1, r1+1, r2+1, 3] -> [s0, s1, s2] in [-1, 1]. */
int mb_resize_bicubic(const float* x, float* y, int planes, int h, int w, int out_h, int out_w, int align_corners, mb_stream stream);
int mb_fir_reflect(const float* x, float* y, int planes, int h, int w, const float* taps, int n_taps, int axis, mb_stream stream);
int mb_std_normalize(float* x, int samples, int64_t per_sample, mb_stream stream);
int mb_perlin_noise(const float* gradients, int s0, int s1, int s2, int r0, int r1, int r2, float* out, mb_stream stream);

/* ---- RRDBNet: the RealESRGAN x4 generator (BASELINE configs[4]) ------------------------------------------------
 * Replaces basicsr.archs.rrdbnet_arch.RRDBNet(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=23 | 6, num_grow_ch=32, scale=4)
 * as maua/super/image/models/realesrgan.py:22-41 builds it and RealESRGANer(half=True).model runs it (:44-49).  basicsr /
 * realesrgan are absent third-party packages: the published architecture is restated (oracle/rrdb.py), parity unpinned.
 * Parameter names = the checkpoint's state-dict keys: conv_first.{weight,bias}, body.<i>.rdb<1..3>.conv<1..5>.{weight,bias},
 * conv_body, conv_up1, conv_up2, conv_hr, conv_last.  All data pointers are device pointers. */
typedef struct mb_rrdb mb_rrdb;
int mb_rrdb_create(int num_in_ch, int num_out_ch, int num_feat, int num_block, int num_grow_ch, int scale, mb_rrdb** out);
void mb_rrdb_destroy(mb_rrdb* net);
int mb_rrdb_set_param(mb_rrdb* net, const char* name, const float* data, const int64_t* shape, int ndim, mb_stream stream);
int mb_rrdb_finalize(mb_rrdb* net, mb_stream stream);                    /* packs the fp16 weight tiles; synchronises */
size_t mb_rrdb_workspace_bytes(const mb_rrdb* net, int batch, int height, int width);
/* x: float32 NCHW in [0, 1] (in_fmt MB_OUT_F32_NCHW) or uint8 NHWC (MB_OUT_U8_NHWC), [batch, C, height, width];
 * out: 4 * height x 4 * width as float32 NCHW (raw: MB_OUT_F32_NCHW, clamped to [0, 1]: MB_OUT_F32_NCHW_01) or uint8 NHWC
 * (clamp, * 255, round: MB_OUT_U8_NHWC).  workspace: caller-owned, 1024-byte aligned, mb_rrdb_workspace_bytes big. */
int mb_rrdb_forward(mb_rrdb* net, const void* x, int in_fmt, int batch, int height, int width, void* out, int out_fmt,
                    void* workspace, size_t workspace_bytes, mb_stream stream);
int mb_rrdb_last_launch_count(const mb_rrdb* net);

#ifdef __cplusplus
}
#endif
#endif /* MAUA_B200_H */
