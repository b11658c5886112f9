// StyleGAN2 network object behind the mb_net handle (sg2.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <functional>

namespace mb {
struct Sg2Net;
int sg2_create(int w_dim, int img_resolution, int img_channels, int channel_base, int channel_max, Sg2Net** out);
void sg2_destroy(Sg2Net* n);
int sg2_set_param(Sg2Net* n, const char* name, const float* data, const int64_t* shape, int ndim, cudaStream_t stream);
int sg2_finalize(Sg2Net* n, cudaStream_t stream);
size_t sg2_workspace_bytes(const Sg2Net* n, int B);
int sg2_num_ws(const Sg2Net* n);
int sg2_resolution(const Sg2Net* n);
int sg2_set_warps(Sg2Net* n, int n_warps, const int32_t* layers, const float* inv_mats, int batch);
// output-size hook (maua/GAN/wrappers/stylegan2.py:104-151): see mb_sg2_set_resize in include/maua_b200.h
int sg2_set_resize(Sg2Net* n, int layer, int mode, int th, int tw, int pad_t, int pad_l, float value, const float* noise, float* stats);
void sg2_output_hw(const Sg2Net* n, int* h, int* w);
int sg2_last_launches(const Sg2Net* n);
void sg2_set_conv_impl(Sg2Net* n, int impl);
void sg2_set_precise(Sg2Net* n, int on);   // 1 (default): fp16 hi + lo operands and conv outputs; 0: plain fp16
// mark(kind, layer): called after every launch group (kind: 0 styles, 1 constant input, 2 modulated conv, 3 fused FIR /
// noise / bias_act / ToRGB kernel, 4 feature warp, 5 skip-image upsample and output conversion) for the per-launch timing
int sg2_forward(Sg2Net* n, const float* ws, int B, void* out, int out_fmt, void* workspace, size_t workspace_bytes,
                int num_sms, cudaStream_t stream, const std::function<void(int, int)>& mark = nullptr);
}  // namespace mb
