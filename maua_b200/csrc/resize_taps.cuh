// Bicubic taps of torch's upsample_bicubic2d (A = -0.75, align_corners=False, taps clamped to the image): shared by the
// output-size hooks of StyleGAN3 (feature_resize.cu) and StyleGAN2 (sg2.cu).
#pragma once

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

struct Taps {
    int iy[4], ix[4];
    float wy[4], wx[4];
};

__device__ __forceinline__ Taps make_taps(int oy, int ox, int h, int w, float sh, float sw) {
    Taps t;
    const float ry = sh * (static_cast<float>(oy) + 0.5f) - 0.5f;
    const float rx = sw * (static_cast<float>(ox) + 0.5f) - 0.5f;
    const float fy = floorf(ry), fx = floorf(rx);
    cubic_w(ry - fy, t.wy);
    cubic_w(rx - fx, t.wx);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        t.iy[j] = min(max(static_cast<int>(fy) - 1 + j, 0), h - 1);
        t.ix[j] = min(max(static_cast<int>(fx) - 1 + j, 0), w - 1);
    }
    return t;
}

}  // namespace
}  // namespace mb
