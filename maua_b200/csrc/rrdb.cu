// RRDBNet (the generator of RealESRGAN x4plus / x4plus-anime) on the stacked pixel-major tcgen05 conv tile: SURVEY §8f row N6,
// BASELINE configs[4] ("StyleGAN2 512^2 + RealESRGAN 4x fused upscale pipeline").
//
// The reference builds basicsr.archs.rrdbnet_arch.RRDBNet(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=23 | 6,
// num_grow_ch=32, scale=4) and runs it through RealESRGANer(half=True) (maua/super/image/models/realesrgan.py:22-49).  basicsr
// and realesrgan are absent third-party packages (maua/submodules/RealESRGAN is an empty submodule): their published
// architecture is what this file computes, restated in oracle/rrdb.py.
//
//     feat = conv_first(x)                                                     3 -> 64
//     body: num_block x RRDB, RRDB = 3 x ResidualDenseBlock, out = rdb3(rdb2(rdb1(x))) * 0.2 + x
//           RDB: x1 = lrelu(conv1(x)); x2 = lrelu(conv2(cat(x, x1))); ... x5 = conv5(cat(x, x1..x4)); out = x5 * 0.2 + x
//     feat = feat + conv_body(body(feat))
//     feat = lrelu(conv_up1(nearest2x(feat))); feat = lrelu(conv_up2(nearest2x(feat)))
//     out  = conv_last(lrelu(conv_hr(feat)))                                   64 -> 3, 4H x 4W
// Every conv is 3x3 'same' with bias; lrelu slope 0.2.  At 512^2 the body is 345 convs = 8.7 TFLOP per frame.
//
// Layout: activations are channels-last fp16.  A dense block lives in ONE [B][H][W][192] buffer: channels 0..63 hold the block
// input, conv k writes its 32 growth channels at 64 + 32 (k - 1) and reads the first 64 + 32 (k - 1) channels through the same
// TMA tensor map (the K loop stops at the real channel count), so torch.cat never happens.  conv5 (64 couts) runs as two
// 32-cout launches whose epilogues fuse the residuals (x5 * 0.2 + x, and for the block's third RDB the RRDB residual on top)
// and write the next block's input into a second buffer; three buffers rotate so the RRDB input stays alive.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "../../include/maua_b200.h"

namespace mb {
namespace {

struct RConv {
    int cin = 0, cout = 0;
    float* w = nullptr;      // [cout][cin][3][3] f32
    float* b = nullptr;      // [cout] f32
    __half* wpk[2] = {nullptr, nullptr};   // packed weights (two 32-cout halves for the 64-cout convs of a dense block)
    bool w_set = false, b_set = false;
};

inline int rgrid(long long total) {
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    return g < 1 ? 1 : static_cast<int>(g);
}

// image in -> channels-last fp16 [B][H][W][16] (3 real channels, 13 zeros: one K = 16 step of the first conv)
__global__ void rrdb_in_kernel(const void* __restrict__ x, int fmt, __half* __restrict__ y, int B, int C, int H, int W) {
    const long long total = static_cast<long long>(B) * H * W;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long px = idx % (static_cast<long long>(H) * W);
        const int b = static_cast<int>(idx / (static_cast<long long>(H) * W));
        __half v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __float2half_rn(0.0f);
        for (int c = 0; c < C; ++c) {
            const float f = fmt == MB_OUT_U8_NHWC ? static_cast<const uint8_t*>(x)[idx * C + c] * (1.0f / 255.0f)
                                                  : static_cast<const float*>(x)[(static_cast<long long>(b) * C + c) * H * W + px];
            v[c] = __float2half_rn(f);
        }
        uint4* o = reinterpret_cast<uint4*>(y + idx * 16);
        o[0] = *reinterpret_cast<const uint4*>(&v[0]);
        o[1] = *reinterpret_cast<const uint4*>(&v[8]);
    }
}

// F.interpolate(scale_factor=2, mode="nearest") on channels-last fp16 [B][H][W][Cp] -> [B][2H][2W][Cp]; one thread = 8 channels
__global__ void rrdb_up2_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W, int Cp) {
    const int g8 = Cp / 8;
    const long long total = static_cast<long long>(B) * 2 * H * 2 * W * g8;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(idx % g8);
        long long q = idx / g8;
        const int ox = static_cast<int>(q % (2 * W)); q /= 2 * W;
        const int oy = static_cast<int>(q % (2 * H));
        const int b = static_cast<int>(q / (2 * H));
        *reinterpret_cast<uint4*>(y + idx * 8) =
            *reinterpret_cast<const uint4*>(x + ((static_cast<long long>(b) * H + (oy >> 1)) * W + (ox >> 1)) * Cp + g * 8);
    }
}

// channels-last fp16 [B][H][W][cp] (first C channels) -> float32 NCHW (raw) or uint8 NHWC (clamp to [0, 1], * 255, round)
__global__ void rrdb_out_kernel(const __half* __restrict__ x, void* __restrict__ out, int fmt, int B, int C, int H, int W, int cp) {
    const long long total = static_cast<long long>(B) * H * W * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        const long long p = idx / C;   // (b, h, w) pixel index
        const float v = __half2float(x[p * cp + c]);
        if (fmt == MB_OUT_U8_NHWC) {
            static_cast<uint8_t*>(out)[idx] = static_cast<uint8_t>(rintf(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f));
        } else {
            const long long px = p % (static_cast<long long>(H) * W);
            const int b = static_cast<int>(p / (static_cast<long long>(H) * W));
            static_cast<float*>(out)[(static_cast<long long>(b) * C + c) * H * W + px] = fmt == MB_OUT_F32_NCHW_01 ? fminf(fmaxf(v, 0.0f), 1.0f) : v;
        }
    }
}

}  // namespace
}  // namespace mb

using namespace mb;

struct mb_rrdb {
    int in_ch = 3, out_ch = 3, feat = 64, blocks = 23, grow = 32, scale = 4;
    RConv conv_first, conv_body, conv_up1, conv_up2, conv_hr, conv_last;
    std::vector<RConv> dense;   // [blocks][3][5]
    std::map<std::string, std::pair<RConv*, int>> by_name;   // name -> (conv, 0 weight / 1 bias)
    bool finalized = false;
    int num_sms = 148;
    int last_launches = 0;
    int cm_stack = 0;   // MB_RRDB_CMS=1: the <= 32-cout convs run the stacked cout-major tile (conv_cms_kernel)
};

static void rconv_free(RConv& c) {
    if (c.w) cudaFree(c.w);
    if (c.b) cudaFree(c.b);
    for (auto& p : c.wpk)
        if (p) cudaFree(p);
    c = RConv();
}

extern "C" void mb_rrdb_destroy(mb_rrdb* n) {
    if (!n) return;
    for (RConv* c : {&n->conv_first, &n->conv_body, &n->conv_up1, &n->conv_up2, &n->conv_hr, &n->conv_last}) rconv_free(*c);
    for (auto& c : n->dense) rconv_free(c);
    delete n;
}

extern "C" int mb_rrdb_create(int num_in_ch, int num_out_ch, int num_feat, int num_block, int num_grow_ch, int scale, mb_rrdb** out) {
    MB_REQUIRE(out, "mb_rrdb_create: null argument");
    MB_REQUIRE(num_in_ch >= 1 && num_in_ch <= 4 && num_out_ch >= 1 && num_out_ch <= 4, "mb_rrdb_create: 1..4 image channels");
    MB_REQUIRE(num_feat == 64 && num_grow_ch == 32, "mb_rrdb_create: num_feat = 64 and num_grow_ch = 32 (every RealESRGAN RRDBNet the reference loads)");
    MB_REQUIRE(scale == 4, "mb_rrdb_create: scale 4 only (x4plus, x4plus-anime; scales 1 / 2 add a pixel-unshuffle in front)");
    MB_REQUIRE(num_block >= 1 && num_block <= 64, "mb_rrdb_create: num_block out of range");
    int dev = 0, sms = 148;
    MB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("mb_rrdb_create: device %d is sm_%d%d; libmaua_b200 is built for sm_100a only", dev, prop.major, prop.minor);
        return MB_ENODEV;
    }
    sms = prop.multiProcessorCount;
    mb_rrdb* n = new mb_rrdb();
    n->in_ch = num_in_ch; n->out_ch = num_out_ch; n->blocks = num_block; n->num_sms = sms;
    if (const char* e = getenv("MB_RRDB_CMS")) n->cm_stack = atoi(e);
    auto add = [&](RConv& c, const std::string& name, int cin, int cout) -> int {
        c.cin = cin; c.cout = cout;
        if (cudaMalloc(&c.w, sizeof(float) * cout * cin * 9) != cudaSuccess || cudaMalloc(&c.b, sizeof(float) * cout) != cudaSuccess) {
            set_error("mb_rrdb_create: cudaMalloc failed");
            return MB_ECUDA;
        }
        n->by_name[name + ".weight"] = {&c, 0};
        n->by_name[name + ".bias"] = {&c, 1};
        return MB_OK;
    };
    int rc = add(n->conv_first, "conv_first", num_in_ch, 64);
    n->dense.resize(static_cast<size_t>(num_block) * 15);
    for (int i = 0; i < num_block && rc == MB_OK; ++i)
        for (int r = 0; r < 3 && rc == MB_OK; ++r)
            for (int k = 0; k < 5 && rc == MB_OK; ++k)
                rc = add(n->dense[(i * 3 + r) * 5 + k], "body." + std::to_string(i) + ".rdb" + std::to_string(r + 1) + ".conv" + std::to_string(k + 1),
                         64 + 32 * k, k < 4 ? 32 : 64);
    if (rc == MB_OK) rc = add(n->conv_body, "conv_body", 64, 64);
    if (rc == MB_OK) rc = add(n->conv_up1, "conv_up1", 64, 64);
    if (rc == MB_OK) rc = add(n->conv_up2, "conv_up2", 64, 64);
    if (rc == MB_OK) rc = add(n->conv_hr, "conv_hr", 64, 64);
    if (rc == MB_OK) rc = add(n->conv_last, "conv_last", 64, num_out_ch);
    if (rc != MB_OK) {
        mb_rrdb_destroy(n);
        return rc;
    }
    *out = n;
    return MB_OK;
}

extern "C" int mb_rrdb_set_param(mb_rrdb* n, const char* name, const float* data, const int64_t* shape, int ndim, mb_stream stream_) {
    MB_REQUIRE(n && name && data && shape, "mb_rrdb_set_param: null argument");
    auto it = n->by_name.find(name);
    if (it == n->by_name.end()) {
        set_error("mb_rrdb_set_param: unknown RRDBNet parameter '%s'", name);
        return MB_EINVAL;
    }
    RConv& c = *it->second.first;
    size_t numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= static_cast<size_t>(shape[i]);
    const size_t want = it->second.second == 0 ? static_cast<size_t>(c.cout) * c.cin * 9 : static_cast<size_t>(c.cout);
    MB_REQUIRE(numel == want, "mb_rrdb_set_param: '%s' has %zu elements, expected %zu", name, numel, want);
    MB_CUDA(cudaMemcpyAsync(it->second.second == 0 ? c.w : c.b, data, sizeof(float) * numel, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream_)));
    (it->second.second == 0 ? c.w_set : c.b_set) = true;
    n->finalized = false;
    return MB_OK;
}

static int rconv_pack(RConv& c, cudaStream_t stream) {
    // convs of a dense block that produce 64 couts run as two 32-cout launches (their weights do not fit shared memory next
    // to two patch stages otherwise: 9 x 3 x 64 x 128 B = 221 KB at 192 input channels)
    const bool split = c.cout == 64 && c.cin > 64;
    const int parts = split ? 2 : 1, co = split ? 32 : c.cout;
    for (int p = 0; p < parts; ++p) {
        const size_t elems = packed_weight_elems(co, c.cin, 3);
        if (!c.wpk[p]) MB_CUDA(cudaMalloc(&c.wpk[p], elems * sizeof(__half)));
        int rc = pack_weights_launch(c.w + static_cast<size_t>(p) * co * c.cin * 9, c.wpk[p], nullptr, co, c.cin, 3, 0, stream, 0);
        if (rc != MB_OK) return rc;
    }
    return MB_OK;
}

extern "C" int mb_rrdb_finalize(mb_rrdb* n, mb_stream stream_) {
    MB_REQUIRE(n, "mb_rrdb_finalize: null argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    for (auto& kv : n->by_name) {
        const RConv& c = *kv.second.first;
        if (!(kv.second.second == 0 ? c.w_set : c.b_set)) {
            set_error("mb_rrdb_finalize: parameter '%s' was never set", kv.first.c_str());
            return MB_ESTATE;
        }
    }
    int rc;
    for (RConv* c : {&n->conv_first, &n->conv_body, &n->conv_up1, &n->conv_up2, &n->conv_hr, &n->conv_last})
        if ((rc = rconv_pack(*c, stream)) != MB_OK) return rc;
    for (auto& c : n->dense)
        if ((rc = rconv_pack(c, stream)) != MB_OK) return rc;
    MB_CUDA(cudaStreamSynchronize(stream));
    n->finalized = true;
    return MB_OK;
}

namespace {
struct RWs {
    size_t in16, feat0, d[3], f1, u1, f2, u2, f3, f4, last, total;
};
RWs rrdb_ws(int B, int H, int W) {
    RWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) / 1024 * 1024; return o; };
    const size_t px = static_cast<size_t>(B) * H * W;
    w.in16 = take(px * 16 * 2);
    w.feat0 = take(px * 64 * 2);
    for (auto& d : w.d) d = take(px * 192 * 2);
    w.f1 = take(px * 64 * 2);
    w.u1 = take(px * 4 * 64 * 2);
    w.f2 = take(px * 4 * 64 * 2);
    w.u2 = take(px * 16 * 64 * 2);
    w.f3 = take(px * 16 * 64 * 2);
    w.f4 = w.u2;            // conv_hr writes over the upsampled map conv_up2 has consumed
    w.last = take(px * 16 * 8 * 2);
    w.total = off;
    return w;
}
}  // namespace

extern "C" size_t mb_rrdb_workspace_bytes(const mb_rrdb* n, int batch, int height, int width) {
    if (!n || batch <= 0 || height <= 0 || width <= 0) return 0;
    return rrdb_ws(batch, height, width).total;
}

extern "C" int mb_rrdb_last_launch_count(const mb_rrdb* n) { return n ? n->last_launches : 0; }

// One 3x3 'same' conv with bias through the stacked pixel-major tile and its channels-last epilogue.
static int rrdb_conv(const mb_rrdb* n, const RConv& c, int part, const __half* x, int cp_in, int B, int H, int W, const ConvTcArgs::NhwcOut& o,
                     cudaStream_t stream) {
    const bool split = c.cout == 64 && c.cin > 64;
    ConvTcArgs ca;
    ca.x = x; ca.wpk = c.wpk[part]; ca.d = nullptr; ca.bias = c.b + (split ? part * 32 : 0); ca.y = o.y;
    ca.B = B; ca.Cin = c.cin < 16 ? 16 : c.cin; ca.Cout = split ? 32 : c.cout; ca.Hin = H; ca.Win = W; ca.Cp_in = cp_in;
    ca.Wp_out = pitch16(W); ca.ksz = 3; ca.pad = 1; ca.tile_w = 32;
    ca.pm_max_cout = 64; ca.pm_stack = 1; ca.cm_stack = n->cm_stack; ca.num_sms = n->num_sms;
    ca.nhwc = o;
    return conv_tc_launch(ca, stream);
}

extern "C" int mb_rrdb_forward(mb_rrdb* n, const void* x, int in_fmt, int B, int H, int W, void* out, int out_fmt, void* workspace,
                               size_t workspace_bytes, mb_stream stream_) {
    MB_REQUIRE(n && x && out && workspace, "mb_rrdb_forward: null argument");
    MB_REQUIRE(n->finalized, "mb_rrdb_forward: call mb_rrdb_finalize after setting parameters");
    MB_REQUIRE(in_fmt == MB_OUT_F32_NCHW || in_fmt == MB_OUT_U8_NHWC, "mb_rrdb_forward: input format must be float32 NCHW in [0, 1] or uint8 NHWC");
    MB_REQUIRE(out_fmt == MB_OUT_F32_NCHW || out_fmt == MB_OUT_F32_NCHW_01 || out_fmt == MB_OUT_U8_NHWC, "mb_rrdb_forward: unknown output format");
    MB_REQUIRE(B >= 1 && H >= 4 && W >= 4, "mb_rrdb_forward: empty input");
    const RWs wl = rrdb_ws(B, H, W);
    MB_REQUIRE(workspace_bytes >= wl.total, "mb_rrdb_forward: workspace too small (%zu < %zu bytes)", workspace_bytes, wl.total);
    MB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "mb_rrdb_forward: the workspace must be 1024-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    auto H16 = [&](size_t off) { return reinterpret_cast<__half*>(base + off); };
    int launches = 0, rc;
    const long long px = static_cast<long long>(B) * H * W;

    rrdb_in_kernel<<<rgrid(px), 256, 0, stream>>>(x, in_fmt, H16(wl.in16), B, n->in_ch, H, W);
    MB_CUDA(cudaGetLastError());
    ++launches;
    // conv_first -> feat0 (kept for the long skip) ... and a copy as the first dense block's input: written twice by the epilogue
    // being cheaper than a copy kernel is not worth a special case; the first RDB reads feat0 through r1 / its own buffer below
    ConvTcArgs::NhwcOut o;
    o.y = H16(wl.feat0); o.cp = 64; o.c_off = 0;
    if ((rc = rrdb_conv(n, n->conv_first, 0, H16(wl.in16), 16, B, H, W, o, stream)) != MB_OK) return rc;
    ++launches;
    __half* D[3] = {H16(wl.d[0]), H16(wl.d[1]), H16(wl.d[2])};
    // the first dense buffer's channels 0..63 = feat0 (strided copy: 128 of every 384 bytes)
    MB_CUDA(cudaMemcpy2DAsync(D[0], 192 * 2, H16(wl.feat0), 64 * 2, 64 * 2, static_cast<size_t>(px), cudaMemcpyDeviceToDevice, stream));
    int in = 0, fa = 1, fb = 2;   // dense buffer holding the RRDB input, and the two free ones
    for (int i = 0; i < n->blocks; ++i) {
        // rdb1: in -> fa, rdb2: fa -> fb, rdb3: fb -> fa (+ RRDB residual from `in`); the next RRDB's input is fa
        const int src[3] = {in, fa, fb}, dst[3] = {fa, fb, fa};
        for (int r = 0; r < 3; ++r) {
            const RConv* cv = &n->dense[(static_cast<size_t>(i) * 3 + r) * 5];
            __half* X = D[src[r]];
            for (int k = 0; k < 4; ++k) {
                ConvTcArgs::NhwcOut g;
                g.y = X; g.cp = 192; g.c_off = 64 + 32 * k; g.slope = 0.2f;
                if ((rc = rrdb_conv(n, cv[k], 0, X, 192, B, H, W, g, stream)) != MB_OK) return rc;
                ++launches;
            }
            for (int part = 0; part < 2; ++part) {
                ConvTcArgs::NhwcOut g;
                g.y = D[dst[r]]; g.cp = 192; g.c_off = 32 * part;
                g.r1 = X; g.r1_cp = 192; g.r1_off = 32 * part;
                if (r < 2) {
                    g.alpha = 0.2f; g.beta = 1.0f;                      // x5 * 0.2 + x
                } else {
                    g.alpha = 0.04f; g.beta = 0.2f;                     // (x5 * 0.2 + x) * 0.2 + rrdb_in
                    g.r2 = D[in]; g.r2_cp = 192; g.r2_off = 32 * part; g.gamma = 1.0f;
                }
                if ((rc = rrdb_conv(n, cv[4], part, X, 192, B, H, W, g, stream)) != MB_OK) return rc;
                ++launches;
            }
        }
        const int nin = fa;
        fa = in; in = nin;   // free buffers: the old input and fb
    }
    // feat = feat0 + conv_body(body)
    o = ConvTcArgs::NhwcOut();
    o.y = H16(wl.f1); o.cp = 64; o.r1 = H16(wl.feat0); o.r1_cp = 64; o.beta = 1.0f;
    if ((rc = rrdb_conv(n, n->conv_body, 0, D[in], 192, B, H, W, o, stream)) != MB_OK) return rc;
    ++launches;
    auto up2 = [&](const __half* src, __half* dstp, int h, int w) -> int {
        rrdb_up2_kernel<<<rgrid(static_cast<long long>(B) * 4 * h * w * 8), 256, 0, stream>>>(src, dstp, B, h, w, 64);
        MB_CUDA(cudaGetLastError());
        ++launches;
        return MB_OK;
    };
    if ((rc = up2(H16(wl.f1), H16(wl.u1), H, W)) != MB_OK) return rc;
    o = ConvTcArgs::NhwcOut();
    o.y = H16(wl.f2); o.cp = 64; o.slope = 0.2f;
    if ((rc = rrdb_conv(n, n->conv_up1, 0, H16(wl.u1), 64, B, 2 * H, 2 * W, o, stream)) != MB_OK) return rc;
    ++launches;
    if ((rc = up2(H16(wl.f2), H16(wl.u2), 2 * H, 2 * W)) != MB_OK) return rc;
    o.y = H16(wl.f3);
    if ((rc = rrdb_conv(n, n->conv_up2, 0, H16(wl.u2), 64, B, 4 * H, 4 * W, o, stream)) != MB_OK) return rc;
    ++launches;
    o.y = H16(wl.f4);
    if ((rc = rrdb_conv(n, n->conv_hr, 0, H16(wl.f3), 64, B, 4 * H, 4 * W, o, stream)) != MB_OK) return rc;
    ++launches;
    o = ConvTcArgs::NhwcOut();
    o.y = H16(wl.last); o.cp = 8;
    if ((rc = rrdb_conv(n, n->conv_last, 0, H16(wl.f4), 64, B, 4 * H, 4 * W, o, stream)) != MB_OK) return rc;
    ++launches;
    rrdb_out_kernel<<<rgrid(px * 16 * n->out_ch), 256, 0, stream>>>(H16(wl.last), out, out_fmt, B, n->out_ch, 4 * H, 4 * W, 8);
    MB_CUDA(cudaGetLastError());
    ++launches;
    n->last_launches = launches;
    return MB_OK;
}
