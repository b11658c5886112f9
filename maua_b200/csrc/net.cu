// libmaua_b200 C ABI: library init, the StyleGAN3 synthesis-network handle and its per-layer
// pipeline (styles -> input -> [modulated conv (tcgen05) -> filtered_lrelu] x N -> ToRGB/out),
// plus op-level entry points that run the very same kernels for the parity tests.
// Reference: maua/GAN/wrappers/stylegan3.py:33,51-60 (ctor + forward call sites); upstream
// training/networks_stylegan3.py SynthesisNetwork.__init__/forward for geometry and data flow.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <utility>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "kernels.h"
#include "sg2.h"

namespace mb {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

static int g_device = -1;
static int g_num_sms = 148;
static int* g_dbg_host = nullptr;
static int* g_dbg_dev = nullptr;
int* debug_words_device() { return g_dbg_dev; }


}  // namespace mb

using namespace mb;

struct Param {
    float* dev = nullptr;
    std::vector<int64_t> shape;
    size_t numel = 0;
    bool set = false;
};

struct LayerState {
    mb_sg3_layer g;
    Param weight, bias, affine_w, affine_b, magnitude_ema, up_filter, down_filter;
    __half* wpk = nullptr;
    float* wsqT = nullptr;
    std::vector<float> fu, fd;  // host copies
};

struct mb_net {
    mb::Sg2Net* sg2 = nullptr;  // set: this handle is a StyleGAN2 network (sg2.cu); the fields below are unused
    mb_sg3_cfg cfg;
    std::vector<LayerState> layers;  // num_layers + 1
    int in_channels = 0, in_size = 0;
    double in_sr = 0, in_bw = 0;
    Param in_freqs, in_phases, in_weight, in_affine_w, in_affine_b, in_transform;
    float* in_weightT = nullptr;
    float* in_w3 = nullptr;      // [C][3C] f32 hi/lo split of weight / sqrt(C) (tensor-core input layer)
    __half* in_wpk = nullptr;    // the same packed for the 1x1 tcgen05 conv
    int input_impl = 1;          // 1: split-fp16 tcgen05 channel mix, 0: fp32 CUDA-core kernel
    std::map<std::string, Param*> by_name;
    bool finalized = false;
    int conv_impl = 0;
    int conv_tile_w = 32;
    int conv_row_il = 0;      // conv outputs of layers with at most this many couts use the row-interleaved layout [B][H][C][Wp]
    int conv_cm_stack = 0;    // stacked cout-major conv tile for 3x3 layers with <= 32 couts (conv_cms_kernel)
    int conv_pm_stack = 1;    // stacked pixel-major conv tile for the narrow 3x3 layers (conv_pms_kernel)
    int conv_epi_groups = 0;  // cout-major conv tile: epilogue warp groups (0 = per layer, see conv_tc_launch)
    int conv_pm_max = 64;  // layers with ceil16(Cout) <= this (and Cin > 32) run the pixel-major conv tile
    int conv_narrow_a = 1; // narrow/resident weight tiles for Cout <= 128 (conv_tc.cu)
    int conv_cm_shift = 0; // cout-major tile with resident weights: one patch load per chunk (conv_tc.cu)
    int conv_pm_shift = 1; // pixel-major tile: one patch load per chunk, kw shift via the descriptor start (conv_tc.cu)
    int flrelu_impl = 0;  // 0 = tensor-core chain where supported, 1 = generic loops, 2 = CUDA-core polyphase kernel
    int debug_stop = 1 << 30;
    int last_launches = 0;
    // output-size hook (mb_net_set_resize): module 0 = input, i = layers[i-1]
    int resize_module = -1, resize_mode = MB_RESIZE_NONE, resize_a = 0, resize_b = 0;
    // optional per-launch timing (CUDA events on the forward's stream), see mb_net_profile_read
    int profile = 0;
    std::vector<cudaEvent_t> ev_pool;
    struct ProfRec { int kind, layer, ev0, ev1; };
    std::vector<ProfRec> prof;
    int ev_used = 0;
    // layout of the last forward (for read_activation)
    void* last_ws = nullptr;
    int last_batch = 0;
    const __half* last_act = nullptr;
    bool last_act_nhwc = false;
    int last_act_c = 0, last_act_h = 0, last_act_w = 0;
};

// ---------------------------------------------------------------------------------------
// geometry (host only)
// ---------------------------------------------------------------------------------------
extern "C" void mb_sg3_default_cfg(mb_sg3_cfg* c, int config_r) {
    memset(c, 0, sizeof(*c));
    c->w_dim = 512;
    c->img_resolution = 1024;
    c->img_channels = 3;
    c->channel_base = config_r ? 65536 : 32768;
    c->channel_max = config_r ? 1024 : 512;
    c->num_layers = 14;
    c->num_critical = 2;
    c->conv_kernel = config_r ? 1 : 3;
    c->filter_size = 6;
    c->lrelu_upsampling = 2;
    c->use_radial_filters = config_r ? 1 : 0;
    c->margin_size = 10;
    c->first_cutoff = 2.0;
    c->first_stopband = pow(2.0, 2.1);
    c->last_stopband_rel = pow(2.0, 0.3);
    c->output_scale = 0.25;
    c->conv_clamp = 256.0;
}

extern "C" int mb_sg3_geometry(const mb_sg3_cfg* c, mb_sg3_layer* L, int32_t* input_channels, int32_t* input_size,
                               double* input_sampling_rate, double* input_bandwidth) {
    MB_REQUIRE(c && L, "mb_sg3_geometry: null argument");
    const int N = c->num_layers;
    MB_REQUIRE(N >= 3 && N + 1 <= kMaxLayers && c->num_critical < N, "mb_sg3_geometry: num_layers %d unsupported", N);
    MB_REQUIRE(c->conv_kernel == 1 || c->conv_kernel == 3, "mb_sg3_geometry: conv_kernel must be 1 or 3");
    const double res = c->img_resolution;
    const double last_cutoff = res / 2;
    const double last_stopband = last_cutoff * c->last_stopband_rel;
    std::vector<double> cut(N + 1), stop(N + 1), sr(N + 1), hw(N + 1), size(N + 1), ch(N + 1);
    for (int i = 0; i <= N; ++i) {
        double e = static_cast<double>(i) / (N - c->num_critical);
        if (e > 1) e = 1;
        cut[i] = c->first_cutoff * pow(last_cutoff / c->first_cutoff, e);
        stop[i] = c->first_stopband * pow(last_stopband / c->first_stopband, e);
        double m = stop[i] * 2 < res ? stop[i] * 2 : res;
        sr[i] = exp2(ceil(log2(m)));
        hw[i] = (stop[i] > sr[i] / 2 ? stop[i] : sr[i] / 2) - cut[i];
        size[i] = sr[i] + c->margin_size * 2;
        double chv = (c->channel_base / 2.0) / cut[i];
        if (chv > c->channel_max) chv = c->channel_max;
        ch[i] = nearbyint(chv);
    }
    size[N] = res;
    size[N - 1] = res;
    ch[N] = c->img_channels;
    for (int idx = 0; idx <= N; ++idx) {
        const int prev = idx > 0 ? idx - 1 : 0;
        mb_sg3_layer& g = L[idx];
        memset(&g, 0, sizeof(g));
        g.idx = idx;
        g.is_torgb = idx == N;
        g.is_critically_sampled = idx >= N - c->num_critical;
        g.use_fp16 = sr[idx] * 16 > res;
        g.in_channels = static_cast<int>(ch[prev]);
        g.out_channels = static_cast<int>(ch[idx]);
        g.in_size = static_cast<int>(size[prev]);
        g.out_size = static_cast<int>(size[idx]);
        g.in_sampling_rate = static_cast<int>(sr[prev]);
        g.out_sampling_rate = static_cast<int>(sr[idx]);
        const int mx = g.in_sampling_rate > g.out_sampling_rate ? g.in_sampling_rate : g.out_sampling_rate;
        g.tmp_sampling_rate = mx * (g.is_torgb ? 1 : c->lrelu_upsampling);
        g.conv_kernel = g.is_torgb ? 1 : c->conv_kernel;
        g.up = static_cast<int>(nearbyint(static_cast<double>(g.tmp_sampling_rate) / g.in_sampling_rate));
        g.down = static_cast<int>(nearbyint(static_cast<double>(g.tmp_sampling_rate) / g.out_sampling_rate));
        g.up_taps = (g.up > 1 && !g.is_torgb) ? c->filter_size * g.up : 1;
        g.down_taps = (g.down > 1 && !g.is_torgb) ? c->filter_size * g.down : 1;
        g.down_radial = c->use_radial_filters && !g.is_critically_sampled;
        int pad_total = (g.out_size - 1) * g.down + 1;
        pad_total -= (g.in_size + g.conv_kernel - 1) * g.up;
        pad_total += g.up_taps + g.down_taps - 2;
        // python floor division
        int num = pad_total + g.up;
        int lo = num / 2;
        if ((num % 2 != 0) && (num < 0)) --lo;
        g.pad_lo = lo;
        g.pad_hi = pad_total - lo;
        g.in_cutoff = cut[prev];
        g.out_cutoff = cut[idx];
        g.in_half_width = hw[prev];
        g.out_half_width = hw[idx];
        snprintf(g.name, sizeof(g.name), "L%d_%d_%d", idx, g.out_size, g.out_channels);
    }
    if (input_channels) *input_channels = static_cast<int>(ch[0]);
    if (input_size) *input_size = static_cast<int>(size[0]);
    if (input_sampling_rate) *input_sampling_rate = sr[0];
    if (input_bandwidth) *input_bandwidth = cut[0];
    return MB_OK;
}

// ---------------------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------------------
extern "C" const char* mb_last_error(void) { return g_err.c_str(); }
extern "C" int mb_abi_version(void) { return 2; }

/* Debug words written by kernels into mapped host memory (enabled by MB_DEBUG=1 before mb_init). */
extern "C" int mb_debug_read(int* out, int n) {
    for (int i = 0; i < n && i < 64; ++i) out[i] = g_dbg_host ? g_dbg_host[i] : -1;
    return g_dbg_host ? MB_OK : MB_ESTATE;
}

extern "C" int mb_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("mb_init: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return MB_ENODEV;
    }
    MB_REQUIRE(device >= 0 && device < n, "mb_init: device %d out of range (%d devices)", device, n);
    cudaDeviceProp prop;
    MB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("mb_init: device %d is sm_%d%d; libmaua_b200 is built for sm_100a only", device, prop.major, prop.minor);
        return MB_ENODEV;
    }
    MB_CUDA(cudaSetDevice(device));
    g_device = device;
    g_num_sms = prop.multiProcessorCount;
    if (!g_dbg_host && getenv("MB_DEBUG")) {
        MB_CUDA(cudaHostAlloc(&g_dbg_host, 64 * sizeof(int), cudaHostAllocMapped));
        memset(g_dbg_host, 0, 64 * sizeof(int));
        MB_CUDA(cudaHostGetDevicePointer(&g_dbg_dev, g_dbg_host, 0));
    }
    return MB_OK;
}

// First use without mb_init(): adopt the thread's CURRENT device (the one the caller's PyTorch selected),
// never force device 0 -- one process per GPU, ranks > 0 must stay on their own device.
static int ensure_init() {
    int dev = 0;
    MB_CUDA(cudaGetDevice(&dev));
    return mb_init(dev);
}

static int alloc_param(Param& p, std::initializer_list<int64_t> shape) {
    p.shape.assign(shape.begin(), shape.end());
    p.numel = 1;
    for (int64_t s : p.shape) p.numel *= static_cast<size_t>(s);
    MB_CUDA(cudaMalloc(&p.dev, sizeof(float) * (p.numel ? p.numel : 1)));
    return MB_OK;
}

extern "C" int mb_sg3_create(const mb_sg3_cfg* cfg, mb_net** out) {
    MB_REQUIRE(cfg && out, "mb_sg3_create: null argument");
    if (g_device < 0) {
        int r = ensure_init();
        if (r != MB_OK) return r;
    }
    mb_net* net = new mb_net();
    net->cfg = *cfg;
    std::vector<mb_sg3_layer> geo(cfg->num_layers + 1);
    int r = mb_sg3_geometry(cfg, geo.data(), &net->in_channels, &net->in_size, &net->in_sr, &net->in_bw);
    if (r != MB_OK) {
        delete net;
        return r;
    }
    const int wd = cfg->w_dim;
    const int C0 = net->in_channels;
#define TRY(x)            \
    do {                  \
        int _r = (x);     \
        if (_r != MB_OK) {\
            mb_net_destroy(net); \
            return _r;    \
        }                 \
    } while (0)
    TRY(alloc_param(net->in_freqs, {C0, 2}));
    TRY(alloc_param(net->in_phases, {C0}));
    TRY(alloc_param(net->in_weight, {C0, C0}));
    TRY(alloc_param(net->in_affine_w, {4, wd}));
    TRY(alloc_param(net->in_affine_b, {4}));
    TRY(alloc_param(net->in_transform, {3, 3}));
    if (cudaMalloc(&net->in_weightT, sizeof(float) * C0 * C0) != cudaSuccess) {
        set_error("cudaMalloc failed");
        mb_net_destroy(net);
        return MB_ECUDA;
    }
    if (C0 % 16 == 0) {
        if (cudaMalloc(&net->in_w3, sizeof(float) * C0 * 3 * C0) != cudaSuccess ||
            cudaMalloc(&net->in_wpk, sizeof(__half) * packed_weight_elems(C0, 3 * C0, 1)) != cudaSuccess) {
            set_error("cudaMalloc failed");
            mb_net_destroy(net);
            return MB_ECUDA;
        }
    }
    net->by_name["input.freqs"] = &net->in_freqs;
    net->by_name["input.phases"] = &net->in_phases;
    net->by_name["input.weight"] = &net->in_weight;
    net->by_name["input.affine.weight"] = &net->in_affine_w;
    net->by_name["input.affine.bias"] = &net->in_affine_b;
    net->by_name["input.transform"] = &net->in_transform;
    {
        const float eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        cudaMemcpy(net->in_transform.dev, eye, sizeof(eye), cudaMemcpyHostToDevice);
        net->in_transform.set = true;
    }
    net->layers.resize(geo.size());
    for (size_t i = 0; i < geo.size(); ++i) {
        LayerState& L = net->layers[i];
        L.g = geo[i];
        const int k = L.g.conv_kernel;
        TRY(alloc_param(L.weight, {L.g.out_channels, L.g.in_channels, k, k}));
        TRY(alloc_param(L.bias, {L.g.out_channels}));
        TRY(alloc_param(L.affine_w, {L.g.in_channels, wd}));
        TRY(alloc_param(L.affine_b, {L.g.in_channels}));
        TRY(alloc_param(L.magnitude_ema, {}));
        {
            const float one = 1.0f;
            cudaMemcpy(L.magnitude_ema.dev, &one, sizeof(one), cudaMemcpyHostToDevice);
            L.magnitude_ema.set = true;
        }
        if (L.g.up_taps > 1) TRY(alloc_param(L.up_filter, {L.g.up_taps}));
        if (L.g.down_taps > 1) {
            if (L.g.down_radial)
                TRY(alloc_param(L.down_filter, {L.g.down_taps, L.g.down_taps}));
            else
                TRY(alloc_param(L.down_filter, {L.g.down_taps}));
        }
        const std::string n = L.g.name;
        net->by_name[n + ".weight"] = &L.weight;
        net->by_name[n + ".bias"] = &L.bias;
        net->by_name[n + ".affine.weight"] = &L.affine_w;
        net->by_name[n + ".affine.bias"] = &L.affine_b;
        net->by_name[n + ".magnitude_ema"] = &L.magnitude_ema;
        if (L.up_filter.dev) net->by_name[n + ".up_filter"] = &L.up_filter;
        if (L.down_filter.dev) net->by_name[n + ".down_filter"] = &L.down_filter;
        if (!L.g.is_torgb) {
            const size_t ne = packed_weight_elems(L.g.out_channels, L.g.in_channels, k);
            if (cudaMalloc(&L.wpk, ne * sizeof(__half)) != cudaSuccess ||
                cudaMalloc(&L.wsqT, sizeof(float) * L.g.in_channels * L.g.out_channels) != cudaSuccess) {
                set_error("cudaMalloc failed for packed weights of %s", L.g.name);
                mb_net_destroy(net);
                return MB_ECUDA;
            }
        }
    }
#undef TRY
    if (const char* e = getenv("MB_CONV_TILE_W")) net->conv_tile_w = atoi(e) == 16 ? 16 : 32;
    if (const char* e = getenv("MB_FLRELU_IMPL")) net->flrelu_impl = atoi(e);
    *out = net;
    return MB_OK;
}

/* StyleGAN2 synthesis network of the reference's in-tree inference module (maua/GAN/wrappers/inference/stylegan2.py:385,
 * ctor call maua/GAN/wrappers/stylegan2.py:34-36).  The handle is used with the same mb_net_* entry points; parameter names
 * are the reference's state-dict keys ("bs.3.conv0.affine.weight", "bs.0.const", "bs.2.conv1.noise_const", ...). */
extern "C" int mb_sg2_create(int w_dim, int img_resolution, int img_channels, int channel_base, int channel_max, mb_net** out) {
    MB_REQUIRE(out, "mb_sg2_create: null argument");
    if (g_device < 0) {
        int r = ensure_init();
        if (r != MB_OK) return r;
    }
    Sg2Net* n = nullptr;
    int r = sg2_create(w_dim, img_resolution, img_channels, channel_base, channel_max, &n);
    if (r != MB_OK) return r;
    mb_net* net = new mb_net();
    net->sg2 = n;
    *out = net;
    return MB_OK;
}

extern "C" int mb_net_num_ws(const mb_net* net) {
    if (!net) return 0;
    return net->sg2 ? sg2_num_ws(net->sg2) : net->cfg.num_layers + 2;
}

extern "C" void mb_net_destroy(mb_net* net) {
    if (!net) return;
    if (net->sg2) {
        sg2_destroy(net->sg2);
        delete net;
        return;
    }
    for (auto& kv : net->by_name)
        if (kv.second->dev) cudaFree(kv.second->dev);
    for (auto& L : net->layers) {
        if (L.wpk) cudaFree(L.wpk);
        if (L.wsqT) cudaFree(L.wsqT);
    }
    if (net->in_weightT) cudaFree(net->in_weightT);
    if (net->in_w3) cudaFree(net->in_w3);
    if (net->in_wpk) cudaFree(net->in_wpk);
    for (cudaEvent_t e : net->ev_pool) cudaEventDestroy(e);
    delete net;
}

extern "C" int mb_net_set_param(mb_net* net, const char* name, const float* data, const int64_t* shape, int ndim,
                                mb_stream stream) {
    MB_REQUIRE(net && name && data, "mb_net_set_param: null argument");
    if (net->sg2) return sg2_set_param(net->sg2, name, data, shape, ndim, static_cast<cudaStream_t>(stream));
    auto it = net->by_name.find(name);
    MB_REQUIRE(it != net->by_name.end(), "mb_net_set_param: unknown parameter '%s'", name);
    Param& p = *it->second;
    size_t numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= static_cast<size_t>(shape[i]);
    bool same = static_cast<size_t>(ndim) == p.shape.size();
    for (int i = 0; same && i < ndim; ++i) same = shape[i] == p.shape[i];
    if (!same && numel == p.numel && numel == 1) same = true;  // scalars: [] vs [1]
    MB_REQUIRE(same, "mb_net_set_param: shape mismatch for '%s' (got %d dims, %zu elements; expected %zu elements)",
               name, ndim, numel, p.numel);
    MB_CUDA(cudaMemcpyAsync(p.dev, data, sizeof(float) * p.numel, cudaMemcpyDeviceToDevice,
                            static_cast<cudaStream_t>(stream)));
    p.set = true;
    // The input layer's affine and the user transform are read by the kernels as they are (no packed copy): the wrapper
    // edits them between forwards (stabilisation trick, per-call transform), so updating them does not ask for a finalize.
    const bool read_directly = &p == &net->in_affine_w || &p == &net->in_affine_b || &p == &net->in_transform;
    if (!read_directly) net->finalized = false;
    return MB_OK;
}

namespace {
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n * n) {
        const int r = idx / n, c = idx % n;
        out[c * n + r] = in[idx];
    }
}
}  // namespace

extern "C" int mb_net_finalize(mb_net* net, mb_stream stream_) {
    MB_REQUIRE(net, "mb_net_finalize: null net");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (net->sg2) return sg2_finalize(net->sg2, stream);
    for (auto& kv : net->by_name) {
        if (!kv.second->set) {
            set_error("mb_net_finalize: parameter '%s' was never set", kv.first.c_str());
            return MB_ESTATE;
        }
    }
    const int C0 = net->in_channels;
    transpose_kernel<<<ceil_div(C0 * C0, 256), 256, 0, stream>>>(net->in_weight.dev, net->in_weightT, C0);
    if (net->in_wpk) {
        int rr = sg3_input_split_weights_launch(net->in_weight.dev, net->in_w3, C0, stream);
        if (rr == MB_OK) rr = pack_weights_launch(net->in_w3, net->in_wpk, nullptr, C0, 3 * C0, 1, 0, stream);
        if (rr != MB_OK) return rr;
    }
    MB_CUDA(cudaGetLastError());
    for (auto& L : net->layers) {
        if (!L.g.is_torgb) {
            int r = pack_weights_launch(L.weight.dev, L.wpk, L.wsqT, L.g.out_channels, L.g.in_channels, L.g.conv_kernel,
                                        /*prenorm=*/1, stream);
            if (r != MB_OK) return r;
        }
        L.fu.assign(32, 0.0f);
        L.fd.assign(144, 0.0f);
        L.fu[0] = 1.0f;
        L.fd[0] = 1.0f;
        if (L.up_filter.dev)
            MB_CUDA(cudaMemcpyAsync(L.fu.data(), L.up_filter.dev, sizeof(float) * L.up_filter.numel, cudaMemcpyDeviceToHost, stream));
        if (L.down_filter.dev)
            MB_CUDA(cudaMemcpyAsync(L.fd.data(), L.down_filter.dev, sizeof(float) * L.down_filter.numel, cudaMemcpyDeviceToHost, stream));
    }
    MB_CUDA(cudaStreamSynchronize(stream));
    net->finalized = true;
    return MB_OK;
}

// ---------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------
namespace {
inline int cpad8(int c) { return (c + 15) / 16 * 16; }  // channel pitch: whole 16-channel (32-byte) groups
struct WsLayout {
    size_t styles_off, d_off, scratch_off, x_off, x2_off, y_off, p_off, total;
    std::vector<size_t> style_l, d_l;  // per-layer float offsets inside styles / d blocks
};
// Feature-map sizes of one forward: the nominal geometry unless an output-size hook is set, in which case every
// module after the hooked one sees the resized map: conv out = in + k - 1, filtered_lrelu out =
// (in*up + pad_lo + pad_hi - (up_taps-1) - (down_taps-1) + down-1) / down per axis (upstream filtered_lrelu.py).
struct LayerSize { int hin, win, hc, wc, hout, wout; };
struct SizePlan {
    int in_h, in_w;              // SynthesisInput output after its hook
    std::vector<LayerSize> l;    // one per layer (ToRGB: hc = hin, hout = hin)
    int out_h, out_w;
    bool ok;
};
struct ResizeSpec { int module, mode, a, b; };
void hook_size(const ResizeSpec& rs, int module, int& h, int& w) {
    if (rs.mode == MB_RESIZE_NONE || rs.module != module) return;
    if (rs.mode == MB_RESIZE_STRETCH) { h = rs.a; w = rs.b; }
    else { h += 2 * rs.a; w += 2 * rs.b; }
}
SizePlan size_plan_geo(const mb_sg3_layer* geo, int n_layers, int in_size, const ResizeSpec& rs) {
    SizePlan sp;
    int h = in_size, w = in_size;
    sp.ok = true;
    hook_size(rs, 0, h, w);
    if (h <= 0 || w <= 0) sp.ok = false;
    sp.in_h = h; sp.in_w = w;
    for (int i = 0; i < n_layers; ++i) {
        const mb_sg3_layer& g = geo[i];
        LayerSize ls;
        ls.hin = h; ls.win = w;
        ls.hc = h + g.conv_kernel - 1; ls.wc = w + g.conv_kernel - 1;
        const int extra = g.pad_lo + g.pad_hi - (g.up_taps - 1) - (g.down_taps - 1) + (g.down - 1);
        ls.hout = (ls.hc * g.up + extra) / g.down;
        ls.wout = (ls.wc * g.up + extra) / g.down;
        if (ls.hin <= 0 || ls.win <= 0 || ls.hc * g.up + extra <= 0 || ls.wc * g.up + extra <= 0 || ls.hout <= 0 || ls.wout <= 0)
            sp.ok = false;
        sp.l.push_back(ls);
        h = ls.hout; w = ls.wout;
        if (!g.is_torgb) hook_size(rs, i + 1, h, w);
        if (h <= 0 || w <= 0) sp.ok = false;
    }
    sp.out_h = h; sp.out_w = w;
    return sp;
}
SizePlan size_plan(const mb_net* net) {
    std::vector<mb_sg3_layer> geo;
    for (const auto& L : net->layers) geo.push_back(L.g);
    return size_plan_geo(geo.data(), static_cast<int>(geo.size()), net->in_size,
                         ResizeSpec{net->resize_module, net->resize_mode, net->resize_a, net->resize_b});
}
// X: channels-last conv input (X2: its resized copy when a hook is set); Y: planar conv output; P: planar filtered_lrelu output.
WsLayout ws_layout(const mb_net* net, int B, const SizePlan& sp) {
    WsLayout w;
    size_t ns = 0, nd = 0;
    size_t max_x = static_cast<size_t>(B) * net->in_size * net->in_size * cpad8(net->in_channels);
    {
        const size_t x0 = static_cast<size_t>(B) * sp.in_h * sp.in_w * cpad8(net->in_channels);
        if (x0 > max_x) max_x = x0;
    }
    // the tensor-core input layer stages its split features in P and its planar 1x1-conv output in Y
    size_t max_y = static_cast<size_t>(B) * net->in_channels * net->in_size * pitch8(net->in_size);
    size_t max_p = static_cast<size_t>(B) * net->in_size * net->in_size * 3 * net->in_channels;
    for (size_t i = 0; i < net->layers.size(); ++i) {
        const auto& L = net->layers[i];
        const LayerSize& ls = sp.l[i];
        w.style_l.push_back(ns);
        w.d_l.push_back(nd);
        ns += static_cast<size_t>(B) * L.g.in_channels;
        nd += static_cast<size_t>(B) * L.g.out_channels;
        if (!L.g.is_torgb) {
            const size_t xin = static_cast<size_t>(B) * ls.hin * ls.win * cpad8(L.g.in_channels);
            if (xin > max_x) max_x = xin;
            // pre-hook output of this layer in channels-last form (what the fused filtered_lrelu writes)
            const size_t xout = static_cast<size_t>(B) * ls.hout * ls.wout * cpad8(L.g.out_channels);
            if (xout > max_x) max_x = xout;
            const size_t y = static_cast<size_t>(B) * L.g.out_channels * ls.hc * pitch16(ls.wc);
            if (y > max_y) max_y = y;
            const size_t po = static_cast<size_t>(B) * L.g.out_channels * ls.hout * pitch8(ls.wout);
            if (po > max_p) max_p = po;
        } else {
            // a hook on the last conv layer resizes the planar map P into Y
            const size_t y = static_cast<size_t>(B) * L.g.in_channels * ls.hin * pitch8(ls.win);
            if (net->resize_mode != MB_RESIZE_NONE && y > max_y) max_y = y;
        }
    }
    size_t off = 0;
    w.styles_off = off; off = round_up_sz(off + ns * sizeof(float), 1024);
    w.d_off = off; off = round_up_sz(off + nd * sizeof(float), 1024);
    // scratch: [B][C][4] floats of the input layer, then one word per layer for the conv's max |y| (clamp guard of the filter)
    w.scratch_off = off; off = round_up_sz(off + static_cast<size_t>(B) * net->in_channels * 4 * sizeof(float) + kMaxLayers * sizeof(unsigned int), 1024);
    w.x_off = off; off = round_up_sz(off + max_x * sizeof(__half), 1024);
    w.x2_off = off;
    if (net->resize_mode != MB_RESIZE_NONE) off = round_up_sz(off + max_x * sizeof(__half), 1024);
    w.y_off = off; off = round_up_sz(off + max_y * sizeof(__half), 1024);
    w.p_off = off; off = round_up_sz(off + max_p * sizeof(__half), 1024);
    w.total = off;
    return w;
}
}  // namespace

extern "C" size_t mb_net_workspace_bytes(const mb_net* net, int batch) {
    if (!net || batch <= 0) return 0;
    if (net->sg2) return sg2_workspace_bytes(net->sg2, batch);
    const SizePlan sp = size_plan(net);
    if (!sp.ok) return 0;
    return ws_layout(net, batch, sp).total;
}

extern "C" int mb_net_set_resize(mb_net* net, int module, int strategy, int a, int b) {
    MB_REQUIRE(net && !net->sg2, "mb_net_set_resize: needs a StyleGAN3 handle");
    if (strategy == MB_RESIZE_NONE) {
        net->resize_module = -1; net->resize_mode = MB_RESIZE_NONE; net->resize_a = net->resize_b = 0;
        return MB_OK;
    }
    MB_REQUIRE(strategy == MB_RESIZE_STRETCH || strategy == MB_RESIZE_PAD_ZERO, "mb_net_set_resize: unknown strategy %d", strategy);
    MB_REQUIRE(module >= 0 && module <= net->cfg.num_layers,
               "mb_net_set_resize: module %d out of range (0 = input .. %d = last layer before ToRGB)", module, net->cfg.num_layers);
    MB_REQUIRE(strategy != MB_RESIZE_STRETCH || (a > 0 && b > 0), "mb_net_set_resize: stretch needs a positive target size, got %dx%d", a, b);
    const int om = net->resize_module, os = net->resize_mode, oa = net->resize_a, ob = net->resize_b;
    net->resize_module = module; net->resize_mode = strategy; net->resize_a = a; net->resize_b = b;
    if (!size_plan(net).ok) {
        net->resize_module = om; net->resize_mode = os; net->resize_a = oa; net->resize_b = ob;
        set_error("mb_net_set_resize: the requested size leaves an empty feature map somewhere in the network");
        return MB_EINVAL;
    }
    return MB_OK;
}

extern "C" int mb_sg3_resized_output(const mb_sg3_cfg* cfg, int module, int strategy, int a, int b, int32_t* height, int32_t* width) {
    MB_REQUIRE(cfg && height && width, "mb_sg3_resized_output: null argument");
    MB_REQUIRE(cfg->num_layers > 0 && cfg->num_layers < kMaxLayers, "mb_sg3_resized_output: bad num_layers");
    std::vector<mb_sg3_layer> geo(cfg->num_layers + 1);
    int32_t ch = 0, in_size = 0;
    const int r = mb_sg3_geometry(cfg, geo.data(), &ch, &in_size, nullptr, nullptr);
    if (r != MB_OK) return r;
    MB_REQUIRE(strategy == MB_RESIZE_NONE || (module >= 0 && module <= cfg->num_layers), "mb_sg3_resized_output: module %d out of range", module);
    const SizePlan sp = size_plan_geo(geo.data(), static_cast<int>(geo.size()), in_size, ResizeSpec{module, strategy, a, b});
    MB_REQUIRE(sp.ok, "mb_sg3_resized_output: the requested size leaves an empty feature map somewhere in the network");
    *height = sp.out_h; *width = sp.out_w;
    return MB_OK;
}

extern "C" int mb_sg2_set_warps(mb_net* net, int n_warps, const int32_t* layers, const float* inv_mats, int batch) {
    MB_REQUIRE(net && net->sg2, "mb_sg2_set_warps: needs a StyleGAN2 handle");
    return sg2_set_warps(net->sg2, n_warps, layers, inv_mats, batch);
}

extern "C" int mb_sg2_set_resize(mb_net* net, int layer, int mode, int target_h, int target_w, int pad_top, int pad_left, float value,
                                 const float* noise, float* stats) {
    MB_REQUIRE(net && net->sg2, "mb_sg2_set_resize: needs a StyleGAN2 handle");
    return sg2_set_resize(net->sg2, layer, mode, target_h, target_w, pad_top, pad_left, value, noise, stats);
}

extern "C" int mb_net_output_shape(const mb_net* net, int32_t* height, int32_t* width) {
    MB_REQUIRE(net && height && width, "mb_net_output_shape: null argument");
    if (net->sg2) {
        int h = 0, w = 0;
        sg2_output_hw(net->sg2, &h, &w);
        *height = h; *width = w;
        return MB_OK;
    }
    const SizePlan sp = size_plan(net);
    *height = sp.out_h; *width = sp.out_w;
    return MB_OK;
}

extern "C" int mb_net_set_conv_impl(mb_net* net, int impl) {
    MB_REQUIRE(net && (impl == 0 || impl == 1), "mb_net_set_conv_impl: bad argument");
    if (net->sg2) sg2_set_conv_impl(net->sg2, impl);
    net->conv_impl = impl;
    return MB_OK;
}

extern "C" int mb_net_set_option(mb_net* net, const char* key, int value) {
    MB_REQUIRE(net && key, "mb_net_set_option: null argument");
    const std::string k = key;
    if (k == "conv_impl") {
        net->conv_impl = value;
        if (net->sg2) sg2_set_conv_impl(net->sg2, value);
    }
    else if (k == "sg2_precise") {
        MB_REQUIRE(net->sg2, "mb_net_set_option: sg2_precise needs a StyleGAN2 handle");
        sg2_set_precise(net->sg2, value);
    }
    else if (k == "conv_tile_w") net->conv_tile_w = value == 16 ? 16 : 32;
    else if (k == "conv_pm_max") net->conv_pm_max = value;
    else if (k == "conv_epi_groups") net->conv_epi_groups = value;
    else if (k == "conv_pm_stack") net->conv_pm_stack = value;
    else if (k == "conv_cm_stack") net->conv_cm_stack = value;
    else if (k == "conv_row_il") net->conv_row_il = value;
    else if (k == "conv_narrow_a") net->conv_narrow_a = value;
    else if (k == "conv_pm_shift") net->conv_pm_shift = value;
    else if (k == "conv_cm_shift") net->conv_cm_shift = value;
    else if (k == "flrelu_impl") net->flrelu_impl = value;
    else if (k == "input_impl") net->input_impl = value;
    else if (k == "debug_stop") net->debug_stop = value;
    else if (k == "profile") net->profile = value;
    else if (k == "profile_reset") { net->prof.clear(); net->ev_used = 0; }
    else {
        set_error("mb_net_set_option: unknown option '%s'", key);
        return MB_EINVAL;
    }
    return MB_OK;
}

extern "C" int mb_net_last_launch_count(const mb_net* net) {
    if (net && net->sg2) return sg2_last_launches(net->sg2);
    return net ? net->last_launches : 0;
}

/* Per-launch device times of the last forward run with option "profile"=1 (call after the stream has
 * been synchronised).  kind: 0 styles, 1 input, 2 modulated conv, 3 filtered_lrelu, 4 layout
 * transpose, 5 torgb/output.  Returns the number of records written (<= cap). */
extern "C" int mb_net_profile_read(mb_net* net, float* ms, int32_t* kind, int32_t* layer, int cap) {
    if (!net) return 0;
    int n = 0;
    for (const auto& rec : net->prof) {
        if (n >= cap) break;
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, net->ev_pool[rec.ev0], net->ev_pool[rec.ev1]) != cudaSuccess) t = -1.0f;
        ms[n] = t; kind[n] = rec.kind; layer[n] = rec.layer;
        ++n;
    }
    return n;
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
static int net_forward(mb_net* net, const float* ws, const float* transform, int transform_stride, int B, void* out, int out_fmt,
                       void* workspace, size_t workspace_bytes, mb_stream stream_);

extern "C" int mb_net_forward(mb_net* net, const float* ws, const float* transform, int B, void* out, int out_fmt,
                              void* workspace, size_t workspace_bytes, mb_stream stream_) {
    return net_forward(net, ws, transform, 0, B, out, out_fmt, workspace, workspace_bytes, stream_);
}

extern "C" int mb_net_forward_xf(mb_net* net, const float* ws, const float* transforms, int B, void* out, int out_fmt,
                                 void* workspace, size_t workspace_bytes, mb_stream stream_) {
    MB_REQUIRE(net && !net->sg2, "mb_net_forward_xf: needs a StyleGAN3 handle");
    MB_REQUIRE(transforms, "mb_net_forward_xf: null transforms");
    return net_forward(net, ws, transforms, 9, B, out, out_fmt, workspace, workspace_bytes, stream_);
}

static int net_forward(mb_net* net, const float* ws, const float* transform, int transform_stride, int B, void* out, int out_fmt,
                       void* workspace, size_t workspace_bytes, mb_stream stream_) {
    MB_REQUIRE(net && ws && out && workspace, "mb_net_forward: null argument");
    MB_REQUIRE(B > 0, "mb_net_forward: batch must be positive");
    MB_REQUIRE(out_fmt == MB_OUT_F32_NCHW || out_fmt == MB_OUT_F32_NCHW_01 || out_fmt == MB_OUT_U8_NHWC || out_fmt == MB_OUT_F32_NCHW_UNIT,
               "mb_net_forward: unknown out_fmt %d", out_fmt);
    if (net->sg2) {
        MB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "mb_net_forward: workspace must be 1024-byte aligned");
        cudaStream_t st = static_cast<cudaStream_t>(stream_);
        if (!net->profile) return sg2_forward(net->sg2, ws, B, out, out_fmt, workspace, workspace_bytes, g_num_sms, st);
        if (net->profile != 2) { net->prof.clear(); net->ev_used = 0; }
        auto ev_next2 = [&]() -> int {
            if (net->ev_used == static_cast<int>(net->ev_pool.size())) {
                cudaEvent_t e;
                cudaEventCreate(&e);
                net->ev_pool.push_back(e);
            }
            cudaEventRecord(net->ev_pool[net->ev_used], st);
            return net->ev_used++;
        };
        int ev_prev2 = ev_next2();
        return sg2_forward(net->sg2, ws, B, out, out_fmt, workspace, workspace_bytes, g_num_sms, st, [&](int kind, int layer) {
            const int e = ev_next2();
            net->prof.push_back({kind, layer, ev_prev2, e});
            ev_prev2 = e;
        });
    }
    if (!net->finalized) {
        set_error("mb_net_forward: call mb_net_finalize after setting parameters");
        return MB_ESTATE;
    }
    MB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "mb_net_forward: workspace must be 1024-byte aligned");
    const SizePlan sp = size_plan(net);
    MB_REQUIRE(sp.ok, "mb_net_forward: the output-size hook leaves an empty feature map");
    const WsLayout wl = ws_layout(net, B, sp);
    if (workspace_bytes < wl.total) {
        set_error("mb_net_forward: workspace too small (%zu < %zu bytes)", workspace_bytes, wl.total);
        return MB_ENOMEM;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    float* styles = reinterpret_cast<float*>(base + wl.styles_off);
    float* dco = reinterpret_cast<float*>(base + wl.d_off);
    float* scratch = reinterpret_cast<float*>(base + wl.scratch_off);
    unsigned int* absmax = reinterpret_cast<unsigned int*>(scratch + static_cast<size_t>(B) * net->in_channels * 4);
    MB_CUDA(cudaMemsetAsync(absmax, 0, kMaxLayers * sizeof(unsigned int), stream));
    __half* X = reinterpret_cast<__half*>(base + wl.x_off);
    __half* X2 = reinterpret_cast<__half*>(base + wl.x2_off);
    __half* Y = reinterpret_cast<__half*>(base + wl.y_off);
    __half* P = reinterpret_cast<__half*>(base + wl.p_off);
    const int nl = static_cast<int>(net->layers.size());
    int launches = 0;
    int r;
    if (net->profile != 2) {  // 2 = accumulate records over several forwards until "profile_reset"
        net->prof.clear();
        net->ev_used = 0;
    }
    auto ev_next = [&]() -> int {
        if (net->ev_used == static_cast<int>(net->ev_pool.size())) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            net->ev_pool.push_back(e);
        }
        cudaEventRecord(net->ev_pool[net->ev_used], stream);
        return net->ev_used++;
    };
    // MB_NVTX=1: one NVTX range per stage of the forward (styles, input, conv Lk, filtered_lrelu Lk, resize, torgb) inside a
    // range for the call, for timeline tools (tracing).  A range is opened when the previous stage's mark is passed.
    static const bool nvtx_on = [] { const char* e = getenv("MB_NVTX"); return e && atoi(e) != 0; }();
    if (nvtx_on) {
        char nm[64];
        snprintf(nm, sizeof(nm), "mb_net_forward B=%d", B);
        nvtxRangePushA(nm);
        nvtxRangePushA("styles");
    }
    auto nvtx_next = [&](int kind, int layer) {
        if (!nvtx_on) return;
        nvtxRangePop();
        char nm[64];
        if (kind == 0) snprintf(nm, sizeof(nm), "input");
        else if (kind == 2) snprintf(nm, sizeof(nm), "filtered_lrelu L%d", layer);
        else if (kind == 5) snprintf(nm, sizeof(nm), "done");
        else snprintf(nm, sizeof(nm), "%s L%d", (layer + 1 < nl && net->layers[layer + 1].g.is_torgb) ? "torgb+output" : "modulated_conv2d", layer + 1);
        nvtxRangePushA(nm);
    };
    struct NvtxClose {
        bool on;
        ~NvtxClose() { if (on) { nvtxRangePop(); nvtxRangePop(); } }
    } nvtx_close{nvtx_on};
    int ev_prev = net->profile ? ev_next() : -1;
    auto prof_mark = [&](int kind, int layer) {
        nvtx_next(kind, layer);
        if (!net->profile) return;
        const int e = ev_next();
        net->prof.push_back({kind, layer, ev_prev, e});
        ev_prev = e;
    };

    // 1. styles + demodulation coefficients of every layer
    {
        StylesArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.num_layers = nl;
        sa.B = B;
        sa.num_ws = net->cfg.num_layers + 2;
        sa.w_dim = net->cfg.w_dim;
        sa.ws = ws;
        for (int i = 0; i < nl; ++i) {
            const LayerState& L = net->layers[i];
            StyleLayerDesc& d = sa.L[i];
            d.affine_w = L.affine_w.dev;
            d.affine_b = L.affine_b.dev;
            d.wsqT = L.wsqT;
            d.magnitude_ema = L.magnitude_ema.dev;
            d.s_out = styles + wl.style_l[i];
            d.d_out = L.g.is_torgb ? nullptr : dco + wl.d_l[i];
            d.Cin = L.g.in_channels;
            d.Cout = L.g.out_channels;
            d.ws_index = i + 1;
            d.demodulate = !L.g.is_torgb;
            d.normalize_style = 1;
            d.style_scale = L.g.is_torgb ? 1.0f / sqrtf(static_cast<float>(L.g.in_channels * L.g.conv_kernel * L.g.conv_kernel)) : 1.0f;
        }
        if ((r = styles_launch(sa, stream)) != MB_OK) return r;
        launches += 1;
        prof_mark(0, -1);
    }
    // 2. Fourier-feature input, pre-multiplied by layer 0's style
    {
        InputArgs ia;
        ia.ws = ws;
        ia.affine_w = net->in_affine_w.dev;
        ia.affine_b = net->in_affine_b.dev;
        ia.transform = transform ? transform : net->in_transform.dev;
        ia.transform_stride = transform ? transform_stride : 0;
        ia.freqs = net->in_freqs.dev;
        ia.phases = net->in_phases.dev;
        ia.weightT = net->in_weightT;
        ia.style = styles + wl.style_l[0];
        ia.scratch = scratch;
        ia.out = X;
        ia.B = B;
        ia.num_ws = net->cfg.num_layers + 2;
        ia.w_dim = net->cfg.w_dim;
        ia.C = net->in_channels;
        ia.size = net->in_size;
        ia.Cp = cpad8(net->in_channels);
        ia.sampling_rate = static_cast<float>(net->in_sr);
        ia.bandwidth = static_cast<float>(net->in_bw);
        if (net->input_impl == 1 && net->in_wpk) {
            // features [fh | fl | fh] -> P (free until the first filtered_lrelu), 1x1 tcgen05 conv with the layer-0 style as
            // the epilogue scale -> planar Y, exact transpose -> channels-last X
            const int C3 = 3 * net->in_channels;
            __half* feat = P;
            if ((r = sg3_input_features_launch(ia, feat, stream)) != MB_OK) return r;
            ConvTcArgs ca;
            ca.x = feat; ca.wpk = net->in_wpk; ca.d = ia.style; ca.bias = nullptr; ca.y = Y;
            ca.B = B; ca.Cin = C3; ca.Cout = net->in_channels; ca.Hin = net->in_size; ca.Win = net->in_size; ca.Cp_in = C3;
            ca.Wp_out = pitch8(net->in_size); ca.ksz = 1; ca.pad = 0; ca.tile_w = net->conv_tile_w;
            ca.pm_max_cout = 0; ca.narrow_a = net->conv_narrow_a; ca.num_sms = g_num_sms;
            if ((r = conv_tc_launch(ca, stream)) != MB_OK) return r;
            if ((r = planar_to_nhwc_launch(Y, X, B, net->in_channels, net->in_size, net->in_size, pitch8(net->in_size),
                                           cpad8(net->in_channels), stream)) != MB_OK) return r;
            launches += 4;
        } else {
            if ((r = sg3_input_launch(ia, stream)) != MB_OK) return r;
            launches += 2;
        }
        prof_mark(1, -1);
    }
    net->last_ws = workspace;
    net->last_batch = B;
    net->last_act = X;
    net->last_act_nhwc = true;
    net->last_act_c = net->in_channels;
    net->last_act_h = net->last_act_w = net->in_size;
    // output-size hook on the module whose channels-last output sits in X (h x w): resize into X2 and swap
    auto hook_nhwc = [&](int module, int channels, int h, int w, int oh, int ow) -> int {
        if (net->resize_mode == MB_RESIZE_NONE || net->resize_module != module) return MB_OK;
        const int rr = resize_nhwc_launch(X, X2, B, h, w, cpad8(channels), oh, ow, net->resize_mode, net->resize_a, net->resize_b,
                                          g_num_sms, stream);
        if (rr != MB_OK) return rr;
        std::swap(X, X2);
        launches += 1;
        prof_mark(4, module - 1);
        net->last_act = X;
        net->last_act_h = oh; net->last_act_w = ow;
        return MB_OK;
    };
    if ((r = hook_nhwc(0, net->in_channels, net->in_size, net->in_size, sp.in_h, sp.in_w)) != MB_OK) return r;
    if (net->debug_stop < 0) {
        net->last_launches = launches;
        return MB_OK;
    }
    // 3. layers
    for (int i = 0; i < nl; ++i) {
        const LayerState& L = net->layers[i];
        const mb_sg3_layer& g = L.g;
        const LayerSize& ls = sp.l[i];
        if (g.is_torgb) {
            ToRgbArgs ta;
            ta.x = P;
            ta.w = L.weight.dev;
            ta.bias = L.bias.dev;
            ta.out = out;
            ta.B = B; ta.Cin = g.in_channels; ta.Cout = g.out_channels;
            ta.H = ls.hin; ta.W = ls.win; ta.Wp = pitch8(ls.win);
            ta.out_fmt = out_fmt;
            ta.clamp = static_cast<float>(net->cfg.conv_clamp);
            ta.output_scale = static_cast<float>(net->cfg.output_scale);
            MB_REQUIRE(g.up == 1 && g.down == 1 && g.conv_kernel == 1 && g.in_size == g.out_size,
                       "forward: unexpected ToRGB geometry");
            if ((r = torgb_out_launch(ta, stream)) != MB_OK) return r;
            launches += 1;
            prof_mark(5, i);
            break;
        }
        ConvTcArgs ca;
        ca.x = X;
        ca.wpk = L.wpk;
        ca.d = dco + wl.d_l[i];
        ca.bias = L.bias.dev;  // the layer bias is added in the conv epilogue (single fp16 rounding)
        ca.y = Y;
        ca.B = B; ca.Cin = g.in_channels; ca.Cout = g.out_channels;
        ca.Hin = ls.hin; ca.Win = ls.win; ca.Cp_in = cpad8(g.in_channels);
        const int hc = ls.hc, wc = ls.wc;
        ca.Wp_out = pitch16(wc);   // 32-byte aligned rows: the narrow layers' epilogue stores 16 pixels per instruction
        ca.ksz = g.conv_kernel;
        ca.pad = g.conv_kernel - 1;
        // 16x16 pixel tiles on the small maps (<= 64^2: a third fewer tiles per wave, measured 0.084 -> 0.061 ms on L0..L2),
        // 32x8 elsewhere (fewer halo rows per TMA box)
        ca.tile_w = (net->conv_tile_w == 32 && hc <= 64 && wc <= 64 && g.conv_kernel == 3) ? 16 : net->conv_tile_w;
        ca.pm_max_cout = net->conv_pm_max;
        ca.epi_groups = net->conv_epi_groups;
        ca.pm_stack = net->conv_pm_stack;
        ca.cm_stack = net->conv_cm_stack;
        ca.narrow_a = net->conv_narrow_a;
        ca.pm_shift = net->conv_pm_shift;
        ca.cm_shift = net->conv_cm_shift;
        ca.num_sms = g_num_sms;
        ca.absmax = net->conv_impl == 0 ? absmax + i : nullptr;
        {
            // row-interleaved conv output: only where the streaming filter kernels (4-D tensor map) consume it
            FlreluArgs pr;
            memset(&pr, 0, sizeof(pr));
            pr.B = B; pr.C = g.out_channels; pr.up = g.up; pr.down = g.down; pr.up_taps = g.up_taps; pr.down_taps = g.down_taps;
            pr.fd_2d = g.down_radial; pr.px0 = g.pad_lo; pr.py0 = g.pad_lo; pr.gain = sqrtf(2.0f); pr.slope = 0.2f;
            memcpy(pr.fd, L.fd.data(), sizeof(pr.fd));
            const char* sev = getenv("MB_FLRELU_STREAM");
            ca.row_interleaved = (net->conv_impl == 0 && net->flrelu_impl == 0 && g.out_channels <= net->conv_row_il &&
                                  !(sev && atoi(sev) == 0) && flrelu_mma_supported(pr)) ? 1 : 0;
        }
        r = net->conv_impl == 0 ? conv_tc_launch(ca, stream) : conv_simt_launch(ca, stream);
        if (r != MB_OK) return r;
        launches += 1;
        prof_mark(2, i);

        FlreluArgs fa;
        memset(&fa, 0, sizeof(fa));
        fa.x = Y;
        fa.bias = nullptr;
        fa.scale = styles + wl.style_l[i + 1];
        fa.y = P;
        memcpy(fa.fu, L.fu.data(), sizeof(fa.fu));
        memcpy(fa.fd, L.fd.data(), sizeof(fa.fd));
        fa.B = B; fa.C = g.out_channels;
        fa.Hin = hc; fa.Win = wc; fa.Wp_in = pitch16(wc);
        fa.Hout = ls.hout; fa.Wout = ls.wout; fa.Wp_out = pitch8(ls.wout);
        fa.up = g.up; fa.down = g.down; fa.up_taps = g.up_taps; fa.down_taps = g.down_taps;
        fa.fd_2d = g.down_radial;
        fa.px0 = g.pad_lo; fa.py0 = g.pad_lo;
        fa.gain = sqrtf(2.0f); fa.slope = 0.2f;
        fa.clamp = static_cast<float>(net->cfg.conv_clamp);
        fa.num_sms = g_num_sms;
        fa.in_absmax = ca.absmax;
        fa.in_row_interleaved = ca.row_interleaved;
        // Layers that feed another conv write channels-last straight from the tensor-core kernel; the
        // fallback kernels (and the last layer, whose consumer is the planar ToRGB kernel) write planar.
        const bool next_is_conv = !net->layers[i + 1].g.is_torgb;
        FlreluArgs probe = fa;
        const bool fused = next_is_conv && net->flrelu_impl == 0 && net->debug_stop > i && flrelu_mma_supported(probe);
        if (fused) {
            fa.y_nhwc = X;
            fa.Cp_out = cpad8(g.out_channels);
        }
        r = flrelu_launch_impl(fa, net->flrelu_impl, stream);
        if (r != MB_OK) return r;
        launches += 1;
        prof_mark(3, i);
        net->last_act = fused ? X : P;
        net->last_act_nhwc = fused;
        net->last_act_c = g.out_channels;
        net->last_act_h = ls.hout; net->last_act_w = ls.wout;
        if (net->debug_stop <= i && !(net->resize_mode != MB_RESIZE_NONE && net->resize_module == i + 1)) break;
        if (next_is_conv && !fused) {
            r = planar_to_nhwc_launch(P, X, B, g.out_channels, ls.hout, ls.wout, pitch8(ls.wout),
                                      cpad8(g.out_channels), stream);
            if (r != MB_OK) return r;
            launches += 1;
            prof_mark(4, i);
            net->last_act = X;
            net->last_act_nhwc = true;
        }
        // output-size hook on this layer (module i + 1)
        if (next_is_conv) {
            if ((r = hook_nhwc(i + 1, g.out_channels, ls.hout, ls.wout, sp.l[i + 1].hin, sp.l[i + 1].win)) != MB_OK) return r;
        } else if (net->resize_mode != MB_RESIZE_NONE && net->resize_module == i + 1) {
            // the last conv layer feeds the planar ToRGB kernel: resize P into Y and hand Y over
            const int oh = sp.l[i + 1].hin, ow = sp.l[i + 1].win;
            r = resize_planar_launch(P, Y, B * g.out_channels, ls.hout, ls.wout, pitch8(ls.wout), oh, ow, pitch8(ow),
                                     net->resize_mode, net->resize_a, net->resize_b, g_num_sms, stream);
            if (r != MB_OK) return r;
            std::swap(P, Y);
            launches += 1;
            prof_mark(4, i);
            net->last_act = P;
            net->last_act_nhwc = false;
            net->last_act_h = oh; net->last_act_w = ow;
        }
        if (net->debug_stop <= i) break;
    }
    net->last_launches = launches;
    return MB_OK;
}

extern "C" int mb_net_read_activation(mb_net* net, int idx, int batch, float* out, mb_stream stream) {
    (void)idx;
    MB_REQUIRE(net && out && net->last_act, "mb_net_read_activation: no forward has run");
    MB_REQUIRE(batch == net->last_batch, "mb_net_read_activation: batch mismatch");
    if (net->last_act_nhwc)
        return nhwc_to_float_launch(net->last_act, out, batch, net->last_act_c, net->last_act_h, net->last_act_w,
                                    cpad8(net->last_act_c), static_cast<cudaStream_t>(stream));
    return half_to_float_launch(net->last_act, out, batch, net->last_act_c, net->last_act_h, net->last_act_w,
                                pitch8(net->last_act_w), static_cast<cudaStream_t>(stream));
}

extern "C" int mb_net_activation_shape(const mb_net* net, int32_t* c, int32_t* h, int32_t* w) {
    MB_REQUIRE(net && c && h && w, "mb_net_activation_shape: null argument");
    *c = net->last_act_c; *h = net->last_act_h; *w = net->last_act_w;
    return MB_OK;
}

// ---------------------------------------------------------------------------------------
// op-level entry points
// ---------------------------------------------------------------------------------------
extern "C" int mb_modulated_conv2d(const float* x, const float* w, const float* s, float* y, int B, int Cin, int Cout,
                                   int H, int W, int k, int demodulate, float input_gain, int impl, mb_stream stream_) {
    MB_REQUIRE(x && w && s && y, "mb_modulated_conv2d: null argument");
    MB_REQUIRE(k >= 1 && k <= 3, "mb_modulated_conv2d: kernel size %d unsupported (1..3)", k);
    if (g_device < 0) {
        int r = ensure_init();
        if (r != MB_OK) return r;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int Ho = H + k - 1, Wo = W + k - 1;
    const int Cp = (Cin + 7) / 8 * 8, Wpo = pitch16(Wo);
    __half *xh = nullptr, *wpk = nullptr, *yh = nullptr;
    float *wsqT = nullptr, *sn = nullptr, *d = nullptr;
    const size_t nx = static_cast<size_t>(B) * H * W * Cp, ny = static_cast<size_t>(B) * Cout * Ho * Wpo;
    MB_CUDA(cudaMalloc(&xh, nx * 2));
    MB_CUDA(cudaMalloc(&yh, ny * 2));
    MB_CUDA(cudaMalloc(&wpk, packed_weight_elems(Cout, Cin, k) * 2));
    MB_CUDA(cudaMalloc(&wsqT, sizeof(float) * Cin * Cout));
    MB_CUDA(cudaMalloc(&sn, sizeof(float) * B * Cin));
    MB_CUDA(cudaMalloc(&d, sizeof(float) * B * Cout));
    int r = pack_weights_launch(w, wpk, wsqT, Cout, Cin, k, demodulate, stream);
    if (r == MB_OK) r = style_demod_launch(s, wsqT, sn, d, B, Cin, Cout, demodulate, input_gain, stream);
    if (r == MB_OK) r = modulate_to_nhwc_launch(x, sn, 1.0f, xh, B, Cin, H, W, Cp, stream);
    if (r == MB_OK) {
        ConvTcArgs ca;
        ca.x = xh; ca.wpk = wpk; ca.d = d; ca.bias = nullptr; ca.y = yh;
        ca.B = B; ca.Cin = Cin; ca.Cout = Cout; ca.Hin = H; ca.Win = W; ca.Cp_in = Cp; ca.Wp_out = Wpo; ca.ksz = k;
        ca.pad = k - 1;
        ca.tile_w = (impl == 2) ? 16 : 32;
        ca.cm_stack = (impl == 12) ? 1 : 0;   // 12: stacked cout-major tile where the layer has <= 32 couts (kw taps along M, TMEM column shifts)
        ca.pm_stack = (impl == 0 || impl == 11) ? 1 : 0;   // 11: stacked pixel-major tile wherever its three kw blocks fit one instruction
        if (impl == 11) ca.pm_max_cout = 80;
        else
        ca.pm_max_cout = (impl >= 4 && impl <= 6) ? 128 : ((impl == 0) ? 64 : 0);  // 4..6: pixel-major tile for every layer up to 128 couts
        ca.pm_shift = (impl == 0 || impl == 5) ? 1 : (impl == 6 ? 2 : 0);  // 4: one patch load per kw; 5: single load, shifted
                                                                           // A descriptors; 6: + base-offset field (WRONG results:
                                                                           // kept as the probe of scripts/conv_shift_probe.py)
        ca.cm_shift = (impl == 7) ? 1 : (impl == 10 ? 2 : 0);  // 7: cout-major tile everywhere, single patch load + shifted B descriptors
                                                                 // 10: the same on a 34-pixel-wide patch (aligned 32-pixel tiles of 7 rows)
        ca.narrow_a = (impl == 3) ? 0 : 1;       // 3: always stream full 128-row weight tiles (the r1 first path)
        ca.epi_groups = (impl == 8 || impl == 10) ? 3 : (impl == 9 ? 1 : 0);  // 8 / 9: cout-major tile everywhere with three / one epilogue warp group(s)
        ca.num_sms = g_num_sms;
        r = (impl == 1) ? conv_simt_launch(ca, stream) : conv_tc_launch(ca, stream);
    }
    if (r == MB_OK) r = half_to_float_launch(yh, y, B, Cout, Ho, Wo, Wpo, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(xh); cudaFree(yh); cudaFree(wpk); cudaFree(wsqT); cudaFree(sn); cudaFree(d);
    if (r == MB_OK && e != cudaSuccess) {
        set_error("mb_modulated_conv2d: %s", cudaGetErrorString(e));
        return MB_ECUDA;
    }
    return r;
}

extern "C" int mb_filtered_lrelu(const float* x, const float* fu, const float* fd, const float* b, float* y, int B,
                                 int C, int H, int W, int up, int down, int up_taps, int down_taps, int fd_2d, int px0,
                                 int px1, int py0, int py1, float gain, float slope, float clamp, mb_stream stream_) {
    MB_REQUIRE(x && y, "mb_filtered_lrelu: null argument");
    MB_REQUIRE(up >= 1 && down >= 1 && up_taps >= 1 && up_taps <= 32 && down_taps >= 1 && down_taps <= 12,
               "mb_filtered_lrelu: unsupported filter sizes");
    MB_REQUIRE((fu != nullptr) == (up_taps > 1) && (fd != nullptr) == (down_taps > 1),
               "mb_filtered_lrelu: filter pointer / tap count mismatch");
    if (g_device < 0) {
        int r = ensure_init();
        if (r != MB_OK) return r;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int Ho = (H * up + py0 + py1 - (up_taps - 1) - (down_taps - 1) + (down - 1)) / down;
    const int Wo = (W * up + px0 + px1 - (up_taps - 1) - (down_taps - 1) + (down - 1)) / down;
    MB_REQUIRE(Ho > 0 && Wo > 0, "mb_filtered_lrelu: empty output");
    FlreluArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.fu[0] = 1.0f;
    fa.fd[0] = 1.0f;
    if (fu) MB_CUDA(cudaMemcpy(fa.fu, fu, sizeof(float) * up_taps, cudaMemcpyDeviceToHost));
    if (fd) MB_CUDA(cudaMemcpy(fa.fd, fd, sizeof(float) * (fd_2d ? down_taps * down_taps : down_taps), cudaMemcpyDeviceToHost));
    const int Wp = pitch8(W), Wpo = pitch8(Wo);
    __half *xh = nullptr, *yh = nullptr;
    MB_CUDA(cudaMalloc(&xh, static_cast<size_t>(B) * C * H * Wp * 2));
    MB_CUDA(cudaMalloc(&yh, static_cast<size_t>(B) * C * Ho * Wpo * 2));
    // The tensor-core kernels take the bias already added (on the product path the conv epilogue adds it before its single
    // fp16 rounding); MB_FLRELU_BIAS_SEPARATE=1 hands it to the kernel instead (tile kernels and CUDA-core kernels).
    const char* impl = getenv("MB_FLRELU_IMPL");
    const char* bsep = getenv("MB_FLRELU_BIAS_SEPARATE");
    const bool fold_bias = b != nullptr && !(impl && atoi(impl) != 0) && !(bsep && atoi(bsep) != 0);
    int r = modulate_to_half_launch(x, nullptr, 1.0f, xh, B, C, H, W, Wp, stream, fold_bias ? b : nullptr);
    fa.x = xh; fa.bias = fold_bias ? nullptr : b; fa.scale = nullptr; fa.y = yh;
    fa.B = B; fa.C = C; fa.Hin = H; fa.Win = W; fa.Wp_in = Wp; fa.Hout = Ho; fa.Wout = Wo; fa.Wp_out = Wpo;
    fa.up = up; fa.down = down; fa.up_taps = up_taps; fa.down_taps = down_taps; fa.fd_2d = fd_2d;
    fa.px0 = px0; fa.py0 = py0;
    fa.gain = gain; fa.slope = slope; fa.clamp = clamp;
    fa.num_sms = g_num_sms;
    // MB_FLRELU_TEST_NHWC=1: run the channels-last variant of the tensor-core kernel (what a layer that feeds a conv uses)
    __half* yn = nullptr;
    const char* tn = getenv("MB_FLRELU_TEST_NHWC");
    const int Cp = (C + 15) / 16 * 16;
    if (tn && atoi(tn) != 0 && fa.bias == nullptr && !(impl && atoi(impl) != 0) && flrelu_mma_supported(fa)) {
        MB_CUDA(cudaMalloc(&yn, static_cast<size_t>(B) * Ho * Wo * Cp * 2));
        MB_CUDA(cudaMemsetAsync(yn, 0xff, static_cast<size_t>(B) * Ho * Wo * Cp * 2, stream));   // NaN pattern: every chunk must be written
        fa.y_nhwc = yn;
        fa.Cp_out = Cp;
    }
    if (r == MB_OK) r = flrelu_launch_impl(fa, impl ? atoi(impl) : 0, stream);
    if (r == MB_OK) r = yn ? nhwc_to_float_launch(yn, y, B, C, Ho, Wo, Cp, stream) : half_to_float_launch(yh, y, B, C, Ho, Wo, Wpo, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    if (yn && r == MB_OK && e == cudaSuccess && Cp > C) {
        // the pad channels feed zero weights of the next conv: they must hold zeros, not stale bytes
        std::vector<uint16_t> host(static_cast<size_t>(B) * Ho * Wo * Cp);
        e = cudaMemcpy(host.data(), yn, host.size() * 2, cudaMemcpyDeviceToHost);
        for (size_t px = 0; e == cudaSuccess && r == MB_OK && px < host.size() / Cp; ++px)
            for (int cpad = C; cpad < Cp; ++cpad)
                if ((host[px * Cp + cpad] & 0x7fff) != 0) {
                    set_error("mb_filtered_lrelu: channels-last pad channel %d of pixel %zu is not zero (0x%04x)", cpad, px, host[px * Cp + cpad]);
                    r = MB_ECUDA;
                    break;
                }
    }
    cudaFree(xh); cudaFree(yh);
    if (yn) cudaFree(yn);
    if (r == MB_OK && e != cudaSuccess) {
        set_error("mb_filtered_lrelu: %s", cudaGetErrorString(e));
        return MB_ECUDA;
    }
    return r;
}
