// filtered_lrelu, register-blocked polyphase kernel for the two layer shapes StyleGAN3 uses at
// down=2 / 12 taps: (up=2, 12 taps) and (up=4, 24 taps).  See flrelu.cu for index conventions.
//
// One CTA = one (b, c) plane x one 64(w) x 32(h) output tile, four shared-memory passes:
//   A  horizontal polyphase up-FIR   s_in [INY][INX]  -> s_uh [INY][WU]
//   B  vertical   polyphase up-FIR + lrelu/gain/clamp -> s_t  [HT][WU]
//   C  horizontal decimating FIR                       -> s_dh [TTY][64]
//   D  vertical   decimating FIR * next-layer style    -> global fp16
// Every pass: a thread owns a short run along the filter axis (inputs held in registers, filter
// taps come from the kernel-parameter constant bank), threads of a warp are spread across the
// other axis with odd row pitches so shared-memory accesses are conflict-free.
// The UP outputs q in (UP(n-1), UP n] of the zero-inserted signal (q = J - pad) all read the same
// six inputs x[n..n+5]; output with phase r = UP n - q uses taps fu[UT-1-r-UP m], m = 0..5.
#include "common.cuh"
#include "kernels.h"

namespace mb {
namespace {

template <int UP>
struct Cfg {
    static constexpr int UT = 6 * UP, DT = 12;
    static constexpr int OTW = 64, OTH = 32;
    static constexpr int TTX = (OTW - 1) * 2 + DT;  // 138
    static constexpr int TTY = (OTH - 1) * 2 + DT;  // 74
    static constexpr int RA = 4;                     // n-groups per thread in passes A and B
    static constexpr int NX = ((TTX - 1) / UP + 2 + RA - 1) / RA * RA;
    static constexpr int NY = ((TTY - 1) / UP + 2 + RA - 1) / RA * RA;
    static constexpr int INX = NX + 5, INY = NY + 5;
    static constexpr int P_IN = INX | 1;             // odd pitches
    static constexpr int WU = UP * NX + UP;
    static constexpr int P_U = WU | 1;
    static constexpr int HT = UP * NY + UP;
    static constexpr int P_D = OTW + 1;
    static constexpr int R1 = (INY * P_IN > HT * P_U) ? INY * P_IN : HT * P_U;   // s_in | s_t
    static constexpr int R2 = (INY * P_U > TTY * P_D) ? INY * P_U : TTY * P_D;   // s_uh | s_dh
    static constexpr int SMEM = (R1 + R2) * 4;
};

template <int UP>
struct SepParams {
    const __half* x;
    const float* bias;
    const float* scale;
    __half* y;
    int C, Hin, Win, Wp_in, Hout, Wout, Wp_out, px0, py0, tiles_x;
    float gain, slope, clamp;
    float fu[6 * UP];  // pre-multiplied by UP
    float fd[12];
};

__device__ __forceinline__ int fdiv(int a, int b) {
    int q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
}

template <int UP>
__global__ void __launch_bounds__(256, 3) flrelu_sep_kernel(const __grid_constant__ SepParams<UP> p) {
    using K = Cfg<UP>;
    extern __shared__ float sm[];
    float* s_in = sm;            // region 1
    float* s_t = sm;             // region 1 (after pass A)
    float* s_uh = sm + K::R1;    // region 2
    float* s_dh = sm + K::R1;    // region 2 (after pass B)

    const int tid = threadIdx.x;
    const int c = blockIdx.y, b = blockIdx.z;
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x % p.tiles_x;
    const int ox0 = tx * K::OTW, oy0 = ty * K::OTH;
    const int J0x = ox0 * 2, J0y = oy0 * 2;
    const int nlx = -fdiv(-(J0x - p.px0), UP);  // ceil
    const int nly = -fdiv(-(J0y - p.py0), UP);
    const int c0x = UP * nlx + p.px0 - J0x;     // in [0, UP)
    const int c0y = UP * nly + p.py0 - J0y;

    // ---- load: global fp16 -> s_in fp32 (+bias), zero outside the image
    {
        const float bias = p.bias ? p.bias[c] : 0.0f;
        const __half* xp = p.x + (static_cast<long long>(b) * p.C + c) * p.Hin * p.Wp_in;
        for (int idx = tid; idx < K::INY * K::INX; idx += 256) {
            const int ly = idx / K::INX, lx = idx - ly * K::INX;
            const int iy = nly + ly, ix = nlx + lx;
            float v = 0.0f;
            if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win)
                v = __half2float(xp[static_cast<long long>(iy) * p.Wp_in + ix]) + bias;
            s_in[ly * K::P_IN + lx] = v;
        }
    }
    __syncthreads();

    // ---- pass A: horizontal up-FIR.  item = (row ly, group-of-RA n values)
    for (int item = tid; item < K::INY * (K::NX / K::RA); item += 256) {
        const int seg = item / K::INY, ly = item - seg * K::INY;
        float xin[K::RA + 5];
#pragma unroll
        for (int i = 0; i < K::RA + 5; ++i) xin[i] = s_in[ly * K::P_IN + seg * K::RA + i];
        float* dst = s_uh + ly * K::P_U + c0x + UP * (seg * K::RA) + (UP - 1);
#pragma unroll
        for (int gi = 0; gi < K::RA; ++gi) {
#pragma unroll
            for (int r = 0; r < UP; ++r) {
                float acc = 0.0f;
#pragma unroll
                for (int m = 0; m < 6; ++m) acc = fmaf(p.fu[K::UT - 1 - r - UP * m], xin[gi + m], acc);
                dst[UP * gi - r] = acc;
            }
        }
    }
    __syncthreads();

    // ---- pass B: vertical up-FIR + activation.  item = (tmp column lc, group-of-RA n values)
    for (int item = tid; item < K::TTX * (K::NY / K::RA); item += 256) {
        const int seg = item / K::TTX, lc = item - seg * K::TTX;
        const int cs = lc + (UP - 1);
        float u[K::RA + 5];
#pragma unroll
        for (int i = 0; i < K::RA + 5; ++i) u[i] = s_uh[(seg * K::RA + i) * K::P_U + cs];
        float* dst = s_t + (c0y + UP * (seg * K::RA) + (UP - 1)) * K::P_U + cs;
#pragma unroll
        for (int gi = 0; gi < K::RA; ++gi) {
#pragma unroll
            for (int r = 0; r < UP; ++r) {
                float acc = 0.0f;
#pragma unroll
                for (int m = 0; m < 6; ++m) acc = fmaf(p.fu[K::UT - 1 - r - UP * m], u[gi + m], acc);
                acc = (acc < 0.0f ? acc * p.slope : acc) * p.gain;
                acc = fminf(fmaxf(acc, -p.clamp), p.clamp);
                dst[(UP * gi - r) * K::P_U] = acc;
            }
        }
    }
    __syncthreads();

    // ---- pass C: horizontal decimating FIR.  item = (tmp row lJ, segment of 8 outputs)
    for (int item = tid; item < K::TTY * (K::OTW / 8); item += 256) {
        const int seg = item / K::TTY, lJ = item - seg * K::TTY;
        const float* src = s_t + (lJ + UP - 1) * K::P_U + (UP - 1) + 16 * seg;
        float tv[26];
#pragma unroll
        for (int i = 0; i < 26; ++i) tv[i] = src[i];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 12; ++k) acc = fmaf(p.fd[11 - k], tv[2 * e + k], acc);
            s_dh[lJ * K::P_D + seg * 8 + e] = acc;
        }
    }
    __syncthreads();

    // ---- pass D: vertical decimating FIR, * next-layer style, fp16 store
    {
        const float oscale = p.scale ? p.scale[b * p.C + c] : 1.0f;
        __half* yp = p.y + (static_cast<long long>(b) * p.C + c) * p.Hout * p.Wp_out;
        for (int item = tid; item < K::OTW * (K::OTH / 8); item += 256) {
            const int seg = item / K::OTW, lx = item - seg * K::OTW;
            float dv[26];
#pragma unroll
            for (int i = 0; i < 26; ++i) dv[i] = s_dh[(16 * seg + i) * K::P_D + lx];
            const int ox = ox0 + lx;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < 12; ++k) acc = fmaf(p.fd[11 - k], dv[2 * e + k], acc);
                const int oy = oy0 + seg * 8 + e;
                if (oy < p.Hout && ox < p.Wout)
                    yp[static_cast<long long>(oy) * p.Wp_out + ox] = __float2half_rn(acc * oscale);
            }
        }
    }
}

template <int UP>
int launch(const FlreluArgs& a, cudaStream_t stream) {
    using K = Cfg<UP>;
    SepParams<UP> p;
    p.x = a.x; p.bias = a.bias; p.scale = a.scale; p.y = a.y;
    p.C = a.C; p.Hin = a.Hin; p.Win = a.Win; p.Wp_in = a.Wp_in;
    p.Hout = a.Hout; p.Wout = a.Wout; p.Wp_out = a.Wp_out;
    p.px0 = a.px0; p.py0 = a.py0;
    p.tiles_x = ceil_div(a.Wout, K::OTW);
    p.gain = a.gain; p.slope = a.slope;
    p.clamp = a.clamp >= 0.0f ? a.clamp : 3.0e38f;
    for (int i = 0; i < 6 * UP; ++i) p.fu[i] = a.fu[i] * UP;
    for (int i = 0; i < 12; ++i) p.fd[i] = a.fd[i];
    static bool attr_done = false;
    if (!attr_done) {
        MB_CUDA(cudaFuncSetAttribute(flrelu_sep_kernel<UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        attr_done = true;
    }
    dim3 grid(p.tiles_x * ceil_div(a.Hout, K::OTH), a.C, a.B);
    flrelu_sep_kernel<UP><<<grid, 256, K::SMEM, stream>>>(p);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

}  // namespace

bool flrelu_sep_supported(const FlreluArgs& a) {
    if (a.fd_2d || a.down != 2 || a.down_taps != 12) return false;
    if (!((a.up == 2 && a.up_taps == 12) || (a.up == 4 && a.up_taps == 24))) return false;
    if (a.C > 65535 || a.B > 65535) return false;
    return true;
}

int flrelu_sep_launch(const FlreluArgs& a, cudaStream_t stream) {
    return a.up == 2 ? launch<2>(a, stream) : launch<4>(a, stream);
}

}  // namespace mb
