// Constant-Q chroma on the GPU: the torch-native chroma_cqt of the reference,
// maua/audiovisual/audioreactive/selfsupervised/features/rosa/constantq.py:13-116 (recursive octave-by-octave
// CQT: kaiser-windowed sinc decimation by 2, rectangular-window STFT, sparse FFT-domain filter bank),
// rosa/spectral.py:286-325 (chroma_cqt: |CQT| -> fold 36 bins/octave onto 12 chroma -> / max) and
// rosa/convert.py:69-117 (cq_to_chroma).
//
// Data flow (one launch each):
//   decimate_kernel x (n_octaves-1)  y_{i+1}[j] = sqrt(2) * sum_k h[k] y_i[2j + k - width]   (28-tap polyphase FIR,
//                                    host-designed exactly as torchaudio's sinc_interp_kaiser kernel)
//   cqt_kernel<NFFT>   grid (T, n_octaves): one CTA stages the NFFT-sample frame of its octave in shared memory
//                      (reflect padding), runs a radix-2 Stockham FFT there and applies the sparse filter bank
//                      (CSR, ~15 non-zeros per bin: the SAME matrix in every octave up to sqrt(2^i), constantq.py:98):
//                      |sum_f B[k,f] X[f]| / sqrt(length_k) -> C[n_bins][T]
//   chroma_fold_kernel chroma[c,t] = sum_b fold[c,b] C[b,t], global max via atomicMax on the (non-negative) bits
//   chroma_norm_kernel chroma /= max
// The audio is a few MB and every stage a handful of microseconds: latency bound, as SURVEY §8d expects.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace mb {
namespace {

__device__ __forceinline__ int reflect_idx(long long i, long long n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return static_cast<int>(i);
}

__global__ void decimate_kernel(const float* __restrict__ x, long long n_in, float* __restrict__ y, long long n_out,
                                const float* __restrict__ h, int taps, int width, float post_scale) {
    __shared__ float hs[64];
    if (threadIdx.x < taps) hs[threadIdx.x] = h[threadIdx.x];
    __syncthreads();
    for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < n_out;
         j += static_cast<long long>(gridDim.x) * blockDim.x) {
        float acc = 0.0f;
        const long long base = 2 * j - width;
        for (int k = 0; k < taps; ++k) {
            const long long i = base + k;
            const float v = (i >= 0 && i < n_in) ? x[i] : 0.0f;
            acc = fmaf(hs[k], v, acc);
        }
        y[j] = acc * post_scale;
    }
}

template <int N>
__device__ float2* fft_pow2(float2* x, float2* y, const float2* tw) {
    for (int ns = 1; ns < N; ns <<= 1) {
        for (int j = threadIdx.x; j < N / 2; j += blockDim.x) {
            const int k = j & (ns - 1);
            const float2 w = tw[k * (N / 2 / ns)];
            const float2 a = x[j], b0 = x[j + N / 2];
            const float2 b = make_float2(w.x * b0.x - w.y * b0.y, w.x * b0.y + w.y * b0.x);
            const int j0 = ((j - k) << 1) + k;
            y[j0] = make_float2(a.x + b.x, a.y + b.y);
            y[j0 + ns] = make_float2(a.x - b.x, a.y - b.y);
        }
        __syncthreads();
        float2* t = x; x = y; y = t;
    }
    return x;
}

struct OctaveSrc {
    const float* y[8];
    long long n[8];
};

template <int N>
__global__ void __launch_bounds__(256) cqt_kernel(OctaveSrc src, int hop0, int T, int bins_per_octave, int n_bins,
                                                  const int* __restrict__ rowptr, const int* __restrict__ col,
                                                  const float2* __restrict__ val, const float* __restrict__ inv_sqrt_len,
                                                  float* __restrict__ C) {
    __shared__ float2 bufA[N];
    __shared__ float2 bufB[N];
    __shared__ float2 tw[N / 2];
    const int t = blockIdx.x, oct = blockIdx.y;
    const float* y = src.y[oct];
    const long long n = src.n[oct];
    const int hop = hop0 >> oct;
    for (int i = threadIdx.x; i < N / 2; i += blockDim.x) {
        float s, c;
        sincospif(static_cast<float>(i) / static_cast<float>(N / 2), &s, &c);
        tw[i] = make_float2(c, -s);
    }
    const long long start = static_cast<long long>(t) * hop - N / 2;
    for (int i = threadIdx.x; i < N; i += blockDim.x) bufA[i] = make_float2(y[reflect_idx(start + i, n)], 0.0f);
    __syncthreads();
    const float2* X = fft_pow2<N>(bufA, bufB, tw);
    const float oct_scale = sqrtf(static_cast<float>(1 << oct));  // constantq.py:98
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = warp; b < bins_per_octave; b += (blockDim.x >> 5)) {
        float re = 0.0f, im = 0.0f;
        for (int e = rowptr[b] + lane; e < rowptr[b + 1]; e += 32) {
            const float2 w = val[e];
            const float2 x = X[col[e]];
            re += w.x * x.x - w.y * x.y;
            im += w.x * x.y + w.y * x.x;
        }
        for (int o = 16; o > 0; o >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, o);
            im += __shfl_xor_sync(0xffffffffu, im, o);
        }
        if (lane == 0) {
            const int gb = n_bins - bins_per_octave * (oct + 1) + b;  // octave 0 = top octave (constantq.py:166-186)
            if (gb >= 0) C[static_cast<long long>(gb) * T + t] = hypotf(re * oct_scale, im * oct_scale) * inv_sqrt_len[gb];
        }
    }
}

__global__ void __launch_bounds__(256) chroma_fold_kernel(const float* __restrict__ C, const float* __restrict__ fold, int n_bins,
                                                          int n_chroma, int T, float threshold, float* __restrict__ chroma,
                                                          int* __restrict__ max_bits) {
    __shared__ float red[8];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.0f;
    if (idx < n_chroma * T) {
        const int c = idx / T, t = idx - c * T;
        for (int b = 0; b < n_bins; ++b) {
            const float f = fold[c * n_bins + b];
            if (f != 0.0f) v = fmaf(f, C[static_cast<long long>(b) * T + t], v);
        }
        if (v < threshold) v = 0.0f;
        chroma[idx] = v;
    }
    float m = v;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
        atomicMax(max_bits, __float_as_int(fmaxf(m, 0.0f)));  // non-negative floats order like their bit patterns
    }
}

__global__ void chroma_norm_kernel(float* __restrict__ chroma, int n, const int* __restrict__ max_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) chroma[i] = chroma[i] / __int_as_float(*max_bits);
}

// ---- tuning estimate: rosa/pitch.py:9-120 -----------------------------------------------------------------------
// piptrack: hann STFT (n_fft 2048, hop 512), parabolic peak interpolation, per-frame relative threshold, band limit;
// one CTA per frame, the spectrum never leaves shared memory.  Only the selected (pitch > 0) bins matter downstream and
// the median / histogram that consume them are order independent, so they are appended to a compact list through one
// atomic counter instead of being written as dense [frames][1025] planes (the dense form made the single-CTA
// selection kernel scan 1.5 M zeros 33 times: 6 ms at configs[1]).
constexpr int kPipFft = 2048, kPipHop = 512, kPipBins = kPipFft / 2 + 1;

__global__ void __launch_bounds__(256) piptrack_kernel(const float* __restrict__ y, long long n, float sr, float fmin, float fmax,
                                                       float threshold, float* __restrict__ pitch, float* __restrict__ mag,
                                                       unsigned int* __restrict__ count, unsigned int cap) {
    __shared__ float2 bufA[kPipFft];
    __shared__ float2 bufB[kPipFft];
    __shared__ float2 tw[kPipFft / 2];
    __shared__ float red[8];
    const int t = blockIdx.x;
    for (int i = threadIdx.x; i < kPipFft / 2; i += blockDim.x) {
        float s, c;
        sincospif(static_cast<float>(i) / static_cast<float>(kPipFft / 2), &s, &c);
        tw[i] = make_float2(c, -s);
    }
    const long long start = static_cast<long long>(t) * kPipHop - kPipFft / 2;
    for (int i = threadIdx.x; i < kPipFft; i += blockDim.x) {
        const float w = 0.5f - 0.5f * cospif(2.0f * static_cast<float>(i) / static_cast<float>(kPipFft));  // periodic hann
        bufA[i] = make_float2(y[reflect_idx(start + i, n)] * w, 0.0f);
    }
    __syncthreads();
    const float2* X = fft_pow2<kPipFft>(bufA, bufB, tw);
    float* S = reinterpret_cast<float*>(X == bufA ? bufB : bufA);   // the other buffer is free now
    float mx = 0.0f;
    for (int k = threadIdx.x; k < kPipBins; k += blockDim.x) {
        const float m = hypotf(X[k].x, X[k].y);
        S[k] = m;
        mx = fmaxf(mx, m);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    const float ref = threshold * mx;
    const float fhi = fminf(fmax, sr * 0.5f), flo = fmaxf(fmin, 0.0f);
    for (int k = threadIdx.x; k < kPipBins; k += blockDim.x) {
        const float s0 = S[k];
        float p = 0.0f, m = 0.0f;
        // torch.linspace(0, sr/2, 1025)[k]
        const float step = (sr * 0.5f) / static_cast<float>(kPipBins - 1);
        const float fk = k < kPipBins / 2 ? step * static_cast<float>(k) : sr * 0.5f - step * static_cast<float>(kPipBins - 1 - k);
        if (fk >= flo && fk < fhi) {
            // localmax(S * (S > ref)): x > left neighbour (0 outside) and x >= right neighbour (0 outside)
            const float x = s0 > ref ? s0 : 0.0f;
            const float xl = k > 0 ? (S[k - 1] > ref ? S[k - 1] : 0.0f) : 0.0f;
            const float xr = k < kPipBins - 1 ? (S[k + 1] > ref ? S[k + 1] : 0.0f) : 0.0f;
            if (x > xl && x >= xr) {
                float avg = 0.0f, shift = 0.0f;
                if (k > 0 && k < kPipBins - 1) {
                    avg = 0.5f * (S[k + 1] - S[k - 1]);
                    float sh = 2.0f * s0 - S[k + 1] - S[k - 1];
                    sh = sh + (fabsf(sh) < 1.17549435e-38f ? 1.0f : 0.0f);
                    shift = avg / sh;
                }
                p = (static_cast<float>(k) + shift) * sr / static_cast<float>(kPipFft);
                m = s0 + 0.5f * avg * shift;
            }
        }
        if (p > 0.0f) {
            const unsigned int slot = atomicAdd(count, 1u);
            if (slot < cap) {
                pitch[slot] = p;
                mag[slot] = m;
            }
        }
    }
}

// Single CTA: lower median of mag over pitch > 0 (bitwise bisection on the non-negative float patterns), then the
// 1/resolution-bin histogram of the pitch residuals relative to the bin grid over mag >= median, first arg-max ->
// tuning = linspace(-0.5, 0.5, bins + 1)[argmax]   (pitch.py:12-24, 100-120)
__global__ void __launch_bounds__(1024) tuning_kernel(const float* __restrict__ pitch, const float* __restrict__ mag,
                                                      const unsigned int* __restrict__ count, unsigned int cap,
                                                      int bins_per_octave, int bins, float* __restrict__ tuning) {
    __shared__ unsigned long long s_cnt;
    __shared__ int hist[512];
    const int tid = threadIdx.x;
    const long long total = *count < cap ? *count : cap;
    auto block_count = [&](unsigned int trial, bool count_all) {
        unsigned long long c = 0;
        for (long long i = tid; i < total; i += blockDim.x)
            if (pitch[i] > 0.0f && (count_all || __float_as_uint(mag[i]) < trial)) ++c;
        __syncthreads();
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((tid & 31) == 0) atomicAdd(&s_cnt, c);
        __syncthreads();
        return s_cnt;
    };
    const unsigned long long K = block_count(0u, true);
    if (K == 0) {
        if (tid == 0) *tuning = 0.0f;
        return;
    }
    const unsigned long long k = (K - 1) / 2;   // torch.median: the lower of the two middle values
    unsigned int prefix = 0;
    for (int bit = 30; bit >= 0; --bit) {
        const unsigned int trial = prefix | (1u << bit);
        if (block_count(trial, false) <= k) prefix = trial;
    }
    const float thr = __uint_as_float(prefix);
    for (int i = tid; i < bins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const float a440 = 440.0f / 16.0f;
    for (long long i = tid; i < total; i += blockDim.x) {
        const float f = pitch[i];
        if (f > 0.0f && mag[i] >= thr) {
            float r = static_cast<float>(bins_per_octave) * log2f(f / a440);
            r = r - floorf(r);                 // python-style % 1.0
            if (r >= 0.5f) r -= 1.0f;
            // torch.histc(bins, min=-0.5, max=0.5): bin = (x - min) / (max - min) * bins, x == max in the last bin
            int b = static_cast<int>((r + 0.5f) * static_cast<float>(bins));
            if (b >= bins) b = bins - 1;
            if (b >= 0) atomicAdd(&hist[b], 1);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int i = 1; i < bins; ++i)
            if (hist[i] > hist[best]) best = i;
        // torch.linspace(-0.5, 0.5, bins + 1)[best]
        const float step = 1.0f / static_cast<float>(bins);
        const int steps = bins + 1;
        *tuning = best < steps / 2 ? -0.5f + step * static_cast<float>(best) : 0.5f - step * static_cast<float>(steps - 1 - best);
    }
}

// ---- CENS post-processing: rosa/spectral.py:239-280 after chroma_cqt(norm=False) -----------------------------------
// L1 normalise each frame -> natural-spline quantiser curve (host-designed knots / coefficients) -> smooth step
__global__ void cens_quantise_kernel(const float* __restrict__ chroma, int nc, int T, const float* __restrict__ kx,
                                     const float* __restrict__ coef, int nk, float* __restrict__ q) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nc * T) return;
    const int t = idx % T;
    float l1 = 0.0f;
    for (int c = 0; c < nc; ++c) l1 += fabsf(chroma[c * T + t]);
    const float v = chroma[idx] / l1;
    // torch.bucketize(v, kx) - 1 (right=False: first index with kx[i] >= v), clamped to [0, nk - 2]
    int lo = 0, hi = nk;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (kx[mid] < v) lo = mid + 1; else hi = mid;
    }
    int i = lo - 1;
    i = i < 0 ? 0 : (i > nk - 2 ? nk - 2 : i);
    const int ni = nk - 1;
    const float f = v - kx[i];
    const float w = coef[i] + (coef[ni + i] + (coef[2 * ni + i] + coef[3 * ni + i] * f) * f) * f;
    // step_function(w, h = 0.25, alpha = 20)
    const float fl = floorf(w - 0.5f);
    const float r = (w - 0.5f) - fl - 0.5f;
    const float m = 1.0f / (1.0f + expf(-20.0f)) - 0.5f;
    q[idx] = 0.25f * (fl + 1.0f / (2.0f * m) * 1.0f / (1.0f + expf(-40.0f * r)));
}

// temporal smoothing (zero-padded 'same' correlation with `win`, L taps) and L2 normalisation per frame
__global__ void cens_smooth_kernel(const float* __restrict__ q, int nc, int T, const float* __restrict__ win, int L,
                                   float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int left = (L - 1) / 2;   // torch 'same' padding: total L - 1, the extra one on the right for even L
    float v[16];
    float ss = 0.0f;
    for (int c = 0; c < nc; ++c) {
        float acc = 0.0f;
        for (int k = 0; k < L; ++k) {
            const int tt = t + k - left;
            if (tt >= 0 && tt < T) acc = fmaf(win[k], q[c * T + tt], acc);
        }
        v[c] = acc;
        ss += acc * acc;
    }
    const float nrm = sqrtf(ss);
    for (int c = 0; c < nc; ++c) out[c * T + t] = v[c] / nrm;
}

size_t up256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" size_t mb_chroma_workspace_bytes(int64_t n_samples, int n_bins, int hop) {
    if (n_samples <= 0 || hop <= 0) return 0;
    const size_t T = static_cast<size_t>(n_samples / hop);
    // decimated signals sum to < n samples; CQT magnitudes; max word
    return up256(sizeof(float) * static_cast<size_t>(n_samples + 64)) + up256(sizeof(float) * n_bins * T) + 256;
}

extern "C" int mb_chroma_cqt(const float* audio, int64_t n, int hop, int n_fft, int n_octaves, int bins_per_octave,
                             const float* decim_kernel, int decim_taps, int decim_width, const int32_t* basis_rowptr,
                             const int32_t* basis_col, const float* basis_val, const float* inv_sqrt_len, const float* fold,
                             int n_chroma, float threshold, int normalize, float* cqt_mag, float* chroma, void* workspace,
                             size_t workspace_bytes, mb_stream stream_) {
    MB_REQUIRE(audio && decim_kernel && basis_rowptr && basis_col && basis_val && inv_sqrt_len && fold && chroma && workspace,
               "mb_chroma_cqt: null argument");
    MB_REQUIRE(n_octaves >= 1 && n_octaves <= 8, "mb_chroma_cqt: 1..8 octaves, got %d", n_octaves);
    MB_REQUIRE(n_fft == 512 || n_fft == 1024 || n_fft == 2048, "mb_chroma_cqt: n_fft %d unsupported (512, 1024, 2048)", n_fft);
    MB_REQUIRE(hop > 0 && (hop % (1 << (n_octaves - 1))) == 0,
               "mb_chroma_cqt: hop_length must be a multiple of 2^%d for a %d-octave CQT (constantq.py:66-72)", n_octaves - 1, n_octaves);
    MB_REQUIRE(n > 0 && n % hop == 0, "mb_chroma_cqt: need a whole number of hops, got %lld samples", static_cast<long long>(n));
    MB_REQUIRE(decim_taps > 0 && decim_taps <= 64, "mb_chroma_cqt: decimation kernel of %d taps unsupported", decim_taps);
    MB_REQUIRE((n >> (n_octaves - 1)) > n_fft / 2, "mb_chroma_cqt: audio too short for reflect padding in the lowest octave");
    const int n_bins = n_octaves * bins_per_octave;
    const int T = static_cast<int>(n / hop);
    const size_t need = mb_chroma_workspace_bytes(n, n_bins, hop);
    if (workspace_bytes < need) {
        set_error("mb_chroma_cqt: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
        return MB_ENOMEM;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    float* ybuf = reinterpret_cast<float*>(base);
    float* Cws = reinterpret_cast<float*>(base + up256(sizeof(float) * static_cast<size_t>(n + 64)));
    int* max_bits = reinterpret_cast<int*>(base + need - 256);
    float* C = cqt_mag ? cqt_mag : Cws;

    OctaveSrc src;
    src.y[0] = audio;
    src.n[0] = n;
    float* next = ybuf;
    for (int i = 1; i < n_octaves; ++i) {
        const long long n_in = src.n[i - 1], n_out = (n_in + 1) / 2;  // ceil(new_freq * length / orig_freq)
        const int blocks = static_cast<int>((n_out + 255) / 256 < 148 * 8 ? (n_out + 255) / 256 : 148 * 8);
        decimate_kernel<<<blocks, 256, 0, stream>>>(src.y[i - 1], n_in, next, n_out, decim_kernel, decim_taps, decim_width,
                                                    1.41421356237309515f);  // my_y *= np.sqrt(2), constantq.py:84
        src.y[i] = next;
        src.n[i] = n_out;
        next += (n_out + 3) / 4 * 4;
    }
    for (int i = n_octaves; i < 8; ++i) { src.y[i] = nullptr; src.n[i] = 0; }
    dim3 grid(T, n_octaves);
    const float2* val = reinterpret_cast<const float2*>(basis_val);
    if (n_fft == 512)
        cqt_kernel<512><<<grid, 256, 0, stream>>>(src, hop, T, bins_per_octave, n_bins, basis_rowptr, basis_col, val, inv_sqrt_len, C);
    else if (n_fft == 1024)
        cqt_kernel<1024><<<grid, 256, 0, stream>>>(src, hop, T, bins_per_octave, n_bins, basis_rowptr, basis_col, val, inv_sqrt_len, C);
    else
        cqt_kernel<2048><<<grid, 256, 0, stream>>>(src, hop, T, bins_per_octave, n_bins, basis_rowptr, basis_col, val, inv_sqrt_len, C);
    MB_CUDA(cudaMemsetAsync(max_bits, 0, sizeof(int), stream));
    const int total = n_chroma * T;
    chroma_fold_kernel<<<(total + 255) / 256, 256, 0, stream>>>(C, fold, n_bins, n_chroma, T, threshold, chroma, max_bits);
    if (normalize) chroma_norm_kernel<<<(total + 255) / 256, 256, 0, stream>>>(chroma, total, max_bits);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" size_t mb_tuning_workspace_bytes(int64_t n_samples) {
    if (n_samples <= 0) return 0;
    const size_t frames = static_cast<size_t>(n_samples / kPipHop);
    return 2 * up256(sizeof(float) * frames * kPipBins) + 256;   // worst case every bin selected; + the counter
}

extern "C" int mb_estimate_tuning(const float* audio, int64_t n, float sr, int bins_per_octave, int bins, float* tuning,
                                  void* workspace, size_t workspace_bytes, mb_stream stream_) {
    MB_REQUIRE(audio && tuning && workspace, "mb_estimate_tuning: null argument");
    MB_REQUIRE(n >= 4 * kPipFft && n % kPipHop == 0, "mb_estimate_tuning: need a multiple of %d samples (>= %d)", kPipHop, 4 * kPipFft);
    MB_REQUIRE(bins >= 1 && bins <= 512, "mb_estimate_tuning: %d histogram bins unsupported (1..512)", bins);
    const size_t need = mb_tuning_workspace_bytes(n);
    if (workspace_bytes < need) {
        set_error("mb_estimate_tuning: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
        return MB_ENOMEM;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int frames = static_cast<int>(n / kPipHop);   // stft(center=True) has n / hop + 1 columns, spectrogram() drops the last
    const size_t plane = (need - 256) / 2;
    float* pitch = static_cast<float*>(workspace);
    float* mag = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + plane);
    unsigned int* count = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) + 2 * plane);
    const unsigned int cap = static_cast<unsigned int>(static_cast<size_t>(frames) * kPipBins);
    MB_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned int), stream));
    piptrack_kernel<<<frames, 256, 0, stream>>>(audio, n, sr, 150.0f, 4000.0f, 0.1f, pitch, mag, count, cap);
    tuning_kernel<<<1, 1024, 0, stream>>>(pitch, mag, count, cap, bins_per_octave, bins, tuning);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

extern "C" int mb_chroma_cens_post(const float* chroma_raw, int n_chroma, int T, const float* knots_x, const float* coef, int n_knots,
                                   const float* smooth_win, int win_len, float* scratch, float* out, mb_stream stream_) {
    MB_REQUIRE(chroma_raw && knots_x && coef && smooth_win && scratch && out, "mb_chroma_cens_post: null argument");
    MB_REQUIRE(n_chroma >= 1 && n_chroma <= 16 && T > 0 && n_knots >= 2 && win_len >= 1, "mb_chroma_cens_post: bad argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int total = n_chroma * T;
    cens_quantise_kernel<<<(total + 255) / 256, 256, 0, stream>>>(chroma_raw, n_chroma, T, knots_x, coef, n_knots, scratch);
    cens_smooth_kernel<<<(T + 127) / 128, 128, 0, stream>>>(scratch, n_chroma, T, smooth_win, win_len, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}
