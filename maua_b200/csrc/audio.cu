// Audio front end of the render path on the GPU: STFT -> HPSS (31-tap medians, soft masks) -> iSTFT ->
// mel power spectrogram -> dB flux onset envelope (+ normalisation, strict-local-maximum peak picking) and
// per-frame RMS.  Replaces the torch-native features of
// maua/audiovisual/audioreactive/selfsupervised/features/audio.py:20-37 (percussive, onsets, rms),
// rosa/spectral.py:10-32,59-110,120-161 (stft, istft, spectrogram, mel, softmask, hpss),
// rosa/beat.py:10-23 (onset_strength), rosa/convert.py:7-12 (power_to_db), processing.py:53-56,75-85
// (normalize, median_filter2d) and the peak rule of audioreactive/signal.py:69-76.
//
// Shapes (n_fft 2048, hop 1024): audio [N] -> T = N / hop frames; spectra are stored frame-major
// [T+1][1025] so both median directions read contiguous or coalesced data.  One CTA per frame stages the
// 2048-sample frame in shared memory (each sample is read from HBM once per pass), runs a radix-2 Stockham
// FFT there, and fuses the consumer (magnitude + RMS, or power -> mel filterbank).  The whole feature pass is
// a handful of launches over a few MB: latency bound, not bandwidth bound.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace mb {
namespace {

constexpr int kNfft = 2048;
constexpr int kHop = 1024;
constexpr int kBins = kNfft / 2 + 1;  // 1025
constexpr int kMedian = 31;
constexpr int kMels = 128;

__device__ __forceinline__ int reflect(long long i, long long n) {
    // torch 'reflect' padding index (no edge repeat); valid for |overshoot| < n
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return static_cast<int>(i);
}

// In-place-on-two-buffers radix-2 Stockham FFT of 2048 complex points held in shared memory.
// tw[k] = exp(-2 pi i k / 2048), k < 1024.  Returns the buffer holding the result.
template <bool INVERSE>
__device__ float2* fft2048(float2* x, float2* y, const float2* tw) {
    for (int ns = 1; ns < kNfft; ns <<= 1) {
        for (int j = threadIdx.x; j < kNfft / 2; j += blockDim.x) {
            const int k = j & (ns - 1);
            float2 w = tw[k * (kNfft / 2 / ns)];
            if (INVERSE) w.y = -w.y;
            const float2 a = x[j], b0 = x[j + kNfft / 2];
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

// mode 0: centered STFT of `audio` -> spec [T+1][1025] (complex), mag [T+1][1025], rms [T] (unwindowed frames)
// mode 1: centered STFT -> power spectrum -> mel [T][128] (frames t < T only)
__global__ void __launch_bounds__(256) stft_kernel(const float* __restrict__ audio, long long n, int T, int mode,
                                                   const float2* __restrict__ tw_g, const float* __restrict__ window,
                                                   float2* __restrict__ spec, float* __restrict__ mag, float* __restrict__ rms,
                                                   const float* __restrict__ melfb, float* __restrict__ mel) {
    __shared__ float2 bufA[kNfft];
    __shared__ float2 bufB[kNfft];
    __shared__ float2 tw[kNfft / 2];
    __shared__ float red[32];
    const int t = blockIdx.x;
    for (int i = threadIdx.x; i < kNfft / 2; i += blockDim.x) tw[i] = tw_g[i];
    float ss = 0.0f;
    const long long start = static_cast<long long>(t) * kHop - kNfft / 2;
    for (int i = threadIdx.x; i < kNfft; i += blockDim.x) {
        const float v = audio[reflect(start + i, n)];
        ss += v * v;
        bufA[i] = make_float2(v * window[i], 0.0f);
    }
    if (mode == 0 && t < T) {  // RMS of the raw frame (features/audio.py:31-37 uses the same framing)
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    }
    __syncthreads();
    if (mode == 0 && t < T && threadIdx.x == 0) {
        float s = 0.0f;
        for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) s += red[i];
        rms[t] = sqrtf(s / static_cast<float>(kNfft));
    }
    float2* X = fft2048<false>(bufA, bufB, tw);
    if (mode == 0) {
        for (int k = threadIdx.x; k < kBins; k += blockDim.x) {
            const float2 v = X[k];
            spec[static_cast<long long>(t) * kBins + k] = v;
            mag[static_cast<long long>(t) * kBins + k] = hypotf(v.x, v.y);
        }
    } else {
        float* pw = reinterpret_cast<float*>(X == bufA ? bufB : bufA);  // the other buffer is free now
        for (int k = threadIdx.x; k < kBins; k += blockDim.x) {
            const float m = hypotf(X[k].x, X[k].y);
            pw[k] = m * m;
        }
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int b = warp; b < kMels; b += (blockDim.x >> 5)) {
            const float* fb = melfb + static_cast<long long>(b) * kBins;
            float acc = 0.0f;
            for (int k = lane; k < kBins; k += 32) acc = fmaf(fb[k], pw[k], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) mel[static_cast<long long>(t) * kMels + b] = acc;
        }
    }
}

__device__ __forceinline__ float median31(const float (&v)[kMedian]) {
    // rank selection: the element with exactly 15 elements ordered before it (ties broken by index)
    float med = v[0];
#pragma unroll
    for (int i = 0; i < kMedian; ++i) {
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < kMedian; ++j) cnt += (v[j] < v[i]) || (v[j] == v[i] && j < i);
        if (cnt == kMedian / 2) med = v[i];
    }
    return med;
}

// which: 0 -> harmonic, 1 -> percussive.  out = D * softmask (power 2) with the reference's margin rule.
__global__ void __launch_bounds__(128) hpss_kernel(const float2* __restrict__ spec, const float* __restrict__ mag, int nT,
                                                   float margin, int which, float2* __restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (f >= kBins) return;
    float v[kMedian];
#pragma unroll
    for (int i = 0; i < kMedian; ++i) v[i] = mag[static_cast<long long>(reflect(t + i - kMedian / 2, nT)) * kBins + f];
    const float harm = median31(v);  // median along time
#pragma unroll
    for (int i = 0; i < kMedian; ++i) v[i] = mag[static_cast<long long>(t) * kBins + reflect(f + i - kMedian / 2, kBins)];
    const float perc = median31(v);  // median along frequency
    const float x = which ? perc : harm;
    const float xr = (which ? harm : perc) * margin;
    float z = fmaxf(x, xr);
    const bool bad = z < 1.17549435e-38f;
    if (bad) z = 1.0f;
    const float m = (x / z) * (x / z), r = (xr / z) * (xr / z);
    const float mask = bad ? 0.0f : m / (m + r);
    const float2 d = spec[static_cast<long long>(t) * kBins + f];
    out[static_cast<long long>(t) * kBins + f] = make_float2(d.x * mask, d.y * mask);
}

// inverse STFT, step 1: per frame Hermitian extension -> inverse FFT -> * window -> frames [T+1][2048]
__global__ void __launch_bounds__(256) istft_frame_kernel(const float2* __restrict__ spec, const float2* __restrict__ tw_g,
                                                          const float* __restrict__ window, float* __restrict__ frames) {
    __shared__ float2 bufA[kNfft];
    __shared__ float2 bufB[kNfft];
    __shared__ float2 tw[kNfft / 2];
    const int t = blockIdx.x;
    for (int i = threadIdx.x; i < kNfft / 2; i += blockDim.x) tw[i] = tw_g[i];
    const float2* s = spec + static_cast<long long>(t) * kBins;
    for (int k = threadIdx.x; k < kNfft; k += blockDim.x) {
        float2 v;
        if (k <= kNfft / 2) {
            v = s[k];
            if (k == 0 || k == kNfft / 2) v.y = 0.0f;  // c2r ignores the imaginary part of DC / Nyquist
        } else {
            v = s[kNfft - k];
            v.y = -v.y;
        }
        bufA[k] = v;
    }
    __syncthreads();
    float2* X = fft2048<true>(bufA, bufB, tw);
    for (int i = threadIdx.x; i < kNfft; i += blockDim.x)
        frames[static_cast<long long>(t) * kNfft + i] = X[i].x * (1.0f / kNfft) * window[i];
}

// step 2: overlap-add of the two frames covering each sample, divided by the window envelope; center crop
__global__ void istft_ola_kernel(const float* __restrict__ frames, const float* __restrict__ window, long long n,
                                 float* __restrict__ y) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long m = i + kNfft / 2;
        const long long t1 = m / kHop, t0 = t1 - 1;
        const int o1 = static_cast<int>(m - t1 * kHop), o0 = o1 + kHop;
        const float num = frames[t0 * kNfft + o0] + frames[t1 * kNfft + o1];
        const float den = window[o0] * window[o0] + window[o1] * window[o1];
        y[i] = num / den;
    }
}

// Single CTA: mel power [T][128] -> dB (global top_db floor) -> positive flux, mean over bands, shifted by 2
// frames -> min/max normalisation -> strict local maxima (index-clamped neighbours), compacted in order.
__global__ void __launch_bounds__(1024) flux_kernel(const float* __restrict__ mel, int T, float* __restrict__ env_tmp,
                                                    float* __restrict__ onsets, int* __restrict__ peak_idx,
                                                    int* __restrict__ n_peaks) {
    __shared__ float red[32];
    __shared__ float s_val[2];
    __shared__ int s_cnt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto block_max = [&](float v) {
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        float r = red[0];
        for (int i = 1; i < 32; ++i) r = fmaxf(r, red[i]);
        return r;
    };
    // 1. global maximum of the dB spectrogram
    float mx = -3.0e38f;
    for (long long i = tid; i < static_cast<long long>(T) * kMels; i += blockDim.x)
        mx = fmaxf(mx, 10.0f * log10f(fmaxf(1e-10f, fabsf(mel[i]))));
    mx = block_max(mx);
    const float floor_db = mx - 80.0f;
    // 2. flux (one warp per frame)
    for (int j = warp; j < T; j += 32) {
        float acc = 0.0f;
        if (j >= 2) {
            const float* a = mel + static_cast<long long>(j - 1) * kMels;
            const float* b = mel + static_cast<long long>(j - 2) * kMels;
            for (int k = lane; k < kMels; k += 32) {
                const float da = fmaxf(10.0f * log10f(fmaxf(1e-10f, fabsf(a[k]))), floor_db);
                const float db = fmaxf(10.0f * log10f(fmaxf(1e-10f, fabsf(b[k]))), floor_db);
                acc += fmaxf(da - db, 0.0f);
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            acc = acc / static_cast<float>(kMels);
        }
        if (lane == 0) env_tmp[j] = acc;
    }
    __syncthreads();
    // 3. normalise: (x - min) / (max(x - min) + 1e-8)
    float lo = 3.0e38f, hi = -3.0e38f;
    for (int j = tid; j < T; j += blockDim.x) {
        lo = fminf(lo, env_tmp[j]);
        hi = fmaxf(hi, env_tmp[j]);
    }
    hi = block_max(hi);
    lo = -block_max(-lo);
    if (tid == 0) { s_val[0] = lo; s_val[1] = (hi - lo) + 1e-8f; s_cnt = 0; }
    __syncthreads();
    for (int j = tid; j < T; j += blockDim.x) onsets[j] = (env_tmp[j] - s_val[0]) / s_val[1];
    __syncthreads();
    // 4. peaks, in index order (chunks of blockDim frames, ballot + prefix inside the block)
    __shared__ int warp_cnt[32];
    for (int base = 0; base < T; base += blockDim.x) {
        const int j = base + tid;
        bool pk = false;
        if (j < T) {
            const float c = onsets[j], l = onsets[j > 0 ? j - 1 : 0], r = onsets[j < T - 1 ? j + 1 : T - 1];
            pk = c > r && c > l;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pk);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = s_cnt;
        for (int w = 0; w < warp; ++w) off += warp_cnt[w];
        if (pk) peak_idx[off + __popc(bal & ((1u << lane) - 1))] = j;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 32; ++w) tot += warp_cnt[w];
            s_cnt += tot;
        }
        __syncthreads();
    }
    if (tid == 0) *n_peaks = s_cnt;
}

size_t align_up(size_t v) { return (v + 255) / 256 * 256; }

struct AudioWs {
    size_t tw, window, spec, mag, perc, frames, yperc, mel, env, total;
};
AudioWs audio_ws(long long n) {
    const long long T = n / kHop;
    AudioWs w;
    size_t off = 0;
    w.tw = off; off = align_up(off + sizeof(float2) * kNfft / 2);
    w.window = off; off = align_up(off + sizeof(float) * kNfft);
    w.spec = off; off = align_up(off + sizeof(float2) * (T + 1) * kBins);
    w.mag = off; off = align_up(off + sizeof(float) * (T + 1) * kBins);
    w.perc = off; off = align_up(off + sizeof(float2) * (T + 1) * kBins);
    w.frames = off; off = align_up(off + sizeof(float) * (T + 1) * kNfft);
    w.yperc = off; off = align_up(off + sizeof(float) * n);
    w.mel = off; off = align_up(off + sizeof(float) * T * kMels);
    w.env = off; off = align_up(off + sizeof(float) * T);
    w.total = off;
    return w;
}

__global__ void audio_tables_kernel(float2* tw, float* window) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kNfft / 2) {
        float s, c;
        sincospif(static_cast<float>(i) / static_cast<float>(kNfft / 2), &s, &c);  // angle = 2 pi i / 2048
        tw[i] = make_float2(c, -s);
    }
    if (i < kNfft) {
        // periodic hann, as torch.hann_window(2048): 0.5 - 0.5 cos(2 pi i / N)
        window[i] = 0.5f - 0.5f * cospif(2.0f * static_cast<float>(i) / static_cast<float>(kNfft));
    }
}


// ---- spectral descriptors of the torch-native feature list (features/audio.py:59-133) ----------------------------
// spectral_flatness (:123-133): per frame exp(mean(log(max(amin, |X|^2)))) / mean(max(amin, |X|^2)) over the 1025 bins.
__global__ void __launch_bounds__(256) flatness_kernel(const float* __restrict__ mag, int T, float amin, float power, float* __restrict__ out) {
    __shared__ float s_log[8], s_sum[8];
    const int t = blockIdx.x;
    const float* m = mag + static_cast<long long>(t) * kBins;
    float lsum = 0.0f, sum = 0.0f;
    for (int k = threadIdx.x; k < kBins; k += blockDim.x) {
        const float v = fmaxf(amin, powf(m[k], power));
        lsum += logf(v);
        sum += v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if ((threadIdx.x & 31) == 0) { s_log[threadIdx.x >> 5] = lsum; s_sum[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.0f, b = 0.0f;
        for (int i = 0; i < 8; ++i) { a += s_log[i]; b += s_sum[i]; }
        out[t] = expf(a / static_cast<float>(kBins)) / (b / static_cast<float>(kBins));
    }
}

// spectral_contrast (:69-120): per frame and band, mean of the `cnt` smallest (valley) and `cnt` largest (peak) magnitudes
// of the band's bins [lo, hi).  Selection by rank counting (ties broken by index): the same multiset the reference's sort
// picks.  peak / valley: [n_bands][T].
struct ContrastBands { int lo[8], hi[8], cnt[8], n; };
__global__ void __launch_bounds__(256) contrast_kernel(const float* __restrict__ mag, int T, ContrastBands bands, float* __restrict__ peak,
                                                        float* __restrict__ valley) {
    __shared__ float row[kBins];
    __shared__ float s_v[8], s_p[8];
    const int t = blockIdx.x;
    for (int k = threadIdx.x; k < kBins; k += blockDim.x) row[k] = mag[static_cast<long long>(t) * kBins + k];
    __syncthreads();
    for (int b = 0; b < bands.n; ++b) {
        const int lo = bands.lo[b], n = bands.hi[b] - lo, cnt = bands.cnt[b];
        float v = 0.0f, pk = 0.0f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float x = row[lo + i];
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const float y = row[lo + j];
                rank += (y < x) || (y == x && j < i);
            }
            if (rank < cnt) v += x;
            if (rank >= n - cnt) pk += x;
        }
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            pk += __shfl_xor_sync(0xffffffffu, pk, o);
        }
        if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = v; s_p[threadIdx.x >> 5] = pk; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.0f, c = 0.0f;
            for (int i = 0; i < 8; ++i) { a += s_v[i]; c += s_p[i]; }
            valley[static_cast<long long>(b) * T + t] = a / static_cast<float>(cnt);
            peak[static_cast<long long>(b) * T + t] = c / static_cast<float>(cnt);
        }
        __syncthreads();
    }
}

// out[t][b] = peak[b][t] - valley[b][t]
__global__ void contrast_diff_kernel(const float* __restrict__ peak, const float* __restrict__ valley, int T, int nb, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * nb) return;
    const int b = idx / T, t = idx - b * T;
    out[static_cast<long long>(t) * nb + b] = peak[idx] - valley[idx];
}

// power_to_db (rosa/convert.py:7-12, ref 1, amin 1e-10, top_db 80 below the GLOBAL maximum): one block, n values in place
__global__ void __launch_bounds__(1024) power_to_db_kernel(float* __restrict__ x, long long n, float top_db) {
    __shared__ float s_max[32];
    float mx = -3.0e38f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float db = 10.0f * log10f(fmaxf(1e-10f, x[i]));
        x[i] = db;
        mx = fmaxf(mx, db);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 32; ++i) mx = fmaxf(mx, s_max[i]);
        s_max[0] = mx;
    }
    __syncthreads();
    const float floor_db = s_max[0] - top_db;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) x[i] = fmaxf(x[i], floor_db);
}

// mfcc (:59-64): orthonormal DCT-II of the dB mel spectrum along the 128 mel bands, first n_mfcc coefficients.
// db: [T][128] frame-major, out: [T][n_mfcc].
__global__ void __launch_bounds__(128) mfcc_dct_kernel(const float* __restrict__ db, int T, int n_mfcc, float* __restrict__ out) {
    __shared__ float row[kMels];
    const int t = blockIdx.x;
    row[threadIdx.x] = db[static_cast<long long>(t) * kMels + threadIdx.x];
    __syncthreads();
    for (int k = threadIdx.x; k < n_mfcc; k += blockDim.x) {
        float acc = 0.0f;
        for (int m = 0; m < kMels; ++m) acc = fmaf(row[m], cospif(static_cast<float>(k * (2 * m + 1)) / static_cast<float>(2 * kMels)), acc);
        out[static_cast<long long>(t) * n_mfcc + k] = acc * (k == 0 ? rsqrtf(static_cast<float>(kMels)) : sqrtf(2.0f / static_cast<float>(kMels)));
    }
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" size_t mb_audio_workspace_bytes(int64_t n_samples) {
    if (n_samples < 16 * kHop) return 0;
    return audio_ws(n_samples).total;
}

extern "C" int mb_audio_onsets_rms(const float* audio, int64_t n, const float* mel_filterbank, float margin,
                                   float* onsets, float* rms, int32_t* peak_idx, int32_t* n_peaks, float* percussive_out,
                                   void* workspace, size_t workspace_bytes, mb_stream stream_) {
    MB_REQUIRE(audio && mel_filterbank && onsets && rms && peak_idx && n_peaks && workspace, "mb_audio_onsets_rms: null argument");
    MB_REQUIRE(n >= 16 * kHop && n % kHop == 0, "mb_audio_onsets_rms: need a multiple of %d samples (>= %d), got %lld", kHop,
               16 * kHop, static_cast<long long>(n));
    MB_REQUIRE(margin != 1.0f, "mb_audio_onsets_rms: margin == 1 (split_zeros branch) is not implemented");
    const AudioWs w = audio_ws(n);
    if (workspace_bytes < w.total) {
        set_error("mb_audio_onsets_rms: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
        return MB_ENOMEM;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int T = static_cast<int>(n / kHop);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    float2* tw = reinterpret_cast<float2*>(base + w.tw);
    float* window = reinterpret_cast<float*>(base + w.window);
    float2* spec = reinterpret_cast<float2*>(base + w.spec);
    float* mag = reinterpret_cast<float*>(base + w.mag);
    float2* perc = reinterpret_cast<float2*>(base + w.perc);
    float* frames = reinterpret_cast<float*>(base + w.frames);
    float* yperc = reinterpret_cast<float*>(base + w.yperc);
    float* mel = reinterpret_cast<float*>(base + w.mel);
    float* env = reinterpret_cast<float*>(base + w.env);

    audio_tables_kernel<<<kNfft / 256, 256, 0, stream>>>(tw, window);
    stft_kernel<<<T + 1, 256, 0, stream>>>(audio, n, T, 0, tw, window, spec, mag, rms, nullptr, nullptr);
    dim3 hg((kBins + 127) / 128, T + 1);
    hpss_kernel<<<hg, 128, 0, stream>>>(spec, mag, T + 1, margin, 1, perc);
    istft_frame_kernel<<<T + 1, 256, 0, stream>>>(perc, tw, window, frames);
    istft_ola_kernel<<<148 * 4, 256, 0, stream>>>(frames, window, n, yperc);
    stft_kernel<<<T, 256, 0, stream>>>(yperc, n, T, 1, tw, window, nullptr, nullptr, nullptr, mel_filterbank, mel);
    flux_kernel<<<1, 1024, 0, stream>>>(mel, T, env, onsets, peak_idx, n_peaks);
    if (percussive_out)
        MB_CUDA(cudaMemcpyAsync(percussive_out, yperc, sizeof(float) * n, cudaMemcpyDeviceToDevice, stream));
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

/* harmonic(audio, margin) / percussive(audio, margin) of features/audio.py:13-24: STFT -> HPSS soft mask -> iSTFT. */
extern "C" int mb_audio_hpss_component(const float* audio, int64_t n, float margin, int which, float* out, void* workspace,
                                       size_t workspace_bytes, mb_stream stream_) {
    MB_REQUIRE(audio && out && workspace, "mb_audio_hpss_component: null argument");
    MB_REQUIRE(which == 0 || which == 1, "mb_audio_hpss_component: which must be 0 (harmonic) or 1 (percussive)");
    MB_REQUIRE(n >= 16 * kHop && n % kHop == 0, "mb_audio_hpss_component: need a multiple of %d samples (>= %d), got %lld", kHop,
               16 * kHop, static_cast<long long>(n));
    MB_REQUIRE(margin != 1.0f, "mb_audio_hpss_component: margin == 1 (split_zeros branch) is not implemented");
    const AudioWs w = audio_ws(n);
    if (workspace_bytes < w.total) {
        set_error("mb_audio_hpss_component: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
        return MB_ENOMEM;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int T = static_cast<int>(n / kHop);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    float2* tw = reinterpret_cast<float2*>(base + w.tw);
    float* window = reinterpret_cast<float*>(base + w.window);
    float2* spec = reinterpret_cast<float2*>(base + w.spec);
    float* mag = reinterpret_cast<float*>(base + w.mag);
    float2* comp = reinterpret_cast<float2*>(base + w.perc);
    float* frames = reinterpret_cast<float*>(base + w.frames);
    float* rms_tmp = reinterpret_cast<float*>(base + w.env);
    audio_tables_kernel<<<kNfft / 256, 256, 0, stream>>>(tw, window);
    stft_kernel<<<T + 1, 256, 0, stream>>>(audio, n, T, 0, tw, window, spec, mag, rms_tmp, nullptr, nullptr);
    dim3 hg((kBins + 127) / 128, T + 1);
    hpss_kernel<<<hg, 128, 0, stream>>>(spec, mag, T + 1, margin, which, comp);
    istft_frame_kernel<<<T + 1, 256, 0, stream>>>(comp, tw, window, frames);
    istft_ola_kernel<<<148 * 4, 256, 0, stream>>>(frames, window, n, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

/* Magnitude spectrogram and mel power spectrogram of the torch-native feature path (rosa/spectral.py:59-70:
 * spectrogram(y)[:, :T] with power 1, melspectrogram(y, sr) with power 2), frame-major: mag [T][1025], mel [T][128]
 * (either may be NULL).  audio: device float32 [n], n a multiple of 1024. */
extern "C" int mb_audio_spectrogram(const float* audio, int64_t n, const float* mel_filterbank, float* mag_out, float* mel_out,
                                    void* workspace, size_t workspace_bytes, mb_stream stream_) {
    MB_REQUIRE(audio && workspace && (mag_out || mel_out), "mb_audio_spectrogram: null argument");
    MB_REQUIRE(!mel_out || mel_filterbank, "mb_audio_spectrogram: mel_out needs a mel filter bank");
    MB_REQUIRE(n >= 16 * kHop && n % kHop == 0, "mb_audio_spectrogram: need a multiple of %d samples (>= %d), got %lld", kHop, 16 * kHop,
               static_cast<long long>(n));
    const AudioWs w = audio_ws(n);
    if (workspace_bytes < w.total) {
        set_error("mb_audio_spectrogram: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
        return MB_ENOMEM;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int T = static_cast<int>(n / kHop);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    float2* tw = reinterpret_cast<float2*>(base + w.tw);
    float* window = reinterpret_cast<float*>(base + w.window);
    audio_tables_kernel<<<kNfft / 256, 256, 0, stream>>>(tw, window);
    if (mag_out) {
        float2* spec = reinterpret_cast<float2*>(base + w.spec);
        float* mag = reinterpret_cast<float*>(base + w.mag);
        float* rms_tmp = reinterpret_cast<float*>(base + w.env);
        stft_kernel<<<T + 1, 256, 0, stream>>>(audio, n, T, 0, tw, window, spec, mag, rms_tmp, nullptr, nullptr);
        MB_CUDA(cudaMemcpyAsync(mag_out, mag, sizeof(float) * static_cast<size_t>(T) * kBins, cudaMemcpyDeviceToDevice, stream));
    }
    if (mel_out) stft_kernel<<<T, 256, 0, stream>>>(audio, n, T, 1, tw, window, nullptr, nullptr, nullptr, mel_filterbank, mel_out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

/* spectral_flatness (features/audio.py:123-133) from mag [T][1025] -> out [T]. */
extern "C" int mb_spectral_flatness(const float* mag, int T, float amin, float power, float* out, mb_stream stream) {
    MB_REQUIRE(mag && out && T > 0, "mb_spectral_flatness: bad argument");
    flatness_kernel<<<T, 256, 0, static_cast<cudaStream_t>(stream)>>>(mag, T, amin, power, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

/* spectral_contrast (features/audio.py:69-120) from mag [T][1025]: band b covers bins [lo[b], hi[b]) and averages its
 * cnt[b] smallest / largest magnitudes (host-designed from the FFT bin frequencies, as the reference's loop does);
 * out [T][n_bands] = power_to_db(peak) - power_to_db(valley) (linear = 0) or peak - valley (linear = 1).
 * scratch: device float32 [2 * n_bands * T]. */
extern "C" int mb_spectral_contrast(const float* mag, int T, int n_bands, const int32_t* lo, const int32_t* hi, const int32_t* cnt,
                                    int linear, float* scratch, float* out, mb_stream stream_) {
    MB_REQUIRE(mag && lo && hi && cnt && scratch && out && T > 0 && n_bands > 0 && n_bands <= 8, "mb_spectral_contrast: bad argument");
    ContrastBands bands;
    bands.n = n_bands;
    for (int b = 0; b < n_bands; ++b) {
        MB_REQUIRE(lo[b] >= 0 && hi[b] <= kBins && hi[b] > lo[b] && cnt[b] >= 1 && cnt[b] <= hi[b] - lo[b], "mb_spectral_contrast: bad band %d", b);
        bands.lo[b] = lo[b]; bands.hi[b] = hi[b]; bands.cnt[b] = cnt[b];
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    float* peak = scratch;
    float* valley = scratch + static_cast<size_t>(n_bands) * T;
    contrast_kernel<<<T, 256, 0, stream>>>(mag, T, bands, peak, valley);
    const long long nn = static_cast<long long>(n_bands) * T;
    if (!linear) {
        power_to_db_kernel<<<1, 1024, 0, stream>>>(peak, nn, 80.0f);
        power_to_db_kernel<<<1, 1024, 0, stream>>>(valley, nn, 80.0f);
    }
    contrast_diff_kernel<<<(static_cast<int>(nn) + 255) / 256, 256, 0, stream>>>(peak, valley, T, n_bands, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

/* mfcc (features/audio.py:59-64) from the mel POWER spectrogram mel [T][128] (modified in place: converted to dB):
 * power_to_db (top_db 80 below the global maximum) -> orthonormal DCT-II over the mel bands -> out [T][n_mfcc]. */
extern "C" int mb_mfcc(float* mel, int T, int n_mfcc, float* out, mb_stream stream_) {
    MB_REQUIRE(mel && out && T > 0 && n_mfcc > 0 && n_mfcc <= kMels, "mb_mfcc: bad argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    power_to_db_kernel<<<1, 1024, 0, stream>>>(mel, static_cast<long long>(T) * kMels, 80.0f);
    mfcc_dct_kernel<<<T, 128, 0, stream>>>(mel, T, n_mfcc, out);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}
