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
