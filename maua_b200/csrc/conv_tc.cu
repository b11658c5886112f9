// modulated_conv2d hot loop as a tcgen05 implicit GEMM (sm_100a only).
//
// Replaces the grouped cuDNN conv of upstream networks_stylegan3.py::modulated_conv2d
// (in-tree twin: maua/GAN/wrappers/inference/ops.py:146-186).  Formulation
//     y[b,o,p] = d[b,o] * sum_{i,kh,kw} Wn[o,i,kh,kw] * (s[b,i] * x[b,i,p+k])
// so ONE packed fp16 weight matrix serves the whole batch: the style is folded into the
// activations by the producing kernel, the demodulation coefficient is applied in the epilogue.
//
// GEMM view per CTA tile:  D[128 couts, 256 pixels] += A[128, K] * B[K, 256]
//   A = packed weights, K-major, SWIZZLE_128B tiles [128 x 64] (TMA 2D)
//   B = activations, channels-last fp16 [B][H][W][Cp], K-major: one TMA 4D box
//       (64 channels, TW pixels, TH+k-1 rows) per (kw, 64-channel chunk) lands as
//       [(TH+k-1)*TW pixel rows][128 B] in the canonical SWIZZLE_128B layout.  The kh taps reuse
//       the same shared-memory patch through row-offset descriptors (kh*TW rows = whole swizzle
//       atoms), so each activation byte crosses L2->smem k times, not k*k times.  The kw taps need
//       their own loads (measured: the TMA unit rejects box starts that are not 16-byte aligned,
//       so a 1-pixel shift is only expressible on an outer dimension -> channels-last).  Zero
//       padding of the convolution = TMA out-of-bounds fill (negative / past-the-end coordinates).
//   D = fp32 in TMEM (2 x 256 columns, double buffered against the epilogue).
// Warp roles (256 threads): warp0 TMA producer, warp1 MMA issuer, warp2 TMEM alloc,
// warps4-7 epilogue (tcgen05.ld -> *d -> fp16 -> row segments of the planar NCHW output that
// filtered_lrelu consumes).
#include "common.cuh"
#include "kernels.h"

namespace mb {

namespace {

constexpr int kMaxStages = 6;
constexpr int kTileM = 128;         // couts per tile (UMMA M)
constexpr int kTileN = 256;         // pixels per tile (UMMA N)
constexpr int kKC = 64;             // channels per pipeline stage
#ifdef MB_WAIT_PROFILE
constexpr int kSmemMax = 232448 - 1024;   // the wait-profile build keeps 64 bytes of static shared memory
#else
constexpr int kSmemMax = 232448;    // 227 KB dynamic shared memory per CTA on sm_100
#endif
constexpr int kEpiPitch = 20;       // words per cout row of the shift-mode epilogue staging tile (16 data + 4 pad)
constexpr int kEpiStageBytes = 4 * 32 * kEpiPitch * 4;  // one [32 couts][kEpiPitch] tile per epilogue warp

template <int TW>
struct Geo {
    static constexpr int TH = kTileN / TW;
    static constexpr int SLAB = TW * 128;            // one h-row of the patch: TW pixels x 64 channels
};

struct KArgs {
    int B, Cin, Cout, Hout, Wout, Wp_out, ksz, nCC, pad;
    int tiles_m, tiles_w, tiles_h, total_tiles;
    const float* d;  // [B][Cout] or nullptr
    const float* bias;  // [Cout] or nullptr
    __half* y;
    long long plane_out;  // Hout * Wp_out
    // output addressing: element (b, co, h, w) lives at y + b * Cout * plane_out + co * cs + h * rs + w.
    // planar [B][C][H][Wp]: cs = plane_out, rs = Wp_out;  row-interleaved [B][H][C][Wp]: cs = Wp_out, rs = Cout * Wp_out
    // (the couts of one image row share a 2 MB page: the narrow layers' epilogues touch 32+ planes per tile)
    long long cs, rs;
    int* dbg;             // debug words (mapped host memory) or nullptr
    // shared-memory plan (conv_tc_launch): see the layout comment in conv_tc_kernel
    int a_tile_bytes;     // one [a_rows x 64] weight tile (a_rows = 128, or ceil8(Cout) when Cout fits one M tile)
    int resident;         // 1: every weight tile stays in shared memory for the whole kernel
    int nstages, stage_bytes;
    // shift = 1 (needs resident weights, TW = 32): ONE patch load per 64-channel chunk serves all k*k taps; tap (kh, kw)
    // reads the 256-row B window that starts kh image rows (whole swizzle atoms) and kw pixels (kw * 128 bytes, inside an
    // atom) into the patch -- the same descriptor-start trick the pixel-major tile uses for its A operand.  A tile then
    // yields tile_wv = TW - (k - 1) valid output columns per row (the window of the last columns wraps into the next
    // patch row; those accumulator columns are never stored) and the activation bytes cross L2 -> smem once, not k times.
    // Bit-correct (parity test impl 7) but OFF by default: measured on B200 it loses to the per-kw loads because the tile
    // origin 30 * wt is only 4-byte aligned and the planar output then needs 4x the store instructions; store instructions
    // into the [B][C][H][Wp] planes (2 MB apart at 1024^2) cost ~40-60 cycles each regardless of the bytes they carry
    // (L13: 0.51 ms product path; shift mode 0.46 ms without its stores, 0.70 ms with every store aimed at two planes,
    // 1.03-1.08 ms with the real 32-plane scatter, aligned or not; 0.90 ms with direct per-lane 4-byte stores).
    int shift, tile_wv;
    // shift = 2 (EG > 1 kernel): the patch is TW + 2 pixels wide, so every shifted window keeps all TW output columns of a row
    // and the tile origin stays a multiple of 32 pixels (aligned vector stores).  Accumulator column n = r * (TW + 2) + c: a tile
    // is th = 7 image rows inside one N = 240 instruction, the epilogue reads row r from TMEM columns [34 r, 34 r + 32).
    int th, slab, colp, nchunks, mma_n, patch_tx;
    int st256;             // 1: the output row pitch is a multiple of 16 pixels -> 32-byte stores (EG > 1 epilogue)
    long long lo_off;      // > 0: split output -- y holds fp16(v), y + lo_off holds fp16(v - fp16(v)) (fp32-class precision for
                           // consumers that add the two; cout-major tile without shift only)
    unsigned int* absmax;  // device word or nullptr: atomicMax of the bits of max |y| over the stored outputs (read by the
                           // following filtered_lrelu to prove its clamp inactive, flrelu_stream.cuh)
};

// running maximum of |y| in packed halves (one HMNMX2 per stored pair) and its hand-over at the end of an epilogue warp
__device__ __forceinline__ void track_abs(__half2& m, const __half2& v) { m = __hmax2(m, __habs2(v)); }
__device__ __forceinline__ void publish_abs(unsigned int* dst, const __half2& m) {
    if (dst == nullptr) return;
    float f = fmaxf(__low2float(m), __high2float(m));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) f = fmaxf(f, __shfl_xor_sync(0xffffffffu, f, s));
    if ((threadIdx.x & 31) == 0) atomicMax(dst, __float_as_uint(f));   // non-negative floats order like their bit patterns
}

// tcgen05.wait::ld with the destination registers as in/out operands: uses of `r` cannot be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// One 32-column accumulator chunk of a cout-major tile (lane = cout): *d + bias -> fp16 -> the lane's row segments of the
// planar output.  ST256: 32-byte stores (row pitch and tile origin are multiples of 16 pixels), else 16-byte stores.
template <int TW, bool ST256>
__device__ __forceinline__ void epi_chunk_planar(const uint32_t (&v)[32], int n0, float scale, float bias, __half* yplane, int h0,
                                                 int w0, int Hout, int Wp_out, long long rs, __half2& amax) {
    uint32_t pk[16];
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
        const __half2 hv = __floats2half2_rn(fmaf(__uint_as_float(v[2 * k2]), scale, bias), fmaf(__uint_as_float(v[2 * k2 + 1]), scale, bias));
        track_abs(amax, hv);
        pk[k2] = *reinterpret_cast<const uint32_t*>(&hv);
    }
    if constexpr (ST256) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int n = n0 + g * 16;
            const int h = h0 + n / TW;
            const int w = w0 + n % TW;
            if (h < Hout && w < Wp_out)
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yplane + h * rs + w),
                             "r"(pk[8 * g]), "r"(pk[8 * g + 1]), "r"(pk[8 * g + 2]), "r"(pk[8 * g + 3]), "r"(pk[8 * g + 4]), "r"(pk[8 * g + 5]),
                             "r"(pk[8 * g + 6]), "r"(pk[8 * g + 7])
                             : "memory");
        }
    } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int n = n0 + g * 8;
            const int h = h0 + n / TW;
            const int w = w0 + n % TW;
            if (h < Hout && w < Wp_out)
                *reinterpret_cast<uint4*>(yplane + h * rs + w) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
    }
}

// EG = number of epilogue warp groups (4 warps each: one per TMEM lane quadrant).  EG = 1 is the general kernel.  With few
// couts only the first quadrant(s) hold real rows and ONE warp per quadrant cannot drain a 256-column accumulator in the time
// the tensor core needs to fill the next (L13 of StyleGAN3-T: 2 304 MMA cycles per tile against ~5 400 epilogue cycles on the
// single warp that owns couts 0..31).  EG > 1 adds warps 8.. whose quadrant is again warp % 4: the groups split the
// tile's eight 32-column chunks between them, each warp double-buffers its tcgen05.ld against the convert / store of the
// previous chunk, and the accumulator is handed back as soon as the last load has landed.
template <int TW, int EG>
__global__ void __launch_bounds__(128 + 128 * EG, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, KArgs a) {
    using G = Geo<TW>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // Layout: [resident weight tiles (ksz*ksz*nCC x a_tile_bytes) | nstages x stage | barriers], a stage being
    // [ksz weight tiles (streaming mode only) | activation patch].  The narrow top layers (Cout <= 64: L11..L13 of
    // StyleGAN3-T) were L2->SM bandwidth bound at ~12 TB/s re-fetching 48 KB of zero-padded weight rows per stage
    // (r1 layer timings); their whole weight matrix now stays in shared memory and the stages only carry patches.
    // UMMA M stays 128: the rows past a_rows read whatever follows in shared memory and land in accumulator lanes
    // the epilogue never stores.
    const int nst = a.nstages;
    const int n_wtiles = a.ksz * a.ksz * a.nCC;
    uint8_t* stages = smem + (a.resident ? n_wtiles * a.a_tile_bytes : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stages + nst * a.stage_bytes);
    uint64_t* full = bars;                       // [kMaxStages]
    uint64_t* empty = bars + kMaxStages;         // [kMaxStages]
    uint64_t* tfull = bars + 2 * kMaxStages;     // [2]
    uint64_t* tempty = bars + 2 * kMaxStages + 2;  // [2]
    uint64_t* wfull = bars + 2 * kMaxStages + 4;   // [1] resident weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);
    uint32_t* epi_stage = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 256);  // shift mode only

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int pad = a.pad;
    const int a_bytes_stage = a.resident ? 0 : a.ksz * a.a_tile_bytes;
    const uint32_t stage_tx = a_bytes_stage + a.patch_tx;
    const int iters = a.ksz * a.nCC;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < nst; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(wfull, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 4 * EG);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp runs the loop and one ELECTED lane issues: a TMA / tcgen05 instruction inside a divergent
        // `if (lane == 0)` region is wrapped by the compiler in a per-thread vote loop (BRA.U.ANY) that costs ~90
        // cycles per instruction (measured, scratch/umma_bench.cu); under an elect.sync predicate it is issued directly.
        {
            const bool leader = elect_one();
            int s = 0;
            uint32_t ph = 0;
            if (a.resident && leader && blockIdx.x < a.total_tiles) {
                mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(n_wtiles * a.a_tile_bytes));
                for (int i = 0; i < n_wtiles; ++i) tma_load_2d(smem + i * a.a_tile_bytes, &tmap_w, wfull, i * kKC, 0);
            }
            __syncwarp();
            for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
                int r = t;
                const int mt = r % a.tiles_m; r /= a.tiles_m;
                const int wt = r % a.tiles_w; r /= a.tiles_w;
                const int ht = r % a.tiles_h; r /= a.tiles_h;
                const int b = r;
                const int h0 = ht * a.th, w0 = wt * a.tile_wv, m0 = mt * kTileM;
                const int nkw = a.shift ? 1 : a.ksz;
                for (int kw = 0; kw < nkw; ++kw) {
                    for (int cc = 0; cc < a.nCC; ++cc) {
                        mbar_wait(&empty[s], ph ^ 1, a.dbg, 1);
                        if (leader) {
                            uint8_t* st = stages + s * a.stage_bytes;
                            mbar_arrive_expect_tx(&full[s], stage_tx);
                            if (!a.resident) {
                                const int kblk = (kw * a.nCC + cc) * a.ksz;
                                for (int kh = 0; kh < a.ksz; ++kh)
                                    tma_load_2d(st + kh * a.a_tile_bytes, &tmap_w, &full[s], (kblk + kh) * kKC, m0);
                            }
                            tma_load_4d(st + a_bytes_stage, &tmap_x, &full[s], cc * kKC, w0 + kw - pad, h0 - pad, b);
                        }
                        __syncwarp();
                        if (++s == nst) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc_f16(kTileM, a.mma_n, /*A K-major*/ 0, /*B K-major*/ 0);
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph = 0;
            if (a.resident && blockIdx.x < a.total_tiles) mbar_wait(wfull, 0, a.dbg, 5);
            for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
                mbar_wait(&tempty[acc], acc_ph ^ 1, a.dbg, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kTileN;
                uint32_t accumulate = 0;
                const int nit = a.shift ? a.nCC : iters;
                for (int it = 0; it < nit; ++it) {
                    const int cc = it % a.nCC;
                    int nk16 = (a.Cin - cc * kKC + 15) / 16;
                    if (nk16 > 4) nk16 = 4;
                    mbar_wait(&full[s], ph, a.dbg, 3);
                    tc_fence_after();
                    const uint32_t st = smem_u32(stages + s * a.stage_bytes);
                    const uint32_t sa = a.resident ? smem_u32(smem) + it * a.ksz * a.a_tile_bytes : st;
                    const uint32_t sb = st + a_bytes_stage;
                    const uint64_t da0 = make_smem_desc(sa, 16, 1024, 2);
                    const uint64_t db0 = make_smem_desc(sb, 16, 1024, 2);
                    if (a.shift) {
                        if (leader) {
                            for (int kw = 0; kw < a.ksz; ++kw) {
                                const uint64_t da = make_smem_desc(smem_u32(smem) + ((kw * a.nCC + cc) * a.ksz) * a.a_tile_bytes, 16, 1024, 2);
                                const uint64_t db = db0 + static_cast<uint64_t>((kw * 128) >> 4);
                                for (int kh = 0; kh < a.ksz; ++kh) {
#pragma unroll 4
                                    for (int j = 0; j < nk16; ++j) {
                                        umma_f16(d_tmem, da + static_cast<uint64_t>((kh * a.a_tile_bytes + j * 32) >> 4),
                                                 db + static_cast<uint64_t>((kh * a.slab + j * 32) >> 4), idesc, accumulate);
                                        accumulate = 1;
                                    }
                                }
                            }
                            umma_commit(&empty[s]);
                        }
                        __syncwarp();
                        if (++s == nst) { s = 0; ph ^= 1; }
                        continue;
                    }
                    if (leader) {
                        for (int kh = 0; kh < a.ksz; ++kh) {
#pragma unroll 4
                            for (int j = 0; j < nk16; ++j) {
                                umma_f16(d_tmem, da0 + static_cast<uint64_t>((kh * a.a_tile_bytes + j * 32) >> 4),
                                         db0 + static_cast<uint64_t>((kh * G::SLAB + j * 32) >> 4), idesc, accumulate);
                                accumulate = 1;
                            }
                        }
                        umma_commit(&empty[s]);
                    }
                    __syncwarp();
                    if (++s == nst) { s = 0; ph ^= 1; }
                }
                if (leader) umma_commit(&tfull[acc]);
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else if (warp >= 4 && EG > 1) {
        // ===================== epilogue, EG groups (no shift mode, no split output) =====================
        const int q = warp & 3;           // TMEM lane quadrant == warp % 4
        const int grp = (warp - 4) >> 2;  // this warp handles chunks grp, grp + EG, ...
        int acc = 0;
        uint32_t acc_ph = 0;
        __half2 amax = __floats2half2_rn(0.0f, 0.0f);
        const int kChunks = a.nchunks;   // one chunk = one image row of the tile (TW = 32)
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            int r = t;
            const int mt = r % a.tiles_m; r /= a.tiles_m;
            const int wt = r % a.tiles_w; r /= a.tiles_w;
            const int ht = r % a.tiles_h; r /= a.tiles_h;
            const int b = r;
            const int h0 = ht * a.th, w0 = wt * TW;
            const int co = mt * kTileM + q * 32 + lane;
            const bool q_ok = mt * kTileM + q * 32 < a.Cout && grp < kChunks;   // warp-uniform
            const bool co_ok = co < a.Cout;
            const float scale = (co_ok && a.d) ? a.d[b * a.Cout + co] : (co_ok ? 1.0f : 0.0f);
            const float bias = (co_ok && a.bias) ? a.bias[co] : 0.0f;
            // lanes past the last cout compute zeros and aim at the last real plane with their stores predicated off
            __half* yplane = a.y + static_cast<long long>(b) * a.Cout * a.plane_out + (co_ok ? co : 0) * a.cs;
            const int hlim = co_ok ? a.Hout : 0;   // 0 switches a lane's stores off
            mbar_wait(&tfull[acc], acc_ph, a.dbg, 4);
            tc_fence_after();
            auto release = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
            };
            if (q_ok) {
                const uint32_t tb = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kTileN;
                uint32_t va[32], vb[32];
                int ch = grp;
                tmem_ld_32x32b_x32(tb + ch * a.colp, va);
                for (;;) {
                    int nx = ch + EG;
                    tmem_ld_wait_dep(va);
                    if (nx < kChunks) tmem_ld_32x32b_x32(tb + nx * a.colp, vb); else release();
                    if (a.st256) epi_chunk_planar<TW, true>(va, ch * 32, scale, bias, yplane, h0, w0, hlim, a.Wp_out, a.rs, amax);
                    else epi_chunk_planar<TW, false>(va, ch * 32, scale, bias, yplane, h0, w0, hlim, a.Wp_out, a.rs, amax);
                    if (nx >= kChunks) break;
                    ch = nx; nx = ch + EG;
                    tmem_ld_wait_dep(vb);
                    if (nx < kChunks) tmem_ld_32x32b_x32(tb + nx * a.colp, va); else release();
                    if (a.st256) epi_chunk_planar<TW, true>(vb, ch * 32, scale, bias, yplane, h0, w0, hlim, a.Wp_out, a.rs, amax);
                    else epi_chunk_planar<TW, false>(vb, ch * 32, scale, bias, yplane, h0, w0, hlim, a.Wp_out, a.rs, amax);
                    if (nx >= kChunks) break;
                    ch = nx;
                }
            } else {
                release();
            }
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
        publish_abs(a.absmax, amax);
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp - 4;  // TMEM lane quadrant == warp % 4
        int acc = 0;
        uint32_t acc_ph = 0;
        __half2 amax = __floats2half2_rn(0.0f, 0.0f);
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            int r = t;
            const int mt = r % a.tiles_m; r /= a.tiles_m;
            const int wt = r % a.tiles_w; r /= a.tiles_w;
            const int ht = r % a.tiles_h; r /= a.tiles_h;
            const int b = r;
            const int h0 = ht * G::TH, w0 = wt * a.tile_wv;
            const int co = mt * kTileM + q * 32 + lane;
            const bool co_ok = co < a.Cout;
            const float scale = (co_ok && a.d) ? a.d[b * a.Cout + co] : 1.0f;
            const float bias = (co_ok && a.bias) ? a.bias[co] : 0.0f;
            __half* yplane = a.y + static_cast<long long>(b) * a.Cout * a.plane_out + (co_ok ? co : 0) * a.cs;

            mbar_wait(&tfull[acc], acc_ph, a.dbg, 4);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < kTileN / 32; ++ch) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kTileN + ch * 32, v);
                tmem_ld_wait();
                const int n0 = ch * 32;
                if (a.shift) {
                    // Tile origin w0 = tile_wv * wt is only 4-byte aligned and a lane (= cout) would scatter 4-byte stores
                    // over 32 planes (measured: L13 0.51 -> 0.90 ms, store-bound).  Transpose through a per-warp staging
                    // tile instead: lane = cout writes its 32 pixels as 4 x STS.128 (row pitch 80 B: conflict-free per
                    // quarter warp), then half a warp reads one cout row back, so one STG.32 covers 2 x 60 contiguous bytes.
                    // chunk ch = accumulator columns [32 ch, 32 ch + 32) = image row h0 + ch (TW = 32).
                    const uint32_t stg = smem_u32(epi_stage) + q * (32 * kEpiPitch * 4);  // shared-space address: STS / LDS
                    uint32_t pk[16];
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        const __half2 hv = __floats2half2_rn(fmaf(__uint_as_float(v[2 * k2]), scale, bias),
                                                             fmaf(__uint_as_float(v[2 * k2 + 1]), scale, bias));
                        if (co_ok) track_abs(amax, hv);
                        pk[k2] = *reinterpret_cast<const uint32_t*>(&hv);
                    }
                    __syncwarp();  // the previous chunk's read-out is done
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (lane * kEpiPitch + g * 4) * 4), "r"(pk[4 * g]),
                                     "r"(pk[4 * g + 1]), "r"(pk[4 * g + 2]), "r"(pk[4 * g + 3])
                                     : "memory");
                    __syncwarp();
                    const int h = h0 + ch;
                    const int kk = lane & 15;
                    const int w = w0 + 2 * kk;
                    const bool px_ok = h < a.Hout && 2 * kk < a.tile_wv && w < a.Wp_out;
                    const int co0 = mt * kTileM + q * 32 + (lane >> 4);
                    __half* yrow = a.y + (static_cast<long long>(b) * a.Cout + co0) * a.plane_out + static_cast<long long>(h) * a.Wp_out + w;
                    // all 16 loads first (independent, pipelined), then the stores: alternating them serialises on the
                    // shared-memory latency, which is long while the tensor core streams its operands (measured 1300 cycles / chunk)
                    uint32_t val[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(val[i]) : "r"(stg + ((2 * i + (lane >> 4)) * kEpiPitch + kk) * 4));
                    if (px_ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (co0 + 2 * i < a.Cout)
                                *reinterpret_cast<uint32_t*>(yrow + static_cast<long long>(2 * i) * a.plane_out) = val[i];
                    }
                } else if (co_ok) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int n = n0 + g * 8;
                        const int h = h0 + n / TW;
                        const int w = w0 + n % TW;
                        if (h < a.Hout && w < a.Wp_out) {
                            uint4 pk;
                            __half2 h0v = __floats2half2_rn(fmaf(__uint_as_float(v[g * 8 + 0]), scale, bias),
                                                            fmaf(__uint_as_float(v[g * 8 + 1]), scale, bias));
                            __half2 h1v = __floats2half2_rn(fmaf(__uint_as_float(v[g * 8 + 2]), scale, bias),
                                                            fmaf(__uint_as_float(v[g * 8 + 3]), scale, bias));
                            __half2 h2v = __floats2half2_rn(fmaf(__uint_as_float(v[g * 8 + 4]), scale, bias),
                                                            fmaf(__uint_as_float(v[g * 8 + 5]), scale, bias));
                            __half2 h3v = __floats2half2_rn(fmaf(__uint_as_float(v[g * 8 + 6]), scale, bias),
                                                            fmaf(__uint_as_float(v[g * 8 + 7]), scale, bias));
                            track_abs(amax, h0v); track_abs(amax, h1v); track_abs(amax, h2v); track_abs(amax, h3v);
                            pk.x = *reinterpret_cast<uint32_t*>(&h0v);
                            pk.y = *reinterpret_cast<uint32_t*>(&h1v);
                            pk.z = *reinterpret_cast<uint32_t*>(&h2v);
                            pk.w = *reinterpret_cast<uint32_t*>(&h3v);
                            *reinterpret_cast<uint4*>(yplane + h * a.rs + w) = pk;
                            if (a.lo_off > 0) {
                                // the part of each value its fp16 rounding dropped, as a second fp16 plane
                                const __half2 hs[4] = {h0v, h1v, h2v, h3v};
                                uint32_t lo[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float2 hf = __half22float2(hs[k]);
                                    const __half2 l2 = __floats2half2_rn(fmaf(__uint_as_float(v[g * 8 + 2 * k]), scale, bias) - hf.x,
                                                                         fmaf(__uint_as_float(v[g * 8 + 2 * k + 1]), scale, bias) - hf.y);
                                    lo[k] = *reinterpret_cast<const uint32_t*>(&l2);
                                }
                                *reinterpret_cast<uint4*>(yplane + a.lo_off + static_cast<long long>(h) * a.Wp_out + w) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&tempty[acc]); if (q == 0) dbg_inc(a.dbg, 2); }
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
        publish_abs(a.absmax, amax);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}


// ---------------------------------------------------------------------------------------------------------
// Pixel-major variant for the narrow top layers (Cout <= 64: L11..L13 of StyleGAN3-T hold 51/32/32 couts).
// The tile above pads couts to UMMA M = 128, so those layers spend 2.5x..4x of their tensor time on zero rows
// (ncu r1: tensor pipe 64..82 % busy at 250..470 useful TFLOP/s).  Here the roles are swapped:
//     D[128 pixels, Np couts] += A[128 pixels, K] * B[K, Np couts],  Np = ceil16(Cout),
// the SAME shared-memory patch serves as the A operand (4 image rows x 32 pixels = 128 K-major rows; the 8-row
// tile is two such halves with their own accumulators) and the SAME packed weight tiles as B (box of Np rows).
// MMA time per 256 pixels and K=16 drops from 128 cycles to max(Np, shared-memory read of A) cycles.  The weight
// tiles stay resident in shared memory whenever they fit (they do for every layer this variant is used for).
// Epilogue: one TMEM lane = one pixel.  Neighbouring lanes swap one cout of every cout pair (one SHFL + PRMT), so a
// lane stores two adjacent pixels of one cout (4 bytes) and a warp store covers 2 x 64-byte row segments; the
// per-cout (demodulation, bias) pairs come from a per-warp shared-memory table refreshed when the frame changes.
// With resident weights ONE patch load per 64-channel chunk serves all nine taps (`shift`): tap (kh, kw) reads the
// 128-row A window that starts kh image rows (whole swizzle atoms) and kw pixels (kw * 128 bytes, INSIDE an atom)
// further on -- the tensor core applies the 128-byte swizzle to absolute shared-memory address bits, so a descriptor
// start that is not 1024-byte aligned needs no base offset (probed on B200, scripts/conv_shift_probe.py).  The tile
// is then 30 pixels wide so every shifted window stays inside the 32-pixel patch; TMA traffic drops 2.8x.
constexpr int kPmTW = 32, kPmTH = 8;
constexpr int kPmSlab = kPmTW * 128;
constexpr int kPmMaxStages = 4;

struct PmArgs {
    KArgs k;
    int Np, stages, stage_bytes_alloc, resident;
    int shift;   // 0: one patch load per (kw, chunk); 1 / 2: ONE patch load per chunk, the kw shift is a +128-byte start
                 // offset of the A descriptor (2: with the descriptor's base-offset field set to the row phase)
    int tile_w;  // valid output pixels per tile row: 32, or 30 when the shifted window must stay inside the 32-px patch
};

template <bool SHIFT>
__global__ void __launch_bounds__(256, 1)
conv_pm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, PmArgs pa) {
    const KArgs& a = pa.k;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nst = pa.stages;
    const int STAGE = pa.stage_bytes_alloc;
    const int Np = pa.Np;
    const int w_tile = Np * 128;                   // bytes of one [Np x 64] weight tile
    const int n_wtiles = a.ksz * a.ksz * a.nCC;
    const int patch_rows = kPmTH + a.ksz - 1;
    const int patch_bytes = patch_rows * kPmSlab;
    // layout: [resident weight tiles | stages: patch (+ ksz weight tiles when streaming) | (scale,bias) tables | barriers]
    uint8_t* stages = smem + (pa.resident ? n_wtiles * w_tile : 0);
    float2* sc_tab = reinterpret_cast<float2*>(stages + nst * STAGE);   // [4 warps][128] (demodulation, bias) of the current frame
    uint64_t* bars = reinterpret_cast<uint64_t*>(sc_tab + 4 * 128);
    uint64_t* full = bars;                         // [kPmMaxStages]
    uint64_t* empty = bars + kPmMaxStages;         // [kPmMaxStages]
    uint64_t* tfull = bars + 2 * kPmMaxStages;     // [2]
    uint64_t* tempty = bars + 2 * kPmMaxStages + 2;  // [2]
    uint64_t* wfull = bars + 2 * kPmMaxStages + 4;   // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPmMaxStages + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int pad = a.pad;
    const uint32_t stage_tx = patch_bytes + (pa.resident ? 0 : a.ksz * w_tile);
    const int iters = a.ksz * a.nCC;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < nst; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 4);
        }
        mbar_init(wfull, 1);
        MB_WAIT_PROFILE_INIT();
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        const bool leader = elect_one();
        int s = 0;
        uint32_t ph = 0;
        if (pa.resident && leader && blockIdx.x < a.total_tiles) {
            mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(n_wtiles * w_tile));
            for (int i = 0; i < n_wtiles; ++i) tma_load_2d(smem + i * w_tile, &tmap_w, wfull, i * kKC, 0);
        }
        __syncwarp();
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            int r = t;
            const int wt = r % a.tiles_w; r /= a.tiles_w;
            const int ht = r % a.tiles_h; r /= a.tiles_h;
            const int b = r;
            const int h0 = ht * kPmTH, w0 = wt * pa.tile_w;
            for (int kw = 0; kw < (SHIFT ? 1 : a.ksz); ++kw) {
                for (int cc = 0; cc < a.nCC; ++cc) {
                    mbar_wait(&empty[s], ph ^ 1, a.dbg, 1);
                    if (leader) {
                        uint8_t* st = stages + s * STAGE;
                        mbar_arrive_expect_tx(&full[s], stage_tx);
                        if (!pa.resident) {
                            const int kblk = (kw * a.nCC + cc) * a.ksz;
                            for (int kh = 0; kh < a.ksz; ++kh)
                                tma_load_2d(st + patch_bytes + kh * w_tile, &tmap_w, &full[s], (kblk + kh) * kKC, 0);
                        }
                        tma_load_4d(st, &tmap_x, &full[s], cc * kKC, w0 + kw - pad, h0 - pad, b);
                    }
                    __syncwarp();
                    if (++s == nst) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_f16(128, Np, 0, 0);
        int s = 0;
        uint32_t ph = 0;
        int acc = 0;
        uint32_t acc_ph = 0;
        if (pa.resident && blockIdx.x < a.total_tiles) mbar_wait(wfull, 0, a.dbg, 5);
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            mbar_wait(&tempty[acc], acc_ph ^ 1, a.dbg, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 2 * Np;
            uint32_t accumulate = 0;
            for (int it = 0; it < (SHIFT ? a.nCC : iters); ++it) {
                const int cc = it % a.nCC;
                int nk16 = (a.Cin - cc * kKC + 15) / 16;
                if (nk16 > 4) nk16 = 4;
                mbar_wait(&full[s], ph, a.dbg, 3);
                tc_fence_after();
                const uint32_t sx = smem_u32(stages + s * STAGE);
                if constexpr (SHIFT) {
                    // one resident patch serves all nine taps: tap (kh, kw) reads the 128-row window that starts
                    // kh image rows (whole swizzle atoms) and kw pixels (kw * 128 bytes, inside an atom) further on
                    if (leader) {
                        for (int kw = 0; kw < a.ksz; ++kw) {
                            const uint32_t sw = smem_u32(smem) + ((kw * a.nCC + cc) * a.ksz) * w_tile;
                            const uint64_t dx = make_smem_desc(sx + kw * 128, 16, 1024, 2, pa.shift == 2 ? kw : 0);
                            const uint64_t dw = make_smem_desc(sw, 16, 1024, 2);
                            for (int kh = 0; kh < a.ksz; ++kh) {
#pragma unroll 4
                                for (int j = 0; j < nk16; ++j) {
                                    const uint64_t dwk = dw + static_cast<uint64_t>((kh * w_tile + j * 32) >> 4);
                                    const uint64_t dxk = dx + static_cast<uint64_t>((kh * kPmSlab + j * 32) >> 4);
                                    umma_f16(d_tmem, dxk, dwk, idesc, accumulate);
                                    umma_f16(d_tmem + Np, dxk + static_cast<uint64_t>((4 * kPmSlab) >> 4), dwk, idesc, accumulate);
                                    accumulate = 1;
                                }
                            }
                        }
                        umma_commit(&empty[s]);
                    }
                    __syncwarp();
                    if (++s == nst) { s = 0; ph ^= 1; }
                    continue;
                }
                const uint32_t sw = pa.resident ? smem_u32(smem) + it * a.ksz * w_tile : sx + patch_bytes;
                const uint64_t dx = make_smem_desc(sx, 16, 1024, 2);
                const uint64_t dw = make_smem_desc(sw, 16, 1024, 2);
                if (leader) {
                    for (int kh = 0; kh < a.ksz; ++kh) {
#pragma unroll 4
                        for (int j = 0; j < nk16; ++j) {
                            const uint64_t dwk = dw + static_cast<uint64_t>((kh * w_tile + j * 32) >> 4);
                            const uint64_t dxk = dx + static_cast<uint64_t>((kh * kPmSlab + j * 32) >> 4);
                            umma_f16(d_tmem, dxk, dwk, idesc, accumulate);
                            umma_f16(d_tmem + Np, dxk + static_cast<uint64_t>((4 * kPmSlab) >> 4), dwk, idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == nst) { s = 0; ph ^= 1; }
            }
            if (leader) umma_commit(&tfull[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;  // TMEM lane quadrant == image row of the half tile
        float2* tab = sc_tab + q * 128;   // per-warp copy (one table shared under a named barrier measured 20 % slower)
        const int par = lane & 1;
        const uint32_t sel = par ? 0x3276u : 0x5410u;   // even lane keeps cout c of a pair, odd lane cout c+1
        __half2 amax = __floats2half2_rn(0.0f, 0.0f);
        int acc = 0;
        uint32_t acc_ph = 0;
        int tab_b = -1;
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            int r = t;
            const int wt = r % a.tiles_w; r /= a.tiles_w;
            const int ht = r % a.tiles_h; r /= a.tiles_h;
            const int b = r;
            if (b != tab_b) {  // warp-uniform
                __syncwarp();
                for (int i = lane; i < Np; i += 32) {
                    const bool ok = i < a.Cout;
                    tab[i] = make_float2((ok && a.d) ? a.d[b * a.Cout + i] : (ok ? 1.0f : 0.0f), (ok && a.bias) ? a.bias[i] : 0.0f);
                }
                __syncwarp();
                tab_b = b;
            }
            const int w = wt * pa.tile_w + lane;
            const bool w_ok = w < a.Wp_out && lane < pa.tile_w;   // Wp_out and tile_w are even: lane pairs are in or out together
            mbar_wait(&tfull[acc], acc_ph, a.dbg, 4);
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int h = ht * kPmTH + half * 4 + q;
                const bool ok = w_ok && h < a.Hout;
                // this lane's cout plane of pair k is (2k + par); it stores pixels (w & ~1, w | 1)
                __half* yp = a.y + static_cast<long long>(b) * a.Cout * a.plane_out + par * a.cs + h * a.rs + (w & ~1);
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 2 * Np + half * Np;
#pragma unroll 1
                for (int c0 = 0; c0 < Np; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(taddr + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 sb = *reinterpret_cast<const float4*>(tab + c0 + 2 * k);  // (d, bias) of couts c, c+1
                        const __half2 mine = __floats2half2_rn(fmaf(__uint_as_float(v[2 * k]), sb.x, sb.y),
                                                               fmaf(__uint_as_float(v[2 * k + 1]), sb.z, sb.w));
                        track_abs(amax, mine);   // padded couts carry (d, bias) = (0, 0): they contribute |0|
                        const uint32_t m = *reinterpret_cast<const uint32_t*>(&mine);
                        const uint32_t o = __shfl_xor_sync(0xffffffffu, m, 1);
                        const uint32_t pr = __byte_perm(m, o, sel);
                        const int co = c0 + 2 * k + par;
                        if (ok && co < a.Cout)
                            *reinterpret_cast<uint32_t*>(yp + (c0 + 2 * k) * a.cs) = pr;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
        publish_abs(a.absmax, amax);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------
// Pixel-major tile with the three kw taps STACKED ALONG N ("pms"), for 3x3 layers with few couts.
// Both orientations above waste the tensor pipe on such layers: the cout-major tile executes M = 128 rows for 32..51
// real couts (L12 / L13 of StyleGAN3-T run at 1.35..1.6 PFLOP/s EXECUTED, the power-capped ceiling, for 310..340 useful),
// the pixel-major tile pays a fixed ~64 cycles of A-operand read per instruction whatever N is (t = 64 + N/2).  Here
//     D[128 patch pixels, (kw, cout)] += A[128 patch pixels, (kh, cin)] * B[(kh, cin), (kw, cout)],   N = 3 * Np,
// so one instruction does the work of three (the A read is amortised over three taps and K shrinks from 9 Cin to 3 Cin),
// every executed row is a real pixel, and the kw shift never touches a descriptor: output pixel p of a row needs
//     y[p, c] = D[p, (0, c)] + D[p + 1, (1, c)] + D[p + 2, (2, c)],
// TMEM lane = pixel, so the epilogue adds its own column of the kw = 0 block to the lane + 1 / lane + 2 values of the
// kw = 1 / 2 blocks (two SHFL.DOWN per value).  Lanes 30 and 31 of a row have no right-hand neighbours: a tile is 30
// output pixels wide on a 32-pixel patch, as in the single-load pixel-major tile.  One patch load per 64-channel chunk,
// weights resident: the stacked B tile of (chunk, kh) is three TMA boxes of the ordinary packed matrix (one per kw)
// landing back to back.  NH = 4-row halves per tile sharing one patch (2 where the accumulators 2 x NH x 3 Np fit the 512
// TMEM columns, i.e. Np <= 32; 1 otherwise).
struct PmsArgs {
    KArgs k;
    int Np, Nst, stages;
    ConvTcArgs::NhwcOut o;   // o.y != nullptr: channels-last epilogue (bias, lrelu, residuals), see kernels.h
};

// kPmsGroups = epilogue warps per TMEM lane quadrant: 3 for the planar epilogue (StyleGAN3's narrow layers: L13 0.77 -> 0.69 ms),
// 2 for the channels-last one (RRDBNet measured 54.8 ms per 4 frames with two, 56.8 with three)
template <int NH, int kPmsGroups>
__global__ void __launch_bounds__(128 + 128 * kPmsGroups, 1)
conv_pms_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, PmsArgs pa) {
    const KArgs& a = pa.k;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int TH = 4 * NH;
    constexpr int kPatch = (TH + 2) * kPmSlab;
    const int nst = pa.stages;
    const int Np = pa.Np, Nst = pa.Nst;
    const int w_tile = Np * 128;                   // bytes of one [Np x 64] weight tile
    const int n_wtiles = 9 * a.nCC;
    // layout: [resident weight tiles, (chunk, kh, kw) order | patch stages | (scale, bias) tables | barriers]
    uint8_t* stages = smem + n_wtiles * w_tile;
    float2* sc_tab = reinterpret_cast<float2*>(stages + nst * kPatch);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sc_tab + 4 * 128);
    uint64_t* full = bars;                         // [kPmMaxStages]
    uint64_t* empty = bars + kPmMaxStages;         // [kPmMaxStages]
    uint64_t* tfull = bars + 2 * kPmMaxStages;     // [2]
    uint64_t* tempty = bars + 2 * kPmMaxStages + 2;  // [2]
    uint64_t* wfull = bars + 2 * kPmMaxStages + 4;   // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPmMaxStages + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int pad = a.pad;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < nst; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 4 * kPmsGroups);
        }
        mbar_init(wfull, 1);
        MB_WAIT_PROFILE_INIT();
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        const bool leader = elect_one();
        int s = 0;
        uint32_t ph = 0;
        if (leader && blockIdx.x < a.total_tiles) {
            mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(n_wtiles * w_tile));
            for (int cc = 0; cc < a.nCC; ++cc)
                for (int kh = 0; kh < 3; ++kh)
                    for (int kw = 0; kw < 3; ++kw)
                        tma_load_2d(smem + ((cc * 3 + kh) * 3 + kw) * w_tile, &tmap_w, wfull, ((kw * a.nCC + cc) * 3 + kh) * kKC, 0);
        }
        __syncwarp();
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            int r = t;
            const int wt = r % a.tiles_w; r /= a.tiles_w;
            const int ht = r % a.tiles_h; r /= a.tiles_h;
            const int b = r;
            for (int cc = 0; cc < a.nCC; ++cc) {
                mbar_wait(&empty[s], ph ^ 1, a.dbg, 1);
                if (leader) {
                    mbar_arrive_expect_tx(&full[s], kPatch);
                    tma_load_4d(stages + s * kPatch, &tmap_x, &full[s], cc * kKC, wt * 30 - pad, ht * TH - pad, b);
                }
                __syncwarp();
                if (++s == nst) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_f16(128, Nst, 0, 0);
        int s = 0;
        uint32_t ph = 0;
        int acc = 0;
        uint32_t acc_ph = 0;
        if (blockIdx.x < a.total_tiles) mbar_wait(wfull, 0, a.dbg, 5);
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            mbar_wait(&tempty[acc], acc_ph ^ 1, a.dbg, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * NH * Nst;
            uint32_t accumulate = 0;
            for (int cc = 0; cc < a.nCC; ++cc) {
                int nk16 = (a.Cin - cc * kKC + 15) / 16;
                if (nk16 > 4) nk16 = 4;
                mbar_wait(&full[s], ph, a.dbg, 3);
                tc_fence_after();
                if (leader) {
                    const uint64_t dx = make_smem_desc(smem_u32(stages + s * kPatch), 16, 1024, 2);
                    const uint64_t dw = make_smem_desc(smem_u32(smem) + cc * 9 * w_tile, 16, 1024, 2);
                    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll 4
                        for (int j = 0; j < nk16; ++j) {
                            const uint64_t dwk = dw + static_cast<uint64_t>((kh * 3 * w_tile + j * 32) >> 4);
                            const uint64_t dxk = dx + static_cast<uint64_t>((kh * kPmSlab + j * 32) >> 4);
#pragma unroll
                            for (int hf = 0; hf < NH; ++hf)
                                umma_f16(d_tmem + hf * Nst, dxk + static_cast<uint64_t>((hf * 4 * kPmSlab) >> 4), dwk, idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == nst) { s = 0; ph ^= 1; }
            }
            if (leader) umma_commit(&tfull[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
    } else if (warp >= 4) {
        // Three warps per TMEM lane quadrant (warps 4..7, 8..11, 12..15; quadrant = warp % 4 = image row of the half tile) share a
        // tile's (half, 16-cout block) units: one warp per quadrant is issue-latency bound at ~1 200 instructions per tile (ncu r2:
        // tensor pipe 23 % busy, the epilogue warps stalled on the SHFL -> FADD / PRMT scoreboard); with two the epilogue warps are
        // still 94 % busy and the MMA warp waits 630..1 400 cycles per tile for an accumulator (wait profile).  The assignment
        // rotates from tile to tile (unit u of the CTA's i-th tile goes to warp (u + i) mod 3), so the warps with two units of one
        // tile are not the ones with two units of the next.
        const int q = warp & 3;
        const int eg = (warp - 4) >> 2;
        float2* tab = sc_tab + q * 128;   // shared by the warps of a quadrant; refreshed by group 0 under a named barrier
        int ti = 0;                        // tiles this CTA has handled
        __half2 amax = __floats2half2_rn(0.0f, 0.0f);
        int acc = 0;
        uint32_t acc_ph = 0;
        int tab_b = -1;
        const int nblk = Np >> 4;
        // tile coordinates advance by the grid stride without divisions (three integer divisions per tile were a fifth of
        // this warp's instructions)
        int wt, ht, b;
        {
            int r = blockIdx.x;
            wt = r % a.tiles_w; r /= a.tiles_w;
            ht = r % a.tiles_h; b = r / a.tiles_h;
        }
        int dwt, dht, db;
        {
            int r = gridDim.x;
            dwt = r % a.tiles_w; r /= a.tiles_w;
            dht = r % a.tiles_h; db = r / a.tiles_h;
        }
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, wt += dwt, ht += dht, b += db) {
            if (wt >= a.tiles_w) { wt -= a.tiles_w; ++ht; }
            if (ht >= a.tiles_h) { ht -= a.tiles_h; ++b; }
            if (b != tab_b) {  // uniform over both warps of the quadrant
                asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * kPmsGroups) : "memory");   // the other warps are done with the old table
                if (eg == 0)
                    for (int i = lane; i < Np; i += 32) {
                        const bool ok = i < a.Cout;
                        tab[i] = make_float2((ok && a.d) ? a.d[b * a.Cout + i] : (ok ? 1.0f : 0.0f), (ok && a.bias) ? a.bias[i] : 0.0f);
                    }
                asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * kPmsGroups) : "memory");
                tab_b = b;
            }
            const int w = wt * 30 + lane;
            const bool w_ok = w < a.Wout && lane < 30;
            __half* ybase = a.y + static_cast<long long>(b) * a.Cout * a.plane_out + w;
            mbar_wait(&tfull[acc], acc_ph, a.dbg, 4);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * NH * Nst;
            int hf = 0, cb = (eg + (kPmsGroups - 1) * ti) % kPmsGroups;   // unit u = hf * nblk + cb; this warp takes u = (eg - i) mod groups, + groups, ...
            ++ti;
            while (cb >= nblk) { cb -= nblk; ++hf; }
#pragma unroll 1
            while (hf < NH) {
                const int c0 = cb << 4;
                const int h = ht * TH + hf * 4 + q;
                const bool ok = w_ok && h < a.Hout;
                const uint32_t taddr = tacc + hf * Nst + c0;
                uint32_t v0[16], v1[16], v2[16];
                tmem_ld_32x32b_x16(taddr, v0);
                tmem_ld_32x32b_x16(taddr + Np, v1);
                tmem_ld_32x32b_x16(taddr + 2 * Np, v2);
                tmem_ld_wait();
                float sum[16];
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    sum[k] = (__uint_as_float(v0[k]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v1[k]), 1)) +
                             __shfl_down_sync(0xffffffffu, __uint_as_float(v2[k]), 2);
                if (pa.o.y) {
                    // channels-last: this lane's 16 couts of its pixel are 32 contiguous bytes
                    if (ok) {
                        const long long pixel = (static_cast<long long>(b) * a.Hout + h) * a.Wout + w;
                        float v[16];
                        const float4* tq = reinterpret_cast<const float4*>(tab + c0);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 sb = tq[k];
                            v[2 * k] = fmaf(sum[2 * k], sb.x, sb.y);
                            v[2 * k + 1] = fmaf(sum[2 * k + 1], sb.z, sb.w);
                        }
#pragma unroll
                        for (int k = 0; k < 16; ++k) v[k] = pa.o.alpha * (v[k] < 0.0f ? pa.o.slope * v[k] : v[k]);
                        auto add_res = [&](const __half* r, int cp, int off, float g) {
                            const uint4* rp = reinterpret_cast<const uint4*>(r + pixel * cp + off + c0);
                            const uint4 q0 = rp[0], q1 = rp[1];
                            const __half2* h0 = reinterpret_cast<const __half2*>(&q0);
                            const __half2* h1 = reinterpret_cast<const __half2*>(&q1);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 f0 = __half22float2(h0[k]), f1 = __half22float2(h1[k]);
                                v[2 * k] = fmaf(g, f0.x, v[2 * k]); v[2 * k + 1] = fmaf(g, f0.y, v[2 * k + 1]);
                                v[8 + 2 * k] = fmaf(g, f1.x, v[8 + 2 * k]); v[8 + 2 * k + 1] = fmaf(g, f1.y, v[8 + 2 * k + 1]);
                            }
                        };
                        if (pa.o.r1) add_res(pa.o.r1, pa.o.r1_cp, pa.o.r1_off, pa.o.beta);
                        if (pa.o.r2) add_res(pa.o.r2, pa.o.r2_cp, pa.o.r2_off, pa.o.gamma);
                        uint4 o0, o1;
                        __half2* p0 = reinterpret_cast<__half2*>(&o0);
                        __half2* p1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            p0[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                            p1[k] = __floats2half2_rn(v[8 + 2 * k], v[8 + 2 * k + 1]);
                        }
                        __half* dst = pa.o.y + pixel * pa.o.cp + pa.o.c_off + c0;
                        const int nco = a.Cout - c0;
                        if (nco >= 16) {
                            reinterpret_cast<uint4*>(dst)[0] = o0;
                            reinterpret_cast<uint4*>(dst)[1] = o1;
                        } else {   // a partial last block (e.g. the 3 couts of the output conv): element stores
#pragma unroll
                            for (int k = 0; k < 16; ++k)
                                if (k < nco) dst[k] = __float2half_rn(v[k]);
                        }
                    }
                    cb += kPmsGroups;
                    while (cb >= nblk) { cb -= nblk; ++hf; }
                    continue;
                }
                // lane = pixel: a warp's 2-byte stores of one cout cover 64 contiguous bytes of its row (no pair-swap shuffles)
                __half* yp = ybase + h * a.rs + c0 * a.cs;
                const float4* tp = reinterpret_cast<const float4*>(tab + c0);
                const int nco = a.Cout - c0;   // real couts in this block (>= 16 except in the last one)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 sb = tp[k];  // (d, bias) of couts c, c+1
                    const __half2 mine = __floats2half2_rn(fmaf(sum[2 * k], sb.x, sb.y), fmaf(sum[2 * k + 1], sb.z, sb.w));
                    if (lane < 30) track_abs(amax, mine);   // lanes 30 / 31 hold incomplete sums; padded couts carry (0, 0)
                    if (ok && 2 * k < nco) *yp = __low2half(mine);
                    if (ok && 2 * k + 1 < nco) *(yp + a.cs) = __high2half(mine);
                    yp += 2 * a.cs;
                }
                cb += kPmsGroups;
                while (cb >= nblk) { cb -= nblk; ++hf; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
        publish_abs(a.absmax, amax);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
    MB_WAIT_PROFILE_FLUSH(a.dbg);
}

// ---------------------------------------------------------------------------------------------------------
// Cout-major tile with the three kw taps STACKED ALONG M ("cms"), for 3x3 layers with at most 32 couts.
// The stacked pixel-major tile above sits at its instruction floor, 64 cycles of A-operand (pixel rows) read per MMA.  Here the
// pixels are the streamed B operand again (0.5 cycles per pixel and K = 16 step) and the M = 128 rows that a 32-cout layer
// wastes in the plain cout-major tile carry the kw taps instead:
//     D[(kw, cout), n] += A[(kw, cout), (kh, cin)] * B[(kh, cin), n],     row = 32 kw + cout  (rows 96..127 unused),
// n = r * 34 + col over a 34-pixel-wide patch of 5 image rows (one TMA load per 64-channel chunk; tap kh = a B window that starts
// kh * 34 pixel rows further on), N = 104 = 3 image rows per instruction, K = 3 Cin.  One MMA does the work of three of the plain
// tile.  The kw shift is a TMEM COLUMN offset: the warp of lane quadrant kw reads row r at columns [34 r + kw, 34 r + kw + 32),
//     y[cout, (r, col)] = D[(0, cout), 34 r + col] + D[(1, cout), 34 r + col + 1] + D[(2, cout), 34 r + col + 2],
// and the three partials of a cout sit in three different quadrants, i.e. three different warps: they meet through a shared-memory
// exchange (each warp reduces one of the tile's three rows, see the epilogue).  Four accumulators in TMEM, one epilogue group of
// four warps per tile, three groups: three tiles drain while the tensor core fills the fourth.  Weights resident: the A tile of
// (chunk, kh) is three 32-row TMA boxes of the ordinary packed matrix (one per kw) landing back to back.  NHWC = channels-last
// epilogue (bias, lrelu, residuals: csrc/rrdb.cu) through a per-warp transposition tile, else planar rows (d, bias, max |y|).
// Measured (B200, 16 frames): bit-correct, L13 0.86 ms and L12 0.99 ms against 0.70 / 0.85 ms for the stacked pixel-major tile,
// RRDBNet 68.6 against 57.5 ms per 4 frames: a group still needs ~4 000 cycles per 3-row tile.  Off by default.
constexpr int kCmsRows = 3;                     // image rows per tile: N = 3 * 34 + 2 = 104 accumulator columns, four accumulators in TMEM
constexpr int kCmsN = 104, kCmsAcc = 4, kCmsAccStride = 128;
constexpr int kCmsPatchAlloc = 22528, kCmsPatchTx = (kCmsRows + 2) * 34 * 128, kCmsATile = 96 * 128, kCmsGroups = 3;
constexpr int kCmsXchg = 3 * 32 * 33 * 4;       // per group: three rows x two source slots x 16 columns x 32 lanes of partial sums (12 KB);
                                                // the channels-last epilogue reuses it as three [32 couts][33] transposition tiles
constexpr int kCmsMaxStages = 6;

struct CmsArgs {
    KArgs k;
    int stages;
    ConvTcArgs::NhwcOut o;
};

template <bool NHWC>
__global__ void __launch_bounds__(128 + 128 * kCmsGroups, 1)
conv_cms_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, CmsArgs ca) {
    const KArgs& a = ca.k;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nst = ca.stages;
    const int n_atiles = 3 * a.nCC;
    // layout: [resident A tiles (chunk, kh): 96 rows each; the M = 128 read of the last one runs 4 KB into the first patch stage,
    //          which feeds accumulator rows nobody reads | patch stages | exchange / transposition tiles | barriers]
    uint8_t* stages = smem + n_atiles * kCmsATile;
    uint8_t* xchg = stages + nst * kCmsPatchAlloc;
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + kCmsGroups * kCmsXchg);
    uint64_t* full = bars;                          // [kCmsMaxStages]
    uint64_t* empty = bars + kCmsMaxStages;         // [kCmsMaxStages]
    uint64_t* tfull = bars + 2 * kCmsMaxStages;                  // [kCmsAcc]
    uint64_t* tempty = bars + 2 * kCmsMaxStages + kCmsAcc;       // [kCmsAcc]
    uint64_t* wfull = bars + 2 * kCmsMaxStages + 2 * kCmsAcc;    // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kCmsMaxStages + 2 * kCmsAcc + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int pad = a.pad;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < nst; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < kCmsAcc; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 4);      // the four warps of the group that owns the tile
        }
        mbar_init(wfull, 1);
        MB_WAIT_PROFILE_INIT();
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        const bool leader = elect_one();
        int s = 0;
        uint32_t ph = 0;
        if (leader && blockIdx.x < a.total_tiles) {
            mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(n_atiles * kCmsATile));
            for (int cc = 0; cc < a.nCC; ++cc)
                for (int kh = 0; kh < 3; ++kh)
                    for (int kw = 0; kw < 3; ++kw)
                        tma_load_2d(smem + (cc * 3 + kh) * kCmsATile + kw * 4096, &tmap_w, wfull, ((kw * a.nCC + cc) * 3 + kh) * kKC, 0);
        }
        __syncwarp();
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            int r = t;
            const int wt = r % a.tiles_w; r /= a.tiles_w;
            const int ht = r % a.tiles_h; r /= a.tiles_h;
            const int b = r;
            for (int cc = 0; cc < a.nCC; ++cc) {
                mbar_wait(&empty[s], ph ^ 1, a.dbg, 1);
                if (leader) {
                    mbar_arrive_expect_tx(&full[s], kCmsPatchTx);
                    tma_load_4d(stages + s * kCmsPatchAlloc, &tmap_x, &full[s], cc * kKC, wt * 32 - pad, ht * kCmsRows - pad, b);
                }
                __syncwarp();
                if (++s == nst) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_f16(128, kCmsN, 0, 0);
        int s = 0;
        uint32_t ph = 0;
        int acc = 0;
        uint32_t acc_ph = 0;
        if (blockIdx.x < a.total_tiles) mbar_wait(wfull, 0, a.dbg, 5);
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            mbar_wait(&tempty[acc], acc_ph ^ 1, a.dbg, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kCmsAccStride;
            uint32_t accumulate = 0;
            for (int cc = 0; cc < a.nCC; ++cc) {
                int nk16 = (a.Cin - cc * kKC + 15) / 16;
                if (nk16 > 4) nk16 = 4;
                mbar_wait(&full[s], ph, a.dbg, 3);
                tc_fence_after();
                if (leader) {
                    const uint64_t db = make_smem_desc(smem_u32(stages + s * kCmsPatchAlloc), 16, 1024, 2);
                    const uint64_t da = make_smem_desc(smem_u32(smem) + cc * 3 * kCmsATile, 16, 1024, 2);
                    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll 4
                        for (int j = 0; j < nk16; ++j) {
                            umma_f16(d_tmem, da + static_cast<uint64_t>((kh * kCmsATile + j * 32) >> 4),
                                     db + static_cast<uint64_t>((kh * 34 * 128 + j * 32) >> 4), idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == nst) { s = 0; ph ^= 1; }
            }
            if (leader) umma_commit(&tfull[acc]);
            __syncwarp();
            if (++acc == kCmsAcc) { acc = 0; acc_ph ^= 1; }
        }
    } else if (warp >= 4) {
        // Epilogue.  A tile is three image rows and belongs to ONE group of four warps (the CTA's i-th tile to group i mod 3, in
        // accumulator i mod 4), so three tiles drain concurrently while the tensor core fills a fourth: the first build shared every
        // tile between the groups and had two accumulators in flight -- each epilogue warp was busy ~2 500 serial cycles per tile
        // and the MMA warp waited 1 400..1 900 cycles per tile for an accumulator (wait profile).  Inside the group the quadrant-q
        // warp REDUCES row q: all three warps hold their kw partials of the three rows, hand the two rows they do not reduce to
        // the shared-memory exchange and take the other quadrants' partials of their own row back, in two halves of 16 columns
        // (tcgen05.ld x16 x 3 rows); every warp converts and stores one row.
        const int q = warp & 3;            // TMEM lane quadrant = kw of this warp's partial sums
        const int grp = (warp - 4) >> 2;
        constexpr int row0 = 0, nr = kCmsRows;
        uint8_t* xg = xchg + grp * kCmsXchg;
        const uint32_t xa = smem_u32(xg);
        float* tr = reinterpret_cast<float*>(xg) + q * (32 * 33);   // this warp's transposition tile (channels-last epilogue)
        const int bar_id = 1 + grp;        // named barrier of the group's quadrants 0..2 (96 threads)
        __half2 amax = __floats2half2_rn(0.0f, 0.0f);
        const bool co_ok = lane < a.Cout;
        const float bias = (co_ok && a.bias) ? a.bias[lane] : 0.0f;
        // division-free walk over this group's tiles: t = blockIdx.x + (grp + 3 j) * gridDim.x
        int wt, ht, b;
        {
            long long r = static_cast<long long>(blockIdx.x) + static_cast<long long>(grp) * gridDim.x;
            wt = static_cast<int>(r % a.tiles_w); r /= a.tiles_w;
            ht = static_cast<int>(r % a.tiles_h); b = static_cast<int>(r / a.tiles_h);
        }
        int dwt, dht, db;
        {
            long long r = 3LL * gridDim.x;
            dwt = static_cast<int>(r % a.tiles_w); r /= a.tiles_w;
            dht = static_cast<int>(r % a.tiles_h); db = static_cast<int>(r / a.tiles_h);
        }
        int li = grp;                      // index of the tile among this CTA's tiles
        for (long long t = static_cast<long long>(blockIdx.x) + static_cast<long long>(grp) * gridDim.x; t < a.total_tiles;
             t += 3LL * gridDim.x, li += 3, wt += dwt, ht += dht, b += db) {
            if (wt >= a.tiles_w) { wt -= a.tiles_w; ++ht; }
            if (ht >= a.tiles_h) { ht -= a.tiles_h; ++b; }
            const int acc = li & (kCmsAcc - 1);
            const uint32_t acc_ph = (li >> 2) & 1;
            const int h0 = ht * kCmsRows, w0 = wt * 32;
            const float scale = (co_ok && a.d) ? a.d[b * a.Cout + lane] : (co_ok ? 1.0f : 0.0f);
            mbar_wait(&tfull[acc], acc_ph, a.dbg, 4);
            tc_fence_after();
            auto release = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
            };
            if (q == 3) {
                release();                     // rows 96..127 of the accumulator carry nothing
            } else {
                const uint32_t tb = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kCmsAccStride + q + 34 * row0;
                const bool reducer = q < nr;   // this warp reduces (and stores) row row0 + q
                float f[32];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[3][16];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (i < nr) tmem_ld_32x32b_x16(tb + 34 * i + 16 * hf, v[i]);
                    tmem_ld_wait();
                    if (hf == 1) release();    // this warp's last load of the tile has landed
                    // the exchange tile is free: every warp is done with its partial loads / transposition reads of the last use
                    asm volatile("bar.sync %0, 96;" ::"r"(bar_id) : "memory");
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        if (i < nr && i != q) {
                            const int sl = q - (q > i ? 1 : 0);   // source slot of this quadrant in row i's pair (ascending kw)
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xa + (((i * 2 + sl) * 4 + j) * 32 + lane) * 16), "r"(v[i][4 * j]),
                                             "r"(v[i][4 * j + 1]), "r"(v[i][4 * j + 2]), "r"(v[i][4 * j + 3])
                                             : "memory");
                        }
                    }
                    asm volatile("bar.sync %0, 96;" ::"r"(bar_id) : "memory");
                    if (reducer) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t p1[4], p2[4];
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(p1[0]), "=r"(p1[1]), "=r"(p1[2]), "=r"(p1[3]) : "r"(xa + (((q * 2 + 0) * 4 + j) * 32 + lane) * 16));
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(p2[0]), "=r"(p2[1]), "=r"(p2[2]), "=r"(p2[3]) : "r"(xa + (((q * 2 + 1) * 4 + j) * 32 + lane) * 16));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                // v[q] with a register index known at compile time: select by the warp-uniform q
                                const uint32_t own = q == 0 ? v[0][4 * j + k] : (q == 1 ? v[1][4 * j + k] : v[2][4 * j + k]);
                                f[16 * hf + 4 * j + k] = (__uint_as_float(own) + __uint_as_float(p1[k])) + __uint_as_float(p2[k]);
                            }
                        }
                    }
                }
                if (reducer) {
                    const int h = h0 + row0 + q;
                    if constexpr (!NHWC) {
                        // planar rows: lane = cout writes its 32 pixels of image row h
                        uint32_t pk[16];
#pragma unroll
                        for (int k2 = 0; k2 < 16; ++k2) {
                            const __half2 hv = __floats2half2_rn(fmaf(f[2 * k2], scale, bias), fmaf(f[2 * k2 + 1], scale, bias));
                            track_abs(amax, hv);
                            pk[k2] = *reinterpret_cast<const uint32_t*>(&hv);
                        }
                        if (co_ok && h < a.Hout) {
                            __half* yrow = a.y + static_cast<long long>(b) * a.Cout * a.plane_out + lane * a.cs + h * a.rs + w0;
                            if (a.st256) {
#pragma unroll
                                for (int g = 0; g < 2; ++g)
                                    if (w0 + 16 * g < a.Wp_out)
                                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yrow + 16 * g), "r"(pk[8 * g]),
                                                     "r"(pk[8 * g + 1]), "r"(pk[8 * g + 2]), "r"(pk[8 * g + 3]), "r"(pk[8 * g + 4]), "r"(pk[8 * g + 5]),
                                                     "r"(pk[8 * g + 6]), "r"(pk[8 * g + 7])
                                                     : "memory");
                            } else {
#pragma unroll
                                for (int g = 0; g < 4; ++g)
                                    if (w0 + 8 * g < a.Wp_out)
                                        *reinterpret_cast<uint4*>(yrow + 8 * g) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                            }
                        }
                    }
                }
                if constexpr (NHWC) {
                    // channels-last: transpose the [32 couts][32 pixels] block through this warp's tile so that a lane owns a pixel.
                    // The tiles alias the exchange tile: wait until every warp of the group has read its partials.
                    asm volatile("bar.sync %0, 96;" ::"r"(bar_id) : "memory");
                    if (reducer) {
                        const int h = h0 + row0 + q;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float y = fmaf(f[j], scale, bias);
                            y = ca.o.alpha * (y < 0.0f ? ca.o.slope * y : y);
                            tr[lane * 33 + j] = y;
                        }
                        __syncwarp();
                        const int w = w0 + lane;
                        float o[32];
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c] = tr[c * 33 + lane];
                        if (h < a.Hout && w < a.Wout) {
                            const long long pixel = (static_cast<long long>(b) * a.Hout + h) * a.Wout + w;
                            auto add_res = [&](const __half* rp, int cp, int off, float gsc) {
                                const uint4* rv = reinterpret_cast<const uint4*>(rp + pixel * cp + off);
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    const uint4 qv = rv[g];
                                    const __half2* hh = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const float2 ff = __half22float2(hh[k]);
                                        o[8 * g + 2 * k] = fmaf(gsc, ff.x, o[8 * g + 2 * k]);
                                        o[8 * g + 2 * k + 1] = fmaf(gsc, ff.y, o[8 * g + 2 * k + 1]);
                                    }
                                }
                            };
                            if (ca.o.r1) add_res(ca.o.r1, ca.o.r1_cp, ca.o.r1_off, ca.o.beta);
                            if (ca.o.r2) add_res(ca.o.r2, ca.o.r2_cp, ca.o.r2_off, ca.o.gamma);
                            __half* dst = ca.o.y + pixel * ca.o.cp + ca.o.c_off;
                            if (a.Cout == 32) {
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    uint4 ov;
                                    __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(o[8 * g + 2 * k], o[8 * g + 2 * k + 1]);
                                    reinterpret_cast<uint4*>(dst)[g] = ov;
                                }
                            } else {
#pragma unroll
                                for (int c = 0; c < 32; ++c)
                                    if (c < a.Cout) dst[c] = __float2half_rn(o[c]);
                            }
                        }
                    }
                }
            }
        }
        if constexpr (!NHWC) publish_abs(a.absmax, amax);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
    MB_WAIT_PROFILE_FLUSH(a.dbg);
}

}  // namespace

int conv_tc_smem_bytes(int /*tw*/) { return kSmemMax; }

int conv_tc_launch(const ConvTcArgs& p, cudaStream_t stream) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return MB_ECUDA;
    }
    const int Np = round_up(p.Cout, 16);
    // pixel-major pays off where the cout-major tile wastes >= half of its M rows and K is deep enough to amortise
    // the per-tile epilogue (B200 A/B, r1: L11 1.38 -> 1.23 ms, L12 0.77 -> 0.66 ms, but L13 with Cin = 32 0.50 -> 0.60 ms)
    // cout-major stacked tile (conv_cms_kernel): 3x3 layers with at most 32 couts whose A tiles stay resident next to two patch stages
    const int cms_fixed = 1024 + 256 + kCmsGroups * kCmsXchg;
    const bool cms = p.cm_stack && p.ksz == 3 && p.Cout <= 32 && p.split_lo_off <= 0 &&
                     3 * ceil_div(p.Cin, kKC) * kCmsATile + 2 * kCmsPatchAlloc + cms_fixed <= kSmemMax;
    // stacked pixel-major tile (conv_pms_kernel): 3x3 layers whose three kw blocks fit one instruction (3 Np <= 256) and whose
    // weights stay resident next to two patch stages
    const int pms_nh = (!cms && p.pm_max_cout > 0 && p.pm_stack && p.ksz == 3 && Np <= p.pm_max_cout && 3 * Np <= 256 && p.split_lo_off <= 0)
                           ? ((12 * Np <= 512 && 9 * ceil_div(p.Cin, kKC) * Np * 128 + 2 * 10 * kPmSlab + 1024 + 4096 + 256 <= kSmemMax) ? 2
                              : (9 * ceil_div(p.Cin, kKC) * Np * 128 + 2 * 6 * kPmSlab + 1024 + 4096 + 256 <= kSmemMax ? 1 : 0))
                           : 0;
    // (1x1 layers, StyleGAN3-R: the cout-major tile with three epilogue groups and 32-byte stores wins, 1.92 -> 1.37 ms on L12 / L13)
    MB_REQUIRE(p.nhwc.y == nullptr || pms_nh > 0 || cms, "conv_tc: the channels-last epilogue lives in the stacked tiles (3x3, <= 80 couts, resident weights)");
    const bool pixel_major = !cms && (pms_nh > 0 || (p.pm_max_cout > 0 && Np <= p.pm_max_cout && Np <= 128 && ((p.Cin > 32 && p.ksz > 1) || p.pm_max_cout > 64)));
    const int tw = pixel_major ? kPmTW : (p.tile_w == 16 ? 16 : 32);
    const int th = kTileN / tw;
    const int pad = p.pad;
    const int halo = p.ksz - 1;
    MB_REQUIRE(pad >= 0 && pad <= halo, "conv_tc: padding %d unsupported for kernel size %d", pad, p.ksz);
    MB_REQUIRE(p.ksz >= 1 && p.ksz <= 3, "conv_tc: kernel size %d unsupported", p.ksz);
    MB_REQUIRE(p.Cp_in % 8 == 0 && p.Wp_out % 8 == 0, "conv_tc: channel / row pitch must be a multiple of 8 elements");
    MB_REQUIRE((reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(p.wpk) & 15) == 0,
               "conv_tc: pointers must be 16-byte aligned");
    const int nCC = ceil_div(p.Cin, kKC);
    const int Mp = round_up(p.Cout, kTileM);
    const long long Ktot = static_cast<long long>(p.ksz) * p.ksz * nCC * kKC;

    // shared-memory plan of the cout-major kernel (layout comment in conv_tc_kernel)
    const int a_rows = (Mp == kTileM && p.narrow_a) ? round_up(p.Cout, 8) : kTileM;
    const int a_tile_bytes = a_rows * kKC * 2;
    const int w_all = p.ksz * p.ksz * nCC * a_tile_bytes;
    // shift 2 (KArgs::th): 34-pixel-wide patch of 9 image rows; the windows of the two unused accumulator columns per row
    // and of columns 238..239 read up to pixel row 309 of the stage
    const int patch2_alloc = 40960, patch2_tx = 9 * 34 * 128;
    const bool shift2 = !cms && !pixel_major && p.cm_shift == 2 && p.ksz == 3 && tw == 32 && p.narrow_a && Mp == kTileM && p.split_lo_off <= 0 &&
                        (p.epi_groups == 0 || p.epi_groups == 3) && w_all + 2 * patch2_alloc + 1024 + 256 <= kSmemMax;
    const int patch_bytes = shift2 ? patch2_alloc : (th + halo) * tw * 128;
    // shift mode (one patch load per chunk, see KArgs) needs resident weights and room for its epilogue staging tiles
    const bool want_shift = !cms && !pixel_major && p.cm_shift == 1 && p.ksz == 3 && tw == 32 && p.narrow_a && Mp == kTileM &&
                            w_all + 2 * patch_bytes + 1024 + 256 + kEpiStageBytes <= kSmemMax;
    MB_REQUIRE(p.split_lo_off <= 0 || (!pixel_major && !want_shift), "conv_tc: the split output needs the plain cout-major tile");
    const int fixed_bytes = 1024 /*align*/ + 256 /*barriers*/ + (want_shift ? kEpiStageBytes : 0);
    const bool resident = p.narrow_a && Mp == kTileM && w_all + 2 * patch_bytes + fixed_bytes <= kSmemMax;
    const int stage_alloc = patch_bytes + (resident ? 0 : p.ksz * a_tile_bytes);
    int nstages = (kSmemMax - fixed_bytes - (resident ? w_all : 0)) / stage_alloc;
    if (nstages > kMaxStages) nstages = kMaxStages;
    MB_REQUIRE(pixel_major || nstages >= 2, "conv_tc: shared-memory plan does not fit (%d-byte stages)", stage_alloc);
    // the UMMA reads 128 A rows from every tile start: the last tile must still end inside the allocation
    const int smem_cm = fixed_bytes + (resident ? w_all : 0) + nstages * stage_alloc;

    CUtensorMap tm_w, tm_x;
    {
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(Ktot), static_cast<cuuint64_t>(Mp)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(Ktot) * 2};
        cuuint32_t box[2] = {kKC, static_cast<cuuint32_t>(cms ? 32 : (pixel_major ? Np : a_rows))};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(p.wpk), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(weights) failed: %d", static_cast<int>(r));
            return MB_ECUDA;
        }
    }
    {
        const cuuint64_t cp = static_cast<cuuint64_t>(p.Cp_in);
        // channel extent = the padded pitch: the pad channels hold real zeros (written by the producers), which keeps
        // the inner box dimension in bounds wherever Cp is a multiple of 64
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(p.Cp_in), static_cast<cuuint64_t>(p.Win),
                              static_cast<cuuint64_t>(p.Hin), static_cast<cuuint64_t>(p.B)};
        cuuint64_t strides[3] = {cp * 2, cp * 2 * p.Win, cp * 2 * p.Win * p.Hin};
        cuuint32_t box[4] = {kKC, static_cast<cuuint32_t>((shift2 || cms) ? 34 : tw),
                             static_cast<cuuint32_t>(pms_nh ? 4 * pms_nh + 2 : (cms ? kCmsRows + 2 : (shift2 ? 9 : th + halo))), 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(p.x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(activations) failed: %d (W=%d C=%d H=%d B=%d Cp=%d)",
                      static_cast<int>(r), p.Win, p.Cin, p.Hin, p.B, p.Cp_in);
            return MB_ECUDA;
        }
    }

    KArgs a;
    a.B = p.B; a.Cin = p.Cin; a.Cout = p.Cout;
    a.Hout = p.Hin + 2 * pad - halo; a.Wout = p.Win + 2 * pad - halo; a.Wp_out = p.Wp_out;
    a.ksz = p.ksz; a.nCC = nCC; a.pad = pad;
    a.tiles_m = Mp / kTileM;
    a.tiles_w = ceil_div(a.Wout, tw);
    a.tiles_h = ceil_div(a.Hout, th);
    a.total_tiles = p.B * a.tiles_h * a.tiles_w * a.tiles_m;
    a.d = p.d; a.bias = p.bias; a.y = p.y;
    a.plane_out = static_cast<long long>(a.Hout) * p.Wp_out;
    a.cs = p.row_interleaved ? p.Wp_out : a.plane_out;
    a.rs = p.row_interleaved ? static_cast<long long>(p.Cout) * p.Wp_out : p.Wp_out;
    MB_REQUIRE(!p.row_interleaved || (p.split_lo_off <= 0 && p.cm_shift != 1), "conv_tc: the row-interleaved output needs a plain epilogue");
    a.dbg = debug_words_device();
    a.absmax = p.absmax;
    a.lo_off = p.split_lo_off;
    a.st256 = 0;
    a.a_tile_bytes = a_tile_bytes;
    a.resident = resident ? 1 : 0;
    a.nstages = nstages;
    a.stage_bytes = stage_alloc;
    a.shift = (want_shift && resident) ? 1 : 0;
    a.tile_wv = a.shift ? tw - halo : tw;
    if (a.shift) {
        a.tiles_w = ceil_div(a.Wout, a.tile_wv);
        a.total_tiles = p.B * a.tiles_h * a.tiles_w * a.tiles_m;
    }
    a.th = th; a.slab = tw * 128; a.colp = 32; a.nchunks = kTileN / 32; a.mma_n = kTileN; a.patch_tx = (th + halo) * tw * 128;
    if (shift2) {
        MB_REQUIRE(resident, "conv_tc: shift 2 needs resident weights");
        a.shift = 2;
        a.th = 7; a.slab = 34 * 128; a.colp = 34; a.nchunks = 7; a.mma_n = 240; a.patch_tx = patch2_tx;
        a.tiles_h = ceil_div(a.Hout, a.th);
        a.total_tiles = p.B * a.tiles_h * a.tiles_w * a.tiles_m;
    }

    if (cms) {
        CmsArgs cg;
        cg.k = a;
        cg.o = p.nhwc;
        cg.k.tiles_m = 1;
        cg.k.tiles_w = ceil_div(a.Wout, 32);
        cg.k.tiles_h = ceil_div(a.Hout, kCmsRows);
        cg.k.total_tiles = p.B * cg.k.tiles_h * cg.k.tiles_w;
        cg.k.st256 = (p.Wp_out % 16 == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0) ? 1 : 0;
        const int a_all = 3 * nCC * kCmsATile;
        cg.stages = (kSmemMax - cms_fixed - a_all) / kCmsPatchAlloc;
        if (cg.stages > kCmsMaxStages) cg.stages = kCmsMaxStages;
        const int smem_bytes = cms_fixed + a_all + cg.stages * kCmsPatchAlloc;
        static bool attr_cms = false;
        if (!attr_cms) {
            MB_CUDA(cudaFuncSetAttribute(conv_cms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            MB_CUDA(cudaFuncSetAttribute(conv_cms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            attr_cms = true;
        }
        int grid = cg.k.total_tiles < p.num_sms ? cg.k.total_tiles : p.num_sms;
        if (grid < 1) grid = 1;
        if (p.nhwc.y) conv_cms_kernel<true><<<grid, 128 + 128 * kCmsGroups, smem_bytes, stream>>>(tm_w, tm_x, cg);
        else conv_cms_kernel<false><<<grid, 128 + 128 * kCmsGroups, smem_bytes, stream>>>(tm_w, tm_x, cg);
        MB_CUDA(cudaGetLastError());
        return MB_OK;
    }
    if (pms_nh) {
        PmsArgs pa;
        pa.k = a;
        pa.Np = Np; pa.Nst = 3 * Np; pa.o = p.nhwc;
        pa.k.tiles_m = 1;
        pa.k.tiles_w = ceil_div(a.Wout, 30);
        pa.k.tiles_h = ceil_div(a.Hout, 4 * pms_nh);
        pa.k.total_tiles = p.B * pa.k.tiles_h * pa.k.tiles_w;
        const int patch = (4 * pms_nh + 2) * kPmSlab;
        const int fixed = 1024 + 4096 + 256;
        const int w_all_pm = 9 * nCC * Np * 128;
        pa.stages = (kSmemMax - fixed - w_all_pm) / patch;
        if (pa.stages > kPmMaxStages) pa.stages = kPmMaxStages;
        MB_REQUIRE(pa.stages >= 2, "conv_tc: stacked pixel-major shared-memory plan does not fit");
        const int smem_bytes = fixed + w_all_pm + pa.stages * patch;
        static bool attr_pms = false;
        if (!attr_pms) {
            MB_CUDA(cudaFuncSetAttribute((conv_pms_kernel<1, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            MB_CUDA(cudaFuncSetAttribute((conv_pms_kernel<2, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            MB_CUDA(cudaFuncSetAttribute((conv_pms_kernel<1, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            MB_CUDA(cudaFuncSetAttribute((conv_pms_kernel<2, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            attr_pms = true;
        }
        int grid = pa.k.total_tiles < p.num_sms ? pa.k.total_tiles : p.num_sms;
        if (grid < 1) grid = 1;
        if (p.nhwc.y) {
            if (pms_nh == 2) conv_pms_kernel<2, 2><<<grid, 384, smem_bytes, stream>>>(tm_w, tm_x, pa);
            else conv_pms_kernel<1, 2><<<grid, 384, smem_bytes, stream>>>(tm_w, tm_x, pa);
        } else {
            if (pms_nh == 2) conv_pms_kernel<2, 3><<<grid, 512, smem_bytes, stream>>>(tm_w, tm_x, pa);
            else conv_pms_kernel<1, 3><<<grid, 512, smem_bytes, stream>>>(tm_w, tm_x, pa);
        }
        MB_CUDA(cudaGetLastError());
        return MB_OK;
    }
    if (pixel_major) {
        PmArgs pa;
        pa.k = a;
        pa.k.tiles_m = 1;
        pa.k.total_tiles = p.B * a.tiles_h * a.tiles_w;
        pa.Np = Np;
        pa.shift = 0;
        pa.tile_w = kPmTW;
        const int w_tile = Np * 128;
        const int pm_patch = (kPmTH + halo) * kPmSlab;
        const int pm_fixed = 1024 /*align*/ + 4 * 128 * 8 /*(scale,bias) tables*/ + 256 /*barriers*/;
        const int pm_w_all = p.ksz * p.ksz * nCC * w_tile;
        pa.resident = (pm_w_all + 2 * pm_patch + pm_fixed <= kSmemMax) ? 1 : 0;
        pa.shift = (p.ksz == 3 && pa.resident) ? p.pm_shift : 0;
        pa.tile_w = pa.shift ? 30 : kPmTW;
        if (pa.shift) {
            pa.k.tiles_w = ceil_div(a.Wout, pa.tile_w);
            pa.k.total_tiles = p.B * a.tiles_h * pa.k.tiles_w;
        }
        pa.stage_bytes_alloc = pm_patch + (pa.resident ? 0 : p.ksz * w_tile);
        pa.stages = (kSmemMax - pm_fixed - (pa.resident ? pm_w_all : 0)) / pa.stage_bytes_alloc;
        if (pa.stages > kPmMaxStages) pa.stages = kPmMaxStages;
        MB_REQUIRE(pa.stages >= 2, "conv_tc: pixel-major shared-memory plan does not fit");
        const int smem_bytes = pm_fixed + (pa.resident ? pm_w_all : 0) + pa.stages * pa.stage_bytes_alloc;
        static bool attr_pm = false;
        if (!attr_pm) {
            MB_CUDA(cudaFuncSetAttribute(conv_pm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            MB_CUDA(cudaFuncSetAttribute(conv_pm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            attr_pm = true;
        }
        int grid = pa.k.total_tiles < p.num_sms ? pa.k.total_tiles : p.num_sms;
        if (grid < 1) grid = 1;
        if (pa.shift) conv_pm_kernel<true><<<grid, 256, smem_bytes, stream>>>(tm_w, tm_x, pa);
        else conv_pm_kernel<false><<<grid, 256, smem_bytes, stream>>>(tm_w, tm_x, pa);
        MB_CUDA(cudaGetLastError());
        return MB_OK;
    }
    int grid = a.total_tiles < p.num_sms ? a.total_tiles : p.num_sms;
    if (grid < 1) grid = 1;
    // Epilogue groups: one warp per TMEM quadrant drains a 256-column accumulator in ~5 400 cycles (measured, L13); the tensor
    // core fills the next one in 128 cycles per K = 16 step.  Three groups where the MMA time per tile is below that, or where
    // few couts leave most quadrants idle.
    int mma_steps = 0;
    for (int cc = 0; cc < nCC; ++cc) mma_steps += p.ksz * p.ksz * ((p.Cin - cc * kKC + 15) / 16 > 4 ? 4 : (p.Cin - cc * kKC + 15) / 16);
    int eg = p.epi_groups;
    if (eg == 0) eg = (tw == 32 && a.shift != 1 && a.lo_off <= 0 && (a.shift == 2 || mma_steps * 128 < 6000 || p.Cout <= 64)) ? 3 : 1;
    MB_REQUIRE((eg == 1 && a.shift != 2) || (eg == 3 && tw == 32 && a.shift != 1 && a.lo_off <= 0), "conv_tc: epilogue groups %d unsupported here", eg);
    a.st256 = (p.Wp_out % 16 == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0) ? 1 : 0;
    if (tw == 32 && eg == 3) {
        static bool attr_done = false;
        if (!attr_done) {
            MB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            attr_done = true;
        }
        conv_tc_kernel<32, 3><<<grid, 512, smem_cm, stream>>>(tm_w, tm_x, a);
    } else if (tw == 32) {
        static bool attr_done = false;
        if (!attr_done) {
            MB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            attr_done = true;
        }
        conv_tc_kernel<32, 1><<<grid, 256, smem_cm, stream>>>(tm_w, tm_x, a);
    } else {
        static bool attr_done = false;
        if (!attr_done) {
            MB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            attr_done = true;
        }
        conv_tc_kernel<16, 1><<<grid, 256, smem_cm, stream>>>(tm_w, tm_x, a);
    }
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

}  // namespace mb
