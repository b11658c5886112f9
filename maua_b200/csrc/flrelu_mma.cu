// filtered_lrelu on the tensor cores: the four separable FIR passes of a StyleGAN3 layer
// (up-H, up-V, [lrelu], down-H, down-V) as a register-resident chain of banded Toeplitz GEMMs.
//
// Why: on CUDA cores this op needs ~72 useful FMAs (154 issued instructions) per output pixel and
// is issue-bound at ~5 % of HBM bandwidth (profiles/r1_ncu_v1_first_path.md).  A 1-D FIR is a
// multiplication by a banded constant matrix, so each pass is D = A_const * B with the filter as the
// A operand of mma.sync.m16n8k16 (fp16 in, fp32 accumulate).  The accumulator fragment of one pass is
// bit-for-bit the B fragment layout of the next pass *transposed*, which is exactly what a separable
// 2-D filter needs (pass k contracts the axis pass k-1 left untouched):
//     S1  A1^T[Jx][iy] = sum_ix Uh[Jx][ix] X[iy][ix]        (B from shared memory, X = conv output tile)
//     S2  T   [Jy][Jx] = sum_iy Uv[Jy][iy] A1^T[Jx][iy]     -> lrelu * gain, clamp
//     S3  O3^T[ox][Jy] = sum_Jx Dh[ox][Jx] T[Jy][Jx]
//     S4  Out [oy][ox] = sum_Jy Dv[oy][Jy] O3^T[ox][Jy]
// so one warp carries a 32x32 output tile of one channel from the fp16 input tile to the fp16 output
// without any intermediate leaving the register file (no shared-memory round trips, no block barriers).
// The (2*32+10)^2 upsampled intermediate is produced and consumed in five 16-row strips.
// Filter taps are single fp16 values.  Rounding every tap to nearest is a *systematic* DC-gain error per polyphase
// branch that costs up to 1e-3 in the final pixels (scratch/emul_mma_fir.py), so the host picks, per branch, the
// round-up/round-down combination of its six taps whose sum is closest to the exact branch gain (tune_branch,
// residual ~1e-6), and the down filter's DC error is one fp32 factor in the output scale.  That leaves the
// activation free of per-lane fp32 multipliers: it runs on packed halves as max(t, slope*t) and a clamp (4 half2
// instructions per 4 values), with the lrelu gain moved behind the (linear) down filter into the output scale.
// Intermediates are rounded to fp16 between passes (measured harmless: 2.8e-4 vs 2.7e-4 max pixel error end to end).
//
// Index conventions as in flrelu.cu.  Supported: down=2/12 taps with up=2/12 taps or up=4/24 taps,
// separable filters (every non-ToRGB layer of StyleGAN3-T and the critically sampled layers of -R).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace mb {
namespace {

__device__ __forceinline__ void mma16816(float (&c)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// first k-step of an accumulator: C = 0 (no register zeroing moves)
__device__ __forceinline__ void mma16816_z(float (&c)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "f"(0.0f));
}
// fp16-accumulator form, C = 0: the two result registers are (row g: cols 2tig, 2tig+1) and (row g+8: same cols) as packed
// halves -- bit for bit the B fragment the next pass wants, so the up-sampling passes (single k-step, result rounded to
// fp16 anyway) need neither fp32 accumulators nor F2FP packs (160 of ~1150 issued instructions per tile).
__device__ __forceinline__ void mma16816_h(uint32_t (&d)[2], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%8,%8};"
        : "=r"(d[0]), "=r"(d[1])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "r"(0u));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// leaky relu (0 <= slope <= 1, gain applied later) and clamp on two packed halves
__device__ __forceinline__ uint32_t lrelu_clamp2(uint32_t v, uint32_t s, uint32_t c) {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    const __half2 sh = *reinterpret_cast<__half2*>(&s), ch = *reinterpret_cast<__half2*>(&c);
    h = __hmax2(h, __hmul2(h, sh));
    h = __hmin2(__hmax2(h, __hneg2(ch)), ch);
    return *reinterpret_cast<uint32_t*>(&h);
}
// leaky relu alone (the clamp is known to be inactive, see StreamParams::in_absmax)
__device__ __forceinline__ uint32_t lrelu2(uint32_t v, uint32_t s) {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    h = __hmax2(h, __hmul2(h, *reinterpret_cast<__half2*>(&s)));
    return *reinterpret_cast<uint32_t*>(&h);
}
// leaky relu as ONE instruction (HFMA2 with an |x| operand): lrelu(t) = a * (|t| + k t) with k = (1 + slope) / (1 - slope) and
// a = (1 - slope) / 2; the factor a rides in the fp32 output scale behind the (linear) down filters.  For slope = 0.2, k = 1.5 is
// exact in fp16: negative inputs become 0.5 t (exact), positive ones 2.5 t (one fp16 rounding), and the realised slope is exactly
// 0.2 (the two-instruction form multiplies by fp16(0.2) = 0.19995).  In general a = 1 / (k + 1) of the ROUNDED k keeps the
// positive gain at exactly one.
__device__ __forceinline__ uint32_t lrelu2_abs(uint32_t v, uint32_t k) {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    h = __hfma2(h, *reinterpret_cast<__half2*>(&k), __habs2(h));
    return *reinterpret_cast<uint32_t*>(&h);
}
// first input sample an axis needs for output o0: ceil((2*o0 - pad) / UP) - e   (UP is 2 or 4: shifts)
template <int UP>
__device__ __forceinline__ int first_in(int o0, int pad, int e) {
    return ((2 * o0 - pad + UP - 1) >> (UP == 2 ? 1 : 2)) - e;
}

constexpr int kOT = 32;      // output tile edge per warp
constexpr int kStrips = 5;   // 16-row strips of the intermediate (2*32+10 = 74 -> 80 rows)
constexpr int kJB = 10;      // n8 blocks across the intermediate (80 columns)
constexpr int kMB = 5;       // m16 blocks across the intermediate
constexpr int kXP = 56;      // shared-memory row pitch of the input tile, in halfs (conflict-free B loads)
constexpr int kWarps = 8;

template <int UP>
struct MC {
    static constexpr int NIB = (UP == 2) ? 6 : 4;   // n8 blocks of input rows/cols a tile needs
    static constexpr int IYT = NIB * 8;
    static constexpr int NVAR = (UP == 2) ? 1 : 2;   // distinct up-filter fragments (window offset 0 / 4)
    static constexpr int NFRAG = NVAR + 3;            // NVAR up fragments and 3 down fragments
    static constexpr int XBYTES = IYT * kXP * 2;
    static constexpr int SMEM = kWarps * 2 * XBYTES;  // double-buffered input tile per warp
    __host__ __device__ static constexpr int wblk(int j) { return UP == 2 ? j : (j >> 1); }
    __host__ __device__ static constexpr int var(int j) { return UP == 2 ? 0 : (j & 1); }
};

struct MmaParams {
    const __half* x;
    const float* bias;
    const float* scale;
    __half* y;
    const uint4* frags;  // [NFRAG][32]
    int C, Hin, Win, Wp_in, Hout, Wout, Wp_out, Cp_out, px0, py0, tiles_x, tiles_y, e, tpw;
    float slope;
    float clamp_pre;  // clamp / gain: the clamp acts before the gain, which is folded into out_gain
    float out_gain;   // gain * (DC correction of the fp16-rounded down filter)^2
    int rho;          // phase of the first intermediate sample of a tile: pad mod UP
    int nt;           // radial down filter: number of separable terms (0 = separable filter, fragments in registers)
    const uint4* rfrags;  // [nt][6][32] fragments of the terms (down-H k-steps 0..2, down-V k-steps 0..2)
};

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// The four-pass chain for one 32x32 output tile; X = fp16 input tile in shared memory (row pitch kXP),
// dx = even column offset of the first needed input sample inside the tile rows.
// sl2 / cl2: slope and clamp (pre-gain) replicated in both halves.
struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};
// The input tile is consumed in two halves of NIB/2 row blocks.  `dead0` / `dead1` are invoked right after the last
// read of the first / second half, `wait1` right before the first read of the second half: a single-buffered caller
// refills each half with the next tile's rows as soon as it is dead (and only blocks on the second half when it
// gets there), so the fetch latency hides behind two to four strips of MMAs.
// Radial (non-separable) down filters of StyleGAN3-R: the 12x12 jinc*kaiser filter is symmetric and numerically of rank
// 2..4 (eigenvalues fall off by 1e-2 per term), so it runs as `nt` separable terms F = sum_k g_k h_k^T whose fragments
// (per term: three down-H and three down-V k-steps) sit in shared memory (`sAD`, [nt][6][32] uint4) and accumulate
// into the same output registers.  nt == 0 selects the separable filter held in registers (AD).
template <int UP, bool RAD = false, class W1 = NoHook, class D0 = NoHook, class D1 = NoHook>
__device__ __forceinline__ void fir_chain(const __half* X, int dx, const uint4 (&AU)[MC<UP>::NVAR], const uint4 (&AD)[3],
                                          uint32_t sl2, uint32_t cl2, int g, int tig,
                                          float (&OUT)[2][4][4], W1 wait1 = W1(), D0 dead0 = D0(), D1 dead1 = D1(),
                                          const uint4* sAD = nullptr, int nt = 0) {
    using K = MC<UP>;
    uint32_t P1[2][kMB][2];  // packed A1^T of the two input-row n8 blocks the strip window covers (slot = block & 1)
    int have0 = -1, have1 = -1;  // compile-time constants after unrolling
    if constexpr (RAD) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int no = 0; no < 4; ++no)
#pragma unroll
                for (int q = 0; q < 4; ++q) OUT[i][no][q] = 0.0f;
    }

    auto stage1 = [&](int blk) {
        if (blk == K::NIB / 2) wait1();
#pragma unroll
        for (int m = 0; m < kMB; ++m) {
            const int w0 = K::wblk(m) * 8;
            const __half* src = X + (blk * 8 + g) * kXP + dx + w0 + 2 * tig;
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(src);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(src + 8);
            mma16816_h(P1[blk & 1][m], AU[K::var(m)], b0, b1);
        }
        if (blk == K::NIB / 2 - 1) dead0();
        if (blk == K::NIB - 1) dead1();
    };

#pragma unroll
    for (int j = 0; j < kStrips; ++j) {
        // ---- S1 for the two input-row blocks this strip's window covers
        const int wb = K::wblk(j);
        if (((wb & 1) ? have1 : have0) != wb) {
            stage1(wb);
            if (wb & 1) have1 = wb; else have0 = wb;
        }
        if ((((wb + 1) & 1) ? have1 : have0) != wb + 1) {
            stage1(wb + 1);
            if ((wb + 1) & 1) have1 = wb + 1; else have0 = wb + 1;
        }
        // ---- S2 (+activation): T[16 rows of strip j][80 cols], packed as B operands of S3
        uint32_t P2[kJB][2];
#pragma unroll
        for (int nb = 0; nb < kJB; ++nb) {
            uint32_t t2[2];
            mma16816_h(t2, AU[K::var(j)], P1[wb & 1][nb >> 1][nb & 1], P1[(wb + 1) & 1][nb >> 1][nb & 1]);
            // activation on the packed halves (an overflow arrives as inf and is clamped all the same)
            P2[nb][0] = lrelu_clamp2(t2[0], sl2, cl2);
            P2[nb][1] = lrelu_clamp2(t2[1], sl2, cl2);
        }
        if constexpr (!RAD) {
            // ---- S3: O3^T[ox m16 block mo][16 rows of strip j], packed as B operands of S4
            uint32_t P3[4][2];
#pragma unroll
            for (int mo = 0; mo < 2; ++mo) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float acc[4];
                    mma16816_z(acc, AD[0], P2[4 * mo][h], P2[4 * mo + 1][h]);
                    mma16816(acc, AD[1], P2[4 * mo + 2][h], P2[4 * mo + 3][h]);
                    mma16816(acc, AD[2], P2[4 * mo + 4][h], P2[4 * mo + 5][h]);
                    P3[2 * mo + 0][h] = pack2(acc[0], acc[1]);
                    P3[2 * mo + 1][h] = pack2(acc[2], acc[3]);
                }
            }
            // ---- S4: strip j is k-step s = j - 2i of output row block i
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int s = j - 2 * i;
                if (s >= 0 && s < 3) {
#pragma unroll
                    for (int no = 0; no < 4; ++no) {
                        if (s == 0) mma16816_z(OUT[i][no], AD[0], P3[no][0], P3[no][1]);
                        else mma16816(OUT[i][no], AD[s], P3[no][0], P3[no][1]);
                    }
                }
            }
        } else {
            const int lane = g * 4 + tig;
#pragma unroll 1
            for (int k = 0; k < nt; ++k) {
                const uint4* fr = sAD + (k * 6) * 32 + lane;
                const uint4 H0 = fr[0], H1 = fr[32], H2 = fr[64];
                uint32_t P3[4][2];
#pragma unroll
                for (int mo = 0; mo < 2; ++mo) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float acc[4];
                        mma16816_z(acc, H0, P2[4 * mo][h], P2[4 * mo + 1][h]);
                        mma16816(acc, H1, P2[4 * mo + 2][h], P2[4 * mo + 3][h]);
                        mma16816(acc, H2, P2[4 * mo + 4][h], P2[4 * mo + 5][h]);
                        P3[2 * mo + 0][h] = pack2(acc[0], acc[1]);
                        P3[2 * mo + 1][h] = pack2(acc[2], acc[3]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int s = j - 2 * i;
                    if (s >= 0 && s < 3) {
                        const uint4 V = fr[(3 + s) * 32];
#pragma unroll
                        for (int no = 0; no < 4; ++no) mma16816(OUT[i][no], V, P3[no][0], P3[no][1]);
                    }
                }
            }
        }
    }
}

// Per-lane setup shared by the kernels: constant fragments and activation constants.
template <int UP>
struct LaneConsts {
    uint4 AU[MC<UP>::NVAR];
    uint4 AD[3];
    uint32_t sl2, cl2;
};
template <int UP>
__device__ __forceinline__ void lane_setup(const uint4* frags, float slope, float clamp_pre, int lane, LaneConsts<UP>& L) {
    using K = MC<UP>;
#pragma unroll
    for (int v = 0; v < K::NVAR; ++v) L.AU[v] = frags[v * 32 + lane];
#pragma unroll
    for (int s = 0; s < 3; ++s) L.AD[s] = frags[(K::NVAR + s) * 32 + lane];
    L.sl2 = pack2(slope, slope);
    const float c = fminf(clamp_pre, 65504.0f);
    L.cl2 = pack2(c, c);
}
template <int UP>
__device__ __forceinline__ void lane_setup(const MmaParams& p, int lane, LaneConsts<UP>& L) {
    lane_setup<UP>(p.frags, p.slope, p.clamp_pre, lane, L);
}

// Fetch the [IYT rows x 56 halfs] input tile of one warp with cp.async (zero fill outside the image = padding).
// Used by the planar-output kernel (double-buffered); the channels-last kernels fetch by TMA (tma_half_tile).
// Lane -> (row mod 4, 16-byte chunk): the chunk column and its byte count are per-tile constants of the lane and
// the row loop is fully unrolled with immediate offsets (this loader was 15 % of the kernel's instructions when
// it derived row and chunk from a running index, ncu r1).
template <int UP>
__device__ __forceinline__ void load_tile_async(const MmaParams& p, const __half* xp, __half* X, int ix0, int iy0, int lane) {
    using K = MC<UP>;
    const int rsub = lane >> 3, ch = lane & 7;
    const int ixc = (ix0 & ~7) + ch * 8;
    const bool col_ok = ch < 7 && ixc >= 0 && ixc < p.Win;
    const int col_bytes = col_ok ? min(8, p.Win - ixc) * 2 : 0;
    const __half* src = xp + static_cast<long long>(iy0 + rsub) * p.Wp_in + ixc;
    __half* dst = X + rsub * kXP + ch * 8;
    if (ch < 7) {
#pragma unroll
        for (int it = 0; it < K::IYT / 4; ++it) {
            const bool ok = static_cast<unsigned>(iy0 + rsub + 4 * it) < static_cast<unsigned>(p.Hin);
            cp_async16_zfill(dst + it * 4 * kXP, (ok && col_ok) ? src + static_cast<long long>(it) * 4 * p.Wp_in : xp, ok ? col_bytes : 0);
        }
    }
    cp_async_commit();
}

// One warp = a run of `tpw` consecutive 32x32 output tiles of one (b, c) plane; the input tile of
// tile t+1 is fetched with cp.async (zero-filled outside the image) while tile t is in the MMA chain.
template <int UP, bool RAD>
__global__ void __launch_bounds__(kWarps * 32, 2) flrelu_mma_kernel(const MmaParams p) {
    using K = MC<UP>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    __half* Xbuf = reinterpret_cast<__half*>(smem_raw + warp * 2 * K::XBYTES);
    const uint4* sAD = reinterpret_cast<const uint4*>(smem_raw + K::SMEM);
    if (RAD) {  // radial down filter: term fragments -> shared memory (before any warp leaves)
        uint4* dst = reinterpret_cast<uint4*>(smem_raw + K::SMEM);
        for (int i = threadIdx.x; i < p.nt * 6 * 32; i += blockDim.x) dst[i] = p.rfrags[i];
        __syncthreads();
    }

    const int ntiles = p.tiles_x * p.tiles_y;
    const int tile0 = (blockIdx.x * kWarps + warp) * p.tpw;
    if (tile0 >= ntiles) return;  // warps are independent: no block-level barrier below
    const int tile_end = min(tile0 + p.tpw, ntiles);
    const int c = blockIdx.y, b = blockIdx.z;
    const __half* xp = p.x + (static_cast<long long>(b) * p.C + c) * p.Hin * p.Wp_in;

    LaneConsts<UP> LC;
    lane_setup<UP>(p, lane, LC);
    const float oscale = (p.scale ? p.scale[b * p.C + c] : 1.0f) * p.out_gain;
    __half* yp = p.y + (static_cast<long long>(b) * p.C + c) * p.Hout * p.Wp_out;

    // first input sample each axis needs: n = ceil((2*o0 - pad)/UP), shifted down by e so it is even;
    // rows are loaded from the 16-byte aligned column ixa <= ix0.
    auto origin = [&](int tile, int& ox0, int& oy0, int& ix0, int& iy0) {
        const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        ox0 = tx * kOT;
        oy0 = ty * kOT;
        ix0 = first_in<UP>(ox0, p.px0, p.e);
        iy0 = first_in<UP>(oy0, p.py0, p.e);
    };
    auto load_tile = [&](int tile, __half* X) {
        int ox0, oy0, ix0, iy0;
        origin(tile, ox0, oy0, ix0, iy0);
        const int ixa = ix0 & ~7;
        if (p.bias == nullptr) {
            load_tile_async<UP>(p, xp, X, ix0, iy0, lane);
            return;
        } else {  // op-level API with a bias: synchronous path, bias added in fp32 and re-rounded
            const float bias = p.bias[c];
            for (int idx = lane; idx < K::IYT * 7; idx += 32) {
                const int row = idx / 7, ch = idx - row * 7;
                const int iy = iy0 + row, ixc = ixa + ch * 8;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (iy >= 0 && iy < p.Hin && ixc >= 0 && ixc < p.Win) {
                    v = *reinterpret_cast<const uint4*>(xp + static_cast<long long>(iy) * p.Wp_in + ixc);
                    __half* hv = reinterpret_cast<__half*>(&v);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        hv[i] = (ixc + i >= p.Win) ? __float2half_rn(0.0f) : __float2half_rn(__half2float(hv[i]) + bias);
                }
                *reinterpret_cast<uint4*>(X + row * kXP + ch * 8) = v;
            }
        }
        cp_async_commit();
    };

    load_tile(tile0, Xbuf);
    int buf = 0;
    for (int tile = tile0; tile < tile_end; ++tile, buf ^= 1) {
        const __half* X = Xbuf + buf * (K::XBYTES / 2);
        if (tile + 1 < tile_end) {
            load_tile(tile + 1, Xbuf + (buf ^ 1) * (K::XBYTES / 2));
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        int ox0, oy0, ix0, iy0;
        origin(tile, ox0, oy0, ix0, iy0);
        const int dx = ix0 & 7;  // even by construction

        float OUT[2][4][4];
        fir_chain<UP, RAD>(X, dx, LC.AU, LC.AD, LC.sl2, LC.cl2, g, tig, OUT, NoHook(), NoHook(), NoHook(), sAD, p.nt);

        // ---- store: * next-layer style, fp16, two adjacent columns per thread
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int no = 0; no < 4; ++no) {
                const int ox = ox0 + no * 8 + 2 * tig;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int oy = oy0 + i * 16 + hh * 8 + g;
                    if (oy < p.Hout && ox < p.Wout) {
                        const uint32_t v = pack2(OUT[i][no][hh * 2 + 0] * oscale, OUT[i][no][hh * 2 + 1] * oscale);
                        // Wp_out is even and ox is even: the pair never straddles the padded row
                        *reinterpret_cast<uint32_t*>(yp + static_cast<long long>(oy) * p.Wp_out + ox) = v;
                    }
                }
            }
        }
        __syncwarp();  // every lane is done reading X[buf] before the next iteration's prefetch overwrites it
    }
}


// ---- channels-last output variant -----------------------------------------------------------------
// The next conv reads channels-last (its TMA box start must be 16-byte aligned, conv_tc.cu), so for every
// layer that feeds a conv the transpose is fused here: one CTA = 16 warps = 16 consecutive channels of
// the same run of tiles.  Each warp stages its finished 32x32 tile in shared memory, then the CTA writes
// 32-byte (16 channel) pixel chunks -- full DRAM sectors -- into [B][H][W][Cp].
constexpr int kCG = 16;          // channels per CTA
constexpr int kTailBytes = 512;  // staging mbarriers (<= 6) at +0, per-warp input-tile mbarriers [16][2] at +64
constexpr int kSP = 40;          // staging row pitch in halfs (conflict-free fragment stores)
constexpr int kStageBytes = kOT * kSP * 2;

// One half (NIB/2 row blocks) of a warp's input tile by TMA: box (56 halfs, IYT/2 rows, 1 plane) of the planar conv
// output viewed as [B*C][Hin][Win] with row pitch Wp_in -- out-of-image rows / columns (the zero padding of the
// upsampler, and the pad columns [Win, Wp_in)) arrive as zeros.  One elected lane, one instruction, no address math.
template <int UP>
__device__ __forceinline__ void tma_half_tile(const CUtensorMap* tmap, __half* X, uint64_t* xbar, int half, int ix0, int iy0,
                                              int plane) {
    using K = MC<UP>;
    constexpr uint32_t kHalfBytes = (K::IYT / 2) * kXP * 2;
    mbar_arrive_expect_tx(&xbar[half], kHalfBytes);
    tma_load_3d(X + half * (K::IYT / 2) * kXP, tmap, &xbar[half], ix0 & ~7, iy0 + half * (K::IYT / 2), plane);
}

// Write-out share of one warp for a finished 32x32 tile held as 16 staged channel planes: image rows 2*warp and
// 2*warp+1, one pixel per lane and row (16 LDS.U16 -> one 32-byte chunk).  A variant reading pixel pairs with
// LDS.32 + PRMT issued fewer instructions but pushed the up=4 kernel into 70+ bytes of spills and measured slower.
__device__ __forceinline__ void write_out_tile(const MmaParams& p, const __half* st, int warp, int lane, int b, int c0,
                                               int ox0, int oy0) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int ly = 2 * warp + k, lx = lane;
        const int oy = oy0 + ly, ox = ox0 + lx;
        if (oy < p.Hout && ox < p.Wout) {
            const __half* sp = st + ly * kSP + lx;
            uint32_t w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint16_t lo = *reinterpret_cast<const uint16_t*>(sp + (2 * q) * (kStageBytes / 2));
                const uint16_t hi = *reinterpret_cast<const uint16_t*>(sp + (2 * q + 1) * (kStageBytes / 2));
                w[q] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
            }
            uint4* dst = reinterpret_cast<uint4*>(p.y + ((static_cast<long long>(b) * p.Hout + oy) * p.Wout + ox) * p.Cp_out + c0);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
}

template <int UP, bool RAD>
__global__ void __launch_bounds__(kCG * 32, 1) flrelu_mma_nhwc_kernel(const __grid_constant__ CUtensorMap tmap_x, const MmaParams p) {
    using K = MC<UP>;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~uintptr_t(127));  // TMA destinations
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    // layout: [16 warps] input tile (single buffer) | [2 buffers][16 planes] staging | mbarriers | radial fragments
    __half* X = reinterpret_cast<__half*>(smem_raw + warp * K::XBYTES);
    __half* stage_base = reinterpret_cast<__half*>(smem_raw + kCG * K::XBYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kCG * K::XBYTES + 2 * kCG * kStageBytes);
    uint64_t* full = bars;       // [2] all 16 planes of a staging buffer written
    uint64_t* empty = bars + 2;  // [2] all 16 write-out shares of a staging buffer done
    uint64_t* xbar = bars + 8 + 2 * warp;   // [2] halves of this warp's input tile landed
    const uint4* sAD = reinterpret_cast<const uint4*>(smem_raw + kCG * K::XBYTES + 2 * kCG * kStageBytes + kTailBytes);
    if (RAD) {  // radial down filter: term fragments -> shared memory (visible after the __syncthreads below)
        uint4* dst = reinterpret_cast<uint4*>(smem_raw + kCG * K::XBYTES + 2 * kCG * kStageBytes + kTailBytes);
        for (int i = threadIdx.x; i < p.nt * 6 * 32; i += blockDim.x) dst[i] = p.rfrags[i];
    }

    const int ntiles = p.tiles_x * p.tiles_y;
    const int tile0 = blockIdx.x * p.tpw;
    const int n = min(p.tpw, ntiles - tile0);
    const int c0 = blockIdx.y * kCG, b = blockIdx.z;
    const int c = c0 + warp;
    const bool valid = c < p.C;
    const int plane = b * p.C + (valid ? c : 0);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        for (int i = 0; i < 4; ++i) mbar_init(&bars[i], kCG);
        for (int i = 0; i < 2 * kCG; ++i) mbar_init(&bars[8 + i], 1);
        fence_barrier_init();
    }
    LaneConsts<UP> LC;
    lane_setup<UP>(p, lane, LC);
    const float oscale = ((valid && p.scale) ? p.scale[b * p.C + c] : 1.0f) * p.out_gain;

    if (!valid) {  // channel padding of the last group: both staging planes of this warp stay zero
        for (int sb = 0; sb < 2; ++sb) {
            uint4* z = reinterpret_cast<uint4*>(stage_base + (sb * kCG + warp) * (kStageBytes / 2));
            for (int i = lane; i < kStageBytes / 16; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    __syncthreads();  // barriers initialised (the only block-wide barrier: warps run decoupled from here on)

    // tile coordinates advance incrementally along the run (one division per CTA, none per tile); tx | ty << 16 in
    // one register each for tile i and tile i-1 (this kernel sits at the 128-register limit)
    uint32_t pos = static_cast<uint32_t>(tile0 % p.tiles_x) | (static_cast<uint32_t>(tile0 / p.tiles_x) << 16);
    uint32_t ppos = pos;
    auto load_half = [&](uint32_t q, int half) {   // one lane issues
        tma_half_tile<UP>(&tmap_x, X, xbar, half, first_in<UP>((q & 0xffff) * kOT, p.px0, p.e),
                          first_in<UP>((q >> 16) * kOT, p.py0, p.e), plane);
    };
    auto write_out = [&](uint32_t q, int sb) {
        write_out_tile(p, stage_base + sb * kCG * (kStageBytes / 2), warp, lane, b, c0, (q & 0xffff) * kOT, (q >> 16) * kOT);
    };

    uint32_t xph = 0;  // parity of this warp's input-tile barriers: both complete once per tile
    if (valid && lane == 0) { load_half(pos, 0); load_half(pos, 1); }
    for (int i = 0; i < n; ++i) {
        const int sb = i & 1;
        const uint32_t npos = ((pos & 0xffff) + 1 == static_cast<uint32_t>(p.tiles_x)) ? (pos & 0xffff0000u) + 0x10000u : pos + 1;
        // the staging buffer is free once every warp has written out its share of tile i-2
        if (i >= 2) mbar_wait(&empty[sb], ((i >> 1) - 1) & 1);
        if (valid) {
            mbar_wait(&xbar[0], xph);
            const int dx = first_in<UP>((pos & 0xffff) * kOT, p.px0, p.e) & 7;
            float OUT[2][4][4];
            const bool more = i + 1 < n;
            auto wait1 = [&]() { mbar_wait(&xbar[1], xph); };
            auto dead0 = [&]() {
                __syncwarp();  // every lane is done reading the first half: refill it with the next tile's rows
                if (more && lane == 0) load_half(npos, 0);
            };
            auto dead1 = [&]() {
                __syncwarp();
                if (more && lane == 0) load_half(npos, 1);
            };
            fir_chain<UP, RAD>(X, dx, LC.AU, LC.AD, LC.sl2, LC.cl2, g, tig, OUT, wait1, dead0, dead1, sAD, p.nt);
            xph ^= 1;
            __half* stage = stage_base + (sb * kCG + warp) * (kStageBytes / 2);
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int no = 0; no < 4; ++no)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
                        *reinterpret_cast<uint32_t*>(stage + (ii * 16 + hh * 8 + g) * kSP + no * 8 + 2 * tig) =
                            pack2(OUT[ii][no][hh * 2 + 0] * oscale, OUT[ii][no][hh * 2 + 1] * oscale);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[sb]);
        // write out this warp's share of the PREVIOUS tile while the other warps are still computing tile i
        if (i >= 1) {
            mbar_wait(&full[sb ^ 1], ((i - 1) >> 1) & 1);
            write_out(ppos, sb ^ 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[sb ^ 1]);
        }
        ppos = pos;
        pos = npos;
    }
    const int last = n - 1;
    mbar_wait(&full[last & 1], (last >> 1) & 1);
    write_out(ppos, last & 1);
}


// ---- persistent channels-last variant --------------------------------------------------------------
// Same arithmetic and staging scheme as flrelu_mma_nhwc_kernel, restructured after the ncu source view of that
// kernel showed ~18 % of all warp samples parked on its two mbarrier waits (every warp had to wait for the slowest
// warp's previous tile before starting its next one) and one exposed first-tile load + last write-out per 8-tile CTA:
//   * one CTA per SM walks a contiguous range of the (frame, tile, channel group) index space -- channel group
//     fastest, so the cheap items of a partial last group are spread evenly over the CTAs -- and the input
//     prefetch chain and the write-out pipeline never drain until the layer is done;
//   * three staging buffers and a write-out that lags TWO tiles behind: a warp may run two tiles ahead of the
//     slowest warp of its CTA before it blocks;
//   * item coordinates advance incrementally (no per-tile integer divisions).
constexpr int kNSB = 3;

struct ItemPos {
    int grp, tx, ty, b;
};

template <int UP>
__global__ void __launch_bounds__(kCG * 32, 1) flrelu_mma_nhwc_p_kernel(const __grid_constant__ CUtensorMap tmap_x, const MmaParams p,
                                                                        int ngroups, int total) {
    using K = MC<UP>;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~uintptr_t(127));  // TMA destinations
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    __half* X = reinterpret_cast<__half*>(smem_raw + warp * K::XBYTES);
    __half* stage_base = reinterpret_cast<__half*>(smem_raw + kCG * K::XBYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kCG * K::XBYTES + kNSB * kCG * kStageBytes);
    uint64_t* full = bars;           // [kNSB] all 16 planes of a staging buffer written
    uint64_t* empty = bars + kNSB;   // [kNSB] all 16 write-out shares of a staging buffer done
    uint64_t* xbar = bars + 8 + 2 * warp;   // [2] halves of this warp's input tile landed

    const int per = (total + gridDim.x - 1) / gridDim.x;
    const int start = blockIdx.x * per;
    const int n = min(per, total - start);
    if (n <= 0) return;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        for (int i = 0; i < 2 * kNSB; ++i) mbar_init(&bars[i], kCG);
        for (int i = 0; i < 2 * kCG; ++i) mbar_init(&bars[8 + i], 1);
        fence_barrier_init();
    }
    LaneConsts<UP> LC;
    lane_setup<UP>(p, lane, LC);
    __syncthreads();  // barriers initialised (the only block-wide barrier: warps run decoupled from here on)

    auto advance = [&](ItemPos& q) {
        if (++q.grp == ngroups) {
            q.grp = 0;
            if (++q.tx == p.tiles_x) {
                q.tx = 0;
                if (++q.ty == p.tiles_y) { q.ty = 0; ++q.b; }
            }
        }
    };
    auto load_half = [&](const ItemPos& q, int half) {   // one lane issues; no-op for the padding channels of a last group
        const int c = q.grp * kCG + warp;
        if (c >= p.C) return;
        tma_half_tile<UP>(&tmap_x, X, xbar, half, first_in<UP>(q.tx * kOT, p.px0, p.e), first_in<UP>(q.ty * kOT, p.py0, p.e),
                          q.b * p.C + c);
    };
    auto write_out = [&](const ItemPos& q, int sb) {
        write_out_tile(p, stage_base + sb * kCG * (kStageBytes / 2), warp, lane, q.b, q.grp * kCG, q.tx * kOT, q.ty * kOT);
    };
    auto drain = [&](const ItemPos& q, int j) {  // write out tile j of this CTA's range (all 16 warps take part)
        const int sb = j % kNSB;
        mbar_wait(&full[sb], (j / kNSB) & 1);
        write_out(q, sb);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[sb]);
    };

    ItemPos cur;
    {
        int r = start;
        cur.grp = r % ngroups; r /= ngroups;
        cur.tx = r % p.tiles_x; r /= p.tiles_x;
        cur.ty = r % p.tiles_y; r /= p.tiles_y;
        cur.b = r;
    }
    // items i-1 and i-2, packed into one register each (grp | tx << 8 | ty << 16 | b << 24: launch() checks the ranges)
    auto pack_pos = [](const ItemPos& q) { return static_cast<uint32_t>(q.grp | (q.tx << 8) | (q.ty << 16) | (q.b << 24)); };
    auto unpack_pos = [](uint32_t v) {
        ItemPos q;
        q.grp = v & 255; q.tx = (v >> 8) & 255; q.ty = (v >> 16) & 255; q.b = v >> 24;
        return q;
    };
    uint32_t old1 = pack_pos(cur), old2 = old1;
    uint32_t xph = 0;  // parity of this warp's input-tile barriers: both complete once per tile this warp computes
    if (lane == 0) { load_half(cur, 0); load_half(cur, 1); }
    for (int i = 0; i < n; ++i) {
        const int sb = i % kNSB;
        if (i >= 2) drain(unpack_pos(old2), i - 2);
        // the staging buffer is free once every warp has written out its share of tile i-3
        if (i >= kNSB) mbar_wait(&empty[sb], ((i / kNSB) - 1) & 1);
        ItemPos nxt = cur;
        advance(nxt);
        __half* stage = stage_base + (sb * kCG + warp) * (kStageBytes / 2);
        const int c = cur.grp * kCG + warp;
        if (c < p.C) {
            const float oscale = (p.scale ? p.scale[cur.b * p.C + c] : 1.0f) * p.out_gain;
            mbar_wait(&xbar[0], xph);
            const int dx = first_in<UP>(cur.tx * kOT, p.px0, p.e) & 7;
            float OUT[2][4][4];
            const bool more = i + 1 < n;
            auto wait1 = [&]() { mbar_wait(&xbar[1], xph); };
            auto dead0 = [&]() {
                __syncwarp();  // every lane is done reading the first half: refill it with the next tile's rows
                if (more && lane == 0) load_half(nxt, 0);
            };
            auto dead1 = [&]() {
                __syncwarp();
                if (more && lane == 0) load_half(nxt, 1);
            };
            fir_chain<UP, false>(X, dx, LC.AU, LC.AD, LC.sl2, LC.cl2, g, tig, OUT, wait1, dead0, dead1);
            xph ^= 1;
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int no = 0; no < 4; ++no)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
                        *reinterpret_cast<uint32_t*>(stage + (ii * 16 + hh * 8 + g) * kSP + no * 8 + 2 * tig) =
                            pack2(OUT[ii][no][hh * 2 + 0] * oscale, OUT[ii][no][hh * 2 + 1] * oscale);
        } else {
            // channel padding of the last group: this warp's plane is zero
            uint4* z = reinterpret_cast<uint4*>(stage);
            for (int k = lane; k < kStageBytes / 16; k += 32) z[k] = make_uint4(0u, 0u, 0u, 0u);
            if (i + 1 < n && lane == 0) { load_half(nxt, 0); load_half(nxt, 1); }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[sb]);
        old2 = old1;
        old1 = pack_pos(cur);
        cur = nxt;
    }
    // after the loop: old1 = item n-1, old2 = item n-2
    if (n >= 2) drain(unpack_pos(old2), n - 2);
    drain(unpack_pos(old1), n - 1);
}

#include "flrelu_stream.cuh"

// ---- host: constant fragments -------------------------------------------------------------------
struct FragCache {
    int up = 0, rho = -1, e = -1, fd_2d = 0;
    float fu[32], fd[144];
    uint4* dev = nullptr;
    int nt = 0;             // radial: number of separable terms, fragments at dev + NFRAG * 32
    double rad_gain = 1.0;  // radial: exact DC gain / DC gain of the fp16 terms
};
FragCache g_cache[8];
int g_cache_n = 0;

inline uint16_t f2h_bits(float f) {
    __half h = __float2half_rn(f);
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}
inline float h2f_bits(uint16_t b) {
    __half h;
    memcpy(&h, &b, 2);
    return __half2float(h);
}

// A[a][c] (16x16) -> per-lane registers in mma.m16n8k16 A-fragment order (taps rounded to fp16)
void emit_fragment(const float (&A)[16][16], uint4* out) {
    for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, tig = lane & 3;
        const int rows[4] = {g, g + 8, g, g + 8};
        const int cols[4] = {2 * tig, 2 * tig, 2 * tig + 8, 2 * tig + 8};
        uint32_t r[4];
        for (int k = 0; k < 4; ++k)
            r[k] = static_cast<uint32_t>(f2h_bits(A[rows[k]][cols[k]])) |
                   (static_cast<uint32_t>(f2h_bits(A[rows[k]][cols[k] + 1])) << 16);
        out[lane] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

// fp16 taps of one polyphase branch with an (almost) exact DC gain: every tap may round down or up; of the 2^n
// combinations the one whose sum is closest to the exact sum wins (ties: least squared tap error).  Rounding all
// taps to nearest instead leaves a systematic gain error of up to ~3e-4 per branch and pass, which showed up as
// 1e-3 in the final pixels (scratch/emul_mma_fir.py).
inline void tune_branch(const float* taps, int n, float* out) {
    float lo[8], hi[8];
    double exact = 0.0;
    for (int i = 0; i < n; ++i) {
        exact += taps[i];
        const uint16_t nb = f2h_bits(taps[i]);
        const float nf = h2f_bits(nb);
        if (nf == taps[i]) {
            lo[i] = hi[i] = nf;
        } else {
            // neighbouring fp16 value on the other side of the exact tap (sign-magnitude bit pattern +-1)
            const bool mag_up = fabsf(nf) < fabsf(taps[i]);
            uint16_t ob = static_cast<uint16_t>(mag_up ? nb + 1 : nb - 1);
            if ((nb & 0x7fff) == 0) ob = static_cast<uint16_t>((taps[i] < 0 ? 0x8000 : 0) | 1);
            const float of = h2f_bits(ob);
            lo[i] = nf < of ? nf : of;
            hi[i] = nf < of ? of : nf;
        }
    }
    double best_err = 1e30, best_sq = 1e30;
    int best = 0;
    for (int m = 0; m < (1 << n); ++m) {
        double sum = 0.0, sq = 0.0;
        for (int i = 0; i < n; ++i) {
            const float v = (m >> i & 1) ? hi[i] : lo[i];
            sum += v;
            sq += (static_cast<double>(v) - taps[i]) * (static_cast<double>(v) - taps[i]);
        }
        const double err = fabs(sum - exact);
        if (err < best_err - 1e-12 || (err < best_err + 1e-12 && sq < best_sq)) {
            best_err = err; best_sq = sq; best = m;
        }
    }
    for (int i = 0; i < n; ++i) out[i] = (best >> i & 1) ? hi[i] : lo[i];
}

// up filter (with the up gain) as fp16-exact floats, branch by branch: fu16[UT - 1 - ph - UP*m] for phase ph
template <int UP>
void tuned_up_taps(const FlreluArgs& a, float (&fu16)[32]) {
    constexpr int UT = 6 * UP;
    for (int ph = 0; ph < UP; ++ph) {
        float t[6], o[6];
        for (int m = 0; m < 6; ++m) t[m] = static_cast<float>(UP) * a.fu[UT - 1 - ph - UP * m];
        tune_branch(t, 6, o);
        for (int m = 0; m < 6; ++m) fu16[UT - 1 - ph - UP * m] = o[m];
    }
}
// down filter: its two polyphase branches are tuned separately; what remains of the total DC error is folded into
// the fp32 output scale (returned as exact / rounded).
inline double tuned_down_taps(const FlreluArgs& a, float (&fd16)[12]) {
    for (int ph = 0; ph < 2; ++ph) {
        float t[6], o[6];
        for (int m = 0; m < 6; ++m) t[m] = a.fd[ph + 2 * m];
        tune_branch(t, 6, o);
        for (int m = 0; m < 6; ++m) fd16[ph + 2 * m] = o[m];
    }
    double exact = 0.0, rounded = 0.0;
    for (int k = 0; k < 12; ++k) { exact += a.fd[k]; rounded += fd16[k]; }
    return exact / rounded;
}

// ---- radial 12x12 down filters (StyleGAN3-R): symmetric eigen-decomposition into separable terms ----------------
constexpr int kMaxTerms = 4;
constexpr double kTermTol = 5e-5;   // terms below this fraction of the leading eigenvalue are dropped

// cyclic Jacobi for a symmetric 12x12 matrix: A = V diag(w) V^T
inline void jacobi12(double (&A)[12][12], double (&V)[12][12], double (&w)[12]) {
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < 12; ++i)
            for (int j = i + 1; j < 12; ++j) off += A[i][j] * A[i][j];
        if (off < 1e-30) break;
        for (int p = 0; p < 12; ++p)
            for (int q = p + 1; q < 12; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 12; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - sn * akq;
                    A[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 12; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - sn * aqk;
                    A[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 12; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 12; ++i) w[i] = A[i][i];
}

struct RadialTerms {
    int nt = 0;
    float h[kMaxTerms][12], g[kMaxTerms][12];  // F[ky][kx] ~= sum_k g_k[ky] h_k[kx]
    bool ok = false;
};

// F row-major [ky][kx]; ok = symmetric and reproduced by <= kMaxTerms terms to kTermTol
inline RadialTerms radial_terms(const float* F) {
    RadialTerms rt;
    double A[12][12], V[12][12], w[12];
    double amax = 0.0;
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) {
            A[i][j] = 0.5 * (static_cast<double>(F[i * 12 + j]) + F[j * 12 + i]);
            amax = fmax(amax, fabs(A[i][j]));
            if (fabs(static_cast<double>(F[i * 12 + j]) - F[j * 12 + i]) > 1e-6 * (fabs(static_cast<double>(F[i * 12 + j])) + 1e-12) + 1e-12) return rt;
        }
    jacobi12(A, V, w);
    int order[12];
    for (int i = 0; i < 12; ++i) order[i] = i;
    for (int i = 0; i < 12; ++i)
        for (int j = i + 1; j < 12; ++j)
            if (fabs(w[order[j]]) > fabs(w[order[i]])) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
    const double w0 = fabs(w[order[0]]);
    if (w0 <= 0.0) return rt;
    int nt = 0;
    while (nt < 12 && fabs(w[order[nt]]) > kTermTol * w0) ++nt;
    if (nt > kMaxTerms) return rt;
    rt.nt = nt;
    for (int k = 0; k < nt; ++k) {
        const double lam = w[order[k]], r = sqrt(fabs(lam));
        for (int i = 0; i < 12; ++i) {
            rt.g[k][i] = static_cast<float>(r * V[i][order[k]]);
            rt.h[k][i] = static_cast<float>((lam < 0 ? -r : r) * V[i][order[k]]);
        }
    }
    rt.ok = true;
    return rt;
}

template <int UP>
int build_frags(const FlreluArgs& a, int rho, int e, uint4** out, FragCache** cache_out = nullptr) {
    using K = MC<UP>;
    for (int i = 0; i < g_cache_n && i < 8; ++i) {
        FragCache& fc = g_cache[i];
        if (fc.up == UP && fc.rho == rho && fc.e == e && fc.fd_2d == a.fd_2d && memcmp(fc.fu, a.fu, sizeof(float) * 6 * UP) == 0 &&
            memcmp(fc.fd, a.fd, sizeof(float) * (a.fd_2d ? 144 : 12)) == 0) {
            *out = fc.dev;
            if (cache_out) *cache_out = &fc;
            return MB_OK;
        }
    }
    std::vector<uint4> host(K::NFRAG * 32);
    int nt = 0;
    double rad_gain = 1.0;
    constexpr int UT = 6 * UP;
    float fu16[32], fd16[12] = {};
    tuned_up_taps<UP>(a, fu16);
    if (!a.fd_2d) tuned_down_taps(a, fd16);
    for (int v = 0; v < K::NVAR; ++v) {
        float A[16][16] = {};
        const int off = (UP == 2) ? 0 : 4 * v;
        for (int r = 0; r < 16; ++r) {
            // ceil((r - rho) / UP) and the phase of row r
            const int num = r - rho;
            int n = num / UP;
            if (num % UP != 0 && num > 0) ++n;  // ceil for positive, trunc == ceil for negative
            const int ph = UP * n - num;
            for (int m = 0; m < 6; ++m) {
                const int col = off + e + n + m;
                if (col >= 0 && col < 16) A[r][col] = fu16[UT - 1 - ph - UP * m];
            }
        }
        emit_fragment(A, &host[v * 32]);
    }
    for (int s = 0; s < 3; ++s) {
        float A[16][16] = {};
        for (int r = 0; r < 16; ++r)
            for (int col = 0; col < 16; ++col) {
                const int k = 16 * s + col - 2 * r;
                if (k >= 0 && k < 12) A[r][col] = fd16[11 - k];
            }
        emit_fragment(A, &host[(K::NVAR + s) * 32]);
    }
    if (a.fd_2d) {
        // separable terms of the radial filter: per term three down-H fragments (taps h_k) and three down-V fragments
        // (taps g_k); the leading term's polyphase branches are DC-tuned like a separable filter, the small terms are
        // rounded to nearest, and what is left of the total DC error goes into the fp32 output scale.
        const RadialTerms rt = radial_terms(a.fd);
        MB_REQUIRE(rt.ok, "filtered_lrelu: 2-D down filter is not a low-rank symmetric (radial) filter");
        nt = rt.nt;
        host.resize((K::NFRAG + nt * 6) * 32);
        double exact = 0.0, rounded = 0.0;
        for (int i = 0; i < 144; ++i) exact += a.fd[i];
        for (int k = 0; k < nt; ++k) {
            float t16[2][12];
            for (int hv = 0; hv < 2; ++hv) {
                const float* taps = hv == 0 ? rt.h[k] : rt.g[k];
                for (int ph = 0; ph < 2; ++ph) {
                    float t[6], o[6];
                    for (int m = 0; m < 6; ++m) t[m] = taps[ph + 2 * m];
                    if (k == 0) tune_branch(t, 6, o);
                    else for (int m = 0; m < 6; ++m) o[m] = h2f_bits(f2h_bits(t[m]));
                    for (int m = 0; m < 6; ++m) t16[hv][ph + 2 * m] = o[m];
                }
                for (int s = 0; s < 3; ++s) {
                    float A[16][16] = {};
                    for (int r = 0; r < 16; ++r)
                        for (int col = 0; col < 16; ++col) {
                            const int kk = 16 * s + col - 2 * r;
                            if (kk >= 0 && kk < 12) A[r][col] = t16[hv][11 - kk];
                        }
                    emit_fragment(A, &host[(K::NFRAG + k * 6 + hv * 3 + s) * 32]);
                }
            }
            double sh = 0.0, sg = 0.0;
            for (int i = 0; i < 12; ++i) { sh += t16[0][i]; sg += t16[1][i]; }
            rounded += sh * sg;
        }
        rad_gain = exact / rounded;
    }
    FragCache& fc = g_cache[g_cache_n % 8];
    if (fc.dev) cudaFree(fc.dev);
    fc.dev = nullptr;
    MB_CUDA(cudaMalloc(&fc.dev, host.size() * sizeof(uint4)));
    MB_CUDA(cudaMemcpy(fc.dev, host.data(), host.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    fc.up = UP; fc.rho = rho; fc.e = e; fc.fd_2d = a.fd_2d; fc.nt = nt; fc.rad_gain = rad_gain;
    memcpy(fc.fu, a.fu, sizeof(float) * 6 * UP);
    memcpy(fc.fd, a.fd, sizeof(float) * (a.fd_2d ? 144 : 12));
    ++g_cache_n;
    *out = fc.dev;
    if (cache_out) *cache_out = &fc;
    return MB_OK;
}

template <int UP>
int launch(const FlreluArgs& a, cudaStream_t stream) {
    using K = MC<UP>;
    // layer constants: rho = pad mod UP (phase of the first intermediate sample of a tile),
    // e = parity of ceil(-pad/UP) (makes the first input column of every tile even)
    const int rho = ((a.px0 % UP) + UP) % UP;
    int q = (-a.px0) / UP;
    if ((-a.px0) % UP != 0 && (-a.px0) > 0) ++q;  // ceil(-pad/UP)
    const int e = ((q % 2) + 2) % 2;
    uint4* frags = nullptr;
    FragCache* fc = nullptr;
    int r = build_frags<UP>(a, rho, e, &frags, &fc);
    if (r != MB_OK) return r;
    const bool radial = a.fd_2d != 0;   // (routing the separable filter through the shared-memory fragment path to free
                                        //  12 registers measured 6 % slower: 8.19 vs 7.71 ms/step)
    MmaParams p;
    p.nt = radial ? fc->nt : 0;
    p.rfrags = radial ? frags + K::NFRAG * 32 : nullptr;
    const int rad_smem = p.nt * 6 * 32 * static_cast<int>(sizeof(uint4));
    p.x = a.x; p.bias = a.bias; p.scale = a.scale; p.y = a.y; p.frags = frags;
    p.C = a.C; p.Hin = a.Hin; p.Win = a.Win; p.Wp_in = a.Wp_in; p.Hout = a.Hout; p.Wout = a.Wout; p.Wp_out = a.Wp_out;
    p.Cp_out = 0;
    p.px0 = a.px0; p.py0 = a.py0;
    p.tiles_x = ceil_div(a.Wout, kOT); p.tiles_y = ceil_div(a.Hout, kOT);
    p.e = e;
    p.rho = rho;
    if (radial) {
        p.out_gain = static_cast<float>(static_cast<double>(a.gain) * fc->rad_gain);
    } else {
        float fd16[12];
        const double cd = tuned_down_taps(a, fd16);
        p.out_gain = static_cast<float>(static_cast<double>(a.gain) * cd * cd);
    }
    const int ntiles = p.tiles_x * p.tiles_y;
    p.tpw = ntiles < 64 ? 1 : (ntiles < 256 ? 2 : 4);
    p.slope = a.slope;
    p.clamp_pre = a.clamp >= 0.0f ? a.clamp / a.gain : 3.0e38f;
    // Streaming kernels (flrelu_stream.cuh): the product path.  MB_FLRELU_STREAM=0 selects the tile kernels below
    // (kept for A/B runs and as the path of the op-level API with a separate bias).
    const char* sev = getenv("MB_FLRELU_STREAM");   // read per call: the A/B tests flip it inside one process
    const int use_stream = sev ? atoi(sev) : 1;
    MB_REQUIRE(!a.in_row_interleaved || (use_stream && a.bias == nullptr), "filtered_lrelu: the row-interleaved input needs the streaming kernels");
    if (use_stream && a.bias == nullptr) {
        using S = SC<UP>;
        const int sms = a.num_sms > 0 ? a.num_sms : 148;
        StreamParams sp;
        sp.scale = a.scale; sp.frags = frags; sp.rfrags = p.rfrags; sp.nt = p.nt;
        sp.C = a.C; sp.Hout = a.Hout; sp.Wout = a.Wout; sp.Wp_out = a.Wp_out; sp.Cp_out = 0;
        sp.px0 = a.px0; sp.e = e; sp.tiles_x = p.tiles_x; sp.B = a.B;
        sp.slope = p.slope; sp.clamp_pre = p.clamp_pre; sp.out_gain = p.out_gain;
        {
            // |t| <= max|x| * L^2 with L the largest absolute tap sum of a polyphase branch of the (gain-carrying) up filter;
            // 1 % head-room covers the fp16 roundings between the two passes.  The lrelu output is never larger than |t|.
            constexpr int UT = 6 * UP;
            double L = 0.0;
            for (int ph = 0; ph < UP; ++ph) {
                double sabs = 0.0;
                for (int m = 0; m < 6; ++m) sabs += fabs(static_cast<double>(UP) * a.fu[UT - 1 - ph - UP * m]);
                L = sabs > L ? sabs : L;
            }
            const char* sev2 = getenv("MB_FLRELU_ASSUME_SAFE");   // per call: a test flips it inside one process
            const bool assume_safe = sev2 && atoi(sev2) != 0;
            static unsigned int* zero_word = nullptr;   // MB_FLRELU_ASSUME_SAFE=1 (tests): a recorded maximum of 0
            if (assume_safe && !a.in_absmax && !zero_word) {
                MB_CUDA(cudaMalloc(&zero_word, 4));
                MB_CUDA(cudaMemset(zero_word, 0, 4));
            }
            const char* gev = getenv("MB_FLRELU_GUARD");          // MB_FLRELU_GUARD=0: ignore the recorded maximum, always clamp
            const bool use_guard = !(gev && atoi(gev) == 0);
            sp.in_absmax = (a.in_absmax && use_guard) ? a.in_absmax : (assume_safe ? zero_word : nullptr);
            const double lim = fmin(static_cast<double>(p.clamp_pre), 60000.0);
            sp.safe_abs = static_cast<float>(lim / (L * L * 1.01));
        }
        sp.y = a.y;
        if (a.y_nhwc) {
            MB_REQUIRE(a.Cp_out % 16 == 0 && a.Cp_out == round_up(a.C, kCG), "filtered_lrelu: Cp_out must be C rounded up to whole 16-channel groups");
            sp.y = a.y_nhwc;
            sp.Cp_out = a.Cp_out;
        }
        const StreamPlan pl = plan_stream(a.B, a.C, a.Hout, p.tiles_x, sms);
        sp.nseg = pl.nseg; sp.R = pl.R; sp.nsc = pl.nsc; sp.Rp = pl.Rp; sp.gfull = pl.gfull; sp.vlast = pl.vlast;
        sp.n_full = pl.n_full; sp.n_total = pl.n_total;
        CUtensorMap tm_x;
        {
            PFN_encodeTiled enc = get_encode_tiled();
            if (!enc) {
                set_error("cuTensorMapEncodeTiled not available from the driver");
                return MB_ECUDA;
            }
            MB_REQUIRE(a.Wp_in % 8 == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0, "filtered_lrelu: input rows must be 16-byte aligned");
            const cuuint64_t row_b = static_cast<cuuint64_t>(a.Wp_in) * 2;
            cuuint64_t dims[4] = {static_cast<cuuint64_t>(a.Win), static_cast<cuuint64_t>(a.Hin), static_cast<cuuint64_t>(a.C), static_cast<cuuint64_t>(a.B)};
            cuuint64_t strides[3] = {a.in_row_interleaved ? row_b * a.C : row_b, a.in_row_interleaved ? row_b : row_b * a.Hin, row_b * a.Hin * a.C};
            cuuint32_t box[4] = {static_cast<cuuint32_t>(kXP), static_cast<cuuint32_t>(S::BOXROWS), 1, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            CUresult cr = enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(a.x), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled(filtered_lrelu input) failed: %d (W=%d H=%d planes=%d pitch=%d)", static_cast<int>(cr),
                          a.Win, a.Hin, a.B * a.C, a.Wp_in);
                return MB_ECUDA;
            }
        }
        const int smem = S::SMEM + (radial ? kMaxTerms * 6 * 32 * static_cast<int>(sizeof(uint4)) : 0);
        const int grid = sp.n_total < sms ? sp.n_total : sms;
        auto go = [&](void (*kern)(CUtensorMap, StreamParams)) -> int {
            static const void* configured[8] = {};   // the kernels whose dynamic shared memory limit is already raised
            int slot = 0;
            while (slot < 8 && configured[slot] && configured[slot] != reinterpret_cast<const void*>(kern)) ++slot;
            if (slot == 8 || !configured[slot]) {
                MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             S::SMEM + kMaxTerms * 6 * 32 * static_cast<int>(sizeof(uint4))));
                if (slot < 8) configured[slot] = reinterpret_cast<const void*>(kern);
            }
            kern<<<grid, kCG * 32, smem, stream>>>(tm_x, sp);
            MB_CUDA(cudaGetLastError());
            return MB_OK;
        };
        if (a.y_nhwc) return radial ? go(flrelu_stream_kernel<UP, true, false>) : go(flrelu_stream_kernel<UP, false, false>);
        return radial ? go(flrelu_stream_kernel<UP, true, true>) : go(flrelu_stream_kernel<UP, false, true>);
    }
    if (a.y_nhwc) {
        // fused transpose: 16 channels per CTA, 32-byte pixel chunks into [B][H][W][Cp_out]
        MB_REQUIRE(a.bias == nullptr, "filtered_lrelu: the channels-last variant takes the bias from the conv epilogue");
        MB_REQUIRE(a.Cp_out % 16 == 0 && a.Cp_out >= round_up(a.C, kCG), "filtered_lrelu: Cp_out must cover whole 16-channel groups");
        p.y = a.y_nhwc;
        p.Cp_out = a.Cp_out;
        p.tpw = ntiles < 16 ? 1 : (ntiles < 128 ? 2 : (ntiles < 600 ? 4 : 8));
        // 0: one CTA per (tile run, channel group, frame); 1: persistent CTAs; default: by layer shape (r1 A/B on B200,
        // scripts/layer_times.py: the persistent kernel wins up to ~300^2 and on up=2 layers without a partial channel
        // group, the grid version on the large up=4 layers and on large maps whose last channel group is mostly padding)
        static int forced = -2;
        if (forced == -2) {
            const char* e = getenv("MB_FLRELU_NHWC");
            forced = e ? atoi(e) : -1;
        }
        const bool packable = ceil_div(a.C, kCG) <= 255 && p.tiles_x <= 255 && p.tiles_y <= 255 && a.B <= 255;
        const bool prefer_p = a.Hout <= 300 || (UP == 2 && a.C % kCG == 0);
        const int variant = (!packable || radial) ? 0 : (forced >= 0 ? forced : (prefer_p ? 1 : 0));
        // the planar input as a 3-D tensor [B*C][Hin][Win] (row pitch Wp_in): half-tile boxes of (56 halfs, IYT/2 rows)
        CUtensorMap tm_x;
        {
            PFN_encodeTiled enc = get_encode_tiled();
            if (!enc) {
                set_error("cuTensorMapEncodeTiled not available from the driver");
                return MB_ECUDA;
            }
            MB_REQUIRE(a.Wp_in % 8 == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0, "filtered_lrelu: input rows must be 16-byte aligned");
            cuuint64_t dims[3] = {static_cast<cuuint64_t>(a.Win), static_cast<cuuint64_t>(a.Hin), static_cast<cuuint64_t>(a.B) * a.C};
            cuuint64_t strides[2] = {static_cast<cuuint64_t>(a.Wp_in) * 2, static_cast<cuuint64_t>(a.Wp_in) * 2 * a.Hin};
            cuuint32_t box[3] = {static_cast<cuuint32_t>(kXP), static_cast<cuuint32_t>(K::IYT / 2), 1};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult cr = enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(a.x), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled(filtered_lrelu input) failed: %d (W=%d H=%d planes=%d pitch=%d)", static_cast<int>(cr),
                          a.Win, a.Hin, a.B * a.C, a.Wp_in);
                return MB_ECUDA;
            }
        }
        if (variant == 1) {
            constexpr int smem_p = kCG * K::XBYTES + kNSB * kCG * kStageBytes + kTailBytes + 128;
            static bool attr_p = false;
            if (!attr_p) {
                MB_CUDA(cudaFuncSetAttribute(flrelu_mma_nhwc_p_kernel<UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p));
                attr_p = true;
            }
            const int ngroups = ceil_div(a.C, kCG);
            const int total = a.B * ngroups * ntiles;
            const int sms = a.num_sms > 0 ? a.num_sms : 148;
            const int grid = total < sms ? total : sms;
            flrelu_mma_nhwc_p_kernel<UP><<<grid, kCG * 32, smem_p, stream>>>(tm_x, p, ngroups, total);
            MB_CUDA(cudaGetLastError());
            return MB_OK;
        }
        constexpr int smem = kCG * K::XBYTES + 2 * kCG * kStageBytes + kTailBytes + 128;
        constexpr int smem_max = smem + kMaxTerms * 6 * 32 * static_cast<int>(sizeof(uint4));
        static bool attr_nhwc = false;
        if (!attr_nhwc) {
            MB_CUDA(cudaFuncSetAttribute(flrelu_mma_nhwc_kernel<UP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            MB_CUDA(cudaFuncSetAttribute(flrelu_mma_nhwc_kernel<UP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
            attr_nhwc = true;
        }
        dim3 grid(ceil_div(ntiles, p.tpw), ceil_div(a.C, kCG), a.B);
        if (radial) flrelu_mma_nhwc_kernel<UP, true><<<grid, kCG * 32, smem + rad_smem, stream>>>(tm_x, p);
        else flrelu_mma_nhwc_kernel<UP, false><<<grid, kCG * 32, smem, stream>>>(tm_x, p);
        MB_CUDA(cudaGetLastError());
        return MB_OK;
    }
    static bool attr_done = false;
    if (!attr_done) {
        MB_CUDA(cudaFuncSetAttribute(flrelu_mma_kernel<UP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        MB_CUDA(cudaFuncSetAttribute(flrelu_mma_kernel<UP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     K::SMEM + kMaxTerms * 6 * 32 * static_cast<int>(sizeof(uint4))));
        attr_done = true;
    }
    dim3 grid(ceil_div(ntiles, kWarps * p.tpw), a.C, a.B);
    if (radial) flrelu_mma_kernel<UP, true><<<grid, kWarps * 32, K::SMEM + rad_smem, stream>>>(p);
    else flrelu_mma_kernel<UP, false><<<grid, kWarps * 32, K::SMEM, stream>>>(p);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

}  // namespace

bool flrelu_mma_supported(const FlreluArgs& a) {
    if (a.down != 2 || a.down_taps != 12) return false;
    if (a.fd_2d && !radial_terms(a.fd).ok) return false;   // radial filters run as a few separable terms
    if (!((a.up == 2 && a.up_taps == 12) || (a.up == 4 && a.up_taps == 24))) return false;
    if (a.px0 != a.py0) return false;
    if (a.C > 65535 || a.B > 65535) return false;
    if (a.slope < 0.0f || a.slope > 1.0f || a.gain <= 0.0f) return false;  // lrelu written as max(t*g, t*g*slope)
    return true;
}

int flrelu_mma_launch(const FlreluArgs& a, cudaStream_t stream) {
    return a.up == 2 ? launch<2>(a, stream) : launch<4>(a, stream);
}

}  // namespace mb
