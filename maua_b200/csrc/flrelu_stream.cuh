// filtered_lrelu, streaming form of the tensor-core chain (included by flrelu_mma.cu inside its namespace).
//
// The tile kernels of flrelu_mma.cu carry one 32x32 output tile per warp through the four-pass chain; every tile
// recomputes a vertical halo (the 74-row intermediate of a 32-row tile is produced as five 16-row strips, the
// input as six 8-row blocks) and pays the tile's fixed costs (coordinates, barriers, write-out of 32 rows at once).
// Here a warp walks DOWN a 32-pixel-wide column of one channel instead: per 16 output rows it computes exactly two
// new intermediate strips and one (up=4) or two (up=2) new 8-row input blocks, so the vertical halo is paid once per
// column segment instead of once per 32 rows (164 -> 132 HMMA per 32x32 outputs at up=2, 154 -> 122 at up=4), the
// input arrives through a per-warp ring of 8-row TMA boxes that is refilled block by block (prefetch distance =
// the ring, across item boundaries), and finished 16x32 blocks leave through a ring of staging buffers.
//
// Round-1 ncu findings this file answers (profiles/r1_ncu_final2.md, source page of flrelu_mma_nhwc_kernel):
//   * every shared-memory access was a generic LD/ST (the 128-byte round-up of the dynamic shared base went through
//     uintptr_t and lost the address space): 9 % of the stall samples were `lg_throttle`, the write-out's generic
//     LD.U16 results sat on the long scoreboard.  All shared accesses below are explicit ld/st.shared on 32-bit
//     shared addresses.
//   * the channels-last write-out cost ~70 instructions per warp and tile (32 LD.U16 + PRMT + address math).  Here
//     it is ldmatrix.x4.trans over the 16 staged channel planes: a thread receives four adjacent channels of one
//     pixel in two registers, four lanes cover a pixel's 32-byte chunk: 2 LDSM + 4 STG.64 per warp and block.
//   * 12 % of the samples were mbarrier spins coupling the 16 warps of a CTA through two staging buffers.  Blocks
//     are half a tile, six staging buffers, write-out lags three blocks.
//   * a partial last channel group (81 = 5*16 + 1, 51 = 3*16 + 3 channels) idled 15 / 13 warps for a whole tile.
//     A partial group with v channels runs 16/v column segments concurrently, one per (segment, channel) warp.
//   * items were walked so that vertically adjacent tiles were far apart in time: halo rows were re-read from DRAM
//     (1.42x the algorithmic bytes at 16 frames per step).  Items are ordered column-fastest and dealt round-robin:
//     at any time the 148 CTAs work on adjacent columns of the same planes, and a column has no vertical halo.
//
// Arithmetic: the MMA sequence per output element is the one of fir_chain() (same fragments, same accumulation
// order), so results are bit-identical to the tile kernels.
#pragma once

constexpr int kRowBlk = 8;                         // input rows per S1 block (one n8 block of the S1 GEMM)
constexpr int kOB = 16;                            // output rows per finished block
constexpr int kPlaneBytes = kOB * kSP * 2 + 16;    // staged 16x32 block of one channel (+16: ldmatrix rows of 8 planes on distinct banks)
constexpr int kBufBytes = kCG * kPlaneBytes;       // 20736 = 162 * 128
constexpr int kSTail = 2048;                       // zero chunk | item table | block descriptors | mbarriers
constexpr int kItemSlots = 8;                      // items whose blocks may still be waiting for their write-out

template <int UP>
struct SC {
    // One TMA box feeds one output block: 16 input rows (two S1 blocks) at up=2, 8 rows at up=4.
    static constexpr int BOXROWS = (UP == 2) ? 16 : 8;
    static constexpr int BOXBYTES = BOXROWS * kXP * 2;
#ifndef MB_FL_NRING2
#define MB_FL_NRING2 3
#define MB_FL_NRING4 4
#endif
    static constexpr int NRING = (UP == 2) ? MB_FL_NRING2 : MB_FL_NRING4;  // boxes per warp: the fetch runs two / three output blocks ahead
#ifndef MB_FL_NSB
#define MB_FL_NSB 6
#define MB_FL_LAG 3
#endif
    static constexpr int NSB = MB_FL_NSB;            // staging buffers
    static constexpr int LAG = MB_FL_LAG;            // the write-out runs this many blocks behind the staging
    static constexpr int RING_BYTES = kCG * NRING * BOXBYTES;
    static constexpr int SMEM = RING_BYTES + NSB * kBufBytes + kSTail + 128;
};

struct StreamParams {
    const float* scale;
    __half* y;
    const uint4* frags;
    const uint4* rfrags;
    int C, Hout, Wout, Wp_out, Cp_out, px0, e, nt;
    int tiles_x, B;
    int gfull;   // full 16-channel groups
    int vlast;   // channels of the partial last group (0 = none)
    int nseg, R;  // full groups: segments per column, 16-row blocks per segment
    int nsc, Rp;  // partial group: concurrent segments per item (16 / vlast), blocks per segment
    int n_full, n_total;
    float slope, clamp_pre, out_gain;
    // Clamp guard: when the producer of the input recorded max |x| (bits of a non-negative float, written by the conv
    // epilogue) and max |x| <= safe_abs, no intermediate can reach the clamp (|t| <= max|x| * (L1 norm of the up filter)^2),
    // and the activation runs without its two clamp instructions.  nullptr = unknown: always clamp.
    const unsigned int* in_absmax;
    float safe_abs;
};

// ---- explicit shared-space accessors (32-bit shared addresses) ----
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t a, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a));
}
__device__ __forceinline__ void stg64(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
    asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void sbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Wait for a phase.  The first probe is the fast path; a waiting warp then parks inside try_wait (suspend-time hint) instead
// of spinning through the issue port its three neighbours on the scheduler need (the r2a capture showed ~12 spins x 6
// instructions per block on the staging-buffer barrier: 12 % of all issued instructions).  Bounded: a broken pipeline
// traps instead of hanging the GPU box.
__device__ __forceinline__ void sbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    uint32_t spins = 0;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return;
        if (++spins > (1u << 20)) __trap();
    }
}
// the input as a 4-D tensor (x, row, channel, frame): planar [B][C][H][Wp] and row-interleaved [B][H][C][Wp] conv outputs
// differ only in the strides of the tensor map
__device__ __forceinline__ void tma_box_4d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// One work item: a 32-pixel column (or `nsc` segments of it) of 16 channels (or of the v channels of the partial group).
struct SItem {
    int b, c0, tx, v, nsc, R, seg0, segrows;
};
__device__ __forceinline__ void decode_item(const StreamParams& p, int idx, SItem& it) {
    if (idx < p.n_full) {
        it.tx = idx % p.tiles_x;
        int r = idx / p.tiles_x;
        it.seg0 = r % p.nseg;
        r /= p.nseg;
        const int grp = r % p.gfull;
        it.b = r / p.gfull;
        it.c0 = grp * kCG;
        it.v = kCG;
        it.nsc = 1;
        it.segrows = p.R * kOB;
        const int left = (p.Hout - it.seg0 * it.segrows + kOB - 1) / kOB;  // blocks down to the last image row
        it.R = left < p.R ? left : p.R;
    } else {
        idx -= p.n_full;
        it.tx = idx % p.tiles_x;
        it.b = idx / p.tiles_x;
        it.seg0 = 0;
        it.c0 = p.gfull * kCG;
        it.v = p.vlast;
        it.nsc = p.nsc;
        it.R = p.Rp;
        it.segrows = p.Rp * kOB;
    }
}
// This warp's share of an item: channel and first output row, or false when it has none.
__device__ __forceinline__ bool warp_share(const StreamParams& p, const SItem& it, int warp, int& c, int& oy0) {
    const int q = (it.v == kCG) ? 0 : warp / it.v;
    const int ch = warp - q * it.v;
    c = it.c0 + ch;
    oy0 = (it.seg0 + q) * it.segrows;
    return q < it.nsc && oy0 < p.Hout;
}

template <int UP, bool RAD, bool PLANAR>
__global__ void __launch_bounds__(kCG * 32, 1) flrelu_stream_kernel(const __grid_constant__ CUtensorMap tmap_x, const StreamParams p) {
    using K = MC<UP>;
    using S = SC<UP>;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 127u) & ~127u;   // TMA destinations
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const uint32_t ring = sbase + warp * (S::NRING * S::BOXBYTES);
    const uint32_t stg = sbase + S::RING_BYTES;
    const uint32_t tail = stg + S::NSB * kBufBytes;
    const uint32_t zero_a = tail;                               // 16 zero bytes: the pad channels of a partial group
    const uint32_t itab_a = tail + 64;                          // [kItemSlots] x 32 bytes: where the blocks of an item go
    const uint32_t desc_a = itab_a + kItemSlots * 32;           // [NSB] x 8 bytes: (item slot, first row) of a staged block
    const uint32_t full_a = desc_a + S::NSB * 8;
    const uint32_t empty_a = full_a + S::NSB * 8;
    const uint32_t xbar_a = empty_a + S::NSB * 8 + warp * (S::NRING * 8);
    const uint32_t rad_a = tail + kSTail;                       // radial down filter: [nt][6][32] uint4
    static_assert(64 + kItemSlots * 32 + SC<UP>::NSB * 24 + kCG * SC<UP>::NRING * 8 <= kSTail, "tail too small");

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        for (int i = 0; i < S::NSB; ++i) {
            sbar_init(full_a + 8 * i, kCG);
            sbar_init(empty_a + 8 * i, kCG);
        }
        for (int i = 0; i < kCG * S::NRING; ++i) sbar_init(empty_a + S::NSB * 8 + 8 * i, 1);
        sts128(zero_a, make_uint4(0u, 0u, 0u, 0u));
        for (int i = 0; i < S::NSB; ++i)   // no block staged yet: the write-out cursor starts LAG buffers behind and skips these
            asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(desc_a + 8 * i), "r"(0xffffffffu), "r"(0u) : "memory");
        fence_barrier_init();
    }
    if (RAD) {
        for (int i = threadIdx.x; i < p.nt * 6 * 32; i += blockDim.x) sts128(rad_a + 16 * i, p.rfrags[i]);
    }
    LaneConsts<UP> LC;
    lane_setup<UP>(p.frags, p.slope, p.clamp_pre, lane, LC);
    __syncthreads();  // the only block-wide barrier: warps run decoupled from here on

    const int G = gridDim.x;
    // layer constant of the column geometry: every column starts at the same offset inside its 16-byte aligned box
    const uint32_t xlane = ring + static_cast<uint32_t>((g * kXP + (first_in<UP>(0, p.px0, p.e) & 7) + 2 * tig) * 2);

    // ---- load cursor (meaningful in lane 0 only): next (item, box) this warp fetches ----
    int l_item = static_cast<int>(blockIdx.x) - G, l_left = 0, l_chan = 0, l_frame = 0, l_ixa = 0, l_row = 0;
    auto refill = [&](int slot) {
        while (l_left == 0) {   // next item in which this warp has a share
            l_item += G;
            if (l_item >= p.n_total) { l_left = -1; return; }
            SItem it;
            decode_item(p, l_item, it);
            int c, oy0;
            if (!warp_share(p, it, warp, c, oy0)) continue;
            l_left = it.R + ((UP == 2) ? 1 : 2);   // up=2: 2R+2 row blocks in 16-row boxes, up=4: R+2 8-row boxes
            l_chan = c; l_frame = it.b;
            l_ixa = first_in<UP>(it.tx * kOT, p.px0, p.e) & ~7;
            l_row = first_in<UP>(oy0, p.px0, p.e);
        }
        sbar_expect_tx(xbar_a + 8 * slot, S::BOXBYTES);
        tma_box_4d(ring + slot * S::BOXBYTES, &tmap_x, xbar_a + 8 * slot, l_ixa, l_row, l_chan, l_frame);
        l_row += S::BOXROWS;
        --l_left;
    };
    if (lane == 0) {
        for (int s = 0; s < S::NRING; ++s) {
            if (l_left < 0) break;
            refill(s);
        }
    }
    int c_slot = 0;
    uint32_t c_par = 0;   // consume cursor of the ring

    // ---- staging cursors: buffer index and parity of its use count.  A fresh mbarrier passes a wait on parity 1, so neither
    // cursor needs a "first round" case: the staging cursor waits for `empty` of use u - 1 (parity s_par ^ 1, passes at u = 0),
    // the write-out cursor starts LAG buffers behind (use -1, parity 1: passes, finds an invalid descriptor, skips).
    int s_sb = 0, w_sb = S::NSB - S::LAG;
    uint32_t s_par = 0, w_par = 1;
    // write-out constants of this lane: its ldmatrix row (channel chl of pixel group lane >> 4) and its 8 bytes of a pixel chunk
    const int chl = 4 * ((lane & 7) >> 1) + (lane & 1) + 2 * ((lane >> 3) & 1);
    const uint32_t pix = static_cast<uint32_t>(p.Cp_out) * 2u;

    // item state of the compute side
    int Rit = 0, blk = 0, islot = 0, oyb0 = 0;
    bool valid = false;
    float oscale = 0.0f;
    __half* yp = nullptr;   // PLANAR: this warp's plane at (first row of the segment, first column of the item)
    int rows_left = 0;      // PLANAR: image rows from the segment's first row down

    // ldmatrix row of a lane: matrix (lane >> 3) = pixel group (lane >> 4), channel set (lane >> 3) & 1; its 8 rows are the
    // channels {0,1,4,5,8,9,12,13} + 2 * set, so that a thread ends up with four ADJACENT channels of one pixel.
    auto writeout = [&]() {
        sbar_wait(full_a + 8 * w_sb, w_par);
        uint32_t dslot, doyb;
        asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(dslot), "=r"(doyb) : "r"(desc_a + 8 * w_sb));
        if (dslot != 0xffffffffu) {
            const uint4 d0 = lds128(itab_a + 32 * dslot);
            uint8_t* base = reinterpret_cast<uint8_t*>((static_cast<unsigned long long>(d0.y) << 32) | d0.x);
            const int dv = d0.z & 0xffff, dnsc = d0.z >> 16, dox0 = d0.w & 0xffff, dsegrows = d0.w >> 16;
            const uint32_t buf = stg + w_sb * kBufBytes;
            // byte offsets inside the frame fit 32 bits (a 1044^2 x 96-channel map is 209 MB): one 64-bit add per store
            const uint32_t lane_off = 8u * (lane & 3) + static_cast<uint32_t>(lane >> 2) * pix;
            const int ox = dox0 + (lane >> 2);
            auto row_out = [&](int oy, uint32_t a, uint32_t hstep) {
                const uint32_t off = static_cast<uint32_t>(oy) * static_cast<uint32_t>(p.Wout) * pix + lane_off;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t r4[4];
                    ldsm_x4_trans(a + h * hstep, r4);
                    if (ox + h * 16 < p.Wout) stg64(base + (off + (h * 16) * pix), r4[0], r4[1]);
                    if (ox + h * 16 + 8 < p.Wout) stg64(base + (off + (h * 16 + 8) * pix), r4[2], r4[3]);
                }
            };
            if (dv == kCG) {
                // a full group: warp w writes image row w of the block, plane = channel
                const int oy = static_cast<int>(doyb) + warp;
                if (oy < p.Hout) row_out(oy, buf + chl * kPlaneBytes + warp * (kSP * 2) + (lane >> 4) * 16, 32u);
            } else {
                const bool real = chl < dv;
                for (int t = 0; t < dnsc; ++t) {
                    const int r = warp * dnsc + t;
                    const int q = r >> 4, yr = r & 15;
                    const int oy = static_cast<int>(doyb) + q * dsegrows + yr;
                    if (oy < p.Hout)   // warp-uniform
                        row_out(oy, real ? buf + (q * dv + chl) * kPlaneBytes + yr * (kSP * 2) + (lane >> 4) * 16 : zero_a, real ? 32u : 0u);
                }
            }
            __syncwarp();
            if (lane == 0) sbar_arrive(empty_a + 8 * w_sb);
        }
        if (++w_sb == S::NSB) { w_sb = 0; w_par ^= 1; }
    };

    // A finished 16x32 block of this warp's channel (fp32 accumulators O) leaves the register file.
    auto finalize = [&](float (&O)[4][4]) {
        if constexpr (PLANAR) {
            const uint32_t pl = stg + warp * kPlaneBytes;
#pragma unroll
            for (int no = 0; no < 4; ++no)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
                    sts32(pl + ((hh * 8 + g) * kSP + no * 8 + 2 * tig) * 2, pack2(O[no][hh * 2 + 0] * oscale, O[no][hh * 2 + 1] * oscale));
            __syncwarp();
            const int row = blk * kOB + (lane >> 1);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int ch = (lane & 1) * 2 + k;
                const uint4 v = lds128(pl + (lane >> 1) * (kSP * 2) + ch * 16);
                if (row < rows_left && ch * 8 < oyb0) stg128(yp + static_cast<long long>(row) * p.Wp_out + ch * 8, v);   // oyb0: columns left
            }
            __syncwarp();
            ++blk;
        } else {
            sbar_wait(empty_a + 8 * s_sb, s_par ^ 1);
            if (valid) {
                const uint32_t pl = stg + s_sb * kBufBytes + warp * kPlaneBytes;
#pragma unroll
                for (int no = 0; no < 4; ++no)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
                        sts32(pl + ((hh * 8 + g) * kSP + no * 8 + 2 * tig) * 2, pack2(O[no][hh * 2 + 0] * oscale, O[no][hh * 2 + 1] * oscale));
            }
            if (threadIdx.x == 0)
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(desc_a + 8 * s_sb), "r"(islot), "r"(oyb0 + blk * kOB) : "memory");
            __syncwarp();
            if (lane == 0) sbar_arrive(full_a + 8 * s_sb);
            if (++s_sb == S::NSB) { s_sb = 0; s_par ^= 1; }
            ++blk;
            writeout();
        }
    };

    // ---- the chain ----
    constexpr int NSET = RAD ? 2 : 1;   // separable filter: one block of accumulators (k-step 2 of block i-1 is finished and
                                        // staged before k-step 0 of block i overwrites it); radial: both accumulate over the terms
    uint32_t x_box = xlane;   // this lane's S1 read address inside the ring box at the consume cursor
    uint32_t P1[2][kMB][2];   // packed A1^T of the two input-row blocks under the current strip
    float OUT[NSET][4][4];

    // S1 of one 8-row input block at byte offset `off` inside the current ring box
    auto s1 = [&](uint32_t (&P)[kMB][2], int off) {
        const uint32_t a = x_box + off;
#pragma unroll
        for (int m = 0; m < kMB; ++m) {
            const uint32_t b0 = lds32(a + K::wblk(m) * 16), b1 = lds32(a + K::wblk(m) * 16 + 16);
            mma16816_h(P[m], LC.AU[K::var(m)], b0, b1);
        }
    };
    auto box_wait = [&]() { sbar_wait(xbar_a + 8 * c_slot, c_par); };
    auto box_next = [&]() {
        x_box += S::BOXBYTES;
        if (++c_slot == S::NRING) { c_slot = 0; c_par ^= 1; x_box = xlane; }
    };
    auto box_release = [&]() {
        __syncwarp();   // every lane has its operands: the slot can take the next box of the sequence
        if (lane == 0 && l_left >= 0) refill(c_slot);
        box_next();
    };
    // S2 (+activation) of one strip
    auto s2 = [&](auto FASTc, const uint4& A, uint32_t (&Pa)[kMB][2], uint32_t (&Pb)[kMB][2], uint32_t (&P2)[kJB][2]) {
#pragma unroll
        for (int n8 = 0; n8 < kJB; ++n8) {
            uint32_t t2[2];
            mma16816_h(t2, A, Pa[n8 >> 1][n8 & 1], Pb[n8 >> 1][n8 & 1]);
            if constexpr (decltype(FASTc)::value && !RAD) {
                P2[n8][0] = lrelu2_abs(t2[0], LC.sl2);   // sl2 holds k = (1 + slope) / (1 - slope) on this path
                P2[n8][1] = lrelu2_abs(t2[1], LC.sl2);
            } else if constexpr (decltype(FASTc)::value) {
                P2[n8][0] = lrelu2(t2[0], LC.sl2);       // radial layers (StyleGAN3-R sits closer to the pixel tolerance): exact positive branch
                P2[n8][1] = lrelu2(t2[1], LC.sl2);
            } else {
                P2[n8][0] = lrelu_clamp2(t2[0], LC.sl2, LC.cl2);
                P2[n8][1] = lrelu_clamp2(t2[1], LC.sl2, LC.cl2);
            }
        }
    };
    // S3: P3 = packed O3^T of the strip, the B operands of S4
    auto s3 = [&](const uint4& H0, const uint4& H1, const uint4& H2, uint32_t (&P2)[kJB][2], uint32_t (&P3)[4][2]) {
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
    };
    // radial: S3 + S4 of one strip over the separable terms (fragments from shared memory).  kstep_lo >= 0: the strip is
    // k-step `kstep_lo` (0 or 1) of block Ocur; fin: it is also k-step 2 of Oprev.
    auto s34_rad = [&](uint32_t (&P2)[kJB][2], float (&Ocur)[4][4], float (&Oprev)[4][4], int kstep_lo, bool fin) {
        if (kstep_lo == 0) {
#pragma unroll
            for (int no = 0; no < 4; ++no)
#pragma unroll
                for (int q = 0; q < 4; ++q) Ocur[no][q] = 0.0f;
        }
#pragma unroll 1
        for (int k = 0; k < p.nt; ++k) {
            const uint32_t fr = rad_a + ((k * 6) * 32 + lane) * 16;
            const uint4 H0 = lds128(fr), H1 = lds128(fr + 512), H2 = lds128(fr + 1024);
            uint32_t P3[4][2];
            s3(H0, H1, H2, P2, P3);
            if (fin) {
                const uint4 V = lds128(fr + 5 * 512);
#pragma unroll
                for (int no = 0; no < 4; ++no) mma16816(Oprev[no], V, P3[no][0], P3[no][1]);
            }
            if (kstep_lo >= 0) {
                const uint4 V = lds128(fr + (3 + kstep_lo) * 512);
#pragma unroll
                for (int no = 0; no < 4; ++no) mma16816(Ocur[no], V, P3[no][0], P3[no][1]);
            }
        }
    };
    // Output block i of the running segment: strips 2i (which also finishes block i-1) and 2i+1.  PAR = i & 1 selects the
    // P1 slots (up=4) and the accumulator set (radial).  Returns true after the last strip of the segment.
    auto block_iter = [&](auto PARc, auto FASTc, int i) -> bool {
        constexpr int PAR = decltype(PARc)::value;
        [[maybe_unused]] constexpr int CUR = RAD ? PAR : 0, PRV = RAD ? (PAR ^ 1) : 0;
        uint32_t P2[kJB][2];
        // even strip 2i
        if constexpr (UP == 2) {
            s1(P1[1], kRowBlk * kXP * 2);                // input block 2i+1: second half of box i
            box_release();
            s2(FASTc, LC.AU[0], P1[0], P1[1], P2);
        } else {
            box_wait();
            s1(P1[PAR ^ 1], 0);                          // input block i+1
            box_release();
            s2(FASTc, LC.AU[0], P1[PAR], P1[PAR ^ 1], P2);
        }
        const bool last = i == Rit;
        if constexpr (!RAD) {
            uint32_t P3[4][2];
            s3(LC.AD[0], LC.AD[1], LC.AD[2], P2, P3);
            if (i > 0) {
#pragma unroll
                for (int no = 0; no < 4; ++no) mma16816(OUT[0][no], LC.AD[2], P3[no][0], P3[no][1]);
                finalize(OUT[0]);
            }
            if (last) return true;
#pragma unroll
            for (int no = 0; no < 4; ++no) mma16816_z(OUT[0][no], LC.AD[0], P3[no][0], P3[no][1]);
        } else {
            s34_rad(P2, OUT[CUR], OUT[PRV], last ? -1 : 0, i > 0);
            if (i > 0) finalize(OUT[PRV]);
            if (last) return true;
        }
        // odd strip 2i+1
        if constexpr (UP == 2) {
            box_wait();
            s1(P1[0], 0);                                // input block 2i+2: first half of box i+1
            s2(FASTc, LC.AU[0], P1[1], P1[0], P2);
        } else {
            s2(FASTc, LC.AU[1], P1[PAR], P1[PAR ^ 1], P2);
        }
        if constexpr (!RAD) {
            uint32_t P3[4][2];
            s3(LC.AD[0], LC.AD[1], LC.AD[2], P2, P3);
#pragma unroll
            for (int no = 0; no < 4; ++no) mma16816(OUT[0][no], LC.AD[1], P3[no][0], P3[no][1]);
        } else {
            s34_rad(P2, OUT[CUR], OUT[PRV], 1, false);
        }
        return false;
    };

    const bool no_clamp = p.in_absmax != nullptr && __uint_as_float(*p.in_absmax) <= p.safe_abs;   // uniform over the grid
    float act_gain = 1.0f;
    if (no_clamp && !RAD) {
        // clamp-free activation as one HFMA2 (lrelu2_abs): its constant replaces the slope, its factor joins the output scale
        const __half k16 = __float2half_rn((1.0f + p.slope) / (1.0f - p.slope));
        LC.sl2 = pack2(__half2float(k16), __half2float(k16));
        act_gain = 1.0f / (__half2float(k16) + 1.0f);
    }
    int nitem = 0;
    for (int item = blockIdx.x; item < p.n_total; item += G, ++nitem) {
        int c, oy0;
        {
            SItem it;
            decode_item(p, item, it);
            valid = warp_share(p, it, warp, c, oy0);
            Rit = it.R;
            blk = 0;
            oscale = (valid && p.scale ? p.scale[it.b * p.C + c] : 1.0f) * p.out_gain * act_gain;
            if constexpr (PLANAR) {
                yp = p.y + ((static_cast<long long>(it.b) * p.C + c) * p.Hout + oy0) * p.Wp_out + it.tx * kOT;
                rows_left = p.Hout - oy0;
                oyb0 = p.Wout - it.tx * kOT;   // columns left (a 16-byte chunk is stored when its first column is inside)
            } else {
                islot = nitem & (kItemSlots - 1);
                oyb0 = it.seg0 * it.segrows;
                if (threadIdx.x == 0) {
                    // where this item's blocks go: channels-last address of (frame, row 0, first column, first channel)
                    const unsigned long long base = reinterpret_cast<unsigned long long>(p.y) +
                        ((static_cast<long long>(it.b) * p.Hout * p.Wout + it.tx * kOT) * p.Cp_out + it.c0) * 2;
                    sts128(itab_a + 32 * islot, make_uint4(static_cast<uint32_t>(base), static_cast<uint32_t>(base >> 32),
                                                           it.v | (it.nsc << 16), (it.tx * kOT) | (it.segrows << 16)));
                }
            }
        }
        if (valid) {
            box_wait();
            s1(P1[0], 0);   // input block 0
            if constexpr (UP == 4) box_release();
            auto column = [&](auto FASTc) {
                if constexpr (UP == 2 && !RAD) {
                    for (int i = 0;; ++i)
                        if (block_iter(std::integral_constant<int, 0>(), FASTc, i)) break;
                } else {
                    for (int i = 0;; i += 2) {
                        if (block_iter(std::integral_constant<int, 0>(), FASTc, i)) break;
                        if (block_iter(std::integral_constant<int, 1>(), FASTc, i + 1)) break;
                    }
                }
            };
            if (no_clamp) column(std::true_type());
            else column(std::false_type());
        } else if (!PLANAR) {
            // no share in this item: keep the staging protocol in step (and take part in the write-out)
            float dummy[4][4];
            for (int i = 0; i < Rit; ++i) finalize(dummy);
        }
    }
    if constexpr (!PLANAR) {
        for (int k = 0; k < S::LAG; ++k) writeout();
    }
}

// Host side: split the columns into segments so that the round-robin deal over `sms` CTAs is even.
struct StreamPlan {
    int nseg, R, nsc, Rp, gfull, vlast, n_full, n_total;
};
inline StreamPlan plan_stream(int B, int C, int Hout, int tiles_x, int sms) {
    StreamPlan pl;
    pl.gfull = C / kCG;
    pl.vlast = C % kCG;
    const int TB = ceil_div(Hout, kOB);
    pl.nsc = pl.vlast ? kCG / pl.vlast : 0;
    pl.Rp = pl.vlast ? ceil_div(TB, pl.nsc) : 0;
    const int n_part = pl.vlast ? B * tiles_x : 0;
    long long best = -1;
    pl.nseg = 1;
    pl.R = TB;
    for (int nseg = 1; nseg <= TB && nseg <= 16; ++nseg) {
        const int R = ceil_div(TB, nseg);
        if ((nseg - 1) * R >= TB) continue;   // an empty last segment
        const long long n_full = static_cast<long long>(B) * pl.gfull * nseg * tiles_x;
        // strips on the busiest CTA: its share of full items, then its share of partial items
        const long long cost = ceil_div(static_cast<int>(n_full), sms) * (2LL * R + 1) + ceil_div(n_part, sms) * (2LL * pl.Rp + 1);
        if (best < 0 || cost < best) { best = cost; pl.nseg = nseg; pl.R = R; }
    }
    pl.n_full = B * pl.gfull * pl.nseg * tiles_x;
    pl.n_total = pl.n_full + n_part;
    return pl;
}
