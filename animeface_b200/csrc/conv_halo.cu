// tcgen05 implicit-GEMM convolution, "halo" variant: the activation patch of a tile is loaded and converted ONCE and
// the k*k taps are shifted windows of it, addressed through the UMMA shared-memory descriptors.
//
// Round 1's first kernels loaded one TMA box per (tap, channel block): every input element crossed L2->SMEM and the
// fp32->split conversion 9 times for a 3x3 kernel, which made the low-channel, high-resolution layers of the path
// (64->64 @256^2 etc.) L2- and conversion-bound.  Here, per CTA tile of 8 x 16 pixels and per block of input channels:
//   TMA      one box [32 ch, 8+2p, 16+2p, 1] (p = k/2; out-of-bounds -> 0 = zero padding) into a SWIZZLE_128B fp32 slot
//   xform    8 warps: (style scale) -> split planes, written as a NON-swizzled K-major operand image
//            plane[chunk c][patch row r][16 B]   (chunk = 16 B of K; LBO = rows*16 B between chunks, 16 B between rows)
//   MMA      tap (dy,dx), K chunk pair kq: A descriptor start = plane + 2kq*LBO + ((dy+p)*PW + dx+p)*16, SBO = PW*16
//            (an 8-pixel tile row is an 8-row core-matrix group, consecutive tile rows are PW patch rows apart), so
//            all taps read the same converted patch; B = pre-packed weight tile of (tap, channel block), SWIZZLE_128B.
//            per (tap, k16) TWO instructions: A0 x [B0 ; B1] (N = 2 BN, the two weight planes are adjacent rows) fills the
//            accumulator halves [A0*B0 | A0*B1]; A1 x B0 (N = BN) adds into the upper half -- the three split products with
//            A fetched from shared memory twice instead of three times (the operand fetch paces the MMAs of this kernel)
// PRECISE = false: bf16x3 (hi/lo planes, kind::f16, 64 channels per block); result = lower + upper half -> data gradients
// PRECISE = true : fp16x3 + promotion: big = rn_f16(v), small = rn_f16((v - big) * 2^11) (22 mantissa bits together, like
//                  a tf32 pair but at the fp16 MMA rate); lower half D1 = big*big, upper half D2 =
//                  big*small + small*big; every `promo_taps` taps (default 2 = 8 chained big*big MMAs -- the tensor core
//                  truncates its accumulator per MMA) the segment is promoted into fp32 registers (acc += D1 + 2^-11 D2)
//                  and its TMEM buffer handed back at once                              -> forward convs, fp32-class
//                  (fp16 range: operands saturate at +-65504; forward activations and weights of this path are O(1))
// Persistent CTAs (one per SM), double-buffered planes, up to 4 TMEM accumulator buffers, dedicated epilogue warps; edge
// tiles of ragged images are masked in the epilogue (TMA zero-fills what hangs over the image).
// Warp roles: 0 patch TMA, 1 MMA (+TMEM alloc), 2..9 transform, 10.. epilogue/promotion, last = weight TMA.
#include <stdlib.h>
#include "tc_common.cuh"
#ifndef SG2_TRACE_EPI
#define SG2_TRACE_EPI 0      // 1: the epilogue warps also time their phases (slots 13-15 of sg2_debug_trace); costs registers
#endif
#include "conv.h"

namespace sg2 {
namespace halo {
using namespace tc;

constexpr int TW = 8, TH = 16, BM = 128;
constexpr int SLOT_BYTES = 24 * 1024;      // 1024-aligned slot pitch
constexpr int MAX_ROWS = (TW + 2) * (TH + 2);            // 180 patch rows
constexpr int PLANE_PITCH = 23 * 1024;                   // 23552

template <int BN, bool PRECISE> struct Cfg {
    static constexpr int EPI_WARPS = PRECISE ? 8 : 4;
    static constexpr int NWARPS = 10 + EPI_WARPS + 1;
    static constexpr int NTHREADS = NWARPS * 32;
    static constexpr int BWARP = NWARPS - 1;                       // weight producer warp
    static constexpr int KB = 64;                                  // input channels per block (128 B of K per operand row)
    static constexpr int BOXES = 2;                                // TMA boxes (32 ch) per block
    static constexpr int SLOTS = BN == 128 ? 1 : 2;                // staging slots (boxes in flight); 1 keeps BN=128 under 227 KB
    static constexpr int BTILE = 2 * BN * 128;                     // two planes of [BN rows x 128 B]
    // weight stages: as deep as the 227 KB allow (the weight tiles stream from L2 once per pixel tile and tap)
    static constexpr int BSTAGES = BN == 128 ? 3 : (BN == 64 ? 5 : 6);
    static constexpr int SMEM = 1024 + SLOTS * SLOT_BYTES + 2 * 2 * PLANE_PITCH + BSTAGES * BTILE + 256;
    // One TMEM accumulator buffer = 2*BN fp32 columns: the first MMA of a k16 step multiplies the big / hi A plane by BOTH
    // weight planes at once (B rows = [plane0 ; plane1], N = 2*BN), so columns [0,BN) collect big*big (hi*hi) and columns
    // [BN,2BN) the cross term big*small (hi*lo); the second MMA adds small*big (lo*hi) into the upper half.  Two MMAs
    // instead of three per k16 step: A is fetched from shared memory twice instead of three times (the MMA rate of this
    // kernel is set by that operand fetch).  PRECISE: lower half = D1 (promoted segment by segment), upper half = D2.
    static constexpr int ACC_COLS = 2 * BN;
    static constexpr int NACC = 512 / ACC_COLS >= 4 ? 4 : 512 / ACC_COLS;       // buffers: promotion / epilogue latency hides behind NACC-1 segments
    static constexpr uint32_t TMEM_COLS = NACC * ACC_COLS;
    // BN = 128, fp32-class: promotion is bound by the TMEM read bandwidth (a [D1|D2] segment is 128 KB), not by the MMAs.  Only
    // D1 (big*big, whose chain length decides the error) needs promoting every few taps; the cross terms are 2^-11 of the result
    // and can accumulate over the whole tile.  Layout [D1_a | D2_x | D1_b | D2_y]: the D1 buffers alternate per segment, the D2
    // buffers per TILE (so reading D2 at the end of a tile overlaps the next tile's MMAs): TMEM reads drop from 128 KB to
    // 64 KB per segment + 64 KB per tile.  The wide MMA A0 x [B0 ; B1] needs D2 right behind D1 -- true for (D1_a, D2_x) and
    // (D1_b, D2_y); the other two pairings, and the first k16 step of every segment (D1 restarts, D2 continues), issue the three
    // products as three N = BN instructions.
    static constexpr bool D2X = PRECISE && BN == 128;
};

struct Params {
    const float* in_scale; const float* out_scale; const float* bias; const float* noise;
    float* y;
    long long ys[4];
    const unsigned char* wp;
    int n, h, w, ci, co, k;
    int tiles_x, tiles_y, m_tiles, n_tiles;
    int nkb;                 // channel blocks
    int act;
    float alpha, gain;
    long long* trace;        // sg2_debug_trace buffer or null
    int promo_taps;          // PRECISE: taps accumulated in TMEM between two promotions (4 chained big*big MMAs per tap)
    int bstages;             // weight stages in use (<= Cfg::BSTAGES)
    int tma_store;           // 1x1 convolutions: staged TMA-store epilogue (see conv_halo_pl.cu)
    int tx_shift, ty_shift;  // log2(tiles_x), log2(tiles_y) when both are powers of two, else -1
};

template <int M> struct Mode { static constexpr int value = M; };      // compile-time store mode of the epilogue

// non-swizzled K-major descriptor: LBO between 16-byte K chunks, SBO between 8-row groups, 16 B between rows
__device__ __forceinline__ uint64_t interleave_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}

template <int BN, bool PRECISE>
__global__ void __launch_bounds__(Cfg<BN, PRECISE>::NTHREADS, 1) conv_halo_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap,
                                                                                  const Params p) {
    using C = Cfg<BN, PRECISE>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t slot_base = base;
    const uint32_t plane_base = slot_base + C::SLOTS * SLOT_BYTES;            // [buf][plane]
    const uint32_t b_base = plane_base + 4 * PLANE_PITCH;
    const int BSTAGES = p.bstages;
    const uint32_t bar_base = b_base + C::BSTAGES * C::BTILE;
    auto st_full = [&](int s) { return bar_base + 8u * s; };                  // 2
    auto st_empty = [&](int s) { return bar_base + 16u + 8u * s; };           // 2
    auto pl_full = [&](int b) { return bar_base + 32u + 8u * b; };            // 2
    auto pl_empty = [&](int b) { return bar_base + 48u + 8u * b; };           // 2
    auto b_full = [&](int s) { return bar_base + 64u + 8u * s; };             // BSTAGES <= 6
    auto b_empty = [&](int s) { return bar_base + 112u + 8u * s; };           // BSTAGES <= 6
    auto acc_full = [&](int b) { return bar_base + 160u + 8u * b; };          // NACC <= 4
    auto acc_empty = [&](int b) { return bar_base + 192u + 8u * b; };         // NACC <= 4
    const uint32_t tmem_slot = bar_base + 224u;
    auto d2_empty = [&](int b) { return bar_base + 232u + 8u * b; };          // 2 (D2X only)
    auto plane = [&](int buf, int pl) { return plane_base + (uint32_t)((buf * 2 + pl) * PLANE_PITCH); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool tr = p.trace != nullptr;
    long long* trow = p.trace + (size_t)blockIdx.x * 32;
    const long long t_begin = tr ? clock64() : 0;
    long long w0 = 0, w1 = 0, w2 = 0;            // wait-cycle accumulators of this thread's role
    const int pad = p.k >> 1, taps = p.k * p.k;
    const int PW = TW + 2 * pad, PH = TH + 2 * pad, PR = PW * PH;
    const uint32_t LBO = (uint32_t)PR * 16u, SBO = (uint32_t)PW * 16u;
    const int total_tiles = p.m_tiles * p.n_tiles;
    auto tile_coords = [&](int tile, int& x0, int& y0, int& b0, int& n0) {
        int nt = 0, mt = tile;                             // n_tiles is 1..4: no division
        while (mt >= p.m_tiles) { mt -= p.m_tiles; ++nt; }
        if (p.tx_shift >= 0) {                             // power-of-two tile grid (every layer of the 256^2 path): shifts and masks
            x0 = (mt & (p.tiles_x - 1)) * TW;
            y0 = ((mt >> p.tx_shift) & (p.tiles_y - 1)) * TH;
            b0 = mt >> (p.tx_shift + p.ty_shift);
        } else {
            x0 = (mt % p.tiles_x) * TW;
            y0 = ((mt / p.tiles_x) % p.tiles_y) * TH;
            b0 = mt / (p.tiles_x * p.tiles_y);
        }
        n0 = nt * BN;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(st_full(s), 1); mbar_init(st_empty(s), 8);
            mbar_init(pl_full(s), 8); mbar_init(pl_empty(s), 1);
        }
        for (int s = 0; s < C::NACC; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), C::EPI_WARPS); }
        for (int s = 0; s < C::BSTAGES; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int s = 0; s < 2; ++s) mbar_init(d2_empty(s), C::EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_d;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot));

    if (warp == 0) {
        // ================= patch producer: one TMA box per (tile, channel block, 32-channel half) =================
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            int bx = 0;                                   // global box counter
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int x0, y0, b0, n0;
                tile_coords(tile, x0, y0, b0, n0);
                for (int kb = 0; kb < p.nkb; ++kb)
                    for (int hb = 0; hb < C::BOXES; ++hb) {
                        if (kb * C::KB + 32 * hb >= p.ci) continue;      // 32-channel tail: no box, no transform, no MMAs for it
                        const int s = (bx++) % C::SLOTS;
                        mbar_wait_t(st_empty(s), (((bx - 1) / C::SLOTS) & 1) ^ 1, tr, w0);
                        mbar_expect_tx(st_full(s), (uint32_t)PR * 128u);
                        tma_load_4d(slot_base + s * SLOT_BYTES, &xmap, st_full(s), kb * C::KB + 32 * hb, x0 - pad, y0 - pad, b0);
                    }
            }
            if (tr) trow[0] = w0;
        }
    } else if (warp == C::BWARP) {
        // ================= weight producer: one pre-packed tile per (tile, channel block, tap) =================
        if (elect_one()) {
            int s = 0;
            uint32_t sph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int nt = 0;
                for (int mt = tile; mt >= p.m_tiles; mt -= p.m_tiles) ++nt;
                const unsigned char* wsrc = p.wp + (size_t)nt * p.nkb * taps * C::BTILE;
                for (int i = 0; i < p.nkb * taps; ++i) {
                    mbar_wait_t(b_empty(s), sph ^ 1, tr, w0);
                    mbar_expect_tx(b_full(s), C::BTILE);
                    bulk_load(b_base + s * C::BTILE, wsrc + (size_t)i * C::BTILE, C::BTILE, b_full(s));
                    if (++s == BSTAGES) { s = 0; sph ^= 1; }
                }
            }
            if (tr) trow[10] = w0;
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // One thread, and it paces the low-channel layers: the loop below is kept free of divisions and 64-bit descriptor
        // arithmetic (descriptor low words advance by adds, stage / buffer indices by wrap-around counters).
        if (elect_one()) {
            constexpr uint32_t idesc = PRECISE ? idesc_f16(BM, BN) : idesc_bf16(BM, BN);              // N = BN
            constexpr uint32_t idesc2 = PRECISE ? idesc_f16(BM, 2 * BN) : idesc_bf16(BM, 2 * BN);     // N = 2 BN (both weight planes)
            const uint32_t a_hi = (SBO >> 4) | (1u << 14);                        // A: interleaved K-major, version bit 46
            const uint32_t a_lo_f = (LBO >> 4) << 16;
            const uint32_t kstep = (2u * LBO) >> 4;                               // one k16 step = two 16-byte K chunks
            constexpr uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);     // B: K-major SWIZZLE_128B, SBO 1024
            constexpr uint32_t b_lo_f = 1u << 16;
            const uint32_t nstages = (uint32_t)BSTAGES, promo = (uint32_t)p.promo_taps;
            const uint32_t kdim = (uint32_t)p.k, row_wrap = (uint32_t)(PW - p.k);
            const uint32_t nflat = (uint32_t)(p.nkb * taps);
            uint32_t s = 0, sph = 0, abuf = 0, aph = 0, kbg = 0, tcount = 0;      // weight stage / accumulator buffer (+ phases), block / tile counters
            long long t_issue = 0, t_commit = 0;          // trace: cycles issuing MMAs / commits
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
                uint32_t f = 0, fseg = 0;                 // flat (kb, tap) index inside the tile / inside the accumulator segment
                const uint32_t tp = tcount & 1;           // D2X: which D2 buffer this tile accumulates its cross terms in
                if (C::D2X) { mbar_wait_t(d2_empty(tp), ((tcount >> 1) & 1) ^ 1, tr, w2); tc_fence_after(); }
                int ci_left = p.ci;
                for (int kb = 0; kb < p.nkb; ++kb, ++kbg, ci_left -= C::KB) {
                    const uint32_t pbuf = kbg & 1;
                    mbar_wait_t(pl_full(pbuf), (kbg >> 1) & 1, tr, w0);
                    const uint32_t a0_base = ((plane(pbuf, 0) & 0x3FFFFu) >> 4) | a_lo_f, a1_base = a0_base + (PLANE_PITCH >> 4);
                    const int kqn = min(4, ci_left >> 4);           // k16 steps that hold real channels (2 for a 32-channel tail)
                    uint32_t arow = 0, dx = 0;                      // (dy * PW + dx): the tap's first patch row
                    for (int t = 0; t < taps; ++t, ++f) {
                        // accumulator segment: PRECISE -> every promo_taps taps, else the whole tile
                        const bool seg_start = PRECISE ? (fseg == 0) : (f == 0);
                        const bool seg_end = PRECISE ? (fseg == promo - 1 || f == nflat - 1) : (f == nflat - 1);
                        fseg = seg_end ? 0 : fseg + 1;
                        if (seg_start) { mbar_wait_t(acc_empty(abuf), aph ^ 1, tr, w2); tc_fence_after(); }
                        mbar_wait_t(b_full(s), sph, tr, w1);
                        const long long ti0 = tr ? clock64() : 0;
                        const uint32_t a0 = a0_base + arow, a1 = a1_base + arow;
                        const uint32_t b0_ = (((b_base + s * C::BTILE) & 0x3FFFFu) >> 4) | b_lo_f;   // plane 0 rows, then plane 1 rows: 2*BN contiguous B rows
                        if (C::D2X) {
                            const uint32_t d1 = tmem_d + (uint32_t)(abuf * 2 * BN), d2 = tmem_d + (uint32_t)(BN + tp * 2 * BN);
                            const bool adjacent = abuf == tp;     // D2 sits right behind this segment's D1
#pragma unroll
                            for (int kq = 0; kq < 4; ++kq) {
                                if (kq >= kqn) break;
                                const uint32_t da0 = a0 + kq * kstep, da1 = a1 + kq * kstep, db = b0_ + kq * 2, db1 = db + (BN * 128 >> 4);
                                const bool first_seg = seg_start && kq == 0, first_tile = f == 0 && kq == 0;
                                if (adjacent && !first_seg) {
                                    mma_f16_words(d1, da0, a_hi, db, b_hi, idesc2, 1);             // [D1 | D2] += A0 * [B0 ; B1]
                                } else {
                                    mma_f16_words(d1, da0, a_hi, db, b_hi, idesc, !first_seg);     // D1 (restarted at a segment start) += A0 * B0
                                    mma_f16_words(d2, da0, a_hi, db1, b_hi, idesc, !first_tile);   // D2 (restarted at a tile start)   += A0 * B1
                                }
                                mma_f16_words(d2, da1, a_hi, db, b_hi, idesc, 1);                  // D2 += A1 * B0
                            }
                        } else {
                            const uint32_t d = tmem_d + (uint32_t)(abuf * C::ACC_COLS);
#pragma unroll
                            for (int kq = 0; kq < 4; ++kq) {
                                if (kq >= kqn) break;
                                const uint32_t da0 = a0 + kq * kstep, da1 = a1 + kq * kstep, db = b0_ + kq * 2;
                                mma_f16_words(d, da0, a_hi, db, b_hi, idesc2, !(seg_start && kq == 0));   // [D1 | D2] += A0 * [B0 ; B1]
                                mma_f16_words(d + BN, da1, a_hi, db, b_hi, idesc, 1);                     //       D2  += A1 * B0
                            }
                        }
                        const long long ti1 = tr ? clock64() : 0;
                        mma_commit(b_empty(s));
                        if (++s == nstages) { s = 0; sph ^= 1; }
                        if (seg_end) {
                            mma_commit(acc_full(abuf));
                            if (++abuf == C::NACC) { abuf = 0; aph ^= 1; }
                        }
                        ++arow;
                        if (++dx == kdim) { dx = 0; arow += row_wrap; }
                        if (tr) { t_issue += ti1 - ti0; t_commit += clock64() - ti1; }
                    }
                    mma_commit(pl_empty(pbuf));
                }
            }
            if (tr) { trow[4] = w0; trow[5] = w1; trow[6] = w2; trow[7] = clock64() - t_begin; trow[16] = t_issue; trow[17] = t_commit; }
        }
    } else if (warp < 10) {
        // ================= transform: fp32 box -> split planes (non-swizzled K-major) =================
        const int tt = threadIdx.x - 64;                  // 0..255
        int bx = 0, kbg = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int x0, y0, b0, n0;
            tile_coords(tile, x0, y0, b0, n0);
            for (int kb = 0; kb < p.nkb; ++kb, ++kbg) {
                const int pbuf = kbg & 1;
                const uint32_t pl0 = plane(pbuf, 0), pl1 = plane(pbuf, 1);
                bool waited = false;
                for (int hb = 0; hb < C::BOXES; ++hb) {
                    if (kb * C::KB + 32 * hb >= p.ci) continue;
                    const int s = bx % C::SLOTS;
                    mbar_wait_t(st_full(s), (bx / C::SLOTS) & 1, tr, w0);
                    if (!waited) { mbar_wait_t(pl_empty(pbuf), ((kbg >> 1) & 1) ^ 1, tr, w1); waited = true; }
                    const uint32_t src0 = slot_base + s * SLOT_BYTES;
                    const int cbase = kb * C::KB + 32 * hb;                      // first channel of this box
                    // items = (row, 16-channel half); consecutive lanes take consecutive rows -> conflict-free both ways
                    for (int item = tt; item < 2 * PR; item += 256) {
                        const int half = item >= PR ? 1 : 0, r = item - half * PR;
                        const uint32_t src = src0 + (uint32_t)r * 128u;
                        const int sw = r & 7;
                        float v[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 t4 = lds4(src + (uint32_t)(((4 * half + q) ^ sw) << 4));
                            v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
                        }
                        if (p.in_scale) {
                            const int c = cbase + 16 * half;
                            if (c < p.ci) {
                                const float* sp = p.in_scale + (long long)b0 * p.ci + c;
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float4 t4 = ldg4(sp + 4 * q);
                                    v[4 * q] *= t4.x; v[4 * q + 1] *= t4.y; v[4 * q + 2] *= t4.z; v[4 * q + 3] *= t4.w;
                                }
                            }
                        }
                        // 16 channels = 2 chunks of 8 halves; chunk index within the 64-channel block = 4*hb + 2*half + q
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            if (PRECISE) split2_f16(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
                            else split2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint32_t off = (uint32_t)(4 * hb + 2 * half + q) * LBO + (uint32_t)r * 16u;
                            sts4(pl0 + off, hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                            sts4(pl1 + off, lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(st_empty(s));
                    ++bx;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(pl_full(pbuf));
            }
        }
        if (tr && tt == 0) { trow[1] = w0; trow[2] = w1; trow[3] = clock64() - t_begin; }
    } else {
        // ================= epilogue warps (PRECISE: promotion of every segment, then epilogue) =================
        const int ew = warp - 10;
        const int q4 = warp & 3;
        const int er = q4 * 32 + lane;                    // accumulator row = pixel (y*8 + x)
        constexpr int COLS = PRECISE ? BN / 2 : BN;        // columns this thread owns
        const int cstart = PRECISE ? (ew >> 2) * COLS : 0;
        const int nflat = p.nkb * taps;
        const int nseg = PRECISE ? (nflat + p.promo_taps - 1) / p.promo_taps : 1;
        const uint32_t stg = b_base + (uint32_t)p.bstages * C::BTILE;         // TMA-store staging: [32-channel block][128 rows][128 B]
        const uint32_t stg_thr = stg + (uint32_t)er * 128u + (uint32_t)(cstart >> 5) * (128u * 128u);
        const uint32_t stg_x = (uint32_t)(((cstart & 31) >> 2) ^ (er & 7));
        constexpr int EPI_THREADS = C::EPI_WARPS * 32;
        const float alpha_eff = p.act == 3 ? p.alpha : 1.f, gain = p.gain;
        const int smode = p.tma_store ? 0 : (p.ys[1] == 1 ? 1 : 2);
        int sg = 0, tl = 0;
#if SG2_TRACE_EPI
        long long t_promo = 0, t_fin = 0, t_stage = 0;     // trace: cycles in promotion / finish+store / staging hand-over
#endif
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
            int x0, y0, b0, n0;
            tile_coords(tile, x0, y0, b0, n0);
            const int ex = x0 + (er & 7), ey = y0 + (er >> 3);
            const bool inside = ex < p.w && ey < p.h;      // ragged images: tiles hang over the right / bottom edge
            const long long pix = ((long long)b0 * p.h + ey) * p.w + ex;
            const float nz = (p.noise && inside) ? __ldg(p.noise + pix) : 0.f;
            float* yrow = p.y + (long long)b0 * p.ys[0] + (long long)ey * p.ys[2] + (long long)ex * p.ys[3];
            const float* osc = p.out_scale ? p.out_scale + (long long)b0 * p.co + n0 + cstart : nullptr;
            const float* bsp = p.bias ? p.bias + n0 + cstart : nullptr;
            float* ydense = yrow + n0 + cstart;                                  // mode 1: this thread's first channel of its pixel
            // out_scale, bias, noise, activation, gain on 4 channels, then the store.  MODE (compile time; chosen once per kernel):
            // 0 = swizzled staging tile for the TMA store, 1 = dense NHWC (16-byte stores), 2 = any strides.
            auto finish4 = [&](auto mode, float (&o)[4], int cbase) {
                constexpr int MODE = decltype(mode)::value;
                const float4 sc = osc ? ldg4(osc + cbase) : make_float4(1.f, 1.f, 1.f, 1.f);
                const float4 bi = bsp ? ldg4(bsp + cbase) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, biv[4] = {bi.x, bi.y, bi.z, bi.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float val = fmaf(o[e], scv[e], biv[e]) + nz;
                    o[e] = (val > 0.f ? val : val * alpha_eff) * gain;
                }
                if (MODE == 0) {
                    // staging address: [32-channel block][row][16-byte chunk ^ (row & 7)]; cstart is a multiple of 16, cbase < COLS
                    sts4(stg_thr + (uint32_t)(cbase >> 5) * (128u * 128u) + ((stg_x ^ (uint32_t)((cbase & 31) >> 2)) << 4),
                         __float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
                } else if (MODE == 1) {
                    if (inside) st4(ydense + cbase, make_float4(o[0], o[1], o[2], o[3]));
                } else {
                    if (inside) {
                        const int cb = n0 + cstart + cbase;
#pragma unroll
                        for (int e = 0; e < 4; ++e) yrow[(long long)(cb + e) * p.ys[1]] = o[e];
                    }
                }
            };
            const uint32_t lane_addr = tmem_d + ((uint32_t)(q4 * 32) << 16);
#if SG2_TRACE_EPI
            long long tq = tr ? clock64() : 0;
#endif
            if (p.tma_store) {
                // the previous tile's TMA stores must have READ the staging tile before it is overwritten
                if (ew == 0 && lane == 0) bulk_wait_read_all();
                named_bar_sync(1, EPI_THREADS);
#if SG2_TRACE_EPI
                if (tr) { const long long t = clock64(); t_stage += t - tq; tq = t; }
#endif
            }
#if SG2_TRACE_EPI
            const long long w_before = w0;
#endif
            if (PRECISE) {
                // every segment (the last one included) is promoted into fp32 registers -- acc += D1 + 2^-11 D2 -- and its
                // TMEM buffer handed back at once; the epilogue then runs from registers while the MMA warp is already
                // filling the next tile's segments
                float racc[COLS];
#pragma unroll
                for (int j = 0; j < COLS; ++j) racc[j] = 0.f;
                constexpr int G = COLS == 32 ? 2 : 1;          // 16-column groups per round: 2G TMEM loads in flight, one wait
                for (int seg = 0; seg < nseg; ++seg, ++sg) {
                    const int abuf = sg % C::NACC;
                    mbar_wait_t(acc_full(abuf), (sg / C::NACC) & 1, tr, w0);
                    tc_fence_after();
                    if (C::D2X) {
                        // D1 only: 4 loads of 16 columns in flight per wait
#pragma unroll
                        for (int c = 0; c < COLS / 16; c += 4) {
                            uint32_t v[4][16];
#pragma unroll
                            for (int g = 0; g < 4; ++g) tmem_ld16_async(lane_addr + (uint32_t)(abuf * 2 * BN + cstart + (c + g) * 16), v[g]);
                            tmem_ld_wait();
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                reg_fence(v[g]);
#pragma unroll
                                for (int j = 0; j < 16; ++j) racc[(c + g) * 16 + j] += __uint_as_float(v[g][j]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < COLS / 16; c += G) {
                            uint32_t v[G][16], v2[G][16];
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                const uint32_t col = (uint32_t)(abuf * C::ACC_COLS + cstart + (c + g) * 16);
                                tmem_ld16_async(lane_addr + col, v[g]);
                                tmem_ld16_async(lane_addr + col + BN, v2[g]);
                            }
                            tmem_ld_wait();
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                reg_fence(v[g]); reg_fence(v2[g]);
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    racc[(c + g) * 16 + j] += fmaf(__uint_as_float(v2[g][j]), 1.f / 2048.f, __uint_as_float(v[g][j]));
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(abuf));
                }
                if (C::D2X) {
                    // the tile's cross terms, once: the last segment's acc_full also covers every MMA into D2
                    const int tp = tl & 1;
#pragma unroll
                    for (int c = 0; c < COLS / 16; c += 4) {
                        uint32_t v[4][16];
#pragma unroll
                        for (int g = 0; g < 4; ++g) tmem_ld16_async(lane_addr + (uint32_t)(BN + tp * 2 * BN + cstart + (c + g) * 16), v[g]);
                        tmem_ld_wait();
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            reg_fence(v[g]);
#pragma unroll
                            for (int j = 0; j < 16; ++j) racc[(c + g) * 16 + j] = fmaf(__uint_as_float(v[g][j]), 1.f / 2048.f, racc[(c + g) * 16 + j]);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(d2_empty(tp));
                }
#if SG2_TRACE_EPI
                if (tr) { const long long t = clock64(); t_promo += t - tq - (w0 - w_before); tq = t; }
#endif
                auto finish_all = [&](auto mode) {
#pragma unroll
                    for (int j = 0; j < COLS; j += 4) {
                        float o[4] = {racc[j], racc[j + 1], racc[j + 2], racc[j + 3]};
                        finish4(mode, o, j);
                    }
                };
                if (smode == 0) finish_all(Mode<0>{}); else if (smode == 1) finish_all(Mode<1>{}); else finish_all(Mode<2>{});
#if SG2_TRACE_EPI
                if (tr) { const long long t = clock64(); t_fin += t - tq; tq = t; }
#endif
            } else {
                // one segment per tile: read both halves (hi*hi | hi*lo + lo*hi), add, finish, store, hand the buffer back
                const int abuf = sg % C::NACC;
                mbar_wait_t(acc_full(abuf), (sg / C::NACC) & 1, tr, w0);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < COLS / 16; ++c) {
                    uint32_t v[16], v2[16];
                    const uint32_t col = (uint32_t)(abuf * C::ACC_COLS + cstart + c * 16);
                    tmem_ld16_async(lane_addr + col, v);
                    tmem_ld16_async(lane_addr + col + BN, v2);
                    tmem_ld_wait();
                    reg_fence(v); reg_fence(v2);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float o[4] = {__uint_as_float(v[j]) + __uint_as_float(v2[j]), __uint_as_float(v[j + 1]) + __uint_as_float(v2[j + 1]),
                                      __uint_as_float(v[j + 2]) + __uint_as_float(v2[j + 2]), __uint_as_float(v[j + 3]) + __uint_as_float(v2[j + 3])};
                        if (smode == 0) finish4(Mode<0>{}, o, c * 16 + j); else if (smode == 1) finish4(Mode<1>{}, o, c * 16 + j);
                        else finish4(Mode<2>{}, o, c * 16 + j);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(abuf));
                ++sg;
            }
            if (p.tma_store) {
                fence_proxy_async();
                named_bar_sync(1, EPI_THREADS);
                if (ew == 0 && lane == 0) {
                    for (int blk = 0; blk < BN / 32; ++blk) tma_store_4d(&ymap, stg + (uint32_t)blk * (128u * 128u), n0 + 32 * blk, x0, y0, b0);
                    bulk_commit();
                }
#if SG2_TRACE_EPI
                if (tr) t_stage += clock64() - tq;
#endif
            }
        }
        if (p.tma_store && ew == 0 && lane == 0) bulk_wait_all();
        if (tr && ew == 0 && lane == 0) {
            trow[8] = w0; trow[9] = clock64() - t_begin;
#if SG2_TRACE_EPI
            trow[13] = t_promo; trow[14] = t_fin; trow[15] = t_stage;
#endif
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tr && threadIdx.x == 0) { trow[11] = clock64() - t_begin; trow[12] = (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x; }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, C::TMEM_COLS);
    }
}

// w[co][ci][k][k] -> per (n-tile, channel block, tap): two planes of [bn rows x 128 B], SWIZZLE_128B, zero beyond ci.
// One thread = one 16-byte chunk (8 consecutive input channels) of one row in both planes: 128-bit stores.
__global__ void conv_pack_halo_kernel(const float* __restrict__ w, unsigned char* __restrict__ wp, int co, int ci, int k,
                                      float coef, int transpose, int bn, int nkb, int kbs, int precise) {
    const int cin = transpose ? co : ci, nout_n = transpose ? ci : co;
    const int kk2 = k * k, chunks = kbs >> 3;
    const long long total = (long long)(nout_n / bn) * nkb * kk2 * bn * chunks;
    const size_t btile = (size_t)2 * bn * 128;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(idx % chunks);
        long long r = idx / chunks;
        const int nl = (int)(r % bn); r /= bn;
        const int t = (int)(r % kk2); r /= kk2;
        const int kb = (int)(r % nkb);
        const int nt = (int)(r / nkb);
        const int nout = nt * bn + nl, ts = transpose ? kk2 - 1 - t : t;
        uint32_t p0[4], p1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kin = kb * kbs + c8 * 8 + 2 * e + h;
                v[h] = 0.f;
                if (kin < cin) {
                    const int o = transpose ? kin : nout, i = transpose ? nout : kin;
                    v[h] = __ldg(w + ((long long)o * ci + i) * kk2 + ts) * coef;
                }
            }
            if (precise) tc::split2_f16(v[0], v[1], p0[e], p1[e]);
            else tc::split2(v[0], v[1], p0[e], p1[e]);
        }
        unsigned char* tile = wp + (((size_t)nt * nkb + kb) * kk2 + t) * btile;
        const size_t off = (size_t)nl * 128 + (size_t)((c8 ^ (nl & 7)) << 4);
        *reinterpret_cast<uint4*>(tile + off) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
        *reinterpret_cast<uint4*>(tile + (size_t)bn * 128 + off) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    }
}

// output-channel tile: every multiple of 32 has one (n_tiles = co / BN)
static int pick_bn(int co) { return co % 128 == 0 ? 128 : (co % 64 == 0 ? 64 : (co % 32 == 0 ? 32 : 0)); }

template <int BN, bool PRECISE>
static int launch(const CUtensorMap& map, const CUtensorMap& ymap, Params& tp, dim3 grid, cudaStream_t st) {
    using C = Cfg<BN, PRECISE>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "conv_halo: cannot opt in to %d B of shared memory: %s", C::SMEM, cudaGetErrorString(e));
        configured = true;
    }
    tp.bstages = C::BSTAGES;
    if (tp.tma_store) {
        // a 1x1 convolution streams one weight tile per channel block: two stages suffice, the rest of the region stages the output
        if ((C::BSTAGES - 2) * C::BTILE >= BN * 512) tp.bstages = 2; else tp.tma_store = 0;
    }
    conv_halo_kernel<BN, PRECISE><<<grid, C::NTHREADS, C::SMEM, st>>>(map, ymap, tp);
    return launched(PRECISE ? "conv_halo_fp16x3" : "conv_halo_bf16x3");
}

}  // namespace halo

bool conv_halo_supported(int n, int h, int w, int ci, int co, int k) {
    if (k != 1 && k != 3) return false;
    if (ci % 32 != 0 || ci < 32) return false;
    const int bn = halo::pick_bn(co);
    if (!bn) return false;
    if (w > 4096 || h > 4096) return false;
    // images that tile exactly by 8 x 16, or large ragged ones (edge tiles are masked; e.g. the 257^2 blurred inputs of
    // the StyleGAN3-style discriminator) ...
    if (((w % halo::TW) == 0 && (h % halo::TH) == 0) || (w >= 32 && h >= 32)) return true;
    // ... or small images (4^2, 8^2) whose tiles, however empty, all fit the machine at once: one 8 x 16 tile per image
    // wastes most of its rows, but every CTA runs the full-rate pipeline in a single wave, which beat round 1's per-tap kernels
    const long long tiles = (long long)((w + halo::TW - 1) / halo::TW) * ((h + halo::TH - 1) / halo::TH) * n * (co / bn);
    return w >= 4 && h >= 4 && tiles <= 2LL * num_sms();
}

long long conv_packed_bytes_halo(int co, int ci, int k) {
    long long best = 0;
    for (int tr = 0; tr < 2; ++tr) {
        const int cin = tr ? co : ci, cout = tr ? ci : co;
        const int bn = halo::pick_bn(cout);
        if (!bn || cin % 32) continue;
        const long long nkb = (cin + 63) / 64;
        const long long b = (long long)(cout / bn) * nkb * k * k * 2 * bn * 128;
        if (b > best) best = b;
    }
    return best;
}

int conv_pack_halo(const float* w, void* wp, int co, int ci, int k, float coef, int transpose, int precise, cudaStream_t st) {
    const int cin = transpose ? co : ci, cout = transpose ? ci : co;
    const int bn = halo::pick_bn(cout);
    if (!bn || cin % 32) return fail(SG2_ENOTSUP, "conv_pack_halo: unsupported shape");
    const int kbs = 64;
    const int nkb = (cin + kbs - 1) / kbs;
    const long long total = (long long)(cout / bn) * nkb * k * k * bn * (kbs / 8);
    const int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 8);
    halo::conv_pack_halo_kernel<<<blocks, 256, 0, st>>>(w, (unsigned char*)wp, co, ci, k, coef, transpose, bn, nkb, kbs, precise);
    return launched("conv_pack_halo");
}

int conv_fwd_halo(const ConvParams& p, int precise, cudaStream_t st) {
    if (!conv_halo_supported(p.n, p.h, p.w, p.ci, p.co, p.k)) return fail(SG2_ENOTSUP, "conv_fwd_halo: unsupported shape");
    const int pad = p.k >> 1;
    CUtensorMap map;
    int rc = tc::make_nhwc_map(&map, p.x, p.n, p.h, p.w, p.ci, halo::TW + 2 * pad, halo::TH + 2 * pad, 1, "conv_fwd_halo");
    if (rc) return rc;
    halo::Params tp;
    tp.in_scale = p.in_scale; tp.out_scale = p.out_scale; tp.bias = p.bias; tp.noise = p.noise;
    tp.y = p.y;
    for (int i = 0; i < 4; ++i) tp.ys[i] = p.ys[i];
    tp.wp = (const unsigned char*)p.wp;
    tp.n = p.n; tp.h = p.h; tp.w = p.w; tp.ci = p.ci; tp.co = p.co; tp.k = p.k;
    tp.tiles_x = (p.w + halo::TW - 1) / halo::TW; tp.tiles_y = (p.h + halo::TH - 1) / halo::TH;
    tp.m_tiles = tp.tiles_x * tp.tiles_y * p.n;
    auto log2_exact = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
    tp.tx_shift = log2_exact(tp.tiles_x); tp.ty_shift = log2_exact(tp.tiles_y);
    if (tp.tx_shift < 0 || tp.ty_shift < 0) tp.tx_shift = tp.ty_shift = -1;
    const int bn = halo::pick_bn(p.co);
    tp.n_tiles = p.co / bn;
    tp.nkb = (p.ci + 63) / 64;
    tp.act = p.act; tp.alpha = p.alpha; tp.gain = p.gain;
    tp.trace = g_trace;
    // promotion interval of the fp32-class kernel (taps accumulated in TMEM between two promotions, 4 chained big*big MMAs
    // each).  Error of one convolution against fp64 (scripts/promo_sweep.py, profiles/r1b_promo_sweep.txt): 2 taps 3.5e-7,
    // 5 taps 5e-7, 9 taps 0.8..1.1e-6 -- torch/cuDNN fp32 itself is 0.3..2e-6 on the same cases.  What decides the setting is the
    // full-width model (scripts/noise_study.py, profiles/r2j_noise_study.txt: 3 seeds, every gradient class, ours and four fp32
    // evaluations of the reference against fp64): at 2 taps ours is 3-5x CLOSER to fp64 than the reference's own fp32 arithmetic
    // (median over a class), at 5 taps on par or better in every class (G gradients 7e-5 vs 8e-5, D gradients 3e-5 vs 5e-5, R1
    // bias gradients 8e-4 vs 1.2e-3), at 9 taps 1.2-4x worse.  Default 5: never less accurate than the reference, 6-9 % faster
    // forward convolutions than 2.  SG2_PROMO_TAPS overrides it for the sweeps.
    static int promo_taps = 0;
    if (!promo_taps) {
        const char* e = getenv("SG2_PROMO_TAPS");
        promo_taps = e ? atoi(e) : 5;
        if (promo_taps < 1 || promo_taps > 9) promo_taps = 5;
    }
    tp.promo_taps = promo_taps;
    // TMA-store epilogue: 1x1 convolutions into a dense NHWC tensor (the skip convolutions of the discriminator blocks)
    CUtensorMap ymap = map;
    tp.tma_store = 0;
    if (p.k == 1 && p.ys[1] == 1 && p.ys[3] == p.co && p.ys[2] == (long long)p.w * p.co && p.ys[0] == (long long)p.h * p.w * p.co &&
        ((uintptr_t)p.y & 15) == 0) {
        rc = tc::make_nhwc_map(&ymap, p.y, p.n, p.h, p.w, p.co, halo::TW, halo::TH, 1, "conv_fwd_halo(y)");
        if (rc) return rc;
        tp.tma_store = 1;
    }
    dim3 grid((unsigned)std::min(tp.m_tiles * tp.n_tiles, num_sms()));
    if (precise) {
        if (bn == 128) return halo::launch<128, true>(map, ymap, tp, grid, st);
        if (bn == 64) return halo::launch<64, true>(map, ymap, tp, grid, st);
        return halo::launch<32, true>(map, ymap, tp, grid, st);
    }
    if (bn == 128) return halo::launch<128, false>(map, ymap, tp, grid, st);
    if (bn == 64) return halo::launch<64, false>(map, ymap, tp, grid, st);
    return halo::launch<32, false>(map, ymap, tp, grid, st);
}

}  // namespace sg2
