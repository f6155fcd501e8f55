// tcgen05 implicit-GEMM convolution on a bf16 pair-planes input ("halo" tiling, bf16x3): TMA -> tensor core, no transform.
//
// The data-gradient convolutions of the first-order backward pass (convolution_backward behind
// implementations/StyleGAN2/model.py:129, 44-53; composed like thirdparty/stylegan3_ops/ops/conv2d_gradfix.py:99-145) take the
// output gradient as pair planes (planes.cu: written by the leaky-ReLU-gradient pass that touches gy anyway).  Compared with
// conv_halo.cu<PRECISE = false>, which loads an fp32 patch and converts it with 8 transform warps, a pipeline stage here is two
// TMA boxes [64 ch, 8 + 2p, 16 + 2p, 1] (hi and lo plane; OOB -> 0 = zero padding) landing as K-major SWIZZLE_128B rows, one
// 128-byte row per patch pixel.  The k*k taps are the SAME tile read through descriptors whose start address is shifted by
// (dy * PW + dx) rows and whose 8-row groups are PW rows apart: tcgen05 applies the 128-byte swizzle on absolute shared-memory
// addresses, so any row-shifted start and any group stride read correctly (scripts/exp_umma_shift.cu,
// profiles/r2a_umma_shift.txt).  MMA scheme, weight tiles, accumulators and epilogue are those of conv_halo.cu:
//   per (tap, k16): [hi*hi | hi*lo] += A_hi x [B_hi ; B_lo] (N = 2 BN),  upper half += A_lo x B_hi;  result = lower + upper.
// Warp roles: 0 patch TMA, 1 MMA (+TMEM alloc), 2..9 epilogue, 10 weight TMA.
#include <stdlib.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"
#include "conv.h"

namespace sg2 {
namespace halopl {
using namespace tc;

constexpr int TW = 8, TH = 16, BM = 128;
constexpr int PLANE_PITCH = 23 * 1024;                  // 180 patch rows x 128 B = 23040, 1024-aligned pitch
constexpr int EPI_WARPS = 8;                           // two per TMEM lane quarter, each owning half of the tile's columns
constexpr int NTHREADS = (3 + EPI_WARPS) * 32;

template <int BN> struct Cfg {
    static constexpr int BTILE = 2 * BN * 128;
    static constexpr int PSTAGES = BN == 32 ? 3 : 2;                        // patch stages (2 planes each)
    static constexpr int BSTAGES = BN == 128 ? 4 : (BN == 64 ? 5 : 6);
    // output staging tile of the TMA-store epilogue, [32-channel block][128 rows][128 B]: BN < 128 has room for a dedicated one
    // (3x3 convolutions too); BN = 128 borrows the weight stages a 1x1 convolution does not need
    static constexpr int STG_BYTES = BN < 128 ? BN * 512 : 0;
    static constexpr int SMEM = 1024 + PSTAGES * 2 * PLANE_PITCH + BSTAGES * BTILE + STG_BYTES + 256;
    static constexpr int ACC_COLS = 2 * BN;
    static constexpr int NACC = 512 / ACC_COLS >= 4 ? 4 : 512 / ACC_COLS;
    static constexpr uint32_t TMEM_COLS = NACC * ACC_COLS;
};

template <int M> struct Mode { static constexpr int value = M; };      // compile-time store mode of the epilogue

struct Params {
    const float* out_scale; const float* bias;
    float* y;
    long long ys[4];
    const unsigned char* wp;
    int n, h, w, ci, co, k;
    int tiles_x, tiles_y, m_tiles, n_tiles;
    int nkb;
    int act;
    float alpha, gain;
    int accumulate;          // y += result (the skip branch's data gradient lands on top of the main branch's)
    int bstages;             // weight stages in use (<= Cfg::BSTAGES)
    int two_term;            // experiment (SG2_GRAD_TERMS=2): drop the A_lo x B_hi product (the input is then effectively bf16)
    int tma_store;           // dense NHWC output: the epilogue stages the tile in shared memory and stores it with TMA (below)
    int stg_off;             // staging tile, bytes after the first weight stage
    int tx_shift, ty_shift;  // log2(tiles_x), log2(tiles_y) when both are powers of two, else -1
};

// Epilogue stores.  A thread owns one pixel (accumulator row); stored straight from registers, one 16-byte store instruction of a
// warp touches 32 different 128-byte lines -- for a 1x1 convolution (8 MMAs per tile) those 2 K store wavefronts per tile were the
// whole kernel (role trace: MMA warp 60 % waiting for a free accumulator; 23 TF/s, 2.3 TB/s).  With tma_store the warps write
// their rows into a SWIZZLE_128B staging tile (conflict-free: 8 lanes hit 8 different 16-byte chunks) and ONE elected thread
// issues a TMA tensor store (or reduce-add, for `accumulate`) per 32-channel block; the staging tile lives in the weight stages
// a 1x1 convolution does not need (it has one tap per channel block).
template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) conv_halo_pl_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap,
                                                                   const Params p) {
    using C = Cfg<BN>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t plane_base = base;                                          // [stage][plane]
    const uint32_t b_base = plane_base + C::PSTAGES * 2 * PLANE_PITCH;
    const uint32_t bar_base = b_base + C::BSTAGES * C::BTILE + C::STG_BYTES;
    auto pl_full = [&](int s) { return bar_base + 8u * s; };                   // <= 4
    auto pl_empty = [&](int s) { return bar_base + 32u + 8u * s; };
    auto b_full = [&](int s) { return bar_base + 64u + 8u * s; };              // <= 6
    auto b_empty = [&](int s) { return bar_base + 112u + 8u * s; };
    auto acc_full = [&](int b) { return bar_base + 160u + 8u * b; };           // <= 4
    auto acc_empty = [&](int b) { return bar_base + 192u + 8u * b; };
    const uint32_t tmem_slot = bar_base + 224u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pad = p.k >> 1, taps = p.k * p.k;
    const int PW = TW + 2 * pad, PH = TH + 2 * pad, PR = PW * PH;
    const int total_tiles = p.m_tiles * p.n_tiles;
    auto tile_coords = [&](int tile, int& x0, int& y0, int& b0, int& n0) {
        int nt = 0, mt = tile;                             // n_tiles is 1..4: no division
        while (mt >= p.m_tiles) { mt -= p.m_tiles; ++nt; }
        if (p.tx_shift >= 0) {                             // power-of-two tile grid: shifts and masks
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
        for (int s = 0; s < C::PSTAGES; ++s) { mbar_init(pl_full(s), 1); mbar_init(pl_empty(s), 1); }
        for (int s = 0; s < C::NACC; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), EPI_WARPS); }
        for (int s = 0; s < C::BSTAGES; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_d;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot));

    if (warp == 0) {
        // ================= patch producer: two TMA boxes (hi, lo plane) per (tile, 64-channel block) =================
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            int kbg = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int x0, y0, b0, n0;
                tile_coords(tile, x0, y0, b0, n0);
                for (int kb = 0; kb < p.nkb; ++kb, ++kbg) {
                    const int s = kbg % C::PSTAGES;
                    mbar_wait(pl_empty(s), ((kbg / C::PSTAGES) & 1) ^ 1);
                    mbar_expect_tx(pl_full(s), (uint32_t)PR * 128u * 2u);
                    const uint32_t dst = plane_base + (uint32_t)(s * 2) * PLANE_PITCH;
                    tma_load_5d(dst, &xmap, pl_full(s), kb * 64, x0 - pad, y0 - pad, b0, 0);
                    tma_load_5d(dst + PLANE_PITCH, &xmap, pl_full(s), kb * 64, x0 - pad, y0 - pad, b0, 1);
                }
            }
        }
    } else if (warp == 2 + EPI_WARPS) {
        // ================= weight producer: one pre-packed tile per (tile, channel block, tap) =================
        if (elect_one()) {
            int s = 0;
            uint32_t sph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int nt = 0;
                for (int mt = tile; mt >= p.m_tiles; mt -= p.m_tiles) ++nt;
                const unsigned char* wsrc = p.wp + (size_t)nt * p.nkb * taps * C::BTILE;
                for (int i = 0; i < p.nkb * taps; ++i) {
                    mbar_wait(b_empty(s), sph ^ 1);
                    mbar_expect_tx(b_full(s), C::BTILE);
                    bulk_load(b_base + s * C::BTILE, wsrc + (size_t)i * C::BTILE, C::BTILE, b_full(s));
                    if (++s == p.bstages) { s = 0; sph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // one thread; kept free of divisions and 64-bit descriptor arithmetic (see conv_halo.cu)
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_bf16(BM, BN), idesc2 = idesc_bf16(BM, 2 * BN);
            const uint32_t a_hi = (((uint32_t)PW * 128u) >> 4) | (1u << 14) | (2u << 29);       // A: K-major SWIZZLE_128B, 8-row groups PW rows apart
            constexpr uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);                     // B: K-major SWIZZLE_128B, SBO 1024
            constexpr uint32_t lo_f = 1u << 16;
            const uint32_t nstages = (uint32_t)p.bstages, kdim = (uint32_t)p.k, row_wrap = (uint32_t)(PW - p.k) * 8u;
            uint32_t s = 0, sph = 0, abuf = 0, aph = 0, ps = 0, pph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(acc_empty(abuf), aph ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_d + abuf * (uint32_t)C::ACC_COLS;
                int ci_left = p.ci;
                for (int kb = 0; kb < p.nkb; ++kb, ci_left -= 64) {
                    mbar_wait(pl_full(ps), pph);
                    const uint32_t pl0 = (((plane_base + ps * 2u * PLANE_PITCH) & 0x3FFFFu) >> 4) | lo_f, pl1 = pl0 + (PLANE_PITCH >> 4);
                    const int kqn = min(4, ci_left >> 4);               // 32-channel tail: the zero-filled half of the box is skipped
                    uint32_t arow = 0, dx = 0;                          // (dy * PW + dx) * 128 B >> 4
                    for (int t = 0; t < taps; ++t) {
                        mbar_wait(b_full(s), sph);
                        const uint32_t b0_ = (((b_base + s * C::BTILE) & 0x3FFFFu) >> 4) | lo_f;
#pragma unroll
                        for (int kq = 0; kq < 4; ++kq) {
                            if (kq >= kqn) break;
                            const uint32_t db = b0_ + kq * 2;
                            mma_f16_words(d, pl0 + arow + kq * 2, a_hi, db, b_hi, idesc2, !(kb == 0 && t == 0 && kq == 0));   // [hi*hi | hi*lo] += A_hi * [B_hi ; B_lo]
                            if (!p.two_term) mma_f16_words(d + BN, pl1 + arow + kq * 2, a_hi, db, b_hi, idesc, 1);          //          upper  += A_lo * B_hi
                        }
                        mma_commit(b_empty(s));
                        if (++s == nstages) { s = 0; sph ^= 1; }
                        arow += 8;
                        if (++dx == kdim) { dx = 0; arow += row_wrap; }
                    }
                    mma_commit(pl_empty(ps));
                    if (++ps == C::PSTAGES) { ps = 0; pph ^= 1; }
                }
                mma_commit(acc_full(abuf));
                if (++abuf == C::NACC) { abuf = 0; aph ^= 1; }
            }
        }
    } else if (warp >= 2 && warp < 2 + EPI_WARPS) {
        // ================= epilogue warps =================
        const int ew = warp - 2, q4 = warp & 3;
        const int er = q4 * 32 + lane;                    // accumulator row = pixel (y*8 + x)
        constexpr int COLS = BN / 2;                      // columns this thread owns
        const int cstart = (ew >> 2) * COLS;
        constexpr int EPI_THREADS = EPI_WARPS * 32;
        const uint32_t stg = b_base + (uint32_t)p.stg_off;                    // staging tile: [32-channel block][128 rows][128 B]
        const uint32_t stg_thr = stg + (uint32_t)er * 128u + (uint32_t)(cstart >> 5) * (128u * 128u);
        const uint32_t stg_x = (uint32_t)(((cstart & 31) >> 2) ^ (er & 7));
        const float alpha_eff = p.act == 3 ? p.alpha : 1.f, gain = p.gain;
        const int smode = p.tma_store ? 0 : (p.ys[1] == 1 ? 1 : 2);
        const uint32_t lane_addr = tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)cstart;
        uint32_t abuf = 0, aph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int x0, y0, b0, n0;
            tile_coords(tile, x0, y0, b0, n0);
            const int ex = x0 + (er & 7), ey = y0 + (er >> 3);
            const bool inside = ex < p.w && ey < p.h;
            float* yrow = p.y + (long long)b0 * p.ys[0] + (long long)ey * p.ys[2] + (long long)ex * p.ys[3];
            float* ydense = yrow + n0 + cstart;
            const float* osc = p.out_scale ? p.out_scale + (long long)b0 * p.co + n0 + cstart : nullptr;
            const float* bsp = p.bias ? p.bias + n0 + cstart : nullptr;
            mbar_wait(acc_full(abuf), aph);
            tc_fence_after();
            // the whole column range in one round of TMEM loads; the accumulator goes back to the MMA warp before the stores
            float racc[COLS];
#pragma unroll
            for (int c = 0; c < COLS / 16; ++c) {
                uint32_t v[16], v2[16];
                const uint32_t col = abuf * (uint32_t)C::ACC_COLS + (uint32_t)(c * 16);
                tmem_ld16_async(lane_addr + col, v);
                tmem_ld16_async(lane_addr + col + BN, v2);
                tmem_ld_wait();
                reg_fence(v); reg_fence(v2);
#pragma unroll
                for (int j = 0; j < 16; ++j) racc[c * 16 + j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(abuf));
            if (++abuf == C::NACC) { abuf = 0; aph ^= 1; }
            if (smode == 0) {
                // the previous tile's TMA stores must have READ the staging tile before it is overwritten
                if (ew == 0 && lane == 0) bulk_wait_read_all();
                named_bar_sync(1, EPI_THREADS);
            }
            // out_scale, bias, activation, gain on 4 channels, then the store.  MODE: 0 = swizzled staging tile for the TMA store
            // (or reduce-add), 1 = dense NHWC 16-byte stores, 2 = any strides.
            auto finish_all = [&](auto mode) {
                constexpr int MODE = decltype(mode)::value;
#pragma unroll
                for (int cbase = 0; cbase < COLS; cbase += 4) {
                    const float4 sc = osc ? ldg4(osc + cbase) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 bi = bsp ? ldg4(bsp + cbase) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, biv[4] = {bi.x, bi.y, bi.z, bi.w};
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float val = fmaf(racc[cbase + e], scv[e], biv[e]);
                        o[e] = (val > 0.f ? val : val * alpha_eff) * gain;
                    }
                    if (MODE == 0) {
                        sts4(stg_thr + (uint32_t)(cbase >> 5) * (128u * 128u) + ((stg_x ^ (uint32_t)((cbase & 31) >> 2)) << 4),
                             __float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
                    } else if (MODE == 1) {
                        if (inside) {
                            if (p.accumulate) { const float4 t = *reinterpret_cast<const float4*>(ydense + cbase); o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w; }
                            st4(ydense + cbase, make_float4(o[0], o[1], o[2], o[3]));
                        }
                    } else if (inside) {
                        const int cb = n0 + cstart + cbase;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float* dst = yrow + (long long)(cb + e) * p.ys[1];
                            *dst = p.accumulate ? *dst + o[e] : o[e];
                        }
                    }
                }
            };
            if (smode == 0) finish_all(Mode<0>{}); else if (smode == 1) finish_all(Mode<1>{}); else finish_all(Mode<2>{});
            if (smode == 0) {
                fence_proxy_async();                      // the staging writes become visible to the async proxy (TMA)
                named_bar_sync(1, EPI_THREADS);
                if (ew == 0 && lane == 0) {
                    for (int blk = 0; blk < BN / 32; ++blk) {
                        if (p.accumulate) tma_reduce_add_4d(&ymap, stg + (uint32_t)blk * (128u * 128u), n0 + 32 * blk, x0, y0, b0);
                        else tma_store_4d(&ymap, stg + (uint32_t)blk * (128u * 128u), n0 + 32 * blk, x0, y0, b0);
                    }
                    bulk_commit();
                }
            }
        }
        if (p.tma_store && ew == 0 && lane == 0) bulk_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, C::TMEM_COLS);
    }
}

// output-channel tile: every multiple of 32 has one (n_tiles = co / BN)
static int pick_bn(int co) { return co % 128 == 0 ? 128 : (co % 64 == 0 ? 64 : (co % 32 == 0 ? 32 : 0)); }

template <int BN>
static int launch(const CUtensorMap& map, const CUtensorMap& ymap, Params& tp, dim3 grid, cudaStream_t st) {
    using C = Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_pl_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "conv_halo_pl: cannot opt in to %d B of shared memory: %s", C::SMEM, cudaGetErrorString(e));
        configured = true;
    }
    tp.bstages = C::BSTAGES;
    tp.stg_off = C::BSTAGES * C::BTILE;
    if (tp.tma_store && !C::STG_BYTES) {
        // BN = 128: only a 1x1 convolution (one weight tile per channel block: two stages are plenty) has room for the staging tile
        if (tp.k == 1 && (C::BSTAGES - 2) * C::BTILE >= BN * 512) { tp.bstages = 2; tp.stg_off = 2 * C::BTILE; } else tp.tma_store = 0;
    }
    conv_halo_pl_kernel<BN><<<grid, NTHREADS, C::SMEM, st>>>(map, ymap, tp);
    return launched("conv_halo_pl");
}

}  // namespace halopl

// ci = channels of the planes input: a multiple of 32.  A TMA box always carries 64 channels (one 128-byte row per pixel); for
// a tensor whose channel count is not a multiple of 64 the box hangs over the channel dimension and TMA fills the overhang
// with zeros (the packed weight is zero there too), so 32-channel tensors need no padded copy in HBM.
bool conv_halo_pl_supported(int n, int h, int w, int ci, int co, int k) {
    if (ci % 32 != 0) return false;
    return conv_halo_supported(n, h, w, ci, co, k);
}

int conv_fwd_halo_pl(const void* x_planes, const ConvParams& p, int accumulate, cudaStream_t st) {
    if (!conv_halo_pl_supported(p.n, p.h, p.w, p.ci, p.co, p.k)) return fail(SG2_ENOTSUP, "conv_fwd_halo_pl: unsupported shape");
    const int pad = p.k >> 1;
    CUtensorMap map;
    int rc = tc::make_planes_map(&map, x_planes, p.n, p.h, p.w, p.ci, halopl::TW + 2 * pad, halopl::TH + 2 * pad, 1, "conv_fwd_halo_pl");
    if (rc) return rc;
    halopl::Params tp;
    tp.out_scale = p.out_scale; tp.bias = p.bias;
    tp.y = p.y;
    for (int i = 0; i < 4; ++i) tp.ys[i] = p.ys[i];
    tp.wp = (const unsigned char*)p.wp;
    tp.n = p.n; tp.h = p.h; tp.w = p.w; tp.ci = p.ci; tp.co = p.co; tp.k = p.k;
    tp.tiles_x = (p.w + halopl::TW - 1) / halopl::TW; tp.tiles_y = (p.h + halopl::TH - 1) / halopl::TH;
    tp.m_tiles = tp.tiles_x * tp.tiles_y * p.n;
    const int bn = halopl::pick_bn(p.co);
    tp.n_tiles = p.co / bn;
    tp.nkb = (p.ci + 63) / 64;
    tp.act = p.act; tp.alpha = p.alpha; tp.gain = p.gain; tp.accumulate = accumulate;
    static int terms = 0;
    if (!terms) { const char* e = getenv("SG2_GRAD_TERMS"); terms = e ? atoi(e) : 3; }
    tp.two_term = terms == 2;
    auto log2_exact = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
    tp.tx_shift = log2_exact(tp.tiles_x); tp.ty_shift = log2_exact(tp.tiles_y);
    if (tp.tx_shift < 0 || tp.ty_shift < 0) tp.tx_shift = tp.ty_shift = -1;
    // TMA-store epilogue: dense NHWC output tensors
    CUtensorMap ymap = map;
    tp.tma_store = 0;
    if (p.ys[1] == 1 && p.ys[3] == p.co && p.ys[2] == (long long)p.w * p.co && p.ys[0] == (long long)p.h * p.w * p.co &&
        ((uintptr_t)p.y & 15) == 0) {
        rc = tc::make_nhwc_map(&ymap, p.y, p.n, p.h, p.w, p.co, halopl::TW, halopl::TH, 1, "conv_fwd_halo_pl(y)");
        if (rc) return rc;
        tp.tma_store = 1;
    }
    dim3 grid((unsigned)std::min(tp.m_tiles * tp.n_tiles, num_sms()));
    if (bn == 128) return halopl::launch<128>(map, ymap, tp, grid, st);
    if (bn == 64) return halopl::launch<64>(map, ymap, tp, grid, st);
    return halopl::launch<32>(map, ymap, tp, grid, st);
}

}  // namespace sg2
