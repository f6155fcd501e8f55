// tcgen05 weight-gradient kernel on bf16 pair planes: TMA -> tensor core, no transform warps, deterministic reduction.
//
// Replaces the wgrad half of ATen's convolution_backward on the reference path (implementations/StyleGAN2/model.py:129, 44-53;
// composed into autograd like Conv2dGradWeight of thirdparty/stylegan3_ops/ops/conv2d_gradfix.py:147-187) for first-order
// training steps.  Round 1's wgrad_tc.cu took fp32 tensors and converted them in the kernel: gy once per (kernel row, ci tile)
// CTA, x once per (kernel row, co tile) CTA -- its MMA warp waited 63 % of the time for converted operands.  Here both
// operands arrive as bf16 pair planes (planes.cu: hi = bf16(v), lo = bf16(v - hi), written once by the elementwise pass that
// touches the tensor anyway), so a pipeline stage is just TMA boxes landing as SWIZZLE_128B tiles:
//
//   dW[co, (dy,dx), ci] = coef * sum_pix gy[pix, co] * x[pix + (dy,dx), ci]
//   GEMM view per CTA: M = co tile (MN-major A, rows = pixels), N = (dx, ci tile of 64) (MN-major B), K = a split of the pixels.
//   K chunk = 32 pixels (one box of the [n,h,w] grid, as wide as possible):
//     A  per plane: box [64 ch, cw, ch, cb] per 64-co block            -> 32 rows x 128 B
//     B  per plane: ONE box [64 ch, cw + 2, ch, cb] shifted by the kernel row dy (x halo; OOB -> 0 = zero padding).  The k
//        taps of the kernel row are the SAME tile read through descriptors whose start is shifted by dx rows: tcgen05 applies
//        the 128-byte swizzle on absolute shared-memory addresses, so a row-shifted start reads correctly, and the three taps
//        form ONE N = 192 operand whose 64-wide MN blocks are LBO = 128 B (one row) apart (scripts/exp_umma_shift.cu).
//   co % 128 == 0: M = 128 output channels; products hi*hi + lo*hi + hi*lo (3 MMAs per k16 step).
//   otherwise (co % 64 == 0): M = [hi plane of 64 channels ; lo plane of the same 64] -- the otherwise idle half of the M = 128
//        tile carries the lo plane, so TWO MMAs (x hi, x lo) give all four products; dW = lower + upper half.
//   Accumulation in TMEM over the CTA's pixel range; every CTA writes its fp32 partial tile to a workspace and a second kernel
//   adds the splits in a fixed order -> run-to-run identical gradients (round 1 used fp32 atomics).
#include <stdlib.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"
#include "conv.h"

namespace sg2 {
namespace wgp {
using namespace tc;

constexpr int CHUNK = 32;
constexpr int ABLK = CHUNK * 128;           // one gy box: 32 rows x 128 B
constexpr int XSLOT = 40 * 128;             // one x box incl. halo: <= 40 rows (8+2)*4; 5120 B keeps every slot 1024-aligned
constexpr int NTHREADS = 192;               // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

struct Params {
    float* ws;                // [split][tile][128][ncols] fp32 partial tiles
    int n, h, w, ci, co, k;
    int cw, ch, cb, chunks_x, chunks_y, total_chunks;
    int co_tiles, ci_tiles, stack;
    int chunks_per_split;
    int ncols;                // 64 * k
    int stages, stage_bytes, a_bytes;
    int two_term;             // experiment (SG2_GRAD_TERMS=2 / SG2_WGRAD_TERMS=2): x taken as its bf16 hi plane only (the x_lo products dropped)
};

template <int KW>
__global__ void __launch_bounds__(NTHREADS, 1) conv_wgrad_pl_kernel(const __grid_constant__ CUtensorMap gmap,
                                                                    const __grid_constant__ CUtensorMap xmap, const Params p) {
    constexpr int N = 64 * KW;
    constexpr uint32_t TMEM_COLS = N <= 64 ? 64 : 256;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + p.stages * p.stage_bytes;
    auto full = [&](int s) { return bar_base + 8u * s; };
    auto empty = [&](int s) { return bar_base + 64u + 8u * s; };
    const uint32_t acc_full = bar_base + 128u;
    const uint32_t tmem_slot = bar_base + 136u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int cot = tile % p.co_tiles;
    const int cit = (tile / p.co_tiles) % p.ci_tiles;
    const int dyi = tile / (p.co_tiles * p.ci_tiles);
    const int mt = p.stack ? 64 : 128;
    const int co0 = cot * mt, ci0 = cit * 64;
    const int pad = p.k >> 1;
    const int xw = p.cw + (KW - 1);
    const int xrows = xw * p.ch * p.cb;
    const int q_begin = blockIdx.y * p.chunks_per_split;
    const int q_end = min(p.total_chunks, q_begin + p.chunks_per_split);
    const int nchunks = max(0, q_end - q_begin);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_d;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot));

    if (warp == 0) {
        if (nchunks > 0 && elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&gmap) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            const int ablocks = p.stack ? 1 : 2;                        // 64-co blocks per plane
            const uint32_t tx = (uint32_t)(2 * ablocks * ABLK + 2 * xrows * 128);
            const int dy = (KW == 1) ? 0 : dyi - pad;
            for (int i = 0, s = 0, ph = 0; i < nchunks; ++i) {
                const int q = q_begin + i;
                const int cx0 = (q % p.chunks_x) * p.cw;
                const int cy0 = ((q / p.chunks_x) % p.chunks_y) * p.ch;
                const int cb0 = (q / (p.chunks_x * p.chunks_y)) * p.cb;
                mbar_wait(empty(s), ph ^ 1);
                mbar_expect_tx(full(s), tx);
                const uint32_t dst = base + s * p.stage_bytes;
                // A region: [plane][block] (stack: [hi block ; lo block]) -- in both cases plane-major, blocks of ABLK bytes
                for (int pl = 0; pl < 2; ++pl)
                    for (int j = 0; j < ablocks; ++j)
                        tma_load_5d(dst + (pl * ablocks + j) * ABLK, &gmap, full(s), co0 + 64 * j, cx0, cy0, cb0, pl);
                for (int pl = 0; pl < 2; ++pl)
                    tma_load_5d(dst + p.a_bytes + pl * XSLOT, &xmap, full(s), ci0, cx0 - (KW >> 1), cy0 + dy, cb0, pl);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (nchunks > 0 && elect_one()) {
            constexpr uint32_t idesc = idesc_bf16_mn(128, N);
            // B rows of k16 step kq: the 16 pixels of the step sit in one image row (cw >= 16) or in two (cw = 8)
            const uint32_t b_sbo = (p.cw >= 16) ? 1024u : (uint32_t)xw * 128u;
            uint32_t b_row[2];
            for (int kq = 0; kq < 2; ++kq) b_row[kq] = (uint32_t)(((16 * kq) / p.cw) * xw + (16 * kq) % p.cw) * 128u;
            const uint32_t b_lbo = KW == 1 ? 4096u : 128u;              // 64-wide MN blocks of B = the dx taps, one row apart
            for (int i = 0, s = 0, ph = 0; i < nchunks; ++i) {
                mbar_wait(full(s), ph);
                tc_fence_after();
                const uint32_t a0 = base + s * p.stage_bytes;
                const uint32_t bh = a0 + p.a_bytes, bl = bh + XSLOT;
#pragma unroll
                for (int kq = 0; kq < 2; ++kq) {
                    const uint32_t ko = kq * 16 * 128;
                    const uint64_t dbh = mnmajor_desc(bh + b_row[kq], b_lbo, b_sbo), dbl = mnmajor_desc(bl + b_row[kq], b_lbo, b_sbo);
                    if (p.stack) {
                        const uint64_t da = mnmajor_desc(a0 + ko, ABLK, 1024);             // rows [hi(64) ; lo(64)]
                        mma_bf16(tmem_d, da, dbh, idesc, (i | kq) != 0);
                        if (!p.two_term) mma_bf16(tmem_d, da, dbl, idesc, 1);
                    } else {
                        const uint64_t dah = mnmajor_desc(a0 + ko, ABLK, 1024), dal = mnmajor_desc(a0 + 2 * ABLK + ko, ABLK, 1024);
                        mma_bf16(tmem_d, dah, dbh, idesc, (i | kq) != 0);
                        mma_bf16(tmem_d, dal, dbh, idesc, 1);
                        if (!p.two_term) mma_bf16(tmem_d, dah, dbl, idesc, 1);
                    }
                }
                mma_commit(empty(s));
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            mma_commit(acc_full);
        }
    } else {
        // ---------------- epilogue: TMEM -> this CTA's partial tile in the workspace ----------------
        const int q4 = warp & 3;
        const int row = q4 * 32 + lane;
        float* out = p.ws + (((size_t)blockIdx.y * gridDim.x + tile) * 128 + row) * N;
        if (nchunks > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < N / 16; ++c) {
                uint32_t acc[16];
                tmem_ld16(tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(c * 16), acc);
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                    st4(out + c * 16 + e, make_float4(__uint_as_float(acc[e]), __uint_as_float(acc[e + 1]), __uint_as_float(acc[e + 2]), __uint_as_float(acc[e + 3])));
            }
            tc_fence_before();
        } else {
            for (int c = 0; c < N; c += 4) st4(out + c, f4zero());
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, TMEM_COLS);
    }
}

// dw[co][ci][k][k] (+)= coef * sum_splits ws[split][tile(dy, ci tile, co tile)][row][dx * 64 + ci % 64]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int co, int ci, int k,
                                                           int co_tiles, int ci_tiles, int stack, int splits, int tiles, float coef, int accumulate) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int kk2 = k * k;
    if (idx >= (long long)co * ci * kk2) return;
    // thread order (c_o, tap, c_i): consecutive threads read consecutive workspace columns (coalesced, `splits` times) and make one
    // scattered write each; the dw order (tap fastest) read 4 bytes out of every 256
    const int c_i = (int)(idx % ci);
    const int tap = (int)((idx / ci) % kk2), c_o = (int)(idx / ((long long)kk2 * ci));
    const long long oidx = ((long long)c_o * ci + c_i) * kk2 + tap;
    const int dyi = tap / k, dxi = tap % k;
    const int mt = stack ? 64 : 128;
    const int tile = (dyi * ci_tiles + c_i / 64) * co_tiles + c_o / mt;
    const int ncols = 64 * k;
    const size_t off = ((size_t)tile * 128 + c_o % mt) * ncols + dxi * 64 + c_i % 64;
    const size_t split_stride = (size_t)tiles * 128 * ncols;
    float s = 0.f;
    // the low-channel layers have few tiles and up to ~100 pixel splits: 8 splits of loads in flight (the sum keeps its order)
    const size_t lo_off = stack ? (size_t)64 * ncols : 0;
    int sp = 0;
    for (; sp + 8 <= splits; sp += 8) {
        float v[8], u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = __ldg(ws + (sp + j) * split_stride + off);
            u[j] = stack ? __ldg(ws + (sp + j) * split_stride + off + lo_off) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j] + u[j];
    }
    for (; sp < splits; ++sp) {
        float v = ws[sp * split_stride + off];
        if (stack) v += ws[sp * split_stride + off + lo_off];
        s += v;
    }
    s *= coef;
    dw[oidx] = accumulate ? dw[oidx] + s : s;
}

struct Plan { Params p; int tiles, splits, smem; bool ok; };

static Plan make_plan(int n, int h, int w, int ci, int co, int k) {
    Plan pl{};
    pl.ok = false;
    Params& p = pl.p;
    if ((k != 1 && k != 3) || ci % 32 != 0 || co % 32 != 0) return pl;      // boxes of 64 channels; TMA zero-fills past the last channel
    if (!pixel_box_ragged(CHUNK, h, w, p.cw, p.ch, p.cb)) return pl;
    if (p.cw < 8) return pl;                                   // an 8-row group of the operands must sit inside one image row
    if ((p.cw + k - 1) * p.ch * p.cb > 40) return pl;
    p.n = n; p.h = h; p.w = w; p.ci = ci; p.co = co; p.k = k;
    p.chunks_x = (w + p.cw - 1) / p.cw; p.chunks_y = (h + p.ch - 1) / p.ch;
    p.total_chunks = p.chunks_x * p.chunks_y * ((n + p.cb - 1) / p.cb);
    p.stack = (co % 128 != 0) ? 1 : 0;
    static int terms = 0;
    if (!terms) {
        const char* e = getenv("SG2_WGRAD_TERMS");
        if (!e) e = getenv("SG2_GRAD_TERMS");
        terms = e ? atoi(e) : 3;
    }
    p.two_term = terms == 2;
    p.co_tiles = (co + (p.stack ? 63 : 127)) / (p.stack ? 64 : 128);
    p.ci_tiles = (ci + 63) / 64;
    p.ncols = 64 * k;
    pl.tiles = p.co_tiles * p.ci_tiles * k;
    int splits = std::max(1, (2 * num_sms()) / pl.tiles);
    splits = std::min(splits, p.total_chunks);
    p.chunks_per_split = (p.total_chunks + splits - 1) / splits;
    pl.splits = (p.total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
    p.a_bytes = (p.stack ? 2 : 4) * ABLK;
    p.stage_bytes = p.a_bytes + 2 * XSLOT;
    p.stages = std::min(8, (227 * 1024 - 1024 - 256) / p.stage_bytes);
    pl.smem = 1024 + p.stages * p.stage_bytes + 256;
    pl.ok = true;
    return pl;
}

}  // namespace wgp

bool wgrad_pl_supported(int n, int h, int w, int ci, int co, int k) { return wgp::make_plan(n, h, w, ci, co, k).ok; }

long long wgrad_pl_workspace_bytes(int n, int h, int w, int ci, int co, int k) {
    const wgp::Plan pl = wgp::make_plan(n, h, w, ci, co, k);
    if (!pl.ok) return -1;
    return (long long)pl.splits * pl.tiles * 128 * pl.p.ncols * (long long)sizeof(float);
}

int conv_wgrad_pl(const void* x_planes, const void* gy_planes, float* dw, void* workspace, int n, int h, int w, int ci, int co, int k,
                  float coef, int accumulate, cudaStream_t st) {
    wgp::Plan pl = wgp::make_plan(n, h, w, ci, co, k);
    if (!pl.ok) return fail(SG2_ENOTSUP, "conv_wgrad_pl: unsupported shape n=%d h=%d w=%d ci=%d co=%d k=%d", n, h, w, ci, co, k);
    wgp::Params& p = pl.p;
    p.ws = (float*)workspace;
    CUtensorMap gmap, xmap;
    int rc = tc::make_planes_map(&gmap, gy_planes, n, h, w, co, p.cw, p.ch, p.cb, "conv_wgrad_pl(gy)");
    if (rc) return rc;
    rc = tc::make_planes_map(&xmap, x_planes, n, h, w, ci, p.cw + (k - 1), p.ch, p.cb, "conv_wgrad_pl(x)");
    if (rc) return rc;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgp::conv_wgrad_pl_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(wgp::conv_wgrad_pl_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "conv_wgrad_pl: cannot opt in to 227 KB of shared memory: %s", cudaGetErrorString(e));
        configured = true;
    }
    dim3 grid((unsigned)pl.tiles, (unsigned)pl.splits);
    if (k == 3) wgp::conv_wgrad_pl_kernel<3><<<grid, wgp::NTHREADS, pl.smem, st>>>(gmap, xmap, p);
    else wgp::conv_wgrad_pl_kernel<1><<<grid, wgp::NTHREADS, pl.smem, st>>>(gmap, xmap, p);
    rc = launched("conv_wgrad_pl");
    if (rc) return rc;
    const long long total = (long long)co * ci * k * k;
    wgp::wgrad_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(p.ws, dw, co, ci, k, p.co_tiles, p.ci_tiles, p.stack, pl.splits,
                                                                             pl.tiles, coef, accumulate);
    return launched("wgrad_reduce");
}

}  // namespace sg2
