// tcgen05 weight-gradient kernel (bf16x3), NHWC fp32 activations.
//
// Replaces the wgrad half of ATen's convolution_backward for the reference path (44 % of the reference's CPU step
// time together with dgrad) -- composed into autograd exactly like Conv2dGradWeight of
// thirdparty/stylegan3_ops/ops/conv2d_gradfix.py:147-187.
//
//   dW[co, tap, ci] = coef * sum_pix gy[pix, co] * x[pix + off(tap), ci]
// GEMM view: M = co (tile 128), N = (dx-tap, ci) (3 x 64 = 192 columns for k = 3, 64 for k = 1), K = pixels.
// Both operands come from NHWC tensors, i.e. the reduction index (pixels) is the STRIDED one: the operands are
// "MN-major" for the tensor core.  A K chunk is a box of 32 pixels; per chunk
//   TMA   : up to 4 boxes of gy (32 co each; boxes past `co` are skipped) + 2 boxes of x (32 ci each) that carry a
//           one-pixel halo in x (box width cw + k - 1, shifted by the tap row dy, OOB -> 0 = zero padding): the k taps of
//           a kernel row read the SAME pixels, so x crosses L2 -> SMEM and the fp32 -> bf16 split once, not k times
//   xform : 14 warps: optional per-sample scales (modulated layers), fp32 -> bf16 hi/lo, written as MN-major
//           SWIZZLE_128B tiles (rows = pixels, 128 B = 64 channels); an x element is stored into the k dx-shifted tiles
//   MMA   : 2 (k16) x 3 (hi*hi, lo*hi, hi*lo) tcgen05.mma 128 x N x 16 into one TMEM accumulator
// Grid: (co tiles x ci tiles x k tap-rows) x pixel splits; partial sums are reduced with fp32 red.global.add.
#include "tc_common.cuh"
#include "conv.h"

namespace sg2 {
namespace wg {
using namespace tc;

constexpr int CHUNK = 32;                  // pixels per pipeline stage
constexpr int BOXB = CHUNK * 128;          // one TMA box: 32 rows x 128 B
constexpr int STAGES = 3;                  // bf16 operand stages (transform -> MMA)
constexpr int FSTAGES = 4;                 // fp32 staging stages (TMA -> transform): deep, the L2 latency hides here
constexpr int XWARPS = 14;                 // transform warps: 448 threads >= 4*64 gy + 4*48 x items of the largest chunk
constexpr int NTHREADS = 64 + 32 * XWARPS;
constexpr int MT = 128, NT_CI = 64;
constexpr int XROWS_MAX = 48;              // rows of a halo'd x box: (cw + 2) * ch * cb <= 48 for every pixel_box(32, ...)
constexpr int XBOXB = XROWS_MAX * 128;     // 6144 B (a multiple of 1024: keeps the SWIZZLE_128B phase of every box)

template <int KW> struct Cfg {
    static constexpr int N = NT_CI * KW;                    // MMA N
    static constexpr int XB_MAX = KW == 1 ? BOXB : XBOXB;   // most bytes an x box takes (the launcher passes the real size)
    static constexpr int A_PLANE = 2 * BOXB;                // 128 co = 2 MN blocks
    static constexpr int B_PLANE = KW * BOXB;               // KW taps x 64 ci
    static constexpr int STAGE_BF = 2 * (A_PLANE + B_PLANE);
    static constexpr int smem(int xb, int fstages) { return 1024 + fstages * (4 * BOXB + 2 * xb) + STAGES * STAGE_BF + 256; }
    static constexpr uint32_t TMEM_COLS = N <= 64 ? 64 : 256;
};

struct Params {
    const float* in_scale;    // [n, ci] or null (scales x)
    const float* out_scale;   // [n, co] or null (scales gy)
    float* dw;                // [co, ci, k, k]
    float* ws;                // partial buffers [split][co, ci, k, k] (deterministic reduction) or null (atomics into dw)
    float coef;
    int n, h, w, ci, co, k;
    int cw, ch, cb, chunks_x, chunks_y, total_chunks;
    int co_tiles, ci_tiles;
    int chunks_per_split;
    int xb;                   // bytes reserved per x box in an fp32 stage (rows * 128 rounded up to 1024)
    int fstages;              // fp32 staging depth that fits next to the operand stages (3 or 4)
    long long* trace;         // sg2_debug_trace buffer or null
};

template <int KW>
__global__ void __launch_bounds__(NTHREADS, 1) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap gmap,
                                                                    const __grid_constant__ CUtensorMap xmap, const Params p) {
    using C = Cfg<KW>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t f32_base = base;
    const int stage_f32 = 4 * BOXB + 2 * p.xb;
    const int FS = p.fstages;
    const uint32_t bf_base = f32_base + FS * stage_f32;
    const uint32_t bar_base = bf_base + STAGES * C::STAGE_BF;
    auto f_full = [&](int s) { return bar_base + 8u * s; };                   // FSTAGES (<= 4)
    auto f_empty = [&](int s) { return bar_base + 32u + 8u * s; };            // FSTAGES
    auto ab_full = [&](int s) { return bar_base + 64u + 8u * s; };            // STAGES (<= 4)
    auto ab_empty = [&](int s) { return bar_base + 96u + 8u * s; };           // STAGES
    const uint32_t acc_full = bar_base + 128u;
    const uint32_t tmem_slot = bar_base + 136u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool tr = p.trace != nullptr;
    long long* trow = p.trace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16;
    const long long t_begin = tr ? clock64() : 0;
    long long w0 = 0, w1 = 0;
    // tile: blockIdx.x = (dyi * ci_tiles + cit) * co_tiles + cot
    const int cot = blockIdx.x % p.co_tiles;
    const int cit = (blockIdx.x / p.co_tiles) % p.ci_tiles;
    const int dyi = blockIdx.x / (p.co_tiles * p.ci_tiles);
    const int co0 = cot * MT, ci0 = cit * NT_CI;
    const int pad = p.k >> 1;
    const int gy_boxes = min(4, (p.co - co0 + 31) / 32);      // gy boxes that hold real channels
    const int xw = p.cw + (KW - 1);                           // x box width incl. halo
    const int xrows = xw * p.ch * p.cb;                       // rows of one x box
    const int q_begin = blockIdx.y * p.chunks_per_split;
    const int q_end = min(p.total_chunks, q_begin + p.chunks_per_split);
    const int nchunks = max(0, q_end - q_begin);

    if (threadIdx.x == 0) {
        for (int s = 0; s < FS; ++s) { mbar_init(f_full(s), 1); mbar_init(f_empty(s), XWARPS); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(ab_full(s), XWARPS); mbar_init(ab_empty(s), 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_d;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot));

    auto chunk_origin = [&](int q, int& cx0, int& cy0, int& cb0) {
        cx0 = (q % p.chunks_x) * p.cw;
        cy0 = ((q / p.chunks_x) % p.chunks_y) * p.ch;
        cb0 = (q / (p.chunks_x * p.chunks_y)) * p.cb;
    };

    if (warp == 0) {
        if (nchunks > 0 && elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&gmap) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            const uint32_t stage_bytes = (uint32_t)(gy_boxes * BOXB + 2 * xrows * 128);
            for (int i = 0, s = 0, ph = 0; i < nchunks; ++i) {
                int cx0, cy0, cb0;
                chunk_origin(q_begin + i, cx0, cy0, cb0);
                mbar_wait_t(f_empty(s), ph ^ 1, tr, w0);
                mbar_expect_tx(f_full(s), stage_bytes);
                const uint32_t dst = f32_base + s * stage_f32;
                for (int j = 0; j < gy_boxes; ++j) tma_load_4d(dst + j * BOXB, &gmap, f_full(s), co0 + 32 * j, cx0, cy0, cb0);
                const int dy = (KW == 1) ? 0 : dyi - pad;
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    tma_load_4d(dst + 4 * BOXB + j * p.xb, &xmap, f_full(s), ci0 + 32 * j, cx0 - (KW >> 1), cy0 + dy, cb0);
                if (++s == FS) { s = 0; ph ^= 1; }
            }
            if (tr) trow[0] = w0;
        }
    } else if (warp == 1) {
        if (nchunks > 0 && elect_one()) {
            constexpr uint32_t idesc = idesc_bf16_mn(MT, C::N);
            for (int i = 0, s = 0, ph = 0; i < nchunks; ++i) {
                mbar_wait_t(ab_full(s), ph, tr, w0);
                tc_fence_after();
                const uint32_t a_hi = bf_base + s * C::STAGE_BF, a_lo = a_hi + C::A_PLANE;
                const uint32_t b_hi = a_lo + C::A_PLANE, b_lo = b_hi + C::B_PLANE;
#pragma unroll
                for (int kq = 0; kq < CHUNK / 16; ++kq) {
                    const uint32_t ko = kq * 16 * 128;           // 16 pixel rows
                    const uint64_t dah = mnmajor_desc(a_hi + ko, BOXB, 1024), dal = mnmajor_desc(a_lo + ko, BOXB, 1024);
                    const uint64_t dbh = mnmajor_desc(b_hi + ko, BOXB, 1024), dbl = mnmajor_desc(b_lo + ko, BOXB, 1024);
                    mma_bf16(tmem_d, dah, dbh, idesc, (i | kq) != 0);
                    mma_bf16(tmem_d, dal, dbh, idesc, 1);
                    mma_bf16(tmem_d, dah, dbl, idesc, 1);
                }
                mma_commit(ab_empty(s));
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            mma_commit(acc_full);
            if (tr) { trow[4] = w0; trow[7] = clock64() - t_begin; }
        }
    } else {
        const int tt = threadIdx.x - 64;                 // 0 .. 32*XWARPS-1
        // One item = 16 channels of one box row; a chunk has gy_boxes*64 + 4*xrows <= 448 of them, so every thread owns at
        // most ONE item whose geometry is the same for all chunks: decode it (the integer divisions) once, up front.
        const int gy_items = gy_boxes * 64;              // (box, 16-channel half, pixel row)
        const int x_items = 2 * 2 * xrows;               // (box, half, halo'd row)
        const bool active = tt < gy_items + x_items;
        const bool is_x = tt >= gy_items;
        // everything below is relative to the stage bases: src_off into the fp32 stage, dst[e][q] into the bf16 stage (hi
        // plane; the lo plane is lo_delta further), with the SWIZZLE_128B chunk permutation already applied
        uint32_t src_off[4] = {0, 0, 0, 0}, dst[3][2] = {{0, 0}, {0, 0}, {0, 0}}, lo_delta = 0;
        int ndst = 0, pb_rel = 0, sc_c = -1;
        if (active) {
            int j, r, half;
            uint32_t src_row;
            if (!is_x) {
                j = tt >> 6; half = (tt >> 5) & 1; r = tt & 31;
                src_row = j * BOXB + r * 128;
                const int sub = j & 1;                     // 32-channel sub-block inside the 64-wide MN block
                for (int q = 0; q < 2; ++q) dst[0][q] = (j >> 1) * BOXB + r * 128 + (((4 * sub + 2 * half + q) ^ (r & 7)) << 4);
                ndst = 1;
                lo_delta = C::A_PLANE;
                pb_rel = (r / p.cw) / p.ch;
                const int c = co0 + 32 * j + 16 * half;
                if (p.out_scale && c < p.co) sc_c = c;
            } else {
                const int it = tt - gy_items;
                j = it / (2 * xrows); half = (it / xrows) & 1; r = it % xrows;
                src_row = 4 * BOXB + j * p.xb + r * 128;
                const int prow = r / xw, pcol = r % xw;
                pb_rel = prow / p.ch;
                // tap dx reads pixel (pcol - dx) of the chunk row: one destination per dx tile whose window holds this pixel
                for (int dxi = 0; dxi < KW; ++dxi) {
                    const int px = pcol - dxi;
                    if (px < 0 || px >= p.cw) continue;
                    const int dr = prow * p.cw + px;
                    for (int q = 0; q < 2; ++q)
                        dst[ndst][q] = 2 * C::A_PLANE + dxi * BOXB + dr * 128 + (((4 * j + 2 * half + q) ^ (dr & 7)) << 4);
                    ++ndst;
                }
                lo_delta = C::B_PLANE;
                const int c = ci0 + 32 * j + 16 * half;
                if (p.in_scale && c < p.ci) sc_c = c;
            }
            for (int q = 0; q < 4; ++q) src_off[q] = src_row + (((4 * half + q) ^ (r & 7)) << 4);
        }
        const int chunks_per_img = p.chunks_x * p.chunks_y;
        for (int i = 0, fs = 0, fph = 0, s = 0, sph = 0; i < nchunks; ++i) {
            mbar_wait_t(f_full(fs), fph, tr, w0);
            mbar_wait_t(ab_empty(s), sph ^ 1, tr, w1);
            if (active) {
                const uint32_t src = f32_base + fs * stage_f32;
                const uint32_t bf = bf_base + s * C::STAGE_BF;
                float v[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t4 = lds4(src + src_off[q]);
                    v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
                }
                if (sc_c >= 0) {
                    const int pb = ((q_begin + i) / chunks_per_img) * p.cb + pb_rel;
                    if (pb < p.n) {
                        const float* sc = is_x ? p.in_scale + (long long)pb * p.ci + sc_c : p.out_scale + (long long)pb * p.co + sc_c;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 t4 = ldg4(sc + 4 * q);
                            v[4 * q] *= t4.x; v[4 * q + 1] *= t4.y; v[4 * q + 2] *= t4.z; v[4 * q + 3] *= t4.w;
                        }
                    }
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) split2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    if (e < ndst) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            sts4(bf + dst[e][q], hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                            sts4(bf + dst[e][q] + lo_delta, lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
                        }
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(ab_full(s)); mbar_arrive(f_empty(fs)); }
            if (++fs == FS) { fs = 0; fph ^= 1; }
            if (++s == STAGES) { s = 0; sph ^= 1; }
        }
        if (tr && tt == 0) { trow[1] = w0; trow[2] = w1; trow[3] = clock64() - t_begin; }
        // ---------------- epilogue: TMEM -> fp32 atomics into dw[co][ci][k][k] ----------------
        if (nchunks > 0 && warp < 10) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int q4 = warp & 3;
            const int co = co0 + q4 * 32 + lane;
            const int chalf = (warp - 2) >> 2;
            const int kk2 = p.k * p.k;
            constexpr int HALF = C::N / 2;
#pragma unroll 1
            for (int c = 0; c < HALF / 16; ++c) {
                const int col0 = chalf * HALF + c * 16;
                uint32_t acc[16];
                tmem_ld16(tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)col0, acc);
                if (co < p.co) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int col = col0 + e;
                        const int dxi = col / NT_CI, ci = ci0 + col % NT_CI;
                        const int tap = (KW == 1) ? 0 : dyi * 3 + dxi;
                        if (ci < p.ci) {
                            const long long o = ((long long)co * p.ci + ci) * kk2 + tap;
                            if (p.ws) p.ws[(long long)blockIdx.y * ((long long)p.co * p.ci * kk2) + o] = __uint_as_float(acc[e]) * p.coef;
                            else atomicAdd(p.dw + o, __uint_as_float(acc[e]) * p.coef);
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (tr && threadIdx.x == 0) { trow[11] = clock64() - t_begin; trow[12] = nchunks; }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, C::TMEM_COLS);
    }
}

constexpr int SMEM_LIMIT = 227 * 1024;

template <int KW>
static int launch(const CUtensorMap& gmap, const CUtensorMap& xmap, Params& p, dim3 grid, cudaStream_t st) {
    using C = Cfg<KW>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "conv_wgrad_tc: cannot opt in to %d B of shared memory: %s", SMEM_LIMIT, cudaGetErrorString(e));
        configured = true;
    }
    p.fstages = FSTAGES;
    while (p.fstages > 2 && C::smem(p.xb, p.fstages) > SMEM_LIMIT) --p.fstages;
    conv_wgrad_tc_kernel<KW><<<grid, NTHREADS, C::smem(p.xb, p.fstages), st>>>(gmap, xmap, p);
    return launched("conv_wgrad_tc");
}

}  // namespace wg

bool wgrad_tc_supported(int n, int h, int w, int ci, int co, int k) {
    (void)n;
    if (k != 1 && k != 3) return false;
    if (ci % 32 != 0 || co % 32 != 0) return false;
    int a, b, c;
    return tc::pixel_box_ragged(wg::CHUNK, h, w, a, b, c);
}

static int tc_wgrad_plan(const WgradParams& wp, wg::Params& p, int& tiles, int& splits) {
    if (!tc::pixel_box_ragged(wg::CHUNK, wp.h, wp.w, p.cw, p.ch, p.cb)) return fail(SG2_ENOTSUP, "conv_wgrad_tc: unsupported shape");
    if (4 * 64 + 4 * (p.cw + wp.k - 1) * p.ch * p.cb > 32 * wg::XWARPS || (p.cw + wp.k - 1) * p.ch * p.cb > wg::XROWS_MAX)
        return fail(SG2_ENOTSUP, "conv_wgrad_tc: chunk %dx%dx%d does not fit the transform", p.cw, p.ch, p.cb);
    p.chunks_x = (wp.w + p.cw - 1) / p.cw; p.chunks_y = (wp.h + p.ch - 1) / p.ch;      // ragged: the last chunk of a row overhangs
    p.total_chunks = p.chunks_x * p.chunks_y * ((wp.n + p.cb - 1) / p.cb);
    p.co_tiles = (wp.co + wg::MT - 1) / wg::MT;
    p.ci_tiles = (wp.ci + wg::NT_CI - 1) / wg::NT_CI;
    tiles = p.co_tiles * p.ci_tiles * wp.k;
    splits = std::max(1, (2 * num_sms()) / tiles);
    splits = std::min(splits, p.total_chunks);
    p.chunks_per_split = (p.total_chunks + splits - 1) / splits;
    splits = (p.total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
    return SG2_OK;
}

int wgrad_parts_tc(const WgradParams& wp) {
    wg::Params p;
    int tiles, splits;
    return tc_wgrad_plan(wp, p, tiles, splits) == SG2_OK ? splits : 0;
}

int conv_wgrad_tc(WgradParams wp, int accumulate, cudaStream_t st) {
    wg::Params p;
    int tiles, splits;
    int rc = tc_wgrad_plan(wp, p, tiles, splits);
    if (rc) return rc;
    const int kk2 = wp.k * wp.k;
    if (!accumulate && !wp.ws) {
        cudaError_t e = cudaMemsetAsync(wp.dw, 0, sizeof(float) * (size_t)wp.co * wp.ci * kk2, st);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "conv_wgrad_tc: memset: %s", cudaGetErrorString(e));
    }
    CUtensorMap gmap, xmap;
    rc = tc::make_nhwc_map(&gmap, wp.gy, wp.n, wp.h, wp.w, wp.co, p.cw, p.ch, p.cb, "conv_wgrad_tc(gy)");
    if (rc) return rc;
    rc = tc::make_nhwc_map(&xmap, wp.x, wp.n, wp.h, wp.w, wp.ci, p.cw + (wp.k - 1), p.ch, p.cb, "conv_wgrad_tc(x)");   // x halo
    if (rc) return rc;
    p.in_scale = wp.in_scale; p.out_scale = wp.out_scale; p.dw = wp.dw; p.ws = wp.ws; p.coef = wp.coef; p.trace = g_trace;
    p.n = wp.n; p.h = wp.h; p.w = wp.w; p.ci = wp.ci; p.co = wp.co; p.k = wp.k;
    p.xb = (((p.cw + wp.k - 1) * p.ch * p.cb * 128) + 1023) / 1024 * 1024;
    dim3 grid((unsigned)tiles, (unsigned)splits);
    if (wp.k == 3) return wg::launch<3>(gmap, xmap, p, grid, st);
    return wg::launch<1>(gmap, xmap, p, grid, st);
}

}  // namespace sg2
