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
//   TMA   : 4 boxes of gy (32 co each) + k x 2 boxes of x (32 ci each, shifted by the tap, OOB -> 0 = zero padding)
//   xform : 8 warps: optional per-sample scales (modulated layers), fp32 -> bf16 hi/lo, written as MN-major
//           SWIZZLE_128B tiles (rows = pixels, 128 B = 64 channels)
//   MMA   : 2 (k16) x 3 (hi*hi, lo*hi, hi*lo) tcgen05.mma 128 x N x 16 into one TMEM accumulator
// Grid: (co tiles x ci tiles x k tap-rows) x pixel splits; partial sums are reduced with fp32 red.global.add.
#include "tc_common.cuh"
#include "conv.h"

namespace sg2 {
namespace wg {
using namespace tc;

constexpr int CHUNK = 32;                  // pixels per pipeline stage
constexpr int BOXB = CHUNK * 128;          // one TMA box: 32 rows x 128 B
constexpr int STAGES = 2;
constexpr int NTHREADS = 320;
constexpr int MT = 128, NT_CI = 64;

template <int KW> struct Cfg {
    static constexpr int NBOX = 4 + 2 * KW;                 // gy boxes + x boxes
    static constexpr int N = NT_CI * KW;                    // MMA N
    static constexpr int STAGE_F32 = NBOX * BOXB;
    static constexpr int A_PLANE = 2 * BOXB;                // 128 co = 2 MN blocks
    static constexpr int B_PLANE = KW * BOXB;               // KW taps x 64 ci
    static constexpr int STAGE_BF = 2 * (A_PLANE + B_PLANE);
    static constexpr int SMEM = 1024 + STAGES * (STAGE_F32 + STAGE_BF) + 256;
    static constexpr uint32_t TMEM_COLS = N <= 64 ? 64 : 256;
};

struct Params {
    const float* in_scale;    // [n, ci] or null (scales x)
    const float* out_scale;   // [n, co] or null (scales gy)
    float* dw;                // [co, ci, k, k]
    float coef;
    int n, h, w, ci, co, k;
    int cw, ch, cb, chunks_x, chunks_y, total_chunks;
    int co_tiles, ci_tiles;
    int chunks_per_split;
};

template <int KW>
__global__ void __launch_bounds__(NTHREADS, 1) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap gmap,
                                                                    const __grid_constant__ CUtensorMap xmap, const Params p) {
    using C = Cfg<KW>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t f32_base = base;
    const uint32_t bf_base = f32_base + STAGES * C::STAGE_F32;
    const uint32_t bar_base = bf_base + STAGES * C::STAGE_BF;
    auto f_full = [&](int s) { return bar_base + 8u * s; };
    auto f_empty = [&](int s) { return bar_base + 16u + 8u * s; };
    auto ab_full = [&](int s) { return bar_base + 32u + 8u * s; };
    auto ab_empty = [&](int s) { return bar_base + 48u + 8u * s; };
    const uint32_t acc_full = bar_base + 64u;
    const uint32_t tmem_slot = bar_base + 72u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile: blockIdx.x = (dyi * ci_tiles + cit) * co_tiles + cot
    const int cot = blockIdx.x % p.co_tiles;
    const int cit = (blockIdx.x / p.co_tiles) % p.ci_tiles;
    const int dyi = blockIdx.x / (p.co_tiles * p.ci_tiles);
    const int co0 = cot * MT, ci0 = cit * NT_CI;
    const int pad = p.k >> 1;
    const int q_begin = blockIdx.y * p.chunks_per_split;
    const int q_end = min(p.total_chunks, q_begin + p.chunks_per_split);
    const int nchunks = max(0, q_end - q_begin);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(f_full(s), 1); mbar_init(f_empty(s), 8); mbar_init(ab_full(s), 8); mbar_init(ab_empty(s), 1);
        }
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
        if (lane == 0 && nchunks > 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&gmap) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            for (int i = 0; i < nchunks; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                int cx0, cy0, cb0;
                chunk_origin(q_begin + i, cx0, cy0, cb0);
                mbar_wait(f_empty(s), ph ^ 1);
                mbar_expect_tx(f_full(s), C::STAGE_F32);
                const uint32_t dst = f32_base + s * C::STAGE_F32;
#pragma unroll
                for (int j = 0; j < 4; ++j) tma_load_4d(dst + j * BOXB, &gmap, f_full(s), co0 + 32 * j, cx0, cy0, cb0);
#pragma unroll
                for (int j = 0; j < 2 * KW; ++j) {
                    const int dxi = j >> 1;
                    const int dx = (KW == 1) ? 0 : dxi - pad;
                    const int dy = (KW == 1) ? 0 : dyi - pad;
                    tma_load_4d(dst + (4 + j) * BOXB, &xmap, f_full(s), ci0 + 32 * (j & 1), cx0 + dx, cy0 + dy, cb0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nchunks > 0) {
            constexpr uint32_t idesc = idesc_bf16_mn(MT, C::N);
            for (int i = 0; i < nchunks; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(ab_full(s), ph);
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
            }
            mma_commit(acc_full);
        }
    } else {
        const int tt = threadIdx.x - 64;                 // 0..255
        for (int i = 0; i < nchunks; ++i) {
            const int s = i % STAGES;
            const uint32_t ph = (i / STAGES) & 1;
            int cx0, cy0, cb0;
            chunk_origin(q_begin + i, cx0, cy0, cb0);
            mbar_wait(f_full(s), ph);
            mbar_wait(ab_empty(s), ph ^ 1);
            const uint32_t src0 = f32_base + s * C::STAGE_F32;
            const uint32_t a_hi = bf_base + s * C::STAGE_BF, a_lo = a_hi + C::A_PLANE;
            const uint32_t b_hi = a_lo + C::A_PLANE, b_lo = b_hi + C::B_PLANE;
            // items: (box j, half, row r); a warp covers 32 rows of one (box, half)
            for (int item = tt; item < C::NBOX * 64; item += 256) {
                const int j = item >> 6, half = (item >> 5) & 1, r = item & 31;
                const int sw = r & 7;
                const int pb = cb0 + r / (p.cw * p.ch);
                const uint32_t src = src0 + j * BOXB + r * 128;
                float v[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t4 = lds4(src + (((4 * half + q) ^ sw) << 4));
                    v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
                }
                const float* sc = nullptr;
                uint32_t dst_hi, dst_lo;
                int sub;                                   // 32-channel sub-block inside the 64-wide MN block
                if (j < 4) {
                    sub = j & 1;
                    dst_hi = a_hi + (j >> 1) * BOXB + r * 128; dst_lo = a_lo + (j >> 1) * BOXB + r * 128;
                    const int c = co0 + 32 * j + 16 * half;
                    if (p.out_scale && pb < p.n && c < p.co) sc = p.out_scale + (long long)pb * p.co + c;
                } else {
                    const int jj = j - 4;
                    sub = jj & 1;
                    dst_hi = b_hi + (jj >> 1) * BOXB + r * 128; dst_lo = b_lo + (jj >> 1) * BOXB + r * 128;
                    const int c = ci0 + 32 * sub + 16 * half;
                    if (p.in_scale && pb < p.n && c < p.ci) sc = p.in_scale + (long long)pb * p.ci + c;
                }
                if (sc) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 t4 = ldg4(sc + 4 * q);
                        v[4 * q] *= t4.x; v[4 * q + 1] *= t4.y; v[4 * q + 2] *= t4.z; v[4 * q + 3] *= t4.w;
                    }
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) split2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint32_t off = (uint32_t)(((4 * sub + 2 * half + q) ^ sw) << 4);
                    sts4(dst_hi + off, hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                    sts4(dst_lo + off, lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(ab_full(s)); mbar_arrive(f_empty(s)); }
        }
        // ---------------- epilogue: TMEM -> fp32 atomics into dw[co][ci][k][k] ----------------
        if (nchunks > 0) {
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
                        if (ci < p.ci) atomicAdd(p.dw + ((long long)co * p.ci + ci) * kk2 + tap, __uint_as_float(acc[e]) * p.coef);
                    }
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, C::TMEM_COLS);
    }
}

template <int KW>
static int launch(const CUtensorMap& gmap, const CUtensorMap& xmap, const Params& p, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<KW>::SMEM);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "conv_wgrad_tc: cannot opt in to %d B of shared memory: %s", Cfg<KW>::SMEM, cudaGetErrorString(e));
        configured = true;
    }
    conv_wgrad_tc_kernel<KW><<<grid, NTHREADS, Cfg<KW>::SMEM, st>>>(gmap, xmap, p);
    return launched("conv_wgrad_tc");
}

}  // namespace wg

bool wgrad_tc_supported(int n, int h, int w, int ci, int co, int k) {
    (void)n;
    if (k != 1 && k != 3) return false;
    if (ci % 32 != 0 || co % 32 != 0) return false;
    int a, b, c;
    return tc::pixel_box(wg::CHUNK, h, w, a, b, c);
}

int conv_wgrad_tc(WgradParams wp, int accumulate, cudaStream_t st) {
    wg::Params p;
    if (!tc::pixel_box(wg::CHUNK, wp.h, wp.w, p.cw, p.ch, p.cb)) return fail(SG2_ENOTSUP, "conv_wgrad_tc: unsupported shape");
    const int kk2 = wp.k * wp.k;
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(wp.dw, 0, sizeof(float) * (size_t)wp.co * wp.ci * kk2, st);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "conv_wgrad_tc: memset: %s", cudaGetErrorString(e));
    }
    CUtensorMap gmap, xmap;
    int rc = tc::make_nhwc_map(&gmap, wp.gy, wp.n, wp.h, wp.w, wp.co, p.cw, p.ch, p.cb, "conv_wgrad_tc(gy)");
    if (rc) return rc;
    rc = tc::make_nhwc_map(&xmap, wp.x, wp.n, wp.h, wp.w, wp.ci, p.cw, p.ch, p.cb, "conv_wgrad_tc(x)");
    if (rc) return rc;
    p.in_scale = wp.in_scale; p.out_scale = wp.out_scale; p.dw = wp.dw; p.coef = wp.coef;
    p.n = wp.n; p.h = wp.h; p.w = wp.w; p.ci = wp.ci; p.co = wp.co; p.k = wp.k;
    p.chunks_x = wp.w / p.cw; p.chunks_y = wp.h / p.ch;
    p.total_chunks = p.chunks_x * p.chunks_y * ((wp.n + p.cb - 1) / p.cb);
    p.co_tiles = (wp.co + wg::MT - 1) / wg::MT;
    p.ci_tiles = (wp.ci + wg::NT_CI - 1) / wg::NT_CI;
    const int tiles = p.co_tiles * p.ci_tiles * wp.k;
    int splits = std::max(1, (2 * num_sms()) / tiles);
    splits = std::min(splits, p.total_chunks);
    p.chunks_per_split = (p.total_chunks + splits - 1) / splits;
    splits = (p.total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
    dim3 grid((unsigned)tiles, (unsigned)splits);
    if (wp.k == 3) return wg::launch<3>(gmap, xmap, p, grid, st);
    return wg::launch<1>(gmap, xmap, p, grid, st);
}

}  // namespace sg2
