// tcgen05 implicit-GEMM convolution, fp32-class precision ("tf32x3 + promotion"): the FORWARD convolutions.
//
// Why a second precision.  Every conv of the reference path feeds a leaky-ReLU (implementations/StyleGAN2/model.py:164,
// 193).  A relative error eps in a layer's input flips the sign of ~0.8*eps of its pre-activations w.r.t. the fp32
// reference, and each flip changes that element's gradient by 80 %: parameter-gradient parity degrades like
// sqrt(eps).  bf16x3 (conv_tc.cu, eps ~ 5e-6 per layer) is ample for data/weight gradients, which are linear in
// the operands, but forward activations must stay at fp32 level (eps ~ 2e-7) for gradients to land within 1e-3.
// Two measured error sources are removed here:
//   1. operand representation: v = big + small, big = rna_tf32(v), small = rna_tf32(v - big)  (22+ mantissa bits);
//      D += big*big + small*big + big*small                                       (3 kind::tf32 MMAs per K=8)
//   2. accumulation: the tensor core truncates the fp32 accumulator on every MMA (measured ~2e-8 relative per
//      instruction, linear in the chain length).  The K loop is therefore cut into segments of SEG sub-blocks
//      (8 big*big MMAs); each segment accumulates from zero into one of two TMEM buffers and is then added, with
//      round-to-nearest fp32 adds, into registers of 8 "promotion" warps while the next segment runs.
//
// Same GEMM view as conv_tc.cu.  One pipeline stage = ONE sub-block of 32 input channels of one tap:
//   A_big   : the TMA fp32 tile itself (128 pixel rows x 128 B, SWIZZLE_128B), rounded (and style-scaled) IN PLACE
//   A_small : written by the transform warps, same layout
//   B_big/B_small : pre-packed, pre-swizzled fp32 weight tiles, one cp.async.bulk per stage
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..9 = transform, 10..17 = promotion + epilogue.
#include "tc_common.cuh"
#include "conv.h"

namespace sg2 {
namespace tc32 {
using namespace tc;

constexpr int BM = 128, SUB = 32, STAGES = 3, NTHREADS = 576, SEG = 2;
constexpr int TILE_A = BM * 128;                                     // 16 KB
__host__ __device__ constexpr int tile_b(int bn) { return bn * 128; }
__host__ __device__ constexpr int stage_bytes(int bn) { return 2 * TILE_A + 2 * tile_b(bn); }
__host__ __device__ constexpr int smem_bytes(int bn) { return 1024 + STAGES * stage_bytes(bn) + 256; }

struct Params {
    const float* in_scale; const float* out_scale; const float* bias; const float* noise;
    float* y;
    long long ys[4];
    const unsigned char* wp;
    int n, h, w, ci, co, k;
    int tw, th, tb, tiles_x, tiles_y, m_tiles, n_tiles;
    int subs, spb;
    int act;
    float alpha, gain;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) conv_fwd_tc32_kernel(const __grid_constant__ CUtensorMap xmap, const Params p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + STAGES * stage_bytes(BN);
    auto st_a_big = [&](int s) { return base + s * stage_bytes(BN); };
    auto st_a_small = [&](int s) { return base + s * stage_bytes(BN) + TILE_A; };
    auto st_b_big = [&](int s) { return base + s * stage_bytes(BN) + 2 * TILE_A; };
    auto st_b_small = [&](int s) { return base + s * stage_bytes(BN) + 2 * TILE_A + tile_b(BN); };
    auto f_full = [&](int s) { return bar_base + 8u * s; };
    auto b_full = [&](int s) { return bar_base + 24u + 8u * s; };
    auto a_full = [&](int s) { return bar_base + 48u + 8u * s; };
    auto empty = [&](int s) { return bar_base + 72u + 8u * s; };
    auto acc_full = [&](int b) { return bar_base + 96u + 8u * b; };
    auto acc_empty = [&](int b) { return bar_base + 112u + 8u * b; };
    const uint32_t tmem_slot = bar_base + 128u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pad = p.k >> 1;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;          // two accumulator buffers
    const int nseg = (p.subs + SEG - 1) / SEG;
    // persistent tile loop: tile = n_tile * m_tiles + m_tile
    const int total_tiles = p.m_tiles * p.n_tiles;
    auto tile_coords = [&](int tile, int& x0, int& y0, int& b0, int& n0) {
        const int nt = tile / p.m_tiles, mt = tile % p.m_tiles;
        x0 = (mt % p.tiles_x) * p.tw;
        y0 = ((mt / p.tiles_x) % p.tiles_y) * p.th;
        b0 = (mt / (p.tiles_x * p.tiles_y)) * p.tb;
        n0 = nt * BN;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(f_full(s), 1); mbar_init(b_full(s), 1); mbar_init(a_full(s), 8); mbar_init(empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_d;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot));

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int x0, y0, b0, n0;
                tile_coords(tile, x0, y0, b0, n0);
                const unsigned char* wsrc = p.wp + (size_t)(n0 / BN) * p.subs * (2 * tile_b(BN));
                for (int t = 0; t < p.subs; ++t, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(empty(s), ph ^ 1);
                    mbar_expect_tx(b_full(s), 2 * tile_b(BN));
                    bulk_load(st_b_big(s), wsrc + (size_t)t * (2 * tile_b(BN)), 2 * tile_b(BN), b_full(s));
                    const int tap = t / p.spb, c0 = (t % p.spb) * SUB;
                    const int dy = tap / p.k - pad, dx = tap % p.k - pad;
                    mbar_expect_tx(f_full(s), TILE_A);
                    tma_load_4d(st_a_big(s), &xmap, f_full(s), c0, x0 + dx, y0 + dy, b0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_tf32(BM, BN);
            int it = 0, sg = 0;                      // global stage / segment counters (continue across tiles)
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int t = 0; t < p.subs; ++t, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const int buf = sg & 1;
                    const bool seg_start = (t % SEG) == 0;
                    if (seg_start) { mbar_wait(acc_empty(buf), ((sg >> 1) & 1) ^ 1); tc_fence_after(); }
                    mbar_wait(b_full(s), ph);
                    mbar_wait(a_full(s), ph);
                    tc_fence_after();
                    const uint32_t d = tmem_d + (uint32_t)(buf * BN);
#pragma unroll
                    for (int kq = 0; kq < SUB / 8; ++kq) {
                        const uint64_t dab = kmajor_desc(st_a_big(s) + kq * 32), das = kmajor_desc(st_a_small(s) + kq * 32);
                        const uint64_t dbb = kmajor_desc(st_b_big(s) + kq * 32), dbs = kmajor_desc(st_b_small(s) + kq * 32);
                        mma_tf32(d, dab, dbb, idesc, !(seg_start && kq == 0));
                        mma_tf32(d, das, dbb, idesc, 1);
                        mma_tf32(d, dab, dbs, idesc, 1);
                    }
                    mma_commit(empty(s));
                    if ((t % SEG) == SEG - 1 || t == p.subs - 1) { mma_commit(acc_full(buf)); ++sg; }
                }
            }
        }
    } else if (warp < 10) {
        // ================= transform: 2 threads per pixel row, 4 x 16 B chunks (16 channels) each =================
        const int tt = threadIdx.x - 64;
        const int r = tt & 127, half = tt >> 7;
        const int sw = r & 7;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int x0, y0, b0, n0;
            tile_coords(tile, x0, y0, b0, n0);
            const int pb = b0 + r / (p.tw * p.th);
            const bool row_ok = pb < p.n;
            for (int t = 0; t < p.subs; ++t, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(f_full(s), ph);
                const uint32_t row_big = st_a_big(s) + r * 128, row_small = st_a_small(s) + r * 128;
                const float* sp = (p.in_scale && row_ok) ? p.in_scale + (long long)pb * p.ci + (t % p.spb) * SUB + 16 * half : nullptr;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t off = (uint32_t)(((4 * half + q) ^ sw) << 4);
                    float4 v = lds4(row_big + off);
                    if (sp) v = mul4(v, ldg4(sp + 4 * q));
                    const uint32_t g0 = rna_tf32(v.x), g1 = rna_tf32(v.y), g2 = rna_tf32(v.z), g3 = rna_tf32(v.w);
                    sts4(row_big + off, g0, g1, g2, g3);
                    sts4(row_small + off, rna_tf32(v.x - __uint_as_float(g0)), rna_tf32(v.y - __uint_as_float(g1)),
                         rna_tf32(v.z - __uint_as_float(g2)), rna_tf32(v.w - __uint_as_float(g3)));
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full(s));
            }
        }
    } else {
        // ================= promotion (fp32 RN adds of the segment sums) + epilogue =================
        const int q4 = warp & 3;                    // TMEM lane quarter this warp may read
        const int chalf = (warp - 10) >> 2;         // which half of the BN columns
        constexpr int HALF = BN / 2;
        const int er = q4 * 32 + lane;
        int sg = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int x0, y0, b0, n0;
            tile_coords(tile, x0, y0, b0, n0);
            float racc[HALF];
#pragma unroll
            for (int j = 0; j < HALF; ++j) racc[j] = 0.f;
            for (int seg = 0; seg < nseg; ++seg, ++sg) {
                const int buf = sg & 1;
                mbar_wait(acc_full(buf), (sg >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < HALF / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld16(tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * BN + chalf * HALF + c * 16), v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) racc[c * 16 + j] += __uint_as_float(v[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(buf));
            }
            const int ex = x0 + er % p.tw, ey = y0 + (er / p.tw) % p.th, eb = b0 + er / (p.tw * p.th);
            if (eb < p.n) {
                const long long pix = ((long long)eb * p.h + ey) * p.w + ex;
                const float nz = p.noise ? __ldg(p.noise + pix) : 0.f;
                float* yrow = p.y + (long long)eb * p.ys[0] + (long long)ey * p.ys[2] + (long long)ex * p.ys[3];
#pragma unroll
                for (int j = 0; j < HALF; j += 4) {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int co = n0 + chalf * HALF + j + e;
                        float val = racc[j + e];
                        if (p.out_scale) val *= __ldg(p.out_scale + (long long)eb * p.co + co);
                        if (p.bias) val += __ldg(p.bias + co);
                        val += nz;
                        if (p.act == 3) val = val > 0.f ? val : val * p.alpha;
                        o[e] = val * p.gain;
                    }
                    const int cbase = n0 + chalf * HALF + j;
                    if (p.ys[1] == 1) st4(yrow + cbase, make_float4(o[0], o[1], o[2], o[3]));
                    else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) yrow[(long long)(cbase + e) * p.ys[1]] = o[e];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, TMEM_COLS);
    }
}

// w[co][ci][k][k] -> per (n-tile, sub-block): {B_big, B_small} fp32 tiles [bn rows x 32 k], 128 B rows, swizzled
__global__ void conv_pack_tc32_kernel(const float* __restrict__ w, unsigned char* __restrict__ wp, int co, int ci, int k,
                                      float coef, int transpose, int bn, int subs, int spb) {
    const int nout_n = transpose ? ci : co;
    const int kk2 = k * k;
    const long long total = (long long)(nout_n / bn) * subs * bn * SUB;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(idx % SUB);
        long long r = idx / SUB;
        const int nl = (int)(r % bn); r /= bn;
        const int t = (int)(r % subs);
        const int nt = (int)(r / subs);
        const int tap = t / spb, kin = (t % spb) * SUB + kk, nout = nt * bn + nl;
        const int o = transpose ? kin : nout, i = transpose ? nout : kin, ts = transpose ? kk2 - 1 - tap : tap;
        const float v = w[((long long)o * ci + i) * kk2 + ts] * coef;
        const uint32_t big = rna_tf32(v), small = rna_tf32(v - __uint_as_float(big));
        unsigned char* tile = wp + ((size_t)nt * subs + t) * (size_t)(2 * tile_b(bn));
        const size_t off = (size_t)nl * 128 + ((((kk * 4) >> 4) ^ (nl & 7)) << 4) + ((kk * 4) & 15);
        *reinterpret_cast<uint32_t*>(tile + off) = big;
        *reinterpret_cast<uint32_t*>(tile + (size_t)tile_b(bn) + off) = small;
    }
}

struct Geometry { int bn, spb, subs, tw, th, tb; };

static bool geometry(int h, int w, int ci, int co, int k, Geometry& g) {
    if (k != 1 && k != 3) return false;
    if (ci % SUB != 0 || ci < SUB) return false;
    g.bn = co % 128 == 0 ? 128 : (co == 64 ? 64 : (co == 32 ? 32 : 0));
    if (!g.bn) return false;
    if (!pixel_box(BM, h, w, g.tw, g.th, g.tb)) return false;
    g.spb = ci / SUB;
    g.subs = k * k * g.spb;
    return true;
}

template <int BN>
static int launch(const CUtensorMap& map, const Params& tp, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    const int smem = smem_bytes(BN);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_fwd_tc32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "conv_fwd_tc32: cannot opt in to %d B of shared memory: %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    conv_fwd_tc32_kernel<BN><<<grid, NTHREADS, smem, st>>>(map, tp);
    return launched("conv_fwd_tc32");
}

}  // namespace tc32

long long conv_packed_bytes_tc32(int co, int ci, int k) { return 2LL * co * ci * k * k * 4; }

int conv_pack_tc32(const float* w, void* wp, int co, int ci, int k, float coef, int transpose, cudaStream_t st) {
    const int cin = transpose ? co : ci, cout = transpose ? ci : co;
    tc32::Geometry g;
    if (!tc32::geometry(16, 16, cin, cout, k, g)) return fail(SG2_ENOTSUP, "conv_pack_tc32: unsupported shape");
    const long long total = (long long)(cout / g.bn) * g.subs * g.bn * tc32::SUB;
    const int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 8);
    tc32::conv_pack_tc32_kernel<<<blocks, 256, 0, st>>>(w, (unsigned char*)wp, co, ci, k, coef, transpose, g.bn, g.subs, g.spb);
    return launched("conv_pack_tc32");
}

int conv_fwd_tc32(const ConvParams& p, cudaStream_t st) {
    tc32::Geometry g;
    if (!tc32::geometry(p.h, p.w, p.ci, p.co, p.k, g)) return fail(SG2_ENOTSUP, "conv_fwd_tc32: unsupported shape");
    CUtensorMap map;
    int rc = tc::make_nhwc_map(&map, p.x, p.n, p.h, p.w, p.ci, g.tw, g.th, g.tb, "conv_fwd_tc32");
    if (rc) return rc;
    tc32::Params tp;
    tp.in_scale = p.in_scale; tp.out_scale = p.out_scale; tp.bias = p.bias; tp.noise = p.noise;
    tp.y = p.y;
    for (int i = 0; i < 4; ++i) tp.ys[i] = p.ys[i];
    tp.wp = (const unsigned char*)p.wp;
    tp.n = p.n; tp.h = p.h; tp.w = p.w; tp.ci = p.ci; tp.co = p.co; tp.k = p.k;
    tp.tw = g.tw; tp.th = g.th; tp.tb = g.tb;
    tp.tiles_x = p.w / g.tw; tp.tiles_y = p.h / g.th;
    const int tiles_b = (p.n + g.tb - 1) / g.tb;
    tp.subs = g.subs; tp.spb = g.spb;
    tp.act = p.act; tp.alpha = p.alpha; tp.gain = p.gain;
    tp.m_tiles = tp.tiles_x * tp.tiles_y * tiles_b;
    tp.n_tiles = p.co / g.bn;
    dim3 grid((unsigned)std::min(tp.m_tiles * tp.n_tiles, num_sms()));      // persistent: one CTA per SM
    if (g.bn == 128) return tc32::launch<128>(map, tp, grid, st);
    if (g.bn == 64) return tc32::launch<64>(map, tp, grid, st);
    return tc32::launch<32>(map, tp, grid, st);
}

}  // namespace sg2
