// tcgen05 implicit-GEMM convolution for sm_100a (stride 1, "same" zero padding, k in {1,3}, NHWC fp32 in HBM).
//
// Replaces the cuDNN/oneDNN convolution calls of the reference path (implementations/StyleGAN2/model.py:106-132
// ModulatedConv2d, :29-53 ELR conv, and their data gradients) with a kernel written for the B200 tensor cores.
//
// Precision: "bf16x3".  fp32 activations and weights are split on the fly into bf16 (hi, lo) pairs and every
// product is evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM: ~2^-16 relative error per
// product (the dropped lo*lo term), i.e. fp32-class results at 1/3 of the bf16 tensor rate -- what the 1e-3
// parity bar against the fp32 reference needs; single-pass TF32 would not hold it (truncation bias ~1e-3/layer).
//
// GEMM view:  D[M = pixels, N = co] = sum_{tap, ci} A[pixel + off(tap), ci] * W[co, tap, ci]
//   CTA tile : 128 pixels (a TW x TH x TB box of the [B,H,W] pixel grid) x BN output channels
//   K step   : 64 k-columns = two "sub-blocks" of 32 input channels of one tap (for Ci = 32 the two sub-blocks
//              are two consecutive taps); K is padded to an even number of sub-blocks with zero weights.
//   A path   : TMA 4-D box [32ch, TW, TH, TB] of the NHWC tensor at the tap-shifted coordinate (OOB -> 0 = padding)
//              -> 128B-swizzled fp32 staging tile -> 8 transform warps (style scale, hi/lo split) -> two bf16
//              K-major SWIZZLE_128B tiles (A_hi, A_lo) -> tcgen05.mma (SS).
//   B path   : weights are pre-packed by conv_pack_tc into the exact shared-memory image (bf16 hi/lo, swizzled),
//              one cp.async.bulk per K step.
//   Epilogue : tcgen05.ld -> out_scale (demod) / bias / noise / leaky-ReLU / gain -> global (NHWC, 128-bit stores).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 = transform, then epilogue.
#include "tc_common.cuh"
#include "conv.h"

namespace sg2 {

namespace tc {

constexpr int BM = 128;            // pixels per CTA tile (UMMA M)
constexpr int KSTEP = 64;          // k-columns per pipeline step (bf16: 128 B rows)
constexpr int SUB = 32;            // channels per TMA box (fp32: 128 B rows)
constexpr int STAGES = 2;
constexpr int NTHREADS = 448;      // 14 warps: TMA, MMA, 8 transform, 4 epilogue
constexpr int STAGE_F32 = 2 * BM * SUB * 4;      // 32 KB: two fp32 sub-block tiles
constexpr int STAGE_A = 2 * BM * KSTEP * 2;      // 32 KB: A_hi + A_lo

__host__ __device__ constexpr int stage_b(int bn) { return 2 * bn * KSTEP * 2; }   // B_hi + B_lo
__host__ __device__ constexpr int smem_bytes(int bn) { return 1024 + STAGES * (STAGE_F32 + STAGE_A + stage_b(bn)) + 256; }

struct TcParams {
    const float* in_scale;   // [n, ci]
    const float* out_scale;  // [n, co]
    const float* bias;       // [co]
    const float* noise;      // [n*h*w]
    float* y;
    long long ys[4];
    const unsigned char* wp; // packed weights
    int n, h, w, ci, co, k;
    int tw, th, tb;          // pixel box of one CTA tile (tw*th*tb == 128)
    int tiles_x, tiles_y, m_tiles, n_tiles;
    int ksteps, subs, spb;   // K steps, real sub-blocks, sub-blocks per tap
    int act;
    float alpha, gain;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap xmap, const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned carve-up (SWIZZLE_128B atoms)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t f32_base = base;
    const uint32_t a_base = f32_base + STAGES * STAGE_F32;
    const uint32_t b_base = a_base + STAGES * STAGE_A;
    const uint32_t bar_base = b_base + STAGES * stage_b(BN);
    auto f_full = [&](int s) { return bar_base + 8u * s; };
    auto f_empty = [&](int s) { return bar_base + 16u + 8u * s; };
    auto a_full = [&](int s) { return bar_base + 32u + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 48u + 8u * s; };
    auto b_full = [&](int s) { return bar_base + 64u + 8u * s; };
    auto b_empty = [&](int s) { return bar_base + 80u + 8u * s; };
    auto acc_full = [&](int b) { return bar_base + 96u + 8u * b; };
    auto acc_empty = [&](int b) { return bar_base + 112u + 8u * b; };
    const uint32_t tmem_slot = bar_base + 128u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pad = p.k >> 1;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;      // two accumulator buffers: epilogue overlaps the next tile
    // persistent tile loop: tile = n_tile * m_tiles + m_tile (all CTAs share an n-tile's weights at any time)
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
            mbar_init(f_full(s), 1); mbar_init(f_empty(s), 8);
            mbar_init(a_full(s), 8); mbar_init(a_empty(s), 1);
            mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 4); }
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
                const unsigned char* wsrc = p.wp + (size_t)(n0 / BN) * p.ksteps * stage_b(BN);
                for (int t = 0; t < p.ksteps; ++t, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(b_empty(s), ph ^ 1);
                    mbar_expect_tx(b_full(s), stage_b(BN));
                    bulk_load(b_base + s * stage_b(BN), wsrc + (size_t)t * stage_b(BN), stage_b(BN), b_full(s));
                    mbar_wait(f_empty(s), ph ^ 1);
                    mbar_expect_tx(f_full(s), STAGE_F32);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int sb = 2 * t + h;
                        int tap = sb / p.spb, c0 = (sb % p.spb) * SUB;
                        int dy = tap / p.k - pad, dx = tap % p.k - pad;
                        if (sb >= p.subs) { c0 = p.ci; dy = 0; dx = 0; }        // K padding: fully out of bounds -> zeros
                        tma_load_4d(f32_base + s * STAGE_F32 + h * (BM * SUB * 4), &xmap, f_full(s), c0, x0 + dx, y0 + dy, b0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_bf16(BM, BN);
            int it = 0, ti = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
                const int buf = ti & 1;
                mbar_wait(acc_empty(buf), ((ti >> 1) & 1) ^ 1);        // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_d + (uint32_t)(buf * BN);
                for (int t = 0; t < p.ksteps; ++t, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(b_full(s), ph);
                    mbar_wait(a_full(s), ph);
                    tc_fence_after();
                    const uint32_t a_hi = a_base + s * STAGE_A, a_lo = a_hi + BM * KSTEP * 2;
                    const uint32_t b_hi = b_base + s * stage_b(BN), b_lo = b_hi + BN * KSTEP * 2;
#pragma unroll
                    for (int kq = 0; kq < KSTEP / 16; ++kq) {
                        const uint64_t dah = kmajor_desc(a_hi + kq * 32), dal = kmajor_desc(a_lo + kq * 32);
                        const uint64_t dbh = kmajor_desc(b_hi + kq * 32), dbl = kmajor_desc(b_lo + kq * 32);
                        mma_bf16(d, dah, dbh, idesc, (t | kq) != 0);
                        mma_bf16(d, dal, dbh, idesc, 1);
                        mma_bf16(d, dah, dbl, idesc, 1);
                    }
                    mma_commit(a_empty(s));
                    mma_commit(b_empty(s));
                }
                mma_commit(acc_full(buf));
            }
        }
    } else if (warp < 10) {
        // ================= transform warps (2..9) =================
        const int tt = threadIdx.x - 64;           // 0..255
        const int r = tt & 127;                    // tile row = pixel
        const int half = tt >> 7;                  // which sub-block of the K step
        const int sw = r & 7;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int x0, y0, b0, n0;
            tile_coords(tile, x0, y0, b0, n0);
            const int pb = b0 + r / (p.tw * p.th);
            const bool row_ok = pb < p.n;
            for (int t = 0; t < p.ksteps; ++t, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(f_full(s), ph);
                const uint32_t src = f32_base + s * STAGE_F32 + half * (BM * SUB * 4) + r * 128;
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 q = lds4(src + ((j ^ sw) << 4));
                    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(f_empty(s));            // staging tile consumed (values are in registers)
                if (p.in_scale) {
                    const int sb = 2 * t + half;
                    if (sb < p.subs && row_ok) {
                        const float* sp = p.in_scale + (long long)pb * p.ci + (sb % p.spb) * SUB;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 q = ldg4(sp + 4 * j);
                            v[4 * j] *= q.x; v[4 * j + 1] *= q.y; v[4 * j + 2] *= q.z; v[4 * j + 3] *= q.w;
                        }
                    }
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
                mbar_wait(a_empty(s), ph ^ 1);
                const uint32_t dst_hi = a_base + s * STAGE_A + r * 128, dst_lo = dst_hi + BM * KSTEP * 2;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t off = (uint32_t)(((4 * half + q) ^ sw) << 4);
                    sts4(dst_hi + off, hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                    sts4(dst_lo + off, lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
                }
                fence_proxy_async();                                // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full(s));
            }
        }
    } else {
        // ================= epilogue warps (10..13): one per TMEM lane quarter, a full accumulator row per thread =====
        const int q4 = warp & 3;
        const int er = q4 * 32 + lane;              // accumulator row = pixel
        int ti = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
            int x0, y0, b0, n0;
            tile_coords(tile, x0, y0, b0, n0);
            const int buf = ti & 1;
            const int ex = x0 + er % p.tw, ey = y0 + (er / p.tw) % p.th, eb = b0 + er / (p.tw * p.th);
            const bool e_ok = eb < p.n;
            const long long pix = ((long long)eb * p.h + ey) * p.w + ex;
            const float nz = (p.noise && e_ok) ? __ldg(p.noise + pix) : 0.f;
            float* yrow = p.y + (long long)eb * p.ys[0] + (long long)ey * p.ys[2] + (long long)ex * p.ys[3];
            mbar_wait(acc_full(buf), (ti >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 16; ++c) {
                const int col0 = c * 16;
                uint32_t acc[16];
                tmem_ld16(tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * BN + col0), acc);
                if (e_ok) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int co = n0 + col0 + j + e;
                            float val = __uint_as_float(acc[j + e]);
                            if (p.out_scale) val *= __ldg(p.out_scale + (long long)eb * p.co + co);
                            if (p.bias) val += __ldg(p.bias + co);
                            val += nz;
                            if (p.act == 3) val = val > 0.f ? val : val * p.alpha;
                            o[e] = val * p.gain;
                        }
                        if (p.ys[1] == 1) st4(yrow + n0 + col0 + j, make_float4(o[0], o[1], o[2], o[3]));
                        else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) yrow[(long long)(n0 + col0 + j + e) * p.ys[1]] = o[e];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, TMEM_COLS);
    }
}

// ---- weight packing: w[co][ci][k][k] -> per (n-tile, K step) shared-memory image {B_hi, B_lo}, bf16, swizzled ----
__global__ void conv_pack_tc_kernel(const float* __restrict__ w, unsigned char* __restrict__ wp, int co, int ci, int k,
                                    float coef, int transpose, int bn, int ksteps, int subs, int spb) {
    // logical GEMM weight: Wg[nout][tap][kin]
    const int kin_n = transpose ? co : ci, nout_n = transpose ? ci : co;
    const int kk2 = k * k;
    const long long total = (long long)(nout_n / bn) * ksteps * bn * KSTEP;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(idx % KSTEP);
        long long r = idx / KSTEP;
        const int nl = (int)(r % bn); r /= bn;
        const int t = (int)(r % ksteps);
        const int nt = (int)(r / ksteps);
        const int sb = 2 * t + kk / SUB;
        float v = 0.f;
        if (sb < subs) {
            const int tap = sb / spb, kin = (sb % spb) * SUB + kk % SUB, nout = nt * bn + nl;
            const int o = transpose ? kin : nout, i = transpose ? nout : kin, ts = transpose ? kk2 - 1 - tap : tap;
            v = w[((long long)o * ci + i) * kk2 + ts] * coef;
        }
        (void)kin_n;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        unsigned char* tile = wp + ((size_t)nt * ksteps + t) * (size_t)stage_b(bn);
        const size_t off = (size_t)nl * 128 + ((((kk * 2) >> 4) ^ (nl & 7)) << 4) + ((kk * 2) & 15);
        *reinterpret_cast<__nv_bfloat16*>(tile + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(tile + (size_t)bn * KSTEP * 2 + off) = l;
    }
}

struct Geometry { int bn, spb, subs, ksteps, tw, th, tb; };

static int pick_bn(int co) { return co % 128 == 0 ? 128 : (co == 64 ? 64 : (co == 32 ? 32 : 0)); }

static bool geometry(int n, int h, int w, int ci, int co, int k, Geometry& g) {
    if (k != 1 && k != 3) return false;
    if (ci % SUB != 0 || ci < SUB) return false;
    g.bn = pick_bn(co);
    if (!g.bn) return false;
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (!pow2(w) || !pow2(h) || w > 256 || h > 256 || w < 4 || h < 4) return false;
    g.tw = w < BM ? w : BM;
    g.th = (BM / g.tw) < h ? (BM / g.tw) : h;
    g.tb = BM / (g.tw * g.th);
    if (g.tw * g.th * g.tb != BM || g.tb > 256) return false;
    g.spb = ci / SUB;
    g.subs = k * k * g.spb;
    g.ksteps = (g.subs + 1) / 2;
    (void)n;
    return true;
}

}  // namespace tc

bool conv_tc_supported(int n, int h, int w, int ci, int co, int k) {
    tc::Geometry g;
    return tc::geometry(n, h, w, ci, co, k, g);
}

long long conv_packed_bytes_tc(int co, int ci, int k) {
    // upper bound over both orientations: n-tiles * ksteps * stage bytes
    long long best = (long long)co * ci * k * k * 4;
    for (int tr = 0; tr < 2; ++tr) {
        const int cin = tr ? co : ci, cout = tr ? ci : co;
        tc::Geometry g;
        if (!tc::geometry(1, 16, 16, cin, cout, k, g)) continue;
        long long b = (long long)(cout / g.bn) * g.ksteps * tc::stage_b(g.bn);
        if (b > best) best = b;
    }
    return best;
}

int conv_pack_tc(const float* w, void* wp, int co, int ci, int k, float coef, int transpose, cudaStream_t st) {
    const int cin = transpose ? co : ci, cout = transpose ? ci : co;
    tc::Geometry g;
    if (!tc::geometry(1, 16, 16, cin, cout, k, g)) return fail(SG2_ENOTSUP, "conv_pack_tc: unsupported shape");
    const long long total = (long long)(cout / g.bn) * g.ksteps * g.bn * tc::KSTEP;
    const int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 8);
    tc::conv_pack_tc_kernel<<<blocks, 256, 0, st>>>(w, (unsigned char*)wp, co, ci, k, coef, transpose, g.bn, g.ksteps, g.subs, g.spb);
    return launched("conv_pack_tc");
}

template <int BN>
static int launch_fwd(const CUtensorMap& map, const tc::TcParams& tp, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    const int smem = tc::smem_bytes(BN);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::conv_fwd_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "conv_fwd_tc: cannot opt in to %d B of shared memory: %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    tc::conv_fwd_tc_kernel<BN><<<grid, tc::NTHREADS, smem, st>>>(map, tp);
    return launched("conv_fwd_tc");
}

int conv_fwd_tc(const ConvParams& p, cudaStream_t st) {
    tc::Geometry g;
    if (!tc::geometry(p.n, p.h, p.w, p.ci, p.co, p.k, g)) return fail(SG2_ENOTSUP, "conv_fwd_tc: unsupported shape");
    tc::EncodeTiledFn enc = tc::encode_fn();
    if (!enc) return fail(SG2_ELAUNCH, "conv_fwd_tc: cuTensorMapEncodeTiled is not available from the driver");
    if (((uintptr_t)p.x & 15) != 0) return fail(SG2_EINVAL, "conv_fwd_tc: x must be 16-byte aligned");
    CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)p.ci, (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.n};
    const cuuint64_t strides[3] = {(cuuint64_t)p.ci * 4, (cuuint64_t)p.w * p.ci * 4, (cuuint64_t)p.h * p.w * p.ci * 4};
    const cuuint32_t box[4] = {(cuuint32_t)tc::SUB, (cuuint32_t)g.tw, (cuuint32_t)g.th, (cuuint32_t)g.tb};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)p.x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SG2_ELAUNCH, "conv_fwd_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    tc::TcParams tp;
    tp.in_scale = p.in_scale; tp.out_scale = p.out_scale; tp.bias = p.bias; tp.noise = p.noise;
    tp.y = p.y;
    for (int i = 0; i < 4; ++i) tp.ys[i] = p.ys[i];
    tp.wp = (const unsigned char*)p.wp;
    tp.n = p.n; tp.h = p.h; tp.w = p.w; tp.ci = p.ci; tp.co = p.co; tp.k = p.k;
    tp.tw = g.tw; tp.th = g.th; tp.tb = g.tb;
    tp.tiles_x = p.w / g.tw; tp.tiles_y = p.h / g.th;
    const int tiles_b = (p.n + g.tb - 1) / g.tb;
    tp.ksteps = g.ksteps; tp.subs = g.subs; tp.spb = g.spb;
    tp.act = p.act; tp.alpha = p.alpha; tp.gain = p.gain;
    tp.m_tiles = tp.tiles_x * tp.tiles_y * tiles_b;
    tp.n_tiles = p.co / g.bn;
    dim3 grid((unsigned)std::min(tp.m_tiles * tp.n_tiles, num_sms()));      // persistent: one CTA per SM
    if (g.bn == 128) return launch_fwd<128>(map, tp, grid, st);
    if (g.bn == 64) return launch_fwd<64>(map, tp, grid, st);
    return launch_fwd<32>(map, tp, grid, st);
}

}  // namespace sg2
