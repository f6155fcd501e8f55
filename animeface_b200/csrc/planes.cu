// "Pair planes": an fp32 NHWC tensor stored as two bf16 tensors hi = bf16(v), lo = bf16(v - hi), back to back ([2][n][hw][c],
// 4 bytes per element like the fp32 it replaces).  hi + lo carries 16 mantissa bits; the three (or four) cross products of two
// such pairs are the bf16x3 arithmetic of the gradient convolutions.  Producing the planes ONCE per tensor, in the
// elementwise pass that touches the tensor anyway, lets the tcgen05 data-gradient and weight-gradient kernels (conv_halo_pl.cu,
// wgrad_pl.cu) take their operands by TMA straight into SWIZZLE_128B tiles: no in-kernel fp32 -> bf16 transform, which was the
// limiter of the round-1 weight-gradient kernel (every CTA of a (kernel row, channel tile) re-converted gy and x).
//
//   split_planes_kernel      planes = split(x * scale[b,c])                       (x of a conv, for its weight gradient)
//   bwd_prep_planes_kernel   the backward prologue of y = lrelu(d*acc + bias + noise) (modconv_aux.cu) with g_acc written as
//                            planes; per-(sample, channel) sums reduced DETERMINISTICALLY (per-slice partials, fixed order).
// HBM-bound, 128-bit accesses, grid = (64-channel chunks, samples, hw slices).
#include <cuda_bf16.h>
#include "common.cuh"

namespace sg2 {

__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

// Thread layout of both kernels: a block covers a chunk of CW channels (64, or 32 for 32-channel tensors so that no thread
// idles) x a slice of the pixels; TPP = CW / 4 threads share a pixel (one float4 each), 256 / TPP pixels per pass.
__device__ __forceinline__ int chunk_width(int c) { return (c % 64 == 0 || c > 32) ? 64 : (c > 16 ? 32 : 16); }

__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                           __nv_bfloat16* __restrict__ planes, long long plane_stride, int hw, int c, int slice) {
    const int cw = chunk_width(c), tpp = cw >> 2, rows = 256 / tpp;
    const int q = threadIdx.x % tpp, pl = threadIdx.x / tpp;
    const int c0 = blockIdx.x * cw + q * 4, b = blockIdx.y;
    const int beg = blockIdx.z * slice, end = min(hw, beg + slice);
    if (c0 >= c) return;
    const long long base = (long long)b * hw * c + c0;
    const float4 sc = scale ? ldg4(scale + (long long)b * c + c0) : make_float4(1.f, 1.f, 1.f, 1.f);
    auto emit = [&](long long o, const float4& v) {
        uint2 hi, lo;
        split4(mul4(v, sc), hi, lo);
        *reinterpret_cast<uint2*>(planes + o) = hi;
        *reinterpret_cast<uint2*>(planes + plane_stride + o) = lo;
    };
    // four pixels per iteration: their loads are in flight together before the first store
    int i = beg + pl;
    const long long step = (long long)rows * c;
    for (; i + 3 * rows < end; i += 4 * rows) {
        const long long o = base + (long long)i * c;
        const float4 v0 = ldg4(x + o), v1 = ldg4(x + o + step), v2 = ldg4(x + o + 2 * step), v3 = ldg4(x + o + 3 * step);
        emit(o, v0); emit(o + step, v1); emit(o + 2 * step, v2); emit(o + 3 * step, v3);
    }
    for (; i < end; i += rows) {
        const long long o = base + (long long)i * c;
        emit(o, ldg4(x + o));
    }
}

// per-block partial sums -> part[(slice, b, c)] (no atomics: the caller's second pass adds the slices in a fixed order)
__device__ __forceinline__ void block_reduce_part(float4 acc, float* sh, float* out_row, int c, int cbase, int cw) {
    const int tpp = cw >> 2, rows = 256 / tpp, pitch = cw + 4;
    const int q = threadIdx.x % tpp, pl = threadIdx.x / tpp;
    float* mine = sh + pl * pitch + q * 4;
    mine[0] = acc.x; mine[1] = acc.y; mine[2] = acc.z; mine[3] = acc.w;
    __syncthreads();
    if ((int)threadIdx.x < cw) {
        float s = 0.f;
        for (int i = 0; i < rows; ++i) s += sh[i * pitch + threadIdx.x];
        const int cc = cbase + threadIdx.x;
        if (cc < c) out_row[cc] = s;
    }
    __syncthreads();
}

// gu = g * lrelu'(y) and u = the pre-activation value, for 4 channels
__device__ __forceinline__ void lrelu_grad4(const float4& g, const float4& yv, float alpha, float inv_alpha, float4& gu, float4& u) {
    gu.x = yv.x > 0.f ? g.x : g.x * alpha; u.x = yv.x > 0.f ? yv.x : yv.x * inv_alpha;
    gu.y = yv.y > 0.f ? g.y : g.y * alpha; u.y = yv.y > 0.f ? yv.y : yv.y * inv_alpha;
    gu.z = yv.z > 0.f ? g.z : g.z * alpha; u.z = yv.z > 0.f ? yv.z : yv.z * inv_alpha;
    gu.w = yv.w > 0.f ? g.w : g.w * alpha; u.w = yv.w > 0.f ? yv.w : yv.w * inv_alpha;
}

// POOLED: gy is the gradient of a 2x2 average pooling's OUTPUT ([n, hw/4, c], full-resolution width pool_w).  The pooling
// adjoint (broadcast over the 2x2 window, times gscale) is applied while reading, so the full-size gradient is never written
// or read: a thread walks POOLED pixels -- one load of g, the four y loads of its window in flight together, four stores.
// `slice` then counts pooled pixels.
template <bool POOLED>
__global__ void __launch_bounds__(256) bwd_prep_planes_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                                              const float* __restrict__ noise, const float* __restrict__ bias,
                                                              const float* __restrict__ d, __nv_bfloat16* __restrict__ planes,
                                                              long long plane_stride, float* __restrict__ part_gb, float* __restrict__ part_gd,
                                                              const unsigned short* __restrict__ signs,
                                                              int n, int hw, int c, int slice, float alpha, int pool_w, float gscale) {
    __shared__ float sh[64 * 36];                          // rows x (CW + 4): 16 x 68 or 32 x 36
    const int cw = chunk_width(c), tpp = cw >> 2, rows = 256 / tpp;
    const int q = threadIdx.x % tpp, pl = threadIdx.x / tpp;
    const int c0 = blockIdx.x * cw + q * 4, b = blockIdx.y;
    const int items = POOLED ? hw >> 2 : hw;
    const int beg = blockIdx.z * slice, end = min(items, beg + slice);
    float4 sgu = f4zero(), sgd = f4zero();
    if (c0 < c) {
        const long long base = (long long)b * hw * c + c0;
        const float4 dv = d ? ldg4(d + (long long)b * c + c0) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 bv = bias ? ldg4(bias + c0) : f4zero();
        const float inv_alpha = 1.f / alpha;
        auto emit = [&](long long o, int pix, const float4& g, const float4& yv) {
            float4 gu = g, u = f4zero();
            if (y) lrelu_grad4(g, yv, alpha, inv_alpha, gu, u);
            uint2 hi, lo;
            split4(mul4(gu, dv), hi, lo);
            *reinterpret_cast<uint2*>(planes + o) = hi;
            *reinterpret_cast<uint2*>(planes + plane_stride + o) = lo;
            sgu = add4(sgu, gu);
            if (part_gd) {
                const float nz = noise ? __ldg(noise + (long long)b * hw + pix) : 0.f;
                sgd.x = fmaf(gu.x, u.x - bv.x - nz, sgd.x); sgd.y = fmaf(gu.y, u.y - bv.y - nz, sgd.y);
                sgd.z = fmaf(gu.z, u.z - bv.z - nz, sgd.z); sgd.w = fmaf(gu.w, u.w - bv.w - nz, sgd.w);
            }
        };
        if (POOLED) {
            const int pw2 = pool_w >> 1;
            const long long gbase = (long long)b * (hw >> 2) * c + c0;
            for (int i = beg + pl; i < end; i += rows) {
                const int py = i / pw2, px = i - py * pw2;
                const float4 g = scale4(ldg4(gy + gbase + (long long)i * c), gscale);
                const int p00 = (2 * py) * pool_w + 2 * px;
                const int pix[4] = {p00, p00 + 1, p00 + pool_w, p00 + pool_w + 1};
                if (signs) {
                    // the window's leaky-ReLU signs from the 16-bit word the pooling kernel wrote (sg2_avgpool2_fwd): y is not read
                    const unsigned m = signs[(gbase + (long long)i * c) >> 2];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const unsigned b4 = m >> (4 * e);
                        float4 gu;
                        gu.x = (b4 & 1) ? g.x : g.x * alpha; gu.y = (b4 & 2) ? g.y : g.y * alpha;
                        gu.z = (b4 & 4) ? g.z : g.z * alpha; gu.w = (b4 & 8) ? g.w : g.w * alpha;
                        const long long o = base + (long long)pix[e] * c;
                        uint2 hi, lo;
                        split4(mul4(gu, dv), hi, lo);
                        *reinterpret_cast<uint2*>(planes + o) = hi;
                        *reinterpret_cast<uint2*>(planes + plane_stride + o) = lo;
                        sgu = add4(sgu, gu);
                    }
                    continue;
                }
                float4 yv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) yv[e] = y ? ldg4(y + base + (long long)pix[e] * c) : f4zero();
#pragma unroll
                for (int e = 0; e < 4; ++e) emit(base + (long long)pix[e] * c, pix[e], g, yv[e]);
            }
        } else {
            // two pixels per iteration: both pairs of loads in flight before the first store
            int i = beg + pl;
            for (; i + rows < end; i += 2 * rows) {
                const long long o0 = base + (long long)i * c, o1 = o0 + (long long)rows * c;
                const float4 g0 = scale4(ldg4(gy + o0), gscale), g1 = scale4(ldg4(gy + o1), gscale);
                const float4 y0 = y ? ldg4(y + o0) : f4zero(), y1 = y ? ldg4(y + o1) : f4zero();
                emit(o0, i, g0, y0);
                emit(o1, i + rows, g1, y1);
            }
            if (i < end) {
                const long long o = base + (long long)i * c;
                emit(o, i, scale4(ldg4(gy + o), gscale), y ? ldg4(y + o) : f4zero());
            }
        }
        if (part_gd) { sgd.x /= dv.x; sgd.y /= dv.y; sgd.z /= dv.z; sgd.w /= dv.w; }
    }
    const long long row = ((long long)blockIdx.z * n + b) * c;
    block_reduce_part(sgu, sh, part_gb + row, c, blockIdx.x * cw, cw);
    if (part_gd) block_reduce_part(sgd, sh, part_gd + row, c, blockIdx.x * cw, cw);
}

// out[i] = sum_s part[s][i] in slice order (deterministic); i over n*c
__global__ void __launch_bounds__(256) sum_slices_kernel(const float* __restrict__ part, float* __restrict__ out, long long nc, int slices) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nc) return;
    float s = 0.f;
    for (int k = 0; k < slices; ++k) s += part[(long long)k * nc + i];
    out[i] = s;
}

// the same for MANY rows and few columns (the bias gradient: rows = slices x samples, up to a few thousand; columns = channels):
// a block owns 32 columns, its 8 warps take every 8th row with 4 rows of loads in flight, and the 8 partial sums meet in
// shared memory in warp order -- a fixed order, so the result is still run-to-run identical
__global__ void __launch_bounds__(256) sum_rows_kernel(const float* __restrict__ part, float* __restrict__ out, int nc, int rows) {
    __shared__ float sh[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < nc) {
        int k = warp;
        for (; k + 24 < rows; k += 32) {
            const float a = part[(long long)k * nc + i], b = part[(long long)(k + 8) * nc + i];
            const float c = part[(long long)(k + 16) * nc + i], d = part[(long long)(k + 24) * nc + i];
            s += a; s += b; s += c; s += d;
        }
        for (; k < rows; k += 8) s += part[(long long)k * nc + i];
    }
    sh[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && i < nc) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += sh[j][lane];
        out[i] = t;
    }
}

static int host_chunk_width(int c) { return (c % 64 == 0 || c > 32) ? 64 : (c > 16 ? 32 : 16); }

// `items` = pixels a thread walks (pooled pixels for the pooled prologue); the slice count is the same for both so that one
// workspace size serves either form
static void pick_grid_pl(int n, int hw, int c, dim3& grid, int& slice, bool pooled = false) {
    const int cw = host_chunk_width(c);
    const int cchunks = (c + cw - 1) / cw;
    long long want = std::max<long long>(1, (4LL * num_sms()) / ((long long)cchunks * n));
    int slices = (int)std::min<long long>(std::min<long long>(want, 64), ceil_div(hw, 64));
    const int items = pooled ? hw / 4 : hw;
    slices = std::max(1, std::min(slices, items));
    slice = (int)ceil_div(items, slices);
    slices = (int)ceil_div(items, slice);
    grid = dim3(cchunks, n, slices);
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_split_planes(const float* x, const float* scale, void* planes, int n, int hw, int c, sg2_stream_t stream) {
    SG2_REQUIRE(x && planes, "split_planes: null pointer");
    SG2_REQUIRE(n > 0 && hw > 0 && c > 0 && c % 4 == 0, "split_planes: need n,hw > 0 and C %% 4 == 0 (C=%d)", c);
    dim3 grid; int slice;
    pick_grid_pl(n, hw, c, grid, slice);
    split_planes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, scale, (__nv_bfloat16*)planes, (long long)n * hw * c, hw, c, slice);
    return launched("split_planes");
}

extern "C" int64_t sg2_bwd_prep_planes_workspace(int n, int hw, int c) {
    if (n <= 0 || hw <= 0 || c <= 0) return -1;
    dim3 grid; int slice;
    pick_grid_pl(n, hw, c, grid, slice);
    return (int64_t)2 * grid.z * n * c * (int64_t)sizeof(float);
}

extern "C" int sg2_bwd_prep_planes(const float* gy, const float* y, const float* noise, const float* bias, const float* d,
                                   void* planes, float* gb, float* gd, void* workspace, const void* signs,
                                   int n, int hw, int c, float alpha, int pool_w, float gscale, sg2_stream_t stream) {
    SG2_REQUIRE(gy && planes && gb && workspace, "bwd_prep_planes: null pointer");
    SG2_REQUIRE(n > 0 && hw > 0 && c > 0 && c % 4 == 0, "bwd_prep_planes: need n,hw > 0 and C %% 4 == 0 (C=%d)", c);
    SG2_REQUIRE(alpha != 0.f, "bwd_prep_planes: alpha must be non-zero (the activation is inverted from y)");
    SG2_REQUIRE(!gd || (d && y), "bwd_prep_planes: gd requested without d / y");
    SG2_REQUIRE(!signs || (pool_w > 0 && !gd && !y), "bwd_prep_planes: window signs replace y in the pooled form only (no gd)");
    SG2_REQUIRE(pool_w == 0 || (pool_w > 0 && pool_w % 2 == 0 && hw % pool_w == 0 && (hw / pool_w) % 2 == 0),
                "bwd_prep_planes: pooled gradient needs an even full-resolution width and height (w=%d, hw=%d)", pool_w, hw);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid; int slice;
    pick_grid_pl(n, hw, c, grid, slice, pool_w > 0);
    const long long nc = (long long)n * c;
    float* part_gb = (float*)workspace;
    float* part_gd = gd ? part_gb + (long long)grid.z * nc : nullptr;
    if (pool_w > 0)
        bwd_prep_planes_kernel<true><<<grid, 256, 0, st>>>(gy, y, noise, bias, d, (__nv_bfloat16*)planes, (long long)n * hw * c, part_gb, part_gd,
                                                           (const unsigned short*)signs, n, hw, c, slice, alpha, pool_w, gscale);
    else
        bwd_prep_planes_kernel<false><<<grid, 256, 0, st>>>(gy, y, noise, bias, d, (__nv_bfloat16*)planes, (long long)n * hw * c, part_gb, part_gd,
                                                            nullptr, n, hw, c, slice, alpha, pool_w, gscale);
    int rc = launched("bwd_prep_planes");
    if (rc) return rc;
    const int blocks = (int)ceil_div(nc, 256);
    // gb: summed over the samples too (rows of the partial buffer are (slice, sample) pairs) -- the bias gradient, no torch pass
    sum_rows_kernel<<<(unsigned)ceil_div(c, 32), 256, 0, st>>>(part_gb, gb, c, (int)grid.z * n);
    rc = launched("sum_slices");
    if (rc || !gd) return rc;
    sum_slices_kernel<<<blocks, 256, 0, st>>>(part_gd, gd, nc, (int)grid.z);
    return launched("sum_slices");
}
