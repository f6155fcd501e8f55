// Per-(sample, channel) reductions and the fused backward prologue of the modulated convolution.
// These replace the autograd graph that PyTorch builds for ModulatedConv2d.forward
// (implementations/StyleGAN2/model.py:106-132): d s[b,i] = sum_hw x * g_xs and
// d d[b,o] = sum_hw acc * g_u are plain reductions over NHWC tensors, and the
// leaky-ReLU / bias / demodulation parts of the backward are fused into ONE pass over (gy, y).
// All kernels are HBM-bound; grid = (channel chunks of 64, samples, hw slices), 128-bit accesses.
#include "common.cuh"

namespace sg2 {

// out_row: the [c] row of this (sample) in the output -- atomics across the hw slices -- or, with `part`, this block's own
// row in the per-slice partial buffer (plain store; aux_sum_slices_kernel adds the slices in a fixed order)
__device__ __forceinline__ void block_reduce_store(float4 acc, float (*sh)[68], float* out_row, int c, int cbase, bool part = false) {
    const int q = threadIdx.x & 15, pl = threadIdx.x >> 4;
    sh[pl][q * 4 + 0] = acc.x; sh[pl][q * 4 + 1] = acc.y; sh[pl][q * 4 + 2] = acc.z; sh[pl][q * 4 + 3] = acc.w;
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += sh[i][threadIdx.x];
        const int cc = cbase + threadIdx.x;
        if (cc < c) { if (part) out_row[cc] = s; else atomicAdd(out_row + cc, s); }
    }
    __syncthreads();
}

// chunk of channels a block covers: 64, or 32 for 32-channel tensors so that no thread idles (planes.cu uses the same layout)
__host__ __device__ __forceinline__ int aux_chunk_width(int c) { return (c % 64 == 0 || c > 32) ? 64 : 32; }

__device__ __forceinline__ void block_reduce_store_cw(float4 acc, float* sh, float* out_row, int c, int cbase, int cw, bool part) {
    const int tpp = cw >> 2, rows = 256 / tpp, pitch = cw + 4;
    const int q = threadIdx.x % tpp, pl = threadIdx.x / tpp;
    float* mine = sh + pl * pitch + q * 4;
    mine[0] = acc.x; mine[1] = acc.y; mine[2] = acc.z; mine[3] = acc.w;
    __syncthreads();
    if ((int)threadIdx.x < cw) {
        float s = 0.f;
        for (int i = 0; i < rows; ++i) s += sh[i * pitch + threadIdx.x];
        const int cc = cbase + threadIdx.x;
        if (cc < c) { if (part) out_row[cc] = s; else atomicAdd(out_row + cc, s); }
    }
    __syncthreads();
}

// out[b,c] += sum_{i in slice} a*bm ; optionally a_out = a * scale[b,c].  Two pixels per iteration: their loads are in flight together.
__global__ void __launch_bounds__(256) scale_reduce_hw_kernel(const float* __restrict__ a, const float* __restrict__ bm,
                                                              const float* __restrict__ scale, float* __restrict__ a_out,
                                                              float* __restrict__ out, float* __restrict__ part, int n, int hw, int c, int slice) {
    __shared__ float sh[32 * 36];                          // rows x (cw + 4): 16 x 68 or 32 x 36
    const int cw = aux_chunk_width(c), tpp = cw >> 2, rows = 256 / tpp;
    const int q = threadIdx.x % tpp, pl = threadIdx.x / tpp;
    const int c0 = blockIdx.x * cw + q * 4, b = blockIdx.y;
    const int beg = blockIdx.z * slice, end = min(hw, beg + slice);
    float4 acc = f4zero();
    if (c0 < c) {
        const long long base = (long long)b * hw * c + c0;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
        if (a_out && scale) sc = ldg4(scale + (long long)b * c + c0);
        auto one = [&](long long o, const float4& v, const float4& m) {
            if (a_out) st4_cs(a_out + o, mul4(v, sc));
            acc = add4(acc, bm ? mul4(v, m) : v);
        };
        int i = beg + pl;
        const long long step = (long long)rows * c;
        for (; i + rows < end; i += 2 * rows) {
            const long long o = base + (long long)i * c;
            const float4 v0 = ldg4(a + o), v1 = ldg4(a + o + step);
            const float4 m0 = bm ? ldg4(bm + o) : f4zero(), m1 = bm ? ldg4(bm + o + step) : f4zero();
            one(o, v0, m0); one(o + step, v1, m1);
        }
        if (i < end) { const long long o = base + (long long)i * c; one(o, ldg4(a + o), bm ? ldg4(bm + o) : f4zero()); }
    }
    if (part) block_reduce_store_cw(acc, sh, part + ((long long)blockIdx.z * n + b) * c, c, blockIdx.x * cw, cw, true);
    else block_reduce_store_cw(acc, sh, out + (long long)b * c, c, blockIdx.x * cw, cw, false);
}

// One pass over (gy, y):  gu = gy * act'(y);  g_acc = gu * d;  gb_part[b,o] += sum gu;
// gd[b,o] += sum gu * (u - bias - noise) / d   with u = act^-1(y).
__global__ void __launch_bounds__(256) modconv_bwd_prep_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                                               const float* __restrict__ noise, const float* __restrict__ bias,
                                                               const float* __restrict__ d, float* __restrict__ g_acc,
                                                               float* __restrict__ gb_part, float* __restrict__ gd,
                                                               float* __restrict__ part, int n, int hw, int c, int slice, float alpha) {
    __shared__ float sh[16][68];
    const int q = threadIdx.x & 15, pl = threadIdx.x >> 4;
    const int c0 = blockIdx.x * 64 + q * 4, b = blockIdx.y;
    const int beg = blockIdx.z * slice, end = min(hw, beg + slice);
    float4 sgu = f4zero(), sgd = f4zero();
    if (c0 < c) {
        const long long base = (long long)b * hw * c + c0;
        const float4 dv = d ? ldg4(d + (long long)b * c + c0) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 bv = bias ? ldg4(bias + c0) : f4zero();
        const float inv_alpha = 1.f / alpha;
        for (int i = beg + pl; i < end; i += 16) {
            const float4 g = ldg4(gy + base + (long long)i * c);
            const float4 yv = ldg4(y + base + (long long)i * c);
            const float nz = noise ? __ldg(noise + (long long)b * hw + i) : 0.f;
            float4 gu, u;
            gu.x = yv.x > 0.f ? g.x : g.x * alpha; u.x = yv.x > 0.f ? yv.x : yv.x * inv_alpha;
            gu.y = yv.y > 0.f ? g.y : g.y * alpha; u.y = yv.y > 0.f ? yv.y : yv.y * inv_alpha;
            gu.z = yv.z > 0.f ? g.z : g.z * alpha; u.z = yv.z > 0.f ? yv.z : yv.z * inv_alpha;
            gu.w = yv.w > 0.f ? g.w : g.w * alpha; u.w = yv.w > 0.f ? yv.w : yv.w * inv_alpha;
            st4_cs(g_acc + base + (long long)i * c, mul4(gu, dv));
            sgu = add4(sgu, gu);
            if (gd) {
                sgd.x = fmaf(gu.x, u.x - bv.x - nz, sgd.x); sgd.y = fmaf(gu.y, u.y - bv.y - nz, sgd.y);
                sgd.z = fmaf(gu.z, u.z - bv.z - nz, sgd.z); sgd.w = fmaf(gu.w, u.w - bv.w - nz, sgd.w);
            }
        }
        if (gd) { sgd.x /= dv.x; sgd.y /= dv.y; sgd.z /= dv.z; sgd.w /= dv.w; }
    }
    if (part) {
        const long long row = ((long long)blockIdx.z * n + b) * c;
        block_reduce_store(sgu, sh, part + row, c, blockIdx.x * 64, true);
        if (gd) block_reduce_store(sgd, sh, part + (long long)gridDim.z * n * c + row, c, blockIdx.x * 64, true);
    } else {
        block_reduce_store(sgu, sh, gb_part + (long long)b * c, c, blockIdx.x * 64);
        if (gd) block_reduce_store(sgd, sh, gd + (long long)b * c, c, blockIdx.x * 64);
    }
}

__global__ void __launch_bounds__(256) aux_sum_slices_kernel(const float* __restrict__ part, float* __restrict__ out, long long nc, int slices) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nc) return;
    float s = 0.f;
    for (int k = 0; k < slices; ++k) s += part[(long long)k * nc + i];
    out[i] = s;
}

static void pick_grid(int n, int hw, int c, dim3& grid, int& slice, bool narrow = false) {
    const int cw = narrow ? aux_chunk_width(c) : 64;
    const int cchunks = (c + cw - 1) / cw;
    long long want = std::max<long long>(1, (4LL * num_sms()) / ((long long)cchunks * n));
    int slices = (int)std::min<long long>(want, ceil_div(hw, 64));
    slice = (int)ceil_div(hw, slices);
    slices = (int)ceil_div(hw, slice);
    grid = dim3(cchunks, n, slices);
}

}  // namespace sg2

using namespace sg2;

extern "C" int64_t sg2_reduce_hw_workspace(int n, int hw, int c) {
    if (n <= 0 || hw <= 0 || c <= 0) return -1;
    dim3 grid; int slice;
    pick_grid(n, hw, c, grid, slice);
    return (int64_t)2 * grid.z * n * c * (int64_t)sizeof(float);
}

static int sum_slices(const float* part, float* out, long long nc, int slices, cudaStream_t st) {
    aux_sum_slices_kernel<<<(unsigned)ceil_div(nc, 256), 256, 0, st>>>(part, out, nc, slices);
    return launched("sum_slices");
}

extern "C" int sg2_scale_reduce_hw(const float* a, const float* bm, const float* scale, float* a_out, float* out,
                                   int n, int hw, int c, void* workspace, sg2_stream_t stream) {
    SG2_REQUIRE(a && out, "scale_reduce_hw: null pointer");
    SG2_REQUIRE(n > 0 && hw > 0 && c > 0 && c % 4 == 0, "scale_reduce_hw: need n,hw > 0 and C %% 4 == 0 (C=%d)", c);
    cudaStream_t st = (cudaStream_t)stream;
    if (!workspace) SG2_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n * c, st));
    dim3 grid; int slice;
    pick_grid(n, hw, c, grid, slice, true);
    scale_reduce_hw_kernel<<<grid, 256, 0, st>>>(a, bm, scale, a_out, out, (float*)workspace, n, hw, c, slice);
    int rc = launched("scale_reduce_hw");
    if (rc || !workspace) return rc;
    return sum_slices((const float*)workspace, out, (long long)n * c, (int)grid.z, st);
}

extern "C" int sg2_reduce_hw(const float* a, const float* bm, float* out, int n, int hw, int c, void* workspace, sg2_stream_t stream) {
    return sg2_scale_reduce_hw(a, bm, nullptr, nullptr, out, n, hw, c, workspace, stream);
}

extern "C" int sg2_modconv_bwd_prep(const float* gy, const float* y, const float* noise, const float* bias,
                                    const float* d, float* g_acc, float* gb_part, float* gd,
                                    int n, int hw, int c, float alpha, void* workspace, sg2_stream_t stream) {
    SG2_REQUIRE(gy && y && g_acc && gb_part, "modconv_bwd_prep: null pointer");
    SG2_REQUIRE(n > 0 && hw > 0 && c > 0 && c % 4 == 0, "modconv_bwd_prep: need n,hw > 0 and C %% 4 == 0 (C=%d)", c);
    SG2_REQUIRE(alpha != 0.f, "modconv_bwd_prep: alpha must be non-zero (the activation is inverted from y)");
    SG2_REQUIRE(!gd || d, "modconv_bwd_prep: gd requested without d");
    cudaStream_t st = (cudaStream_t)stream;
    if (!workspace) {
        SG2_CUDA(cudaMemsetAsync(gb_part, 0, sizeof(float) * (size_t)n * c, st));
        if (gd) SG2_CUDA(cudaMemsetAsync(gd, 0, sizeof(float) * (size_t)n * c, st));
    }
    dim3 grid; int slice;
    pick_grid(n, hw, c, grid, slice);
    modconv_bwd_prep_kernel<<<grid, 256, 0, st>>>(gy, y, noise, bias, d, g_acc, gb_part, gd, (float*)workspace, n, hw, c, slice, alpha);
    int rc = launched("modconv_bwd_prep");
    if (rc || !workspace) return rc;
    const long long nc = (long long)n * c;
    rc = sum_slices((const float*)workspace, gb_part, nc, (int)grid.z, st);
    if (rc || !gd) return rc;
    return sum_slices((const float*)workspace + (long long)grid.z * nc, gd, nc, (int)grid.z, st);
}
