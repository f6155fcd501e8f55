// DiffAugment 'color,translation,cutout' as ONE pass forward and ONE pass backward over the image batch (plus a small
// per-sample mean reduction), NCHW fp32.
//
// Replaces thirdparty/diffaugment/DiffAugment.py:10-77 on the training path (3 calls per step on [32,3,256,256]): the
// reference runs ~10 elementwise / reduction / gather passes (brightness :23-25, saturation :28-31, contrast :34-37,
// translation :40-53 with a pad + NHWC permute + 3-D advanced-index gather, cutout :56-72).  Per sample b with the
// reference's draws r_b, r_s, r_c in [0,1), integer shifts (ty, tx) and cut window:
//   x1 = x + (r_b - 0.5);  x2 = (x1 - mean_c x1) * 2 r_s + mean_c x1;  x3 = (x2 - M) * (r_c + 0.5) + M,  M = mean_chw x2
//   y[i,j] = x3[i + ty, j + tx] (zero outside the image), then zero inside the cut window.
// Saturation preserves every pixel's channel mean, so M = mean_chw(x) + (r_b - 0.5): one reduction over x, then one pass.
// The map is affine in x; the backward kernel applies its transpose, and `linear_only` (no brightness constant) makes the
// forward kernel the backward's own derivative, so the op is closed under differentiation.  No atomics (deterministic).
#include "common.cuh"

namespace sg2 {
namespace da {

constexpr int MAXC = 8;

struct Params {
    const float* x; float* y;          // fwd: x -> y ; bwd: gy -> gx
    const float* rb; const float* rs; const float* rc;          // [B] raw uniform draws or null (no colour policy)
    const long long* ty; const long long* tx;                   // [B] integer shifts or null
    const long long* cy; const long long* cx;                   // [B] cutout centres (reference offset_x / offset_y) or null
    const float* sums;                 // [B] per-sample sum (fwd: of x; bwd: of the masked, shifted gy)
    float* part;                       // [B][slices] partial sums (reduction kernels)
    int B, C, H, W, slices;
    int cut_h, cut_w;
    int linear_only;
};

__device__ __forceinline__ bool in_cut(const Params& p, int b, int i, int j) {
    if (!p.cy) return false;
    // reference: grid = clamp(arange(cut) + offset - cut // 2, 0, size - 1)  -> window clipped at the borders
    const int oy = (int)p.cy[b], ox = (int)p.cx[b];
    const int y0 = min(max(oy - p.cut_h / 2, 0), p.H - 1), y1 = min(max(oy - p.cut_h / 2 + p.cut_h - 1, 0), p.H - 1);
    const int x0 = min(max(ox - p.cut_w / 2, 0), p.W - 1), x1 = min(max(ox - p.cut_w / 2 + p.cut_w - 1, 0), p.W - 1);
    return i >= y0 && i <= y1 && j >= x0 && j <= x1;
}

// part[b][slice] = sum over the slice's pixels (all channels) of: fwd -> x ; bwd -> gy at valid (un-cut, in-range source) outputs
template <bool BWD>
__global__ void __launch_bounds__(256) sum_kernel(Params p) {
    __shared__ float sh[32];
    const int b = blockIdx.y, hw = p.H * p.W;
    const int per = (hw + p.slices - 1) / p.slices;
    const int beg = blockIdx.x * per, end = min(hw, beg + per);
    float acc = 0.f;
    const int ty = p.ty ? (int)p.ty[b] : 0, tx = p.tx ? (int)p.tx[b] : 0;
    for (int pix = beg + threadIdx.x; pix < end; pix += blockDim.x) {
        if (BWD) {
            const int i = pix / p.W, j = pix % p.W, si = i + ty, sj = j + tx;
            if (si < 0 || si >= p.H || sj < 0 || sj >= p.W || in_cut(p, b, i, j)) continue;
        }
        for (int c = 0; c < p.C; ++c) acc += __ldg(p.x + ((long long)b * p.C + c) * hw + pix);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
        p.part[(long long)b * p.slices + blockIdx.x] = s;
    }
}

__global__ void finish_sum_kernel(const float* __restrict__ part, float* __restrict__ sums, int B, int slices) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float s = 0.f;
    for (int k = 0; k < slices; ++k) s += part[(long long)b * slices + k];
    sums[b] = s;
}

__global__ void __launch_bounds__(256) fwd_kernel(Params p) {
    const int hw = p.H * p.W;
    const long long total = (long long)p.B * hw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / hw), pix = (int)(idx % hw), i = pix / p.W, j = pix % p.W;
        const int si = i + (p.ty ? (int)p.ty[b] : 0), sj = j + (p.tx ? (int)p.tx[b] : 0);
        float* yo = p.y + (long long)b * p.C * hw + pix;
        if (si < 0 || si >= p.H || sj < 0 || sj >= p.W || in_cut(p, b, i, j)) {
            for (int c = 0; c < p.C; ++c) yo[(long long)c * hw] = 0.f;
            continue;
        }
        const float* xi = p.x + (long long)b * p.C * hw + si * p.W + sj;
        float v[MAXC];
        float m = 0.f;
        const float br = (p.rb && !p.linear_only) ? __ldg(p.rb + b) - 0.5f : 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < p.C) { v[c] = __ldg(xi + (long long)c * hw) + br; m += v[c]; }
        if (p.rb) {
            m /= (float)p.C;
            const float sa = __ldg(p.rs + b) * 2.f, co = __ldg(p.rc + b) + 0.5f;
            const float M = __ldg(p.sums + b) / ((float)p.C * (float)hw) + br;
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
                if (c < p.C) { const float x2 = (v[c] - m) * sa + m; v[c] = (x2 - M) * co + M; }
        }
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < p.C) yo[(long long)c * hw] = v[c];
    }
}

// gx[si,sj] = sa * g2 + (1 - sa) * mean_c g2,  g2 = co * g3 + (1 - co) * S / CHW,  g3 = gy at the output that read (si,sj) (or 0)
__global__ void __launch_bounds__(256) bwd_kernel(Params p) {
    const int hw = p.H * p.W;
    const long long total = (long long)p.B * hw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / hw), pix = (int)(idx % hw), si = pix / p.W, sj = pix % p.W;
        const int i = si - (p.ty ? (int)p.ty[b] : 0), j = sj - (p.tx ? (int)p.tx[b] : 0);
        const bool hit = i >= 0 && i < p.H && j >= 0 && j < p.W && !in_cut(p, b, i, j);
        const float* gi = p.x + (long long)b * p.C * hw + i * p.W + j;
        float g[MAXC];
        float co = 1.f, sa = 1.f, base = 0.f;
        if (p.rb) {
            co = __ldg(p.rc + b) + 0.5f; sa = __ldg(p.rs + b) * 2.f;
            base = (1.f - co) * __ldg(p.sums + b) / ((float)p.C * (float)hw);
        }
        float m = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < p.C) { g[c] = co * (hit ? __ldg(gi + (long long)c * hw) : 0.f) + base; m += g[c]; }
        m /= (float)p.C;
        float* go = p.y + (long long)b * p.C * hw + pix;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < p.C) go[(long long)c * hw] = p.rb ? sa * g[c] + (1.f - sa) * m : g[c];
    }
}

}  // namespace da
}  // namespace sg2

using namespace sg2;

extern "C" int64_t sg2_diffaugment_workspace(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return -1;
    return (int64_t)B * (64 + 1) * (int64_t)sizeof(float);
}

extern "C" int sg2_diffaugment(const float* x, float* y, const float* rb, const float* rs, const float* rc,
                               const int64_t* ty, const int64_t* tx, const int64_t* cy, const int64_t* cx, int cut_h, int cut_w,
                               int B, int C, int H, int W, int backward, int linear_only, void* workspace, sg2_stream_t stream) {
    SG2_REQUIRE(x && y && workspace, "diffaugment: null pointer");
    SG2_REQUIRE(B > 0 && C > 0 && C <= da::MAXC && H > 0 && W > 0, "diffaugment: need 1 <= C <= %d channels (C=%d)", da::MAXC, C);
    SG2_REQUIRE((rb == nullptr) == (rs == nullptr) && (rb == nullptr) == (rc == nullptr), "diffaugment: colour draws come as a triple");
    SG2_REQUIRE((ty == nullptr) == (tx == nullptr) && (cy == nullptr) == (cx == nullptr), "diffaugment: shifts / cut centres come in pairs");
    cudaStream_t st = (cudaStream_t)stream;
    da::Params p;
    p.x = x; p.y = y; p.rb = rb; p.rs = rs; p.rc = rc;
    p.ty = (const long long*)ty; p.tx = (const long long*)tx; p.cy = (const long long*)cy; p.cx = (const long long*)cx;
    p.B = B; p.C = C; p.H = H; p.W = W; p.cut_h = cut_h; p.cut_w = cut_w; p.linear_only = linear_only;
    p.slices = (int)std::min<long long>(64, ceil_div((long long)H * W, 1024));
    p.part = (float*)workspace;
    float* sums = (float*)workspace + (size_t)B * 64;
    p.sums = sums;
    int rc_ = SG2_OK;
    if (rb) {
        dim3 grid((unsigned)p.slices, (unsigned)B);
        if (backward) da::sum_kernel<true><<<grid, 256, 0, st>>>(p); else da::sum_kernel<false><<<grid, 256, 0, st>>>(p);
        rc_ = launched("diffaugment_sum");
        if (rc_) return rc_;
        da::finish_sum_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, st>>>(p.part, sums, B, p.slices);
        rc_ = launched("diffaugment_finish_sum");
        if (rc_) return rc_;
    }
    const int blocks = (int)std::min<long long>(ceil_div((long long)B * H * W, 256), (long long)num_sms() * 16);
    if (backward) da::bwd_kernel<<<blocks, 256, 0, st>>>(p); else da::fwd_kernel<<<blocks, 256, 0, st>>>(p);
    return launched(backward ? "diffaugment_bwd" : "diffaugment_fwd");
}
