// Minibatch standard deviation (fp32), forward and first-order backward.
// Semantics: implementations/StyleGAN2/model.py:215-236 (MiniBatchStdDev.forward):
//   view [G, M, C, H, W]; biased variance over G; sqrt(var + eps); mean over (C,H,W) -> one value per
//   column m; replicated over the group and all pixels and concatenated as channel C.
// Sample i = g*M + m.  Tiny tensor ([32,512,4,4] on the hot path) -> latency-bound; the point of the
// kernel is replacing ~8 ATen launches + torch.cat by 2 launches and supporting arbitrary strides so
// the channels_last pipeline needs no layout change.
#include "common.cuh"

namespace sg2 {

struct MbstdGeom {
    int n, c, h, w, G, M;
    long long xs[4], ys[4], gs[4];
    float eps;
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
    if (wid == 0) r = warp_sum(r);
    if (threadIdx.x == 0) sh[0] = r;
    __syncthreads();
    r = sh[0];
    __syncthreads();
    return r;
}

// One block per column m: stat[m] = mean_{c,h,w} sqrt(var_g + eps).
__global__ void __launch_bounds__(512) mbstd_stat_kernel(const float* __restrict__ x, float* __restrict__ stat, MbstdGeom g) {
    __shared__ float sh[32];
    const int m = blockIdx.x;
    const int chw = g.c * g.h * g.w, hw = g.h * g.w;
    float acc = 0.f;
    for (int pos = threadIdx.x; pos < chw; pos += blockDim.x) {
        // pos enumerates (h, w, c) with c fastest when x is channels_last, (c, h, w) otherwise:
        int ch, py, px;
        if (g.xs[1] == 1) { ch = pos % g.c; int r = pos / g.c; px = r % g.w; py = r / g.w; }
        else              { px = pos % g.w; int r = pos / g.w; py = r % g.h; ch = r / g.h; }
        const float* xp = x + ch * g.xs[1] + py * g.xs[2] + px * g.xs[3];
        float mean = 0.f;
        for (int k = 0; k < g.G; ++k) mean += xp[(long long)(k * g.M + m) * g.xs[0]];
        mean /= (float)g.G;
        float var = 0.f;
        for (int k = 0; k < g.G; ++k) { float d = xp[(long long)(k * g.M + m) * g.xs[0]] - mean; var = fmaf(d, d, var); }
        var /= (float)g.G;
        acc += sqrtf(var + g.eps);
    }
    (void)hw;
    float tot = block_sum(acc, sh);
    if (threadIdx.x == 0) stat[m] = tot / (float)chw;
}

// y[:, :c] = x ; y[:, c] = stat[i % M]
__global__ void __launch_bounds__(256) mbstd_write_kernel(const float* __restrict__ x, const float* __restrict__ stat,
                                                          float* __restrict__ y, MbstdGeom g) {
    const int c1 = g.c + 1;
    const long long total = (long long)g.n * c1 * g.h * g.w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int ch, py, px, i;
        long long r = idx;
        if (g.ys[1] == 1) { ch = (int)(r % c1); r /= c1; px = (int)(r % g.w); r /= g.w; py = (int)(r % g.h); i = (int)(r / g.h); }
        else              { px = (int)(r % g.w); r /= g.w; py = (int)(r % g.h); r /= g.h; ch = (int)(r % c1); i = (int)(r / c1); }
        float v = (ch < g.c) ? x[i * g.xs[0] + ch * g.xs[1] + py * g.xs[2] + px * g.xs[3]] : stat[i % g.M];
        y[i * g.ys[0] + ch * g.ys[1] + py * g.ys[2] + px * g.ys[3]] = v;
    }
}

// grid (chunks, M).  gx[g,m,pos] = gy[g,m,pos] + gf[m] * (x_g - mean) / (G * CHW * sd)
__global__ void __launch_bounds__(256) mbstd_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                        float* __restrict__ gx, MbstdGeom g) {
    __shared__ float sh[32];
    const int m = blockIdx.y;
    const int chw = g.c * g.h * g.w, hw = g.h * g.w;
    float part = 0.f;
    for (int t = threadIdx.x; t < g.G * hw; t += blockDim.x) {
        int k = t / hw, r = t % hw, py = r / g.w, px = r % g.w;
        part += gy[(long long)(k * g.M + m) * g.gs[0] + g.c * g.gs[1] + py * g.gs[2] + px * g.gs[3]];
    }
    const float gf = block_sum(part, sh);
    const float k0 = gf / ((float)g.G * (float)chw);
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < chw; pos += gridDim.x * blockDim.x) {
        int ch, py, px;
        if (g.xs[1] == 1) { ch = pos % g.c; int r = pos / g.c; px = r % g.w; py = r / g.w; }
        else              { px = pos % g.w; int r = pos / g.w; py = r % g.h; ch = r / g.h; }
        const long long xo = ch * g.xs[1] + py * g.xs[2] + px * g.xs[3];
        float mean = 0.f;
        for (int k = 0; k < g.G; ++k) mean += x[(long long)(k * g.M + m) * g.xs[0] + xo];
        mean /= (float)g.G;
        float var = 0.f;
        for (int k = 0; k < g.G; ++k) { float d = x[(long long)(k * g.M + m) * g.xs[0] + xo] - mean; var = fmaf(d, d, var); }
        var /= (float)g.G;
        const float inv_sd = rsqrtf(var + g.eps);
        for (int k = 0; k < g.G; ++k) {
            const int i = k * g.M + m;
            float d = x[(long long)i * g.xs[0] + xo] - mean;
            float gin = gy[(long long)i * g.gs[0] + ch * g.gs[1] + py * g.gs[2] + px * g.gs[3]];
            gx[(long long)i * g.ys[0] + ch * g.ys[1] + py * g.ys[2] + px * g.ys[3]] = fmaf(k0 * inv_sd, d, gin);
        }
    }
}

static int check(const char* who, int n, int c, int h, int w, int groups) {
    if (n <= 0 || c <= 0 || h <= 0 || w <= 0) return fail(SG2_EINVAL, "%s: empty tensor", who);
    if (groups <= 0 || n % groups != 0) return fail(SG2_EINVAL, "%s: groups=%d must divide n=%d", who, groups, n);
    return SG2_OK;
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_mbstd_fwd(const float* x, const int64_t x_strides[4], float* y, const int64_t y_strides[4],
                             float* stat, int n, int c, int h, int w, int groups, float eps, sg2_stream_t stream) {
    SG2_REQUIRE(x && y && stat, "mbstd_fwd: null pointer");
    int rc = check("mbstd_fwd", n, c, h, w, groups);
    if (rc) return rc;
    MbstdGeom g;
    g.n = n; g.c = c; g.h = h; g.w = w; g.G = groups; g.M = n / groups; g.eps = eps;
    for (int i = 0; i < 4; ++i) { g.xs[i] = x_strides[i]; g.ys[i] = y_strides[i]; g.gs[i] = 0; }
    cudaStream_t st = (cudaStream_t)stream;
    mbstd_stat_kernel<<<g.M, 512, 0, st>>>(x, stat, g);
    rc = launched("mbstd_stat");
    if (rc) return rc;
    const long long total = (long long)n * (c + 1) * h * w;
    const int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 8);
    mbstd_write_kernel<<<blocks, 256, 0, st>>>(x, stat, y, g);
    return launched("mbstd_write");
}

extern "C" int sg2_mbstd_bwd(const float* x, const int64_t x_strides[4], const float* gy, const int64_t gy_strides[4],
                             float* gx, const int64_t gx_strides[4], int n, int c, int h, int w, int groups,
                             float eps, sg2_stream_t stream) {
    SG2_REQUIRE(x && gy && gx, "mbstd_bwd: null pointer");
    int rc = check("mbstd_bwd", n, c, h, w, groups);
    if (rc) return rc;
    MbstdGeom g;
    g.n = n; g.c = c; g.h = h; g.w = w; g.G = groups; g.M = n / groups; g.eps = eps;
    for (int i = 0; i < 4; ++i) { g.xs[i] = x_strides[i]; g.gs[i] = gy_strides[i]; g.ys[i] = gx_strides[i]; }
    const int chw = c * h * w;
    dim3 grid((unsigned)std::min<long long>(ceil_div(chw, 256), 64), (unsigned)g.M);
    mbstd_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, gy, gx, g);
    return launched("mbstd_bwd");
}
