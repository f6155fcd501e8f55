// Kernels of the ADA augmentation pipeline (thirdparty/ada/augment.py:115-427) that have no counterpart in the StyleGAN2 step:
//   reflect_pad        torch.nn.functional.pad(mode='reflect') of augment.py:284, forward and exact adjoint (gather form)
//   affine_sample      affine_grid + grid_sample(bilinear, zeros, align_corners=False) of augment.py:293-295 as ONE pass: the
//                      sampling grid is an affine function of the output pixel, so it is evaluated in registers instead of
//                      being materialised ([B, H, W, 2] floats = 2/3 of the image bytes at 3 channels); the adjoint scatters
//                      with fp32 atomics like ATen's grid_sampler_2d_backward (thirdparty/stylegan3_ops/ops/grid_sample_gradfix.py)
//   color_affine       the per-sample 3x4 colour matrix of augment.py:352-361, forward and transpose
// All three are linear in the image, each adjoint's own derivative is the forward kernel: the ops are closed under
// differentiation (what grid_sample_gradfix exists for in the reference).  NCHW fp32, HBM-bound.
#include "common.cuh"

namespace sg2 {
namespace ada {

__device__ __forceinline__ int reflect(int i, int n) {          // index into [0, n) by reflection without edge repeat
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(256) reflect_pad_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes,
                                                          int h, int w, int px0, int px1, int py0, int py1) {
    const int oh = h + py0 + py1, ow = w + px0 + px1;
    const long long total = planes * oh * ow;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % ow), i = (int)((idx / ow) % oh);
        const long long pl = idx / ((long long)ow * oh);
        y[idx] = __ldg(x + (pl * h + reflect(i - py0, h)) * w + reflect(j - px0, w));
    }
}

// adjoint: every input pixel gathers the (up to 3 x 3) padded positions that read it
__global__ void __launch_bounds__(256) reflect_pad_adj_kernel(const float* __restrict__ gy, float* __restrict__ gx, long long planes,
                                                              int h, int w, int px0, int px1, int py0, int py1) {
    const int oh = h + py0 + py1, ow = w + px0 + px1;
    const long long total = planes * h * w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % w), i = (int)((idx / w) % h);
        const long long pl = idx / ((long long)w * h);
        int oy[3], ox[3], ny = 0, nx = 0;
        oy[ny++] = i + py0;
        if (i >= 1 && py0 - i >= 0) oy[ny++] = py0 - i;
        if (i <= h - 2 && py0 + 2 * (h - 1) - i < oh) oy[ny++] = py0 + 2 * (h - 1) - i;
        ox[nx++] = j + px0;
        if (j >= 1 && px0 - j >= 0) ox[nx++] = px0 - j;
        if (j <= w - 2 && px0 + 2 * (w - 1) - j < ow) ox[nx++] = px0 + 2 * (w - 1) - j;
        float s = 0.f;
        for (int a = 0; a < ny; ++a)
            for (int b = 0; b < nx; ++b) s += __ldg(gy + (pl * oh + oy[a]) * ow + ox[b]);
        gx[idx] = s;
    }
}

// torch.linspace(-1, 1, n) * (n - 1) / n, evaluated the way ATen does (two-sided, so the grid is symmetric)
__device__ __forceinline__ float base_coord(int i, int n) {
    if (n <= 1) return 0.f;
    const float step = 2.f / (float)(n - 1);
    const float v = i < n / 2 ? fmaf(step, (float)i, -1.f) : 1.f - step * (float)(n - 1 - i);
    return v * (float)(n - 1) / (float)n;
}

struct SampleGeom { int n, c, ih, iw, oh, ow; };

template <bool ADJ>
__global__ void __launch_bounds__(256) affine_sample_kernel(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ theta, SampleGeom g) {
    const long long total = (long long)g.n * g.oh * g.ow;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % g.ow), i = (int)((idx / g.ow) % g.oh), b = (int)(idx / ((long long)g.ow * g.oh));
        const float* t = theta + b * 6;
        const float xn = base_coord(j, g.ow), yn = base_coord(i, g.oh);
        const float sx = fmaf(__ldg(t + 0), xn, fmaf(__ldg(t + 1), yn, __ldg(t + 2)));
        const float sy = fmaf(__ldg(t + 3), xn, fmaf(__ldg(t + 4), yn, __ldg(t + 5)));
        const float fx = ((sx + 1.f) * (float)g.iw - 1.f) * 0.5f, fy = ((sy + 1.f) * (float)g.ih - 1.f) * 0.5f;
        const float x0f = floorf(fx), y0f = floorf(fy);
        const int x0 = (int)x0f, y0 = (int)y0f;
        const float wx1 = fx - x0f, wy1 = fy - y0f, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
        const bool vx0 = x0 >= 0 && x0 < g.iw, vx1 = x0 + 1 >= 0 && x0 + 1 < g.iw;
        const bool vy0 = y0 >= 0 && y0 < g.ih, vy1 = y0 + 1 >= 0 && y0 + 1 < g.ih;
        const long long ipl = (long long)g.ih * g.iw, opl = (long long)g.oh * g.ow;
        for (int ch = 0; ch < g.c; ++ch) {
            const long long ibase = ((long long)b * g.c + ch) * ipl, o = ((long long)b * g.c + ch) * opl + (long long)i * g.ow + j;
            if (!ADJ) {
                const float* p = src + ibase;
                float v = 0.f;
                if (vy0 && vx0) v = fmaf(wy0 * wx0, __ldg(p + (long long)y0 * g.iw + x0), v);
                if (vy0 && vx1) v = fmaf(wy0 * wx1, __ldg(p + (long long)y0 * g.iw + x0 + 1), v);
                if (vy1 && vx0) v = fmaf(wy1 * wx0, __ldg(p + (long long)(y0 + 1) * g.iw + x0), v);
                if (vy1 && vx1) v = fmaf(wy1 * wx1, __ldg(p + (long long)(y0 + 1) * g.iw + x0 + 1), v);
                dst[o] = v;
            } else {
                const float gv = __ldg(src + o);
                float* p = dst + ibase;
                if (vy0 && vx0) atomicAdd(p + (long long)y0 * g.iw + x0, wy0 * wx0 * gv);
                if (vy0 && vx1) atomicAdd(p + (long long)y0 * g.iw + x0 + 1, wy0 * wx1 * gv);
                if (vy1 && vx0) atomicAdd(p + (long long)(y0 + 1) * g.iw + x0, wy1 * wx0 * gv);
                if (vy1 && vx1) atomicAdd(p + (long long)(y0 + 1) * g.iw + x0 + 1, wy1 * wx1 * gv);
            }
        }
    }
}

// y[b, :, p] = M[b] x[b, :, p] + t[b]  (M = C[:, :3, :3], t = C[:, :3, 3]; transpose: y = M^T x, no offset); C: [B, 4, 4]
__global__ void __launch_bounds__(256) color_affine_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ cm,
                                                           int n, long long hw, int transpose) {
    const long long total = (long long)n * hw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / hw);
        const long long p = idx % hw;
        const float* c = cm + b * 16;
        const float* xi = x + (long long)b * 3 * hw + p;
        const float r = __ldg(xi), gch = __ldg(xi + hw), bl = __ldg(xi + 2 * hw);
        float* yo = y + (long long)b * 3 * hw + p;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v;
            if (!transpose) v = fmaf(__ldg(c + 4 * k), r, fmaf(__ldg(c + 4 * k + 1), gch, fmaf(__ldg(c + 4 * k + 2), bl, __ldg(c + 4 * k + 3))));
            else v = fmaf(__ldg(c + k), r, fmaf(__ldg(c + 4 + k), gch, __ldg(c + 8 + k) * bl));
            yo[(long long)k * hw] = v;
        }
    }
}

static int grid_for(long long total) { return (int)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 16); }

}  // namespace ada
}  // namespace sg2

using namespace sg2;

extern "C" int sg2_reflect_pad(const float* x, float* y, int64_t planes, int h, int w, int px0, int px1, int py0, int py1,
                               int adjoint, sg2_stream_t stream) {
    SG2_REQUIRE(x && y && planes > 0 && h > 0 && w > 0, "reflect_pad: bad arguments");
    SG2_REQUIRE(px0 >= 0 && px1 >= 0 && py0 >= 0 && py1 >= 0 && px0 < w && px1 < w && py0 < h && py1 < h,
                "reflect_pad: padding (%d,%d,%d,%d) must be non-negative and smaller than the image %dx%d", px0, px1, py0, py1, h, w);
    cudaStream_t st = (cudaStream_t)stream;
    if (adjoint) {   // x = gy [planes, h+py, w+px], y = gx [planes, h, w]
        ada::reflect_pad_adj_kernel<<<ada::grid_for(planes * h * w), 256, 0, st>>>(x, y, planes, h, w, px0, px1, py0, py1);
        return launched("reflect_pad_adj");
    }
    ada::reflect_pad_kernel<<<ada::grid_for(planes * (h + py0 + py1) * (long long)(w + px0 + px1)), 256, 0, st>>>(x, y, planes, h, w, px0, px1, py0, py1);
    return launched("reflect_pad");
}

extern "C" int sg2_affine_sample(const float* x, float* y, const float* theta, int n, int c, int ih, int iw, int oh, int ow,
                                 int adjoint, sg2_stream_t stream) {
    SG2_REQUIRE(x && y && theta, "affine_sample: null pointer");
    SG2_REQUIRE(n > 0 && c > 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0, "affine_sample: empty tensor");
    cudaStream_t st = (cudaStream_t)stream;
    ada::SampleGeom g{n, c, ih, iw, oh, ow};
    const int blocks = ada::grid_for((long long)n * oh * ow);
    if (adjoint) {   // x = gy [n,c,oh,ow], y = gx [n,c,ih,iw] (zeroed here, then scattered into)
        SG2_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)n * c * ih * iw, st));
        ada::affine_sample_kernel<true><<<blocks, 256, 0, st>>>(x, y, theta, g);
        return launched("affine_sample_adj");
    }
    ada::affine_sample_kernel<false><<<blocks, 256, 0, st>>>(x, y, theta, g);
    return launched("affine_sample");
}

extern "C" int sg2_color_affine(const float* x, float* y, const float* cmat, int n, int64_t hw, int transpose, sg2_stream_t stream) {
    SG2_REQUIRE(x && y && cmat && n > 0 && hw > 0, "color_affine: bad arguments");
    ada::color_affine_kernel<<<ada::grid_for((long long)n * hw), 256, 0, (cudaStream_t)stream>>>(x, y, cmat, n, hw, transpose);
    return launched("color_affine");
}
