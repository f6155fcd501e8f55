// StyleGAN2 resampling kernels (fp32): fused bilinear-x2 (+3x3 binomial blur) and 2x2 average pool.
//
// Reference path being replaced (two ATen kernels each way, 13x|x| bytes of traffic vs 5x|x| fused):
//   implementations/StyleGAN2/model.py:56-58   nn.Upsample(scale_factor=2, 'bilinear', align_corners=False)
//   implementations/StyleGAN2/model.py:138-149 Blur2d: depthwise [1,2,1]x[1,2,1]/16, zero padding 1
//   implementations/StyleGAN2/model.py:61-63   nn.AvgPool2d(2);  model.py:209-212  (x + t)/sqrt(2)
//
// Per axis (length n -> 2n) the fused operator is, for output j = 2i + p:
//   bilinear:  U(2i) = .25 x[c(i-1)] + .75 x[i],  U(2i+1) = .75 x[i] + .25 x[c(i+1)],  c = clamp to [0,n-1]
//   blur:      z[j] = (U(j-1) + 2 U(j) + U(j+1)) / 4,  U(m) = 0 for m outside [0, 2n)   (zero padding)
// which collapses to 3 taps on (x[c(i-1)], x[i], x[c(i+1)]):
//   p=0: (5,10,1)/16   p=1: (1,10,5)/16   except j=0: (0,11,1)/16 and j=2n-1: (1,11,0)/16
// (the replicate-then-zero border).  2-D weights are the outer product.  HBM-bound: bytes = |x| + 4|x|.
#include "common.cuh"

namespace sg2 {

__device__ __forceinline__ void axis_w(int j, int n, int blur, float w[3]) {
    const int p = j & 1;
    if (!blur) { w[0] = p ? 0.f : 0.25f; w[1] = 0.75f; w[2] = p ? 0.25f : 0.f; return; }
    if (j == 0)              { w[0] = 0.f;           w[1] = 11.f / 16.f; w[2] = 1.f / 16.f; }
    else if (j == 2 * n - 1) { w[0] = 1.f / 16.f;    w[1] = 11.f / 16.f; w[2] = 0.f; }
    else if (p == 0)         { w[0] = 5.f / 16.f;    w[1] = 10.f / 16.f; w[2] = 1.f / 16.f; }
    else                     { w[0] = 1.f / 16.f;    w[1] = 10.f / 16.f; w[2] = 5.f / 16.f; }
}

// Adjoint weights: input i receives from outputs j = 2i-2 .. 2i+3 (clamped reads fold onto the border).
__device__ __forceinline__ void axis_w_adj(int i, int n, int blur, float A[6]) {
#pragma unroll
    for (int jj = 0; jj < 6; ++jj) {
        const int j = 2 * i - 2 + jj;
        float a = 0.f;
        if (j >= 0 && j < 2 * n) {
            float w[3];
            axis_w(j, n, blur, w);
            const int ij = j >> 1;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                int src = min(max(ij + t - 1, 0), n - 1);
                if (src == i) a += w[t];
            }
        }
        A[jj] = a;
    }
}

struct F1 {  // scalar lane (NCHW planes)
    typedef float V;
    static __device__ __forceinline__ V ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void st(float* p, V v) { __stcs(p, v); }
    static __device__ __forceinline__ V zero() { return 0.f; }
    static __device__ __forceinline__ void fma(V& a, float s, V v) { a = fmaf(s, v, a); }
    static __device__ __forceinline__ V mul(V a, V b) { return a * b; }
};
struct F4 {  // 4 channels (NHWC)
    typedef float4 V;
    static __device__ __forceinline__ V ld(const float* p) { return ldg4(p); }
    static __device__ __forceinline__ void st(float* p, V v) { st4_cs(p, v); }
    static __device__ __forceinline__ V zero() { return f4zero(); }
    static __device__ __forceinline__ void fma(V& a, float s, V v) { fma4(a, s, v); }
    static __device__ __forceinline__ V mul(V a, V b) { return mul4(a, b); }
};

// Geometry shared by both layouts: element offset = img*img_stride + (y*w + x)*pix_stride + lane*lane_stride
//   NHWC: lanes = c/4 quads, lane_stride 4, pix_stride c, img_stride h*w*c
//   NCHW: lanes = 1, planes = n*c (img index = plane), pix_stride 1, img_stride h*w
struct ResampleGeom {
    int imgs, lanes, h, w;           // input spatial size h x w
    long long pix_stride, lane_stride;
    int c;                           // channels (for scale lookup)
    int nhwc;
};

template <class L>
__global__ void __launch_bounds__(256) up2x_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                       const float* __restrict__ scale, ResampleGeom g, int blur) {
    typedef typename L::V V;
    const long long total = (long long)g.imgs * g.h * g.w * g.lanes;
    const int H = g.h, W = g.w, OW = 2 * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int lane, ix, iy, img;
        long long r = idx;
        if (g.nhwc) { lane = (int)(r % g.lanes); r /= g.lanes; ix = (int)(r % W); r /= W; iy = (int)(r % H); img = (int)(r / H); }
        else        { lane = 0; ix = (int)(r % W); r /= W; iy = (int)(r % H); img = (int)(r / H); }
        const float* xi = x + (long long)img * H * W * g.pix_stride + lane * g.lane_stride;
        float* yi = y + (long long)img * 4 * H * W * g.pix_stride + lane * g.lane_stride;
        const int ym = max(iy - 1, 0), yp = min(iy + 1, H - 1), xm = max(ix - 1, 0), xp = min(ix + 1, W - 1);
        const int rows[3] = {ym, iy, yp}, cols[3] = {xm, ix, xp};
        float wx0[3], wx1[3], wy0[3], wy1[3];
        axis_w(2 * ix, W, blur, wx0); axis_w(2 * ix + 1, W, blur, wx1);
        axis_w(2 * iy, H, blur, wy0); axis_w(2 * iy + 1, H, blur, wy1);
        V h0[3], h1[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            h0[a] = L::zero(); h1[a] = L::zero();
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                V v = L::ld(xi + ((long long)rows[a] * W + cols[b]) * g.pix_stride);
                L::fma(h0[a], wx0[b], v); L::fma(h1[a], wx1[b], v);
            }
        }
        V o00 = L::zero(), o01 = L::zero(), o10 = L::zero(), o11 = L::zero();
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            L::fma(o00, wy0[a], h0[a]); L::fma(o01, wy0[a], h1[a]);
            L::fma(o10, wy1[a], h0[a]); L::fma(o11, wy1[a], h1[a]);
        }
        if (scale) {
            V s;
            if (g.nhwc) s = L::ld(scale + (long long)img * g.c + lane * g.lane_stride);
            else        s = L::ld(scale + img);      // plane index == n*c + c
            o00 = L::mul(o00, s); o01 = L::mul(o01, s); o10 = L::mul(o10, s); o11 = L::mul(o11, s);
        }
        float* o = yi + ((long long)(2 * iy) * OW + 2 * ix) * g.pix_stride;
        L::st(o, o00); L::st(o + g.pix_stride, o01);
        o += (long long)OW * g.pix_stride;
        L::st(o, o10); L::st(o + g.pix_stride, o11);
    }
}

template <class L>
__global__ void __launch_bounds__(256) up2x_adj_kernel(const float* __restrict__ gy, float* __restrict__ gx,
                                                       const float* __restrict__ scale, ResampleGeom g, int blur) {
    typedef typename L::V V;
    const long long total = (long long)g.imgs * g.h * g.w * g.lanes;
    const int H = g.h, W = g.w, OW = 2 * W, OH = 2 * H;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int lane, ix, iy, img;
        long long r = idx;
        if (g.nhwc) { lane = (int)(r % g.lanes); r /= g.lanes; ix = (int)(r % W); r /= W; iy = (int)(r % H); img = (int)(r / H); }
        else        { lane = 0; ix = (int)(r % W); r /= W; iy = (int)(r % H); img = (int)(r / H); }
        const float* gi = gy + (long long)img * 4 * H * W * g.pix_stride + lane * g.lane_stride;
        float Ay[6], Ax[6];
        axis_w_adj(iy, H, blur, Ay); axis_w_adj(ix, W, blur, Ax);
        V acc = L::zero();
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const int jy = 2 * iy - 2 + a;
            if (jy < 0 || jy >= OH || Ay[a] == 0.f) continue;
            V row = L::zero();
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                const int jx = 2 * ix - 2 + b;
                if (jx < 0 || jx >= OW || Ax[b] == 0.f) continue;
                L::fma(row, Ax[b], L::ld(gi + ((long long)jy * OW + jx) * g.pix_stride));
            }
            L::fma(acc, Ay[a], row);
        }
        if (scale) {
            V s;
            if (g.nhwc) s = L::ld(scale + (long long)img * g.c + lane * g.lane_stride);
            else        s = L::ld(scale + img);
            acc = L::mul(acc, s);
        }
        L::st(gx + (long long)img * H * W * g.pix_stride + lane * g.lane_stride + ((long long)iy * W + ix) * g.pix_stride, acc);
    }
}


// ---- channels_last strip kernels: a block owns (image, an x range, a strip of ROWS input rows); a thread owns one
// x position x 4 channels and walks down the strip keeping the horizontally filtered rows in registers, so every
// input row is loaded once (3 x 128-bit loads per input pixel forward, 12 per output pixel for the adjoint) and all
// index arithmetic is 32-bit and hoisted out of the loop.
constexpr int kStripRows = 16;

struct H2 { float4 a, b; };     // the two horizontal phases (outputs 2ix, 2ix+1) of one input row

__global__ void __launch_bounds__(256) up2x_fwd_strip_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                             const float* __restrict__ scale, int c, int h, int w, int blur) {
    const int cq = c >> 2, xs = 256 / cq;
    const int q = threadIdx.x % cq, ix = blockIdx.x * xs + threadIdx.x / cq;
    if (ix >= w) return;
    const int img = blockIdx.z, iy0 = blockIdx.y * kStripRows, iy1 = min(h, iy0 + kStripRows);
    const int xm = max(ix - 1, 0), xp = min(ix + 1, w - 1);
    float wx0[3], wx1[3];
    axis_w(2 * ix, w, blur, wx0); axis_w(2 * ix + 1, w, blur, wx1);
    const float* xi = x + (size_t)img * h * w * c + 4 * q;
    float* yi = y + (size_t)img * 4 * h * w * c + 4 * q;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (scale) sc = ldg4(scale + (size_t)img * c + 4 * q);
    auto hrow = [&](int iy) {
        const float* r = xi + (size_t)iy * w * c;
        const float4 vm = ldg4(r + xm * c), v0 = ldg4(r + ix * c), vp = ldg4(r + xp * c);
        H2 o;
        o.a = scale4(vm, wx0[0]); fma4(o.a, wx0[1], v0); fma4(o.a, wx0[2], vp);
        o.b = scale4(vm, wx1[0]); fma4(o.b, wx1[1], v0); fma4(o.b, wx1[2], vp);
        o.a = mul4(o.a, sc); o.b = mul4(o.b, sc);
        return o;
    };
    H2 A = hrow(max(iy0 - 1, 0)), B = hrow(iy0);
    const int ow_c = 2 * w * c;
    for (int iy = iy0; iy < iy1; ++iy) {
        const H2 Cn = hrow(min(iy + 1, h - 1));
        float wy0[3], wy1[3];
        axis_w(2 * iy, h, blur, wy0); axis_w(2 * iy + 1, h, blur, wy1);
        float4 o00 = scale4(A.a, wy0[0]), o01 = scale4(A.b, wy0[0]), o10 = scale4(A.a, wy1[0]), o11 = scale4(A.b, wy1[0]);
        fma4(o00, wy0[1], B.a); fma4(o01, wy0[1], B.b); fma4(o10, wy1[1], B.a); fma4(o11, wy1[1], B.b);
        fma4(o00, wy0[2], Cn.a); fma4(o01, wy0[2], Cn.b); fma4(o10, wy1[2], Cn.a); fma4(o11, wy1[2], Cn.b);
        float* o = yi + (size_t)(2 * iy) * ow_c + (size_t)(2 * ix) * c;
        st4_cs(o, o00); st4_cs(o + c, o01);
        st4_cs(o + ow_c, o10); st4_cs(o + ow_c + c, o11);
        A = B; B = Cn;
    }
}

__global__ void __launch_bounds__(256) up2x_adj_strip_kernel(const float* __restrict__ gy, float* __restrict__ gx,
                                                             const float* __restrict__ scale, int c, int h, int w, int blur) {
    const int cq = c >> 2, xs = 256 / cq;
    const int q = threadIdx.x % cq, ix = blockIdx.x * xs + threadIdx.x / cq;
    if (ix >= w) return;
    const int img = blockIdx.z, iy0 = blockIdx.y * kStripRows, iy1 = min(h, iy0 + kStripRows);
    const int OH = 2 * h, OW = 2 * w;
    float Ax[6];
    axis_w_adj(ix, w, blur, Ax);
    const float* gi = gy + (size_t)img * 4 * h * w * c + 4 * q;
    // column offsets clamped into the image (weights of out-of-range columns are zeroed) so that the 6 loads of a row
    // are unconditional and can all be in flight together
    int jxo[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const int jx = 2 * ix - 2 + b;
        if (jx < 0 || jx >= OW) Ax[b] = 0.f;
        jxo[b] = min(max(jx, 0), OW - 1) * c;
    }
    // horizontally reduced row jy of gy (zero outside the image)
    auto hrow = [&](int jy) {
        const float keep = (jy >= 0 && jy < OH) ? 1.f : 0.f;
        const float* row = gi + (size_t)min(max(jy, 0), OH - 1) * OW * c;
        float4 v[6];
#pragma unroll
        for (int b = 0; b < 6; ++b) v[b] = ldg4(row + jxo[b]);
        float4 r = scale4(v[0], Ax[0] * keep);
#pragma unroll
        for (int b = 1; b < 6; ++b) fma4_x2(r, Ax[b] * keep, v[b]);
        return r;
    };
    float4 H[6];
#pragma unroll
    for (int a = 0; a < 4; ++a) H[a + 2] = hrow(2 * iy0 - 2 + a);      // rows 2iy0-2 .. 2iy0+1 sit in H[2..5] before the first shift
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (scale) sc = ldg4(scale + (size_t)img * c + 4 * q);
    float* go = gx + (size_t)img * h * w * c + 4 * q;
    for (int iy = iy0; iy < iy1; ++iy) {
        H[0] = H[2]; H[1] = H[3]; H[2] = H[4]; H[3] = H[5];
        H[4] = hrow(2 * iy + 2); H[5] = hrow(2 * iy + 3);
        float Ay[6];
        axis_w_adj(iy, h, blur, Ay);
        float4 acc = f4zero();
#pragma unroll
        for (int a = 0; a < 6; ++a) fma4_x2(acc, Ay[a], H[a]);
        st4_cs(go + ((size_t)iy * w + ix) * c, mul4(acc, sc));
    }
}

// y = alpha * (avg2x2(x) + avg2x2(t)),  NHWC, one thread = one output pixel x 4 channels.
// t_pooled: t already has the output resolution and is added as it is: y = alpha * (avg2x2(x) + t).
// signs (optional): one uint16 per (output pixel, 4-channel group) = the signs (x > 0) of the 2x2 window, bit 4*p + e for window
// pixel p (row-major) and channel e -- what the leaky-ReLU gradient pass of the layer that produced x needs of it (planes.cu).
__global__ void __launch_bounds__(256) avgpool2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                           float* __restrict__ y, unsigned short* __restrict__ signs, float alpha,
                                                           int n, int c, int h, int w, int t_pooled) {
    const int cq = c >> 2, oh = h >> 1, ow = w >> 1;
    const long long total = (long long)n * oh * ow * cq;
    const float k = 0.25f * alpha;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int q = (int)(idx % cq);
        long long r = idx / cq;
        int ox = (int)(r % ow); r /= ow;
        int oy = (int)(r % oh);
        int b = (int)(r / oh);
        const long long base = (((long long)b * h + 2 * oy) * w + 2 * ox) * c + 4 * q;
        const long long rs = (long long)w * c;
        const float4 x00 = ldg4(x + base), x01 = ldg4(x + base + c), x10 = ldg4(x + base + rs), x11 = ldg4(x + base + rs + c);
        float4 a = add4(add4(x00, x01), add4(x10, x11));
        const long long o = (((long long)b * oh + oy) * ow + ox) * c + 4 * q;
        if (signs) {
            auto bits = [](const float4& v) { return (unsigned)(v.x > 0.f) | ((unsigned)(v.y > 0.f) << 1) | ((unsigned)(v.z > 0.f) << 2) | ((unsigned)(v.w > 0.f) << 3); };
            signs[o >> 2] = (unsigned short)(bits(x00) | (bits(x01) << 4) | (bits(x10) << 8) | (bits(x11) << 12));
        }
        if (t && !t_pooled) a = add4(a, add4(add4(ldg4(t + base), ldg4(t + base + c)), add4(ldg4(t + base + rs), ldg4(t + base + rs + c))));
        a = scale4(a, k);
        if (t && t_pooled) { const float4 tv = ldg4(t + o); a.x = fmaf(alpha, tv.x, a.x); a.y = fmaf(alpha, tv.y, a.y); a.z = fmaf(alpha, tv.z, a.z); a.w = fmaf(alpha, tv.w, a.w); }
        st4_cs(y + o, a);
    }
}

// gx[b, y, x, :] = alpha/4 * gy[b, y/2, x/2, :],  one thread = one OUTPUT (pooled) pixel, writes its 2x2 window.
__global__ void __launch_bounds__(256) avgpool2_adj_kernel(const float* __restrict__ gy, float* __restrict__ gx,
                                                           float alpha, int n, int c, int h, int w) {
    const int cq = c >> 2, oh = h >> 1, ow = w >> 1;
    const long long total = (long long)n * oh * ow * cq;
    const float k = 0.25f * alpha;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int q = (int)(idx % cq);
        long long r = idx / cq;
        int ox = (int)(r % ow); r /= ow;
        int oy = (int)(r % oh);
        int b = (int)(r / oh);
        float4 g = scale4(ldg4(gy + (((long long)b * oh + oy) * ow + ox) * c + 4 * q), k);
        const long long base = (((long long)b * h + 2 * oy) * w + 2 * ox) * c + 4 * q;
        const long long rs = (long long)w * c;
        st4_cs(gx + base, g); st4_cs(gx + base + c, g); st4_cs(gx + base + rs, g); st4_cs(gx + base + rs + c, g);
    }
}

static int up2x_launch(bool adj, const float* a, float* b, const float* scale, int n, int c, int h, int w,
                       int nhwc, int blur, cudaStream_t st) {
    ResampleGeom g;
    g.h = h; g.w = w; g.c = c; g.nhwc = nhwc;
    const bool vec = nhwc && (c % 4 == 0) && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) &&
                     (!scale || (uintptr_t)scale % 16 == 0);
    if (nhwc && !vec) return fail(SG2_ENOTSUP, "up2x: channels_last path needs C %% 4 == 0 and 16-byte aligned pointers (C=%d)", c);
    if (nhwc) { g.imgs = n; g.lanes = c / 4; g.pix_stride = c; g.lane_stride = 4; }
    else      { g.imgs = n * c; g.lanes = 1; g.pix_stride = 1; g.lane_stride = 0; }
    const int cq = c / 4;
    if (nhwc && cq >= 1 && cq <= 256 && (256 % cq) == 0) {
        const int xs = 256 / cq;
        dim3 grid((unsigned)ceil_div(w, xs), (unsigned)ceil_div(h, kStripRows), (unsigned)n);
        if (adj) up2x_adj_strip_kernel<<<grid, 256, 0, st>>>(a, b, scale, c, h, w, blur);
        else     up2x_fwd_strip_kernel<<<grid, 256, 0, st>>>(a, b, scale, c, h, w, blur);
        return launched(adj ? "up2x_adj_strip" : "up2x_fwd_strip");
    }
    const long long total = (long long)g.imgs * h * w * g.lanes;
    const int threads = 256;
    const int blocks = (int)std::min<long long>(ceil_div(total, threads), (long long)num_sms() * 16);
    if (nhwc) {
        if (adj) up2x_adj_kernel<F4><<<blocks, threads, 0, st>>>(a, b, scale, g, blur);
        else     up2x_fwd_kernel<F4><<<blocks, threads, 0, st>>>(a, b, scale, g, blur);
    } else {
        if (adj) up2x_adj_kernel<F1><<<blocks, threads, 0, st>>>(a, b, scale, g, blur);
        else     up2x_fwd_kernel<F1><<<blocks, threads, 0, st>>>(a, b, scale, g, blur);
    }
    return launched(adj ? "up2x_adj" : "up2x_fwd");
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_up2x_fwd(const float* x, float* y, const float* scale, int n, int c, int h, int w,
                            int nhwc, int blur, sg2_stream_t stream) {
    SG2_REQUIRE(x && y, "up2x_fwd: null pointer");
    SG2_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "up2x_fwd: empty tensor");
    SG2_REQUIRE((long long)n * c * h * w * 4 <= 2147483647LL, "up2x_fwd: tensor is too large");
    return up2x_launch(false, x, y, scale, n, c, h, w, nhwc, blur, (cudaStream_t)stream);
}

extern "C" int sg2_up2x_adj(const float* gy, float* gx, const float* scale, int n, int c, int h, int w,
                            int nhwc, int blur, sg2_stream_t stream) {
    SG2_REQUIRE(gy && gx, "up2x_adj: null pointer");
    SG2_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "up2x_adj: empty tensor");
    SG2_REQUIRE((long long)n * c * h * w * 4 <= 2147483647LL, "up2x_adj: tensor is too large");
    return up2x_launch(true, gy, gx, scale, n, c, h, w, nhwc, blur, (cudaStream_t)stream);
}

extern "C" int sg2_avgpool2_fwd(const float* x, const float* t, float* y, void* signs, float alpha,
                                int n, int c, int h, int w, int t_pooled, sg2_stream_t stream) {
    SG2_REQUIRE(x && y, "avgpool2_fwd: null pointer");
    SG2_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "avgpool2_fwd: h and w must be even and positive");
    SG2_REQUIRE(c % 4 == 0, "avgpool2_fwd: C %% 4 != 0 (C=%d)", c);
    const long long total = (long long)n * (h / 2) * (w / 2) * (c / 4);
    const int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 16);
    avgpool2_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, t, y, (unsigned short*)signs, alpha, n, c, h, w, t_pooled);
    return launched("avgpool2_fwd");
}

extern "C" int sg2_avgpool2_adj(const float* gy, float* gx, float alpha,
                                int n, int c, int h, int w, sg2_stream_t stream) {
    SG2_REQUIRE(gy && gx, "avgpool2_adj: null pointer");
    SG2_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "avgpool2_adj: h and w must be even and positive");
    SG2_REQUIRE(c % 4 == 0, "avgpool2_adj: C %% 4 != 0 (C=%d)", c);
    const long long total = (long long)n * (h / 2) * (w / 2) * (c / 4);
    const int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 16);
    avgpool2_adj_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(gy, gx, alpha, n, c, h, w);
    return launched("avgpool2_adj");
}
