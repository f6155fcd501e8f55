// Small-batch fully connected layers (fp32 FMA), fused with the equalised-learning-rate scale, bias, gain and leaky-ReLU.
//
// Replaces the cuBLAS/cutlass sgemm + 3 elementwise kernels that every nn.Linear of the reference path turns into:
//   * Mapping: 8 x [ELR-Linear(512,512) * lr -> LeakyReLU]            implementations/StyleGAN2/model.py:71-78, 263-282
//   * ModulatedConv2d.affine: ELR-Linear(style_dim, Ci) (20 per G pass)  model.py:102, 110
//   * Discriminator epilogue: ELR-Linear(8192,512) -> LeakyReLU -> ELR-Linear(512,1)   model.py:392-396
// plus PixelNorm (model.py:253-256).  The batch is at most a few dozen rows, so these are weight-streaming GEMV-like
// problems (HBM/L2-bound on W, never tensor-core shaped): one warp per output feature, coalesced 128-bit weight loads, the
// batch rows as register accumulators.  The three kernels form a closed family under differentiation
//   F(x, W) = x W^T,   Dx(g, W) = g W,   Dw(g, x) = g^T x
// (each one's gradients are the other two), which is what R1's double backward through the discriminator epilogue needs;
// the fused prologue/epilogue arguments (bias, gain, slope, y) are optional.
//   fwd:        y[b,n]  = act( gain * (coef * sum_k x[b,k] W[n,k] + bias[n]) ),  act = lrelu(slope) (slope 1 = linear)
//   bwd_data:   gx[b,k] = coef * sum_n gu[b,n] W[n,k],    gu = gy * gain * (y > 0 ? 1 : slope)   (y NULL: gu = gy * gain)
//   bwd_weight: gw[n,k] = coef * sum_b gu[b,n] x[b,k],    gb[n] = sum_b gu[b,n]
// No atomics: results are run-to-run identical.
#include "common.cuh"

namespace sg2 {
namespace lin {


// CTA = 16 output features (8 warps x 2) x all batch rows; the x tile [32 rows][256 k] is staged in shared memory so x is read
// once per CTA (not once per feature) and every weight row once per 32 batch rows; lanes stride k with 128-bit loads
constexpr int FWD_NT = 16, FWD_RB = 32, FWD_KC = 256;

__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ y, int B, int K, int N, float coef, float gain, float slope) {
    __shared__ __align__(16) float xs[FWD_RB][FWD_KC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n0 = blockIdx.x * FWD_NT + warp * 2;
    const bool vec = (K & 3) == 0 && (((uintptr_t)x | (uintptr_t)w) & 15) == 0;
    for (int b0 = 0; b0 < B; b0 += FWD_RB) {
        const int rows = min(FWD_RB, B - b0);
        float acc[2][FWD_RB];
#pragma unroll
        for (int r = 0; r < FWD_RB; ++r) acc[0][r] = acc[1][r] = 0.f;
        for (int k0 = 0; k0 < K; k0 += FWD_KC) {
            const int kc = min(FWD_KC, K - k0);
            __syncthreads();
            if (vec) {
                for (int i = threadIdx.x; i < FWD_RB * (FWD_KC / 4); i += 256) {
                    const int r = i / (FWD_KC / 4), c4 = (i % (FWD_KC / 4)) * 4;
                    float4 v = f4zero();
                    if (r < rows && c4 < kc) v = ldg4(x + (long long)(b0 + r) * K + k0 + c4);
                    *reinterpret_cast<float4*>(&xs[r][c4]) = v;
                }
            } else {
                for (int i = threadIdx.x; i < FWD_RB * FWD_KC; i += 256) {
                    const int r = i / FWD_KC, c = i % FWD_KC;
                    xs[r][c] = (r < rows && c < kc) ? __ldg(x + (long long)(b0 + r) * K + k0 + c) : 0.f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const int n = n0 + f;
                if (n >= N) continue;
                const float* wr = w + (long long)n * K + k0;
                for (int c = lane * 4; c < kc; c += 128) {
                    float4 wv;
                    if (vec) wv = ldg4(wr + c);
                    else {
                        wv.x = __ldg(wr + c); wv.y = c + 1 < kc ? __ldg(wr + c + 1) : 0.f;
                        wv.z = c + 2 < kc ? __ldg(wr + c + 2) : 0.f; wv.w = c + 3 < kc ? __ldg(wr + c + 3) : 0.f;
                    }
#pragma unroll
                    for (int r = 0; r < FWD_RB; ++r) {
                        const float4 xv = *reinterpret_cast<const float4*>(&xs[r][c]);
                        acc[f][r] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[f][r]))));
                    }
                }
            }
        }
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const int n = n0 + f;
            float mine = 0.f;
#pragma unroll
            for (int r = 0; r < FWD_RB; ++r) {
                const float sum = warp_sum(acc[f][r]);
                if (lane == r) mine = sum;
            }
            if (n < N && lane < rows) {
                float t = (coef * mine + (bias ? __ldg(bias + n) : 0.f)) * gain;
                t = t > 0.f ? t : t * slope;
                y[(long long)(b0 + lane) * N + n] = t;
            }
        }
    }
}

__device__ __forceinline__ float gu_of(const float* gy, const float* y, long long i, float gain, float slope) {
    const float g = __ldg(gy + i) * gain;
    return (y && !(__ldg(y + i) > 0.f)) ? g * slope : g;
}

// grid (k slices of 32 * KV, row groups of 8); 4 warps split n, reduced through shared memory.  KV = 4 (128-bit weight loads)
// when K is large enough to fill the machine that way, KV = 1 (more, narrower CTAs) for the 512-wide mapping / affine layers.
template <int KV>
__global__ void __launch_bounds__(128) linear_bwd_data_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ w,
                                                              float* __restrict__ gx, int B, int K, int N, float coef, float gain, float slope) {
    constexpr int R = 8, NC = 512, KS = 32 * KV;
    __shared__ float gu[R][NC];
    __shared__ float red[4][R][KS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = blockIdx.y * R, rows = min(R, B - b0);
    const int k0 = blockIdx.x * KS + lane * KV;
    float acc[R][KV];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int e = 0; e < KV; ++e) acc[r][e] = 0.f;
    const bool vec = KV == 4 && (K & 3) == 0 && ((uintptr_t)w & 15) == 0;
    for (int n0 = 0; n0 < N; n0 += NC) {
        const int nn = min(NC, N - n0);
        __syncthreads();
        for (int i = threadIdx.x; i < R * nn; i += 128) {
            const int r = i / nn, c = i % nn;
            gu[r][c] = r < rows ? gu_of(gy, y, (long long)(b0 + r) * N + n0 + c, gain, slope) : 0.f;
        }
        __syncthreads();
        for (int c = warp; c < nn; c += 4) {
            const float* wr = w + (long long)(n0 + c) * K;
            float wv[KV];
            if (vec) {
                const float4 t = k0 < K ? ldg4(wr + k0) : f4zero();
                wv[0] = t.x; if (KV == 4) { wv[1] = t.y; wv[2] = t.z; wv[3] = t.w; }
            } else {
#pragma unroll
                for (int e = 0; e < KV; ++e) wv[e] = k0 + e < K ? __ldg(wr + k0 + e) : 0.f;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float g = gu[r][c];
#pragma unroll
                for (int e = 0; e < KV; ++e) acc[r][e] = fmaf(g, wv[e], acc[r][e]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int e = 0; e < KV; ++e) red[warp][r][lane * KV + e] = acc[r][e];
    __syncthreads();
    for (int i = threadIdx.x; i < R * KS; i += 128) {
        const int r = i / KS, kk = i % KS;
        const int k = blockIdx.x * KS + kk;
        if (r < rows && k < K) gx[(long long)(b0 + r) * K + k] = coef * (red[0][r][kk] + red[1][r][kk] + red[2][r][kk] + red[3][r][kk]);
    }
}

// grid (k slices of 128, feature groups of 4): warp = feature n, lane = 4 consecutive k
__global__ void __launch_bounds__(128) linear_bwd_weight_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ x,
                                                                float* __restrict__ gw, float* __restrict__ gb, int B, int K, int N,
                                                                float coef, float gain, float slope) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.y * 4 + (threadIdx.x >> 5);
    if (n >= N) return;
    const int k0 = blockIdx.x * 128 + lane * 4;
    const bool vec = (K & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)gw & 15) == 0;
    float4 acc = f4zero();
    float sb = 0.f;
    for (int b = 0; b < B; ++b) {
        const float g = gu_of(gy, y, (long long)b * N + n, gain, slope);
        sb += g;
        const float* xr = x + (long long)b * K;
        float4 xv = f4zero();
        if (vec) { if (k0 < K) xv = ldg4(xr + k0); }
        else {
            xv.x = k0 < K ? __ldg(xr + k0) : 0.f; xv.y = k0 + 1 < K ? __ldg(xr + k0 + 1) : 0.f;
            xv.z = k0 + 2 < K ? __ldg(xr + k0 + 2) : 0.f; xv.w = k0 + 3 < K ? __ldg(xr + k0 + 3) : 0.f;
        }
        fma4(acc, g, xv);
    }
    float* dst = gw + (long long)n * K + k0;
    if (vec) { if (k0 < K) st4(dst, scale4(acc, coef)); }
    else {
        if (k0 < K) dst[0] = acc.x * coef;
        if (k0 + 1 < K) dst[1] = acc.y * coef;
        if (k0 + 2 < K) dst[2] = acc.z * coef;
        if (k0 + 3 < K) dst[3] = acc.w * coef;
    }
    if (gb && blockIdx.x == 0 && lane == 0) gb[n] = sb;
}

// y = x / (sqrt(mean_k x^2) + eps): one warp per row
__global__ void __launch_bounds__(128) pixelnorm_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int K, float eps) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const float* xr = x + (long long)b * K;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) { const float v = __ldg(xr + k); s = fmaf(v, v, s); }
    s = warp_sum(s);
    const float inv = 1.f / (sqrtf(s / (float)K) + eps);
    for (int k = lane; k < K; k += 32) y[(long long)b * K + k] = __ldg(xr + k) * inv;
}

}  // namespace lin
}  // namespace sg2

using namespace sg2;

extern "C" int sg2_linear_fwd(const float* x, const float* w, const float* bias, float* y, int B, int K, int N,
                              float coef, float gain, float slope, sg2_stream_t stream) {
    SG2_REQUIRE(x && w && y, "linear_fwd: null pointer");
    SG2_REQUIRE(B > 0 && K > 0 && N > 0, "linear_fwd: empty tensor");
    lin::linear_fwd_kernel<<<(unsigned)ceil_div(N, lin::FWD_NT), 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, B, K, N, coef, gain, slope);
    return launched("linear_fwd");
}

extern "C" int sg2_linear_bwd_data(const float* gy, const float* y, const float* w, float* gx, int B, int K, int N,
                                   float coef, float gain, float slope, sg2_stream_t stream) {
    SG2_REQUIRE(gy && w && gx, "linear_bwd_data: null pointer");
    SG2_REQUIRE(B > 0 && K > 0 && N > 0, "linear_bwd_data: empty tensor");
    if (ceil_div(K, 128) * ceil_div(B, 8) >= num_sms()) {
        dim3 grid((unsigned)ceil_div(K, 128), (unsigned)ceil_div(B, 8));
        lin::linear_bwd_data_kernel<4><<<grid, 128, 0, (cudaStream_t)stream>>>(gy, y, w, gx, B, K, N, coef, gain, slope);
    } else {
        dim3 grid((unsigned)ceil_div(K, 32), (unsigned)ceil_div(B, 8));
        lin::linear_bwd_data_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(gy, y, w, gx, B, K, N, coef, gain, slope);
    }
    return launched("linear_bwd_data");
}

extern "C" int sg2_linear_bwd_weight(const float* gy, const float* y, const float* x, float* gw, float* gb, int B, int K, int N,
                                     float coef, float gain, float slope, sg2_stream_t stream) {
    SG2_REQUIRE(gy && x && gw, "linear_bwd_weight: null pointer");
    SG2_REQUIRE(B > 0 && K > 0 && N > 0, "linear_bwd_weight: empty tensor");
    dim3 grid((unsigned)ceil_div(K, 128), (unsigned)ceil_div(N, 4));
    SG2_REQUIRE(grid.y <= 65535, "linear_bwd_weight: too many output features (%d)", N);
    lin::linear_bwd_weight_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(gy, y, x, gw, gb, B, K, N, coef, gain, slope);
    return launched("linear_bwd_weight");
}

extern "C" int sg2_pixelnorm(const float* x, float* y, int B, int K, float eps, sg2_stream_t stream) {
    SG2_REQUIRE(x && y && B > 0 && K > 0, "pixelnorm: bad arguments");
    lin::pixelnorm_kernel<<<(unsigned)ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(x, y, B, K, eps);
    return launched("pixelnorm");
}
