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


// CTA = FPW * 8 output features (8 warps, FPW features each) x a slice of FWD_KS input features x all batch rows (32 at a time).
// A warp keeps its weight-row slices in registers (loaded first, so their latency hides behind the staging of x); x is staged
// in shared memory as [32 rows][256 k] tiles; the 32 row sums of a feature are reduced across the lanes by a halving butterfly
// (31 shuffles instead of 32 full reductions).  K larger than one slice is split over blockIdx.y: the CTAs write partial sums
// and linear_fwd_finish_kernel adds them in slice order (fixed order -> run-to-run identical) and applies the epilogue.
constexpr int FWD_RB = 32, FWD_KC = 256, FWD_KS = 512;

// v[r] (r = 0..31) holds this lane's partial of row r; returns the full sum of row `lane`
__device__ __forceinline__ float row_sums_to_lanes(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            // keep the half of the values whose row index has bit `off` equal to this lane's; hand the other half over
            const float keep = upper ? v[i + off] : v[i], send = upper ? v[i] : v[i + off];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

template <int FPW>
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ y, float* __restrict__ partial, int B, int K, int N, float coef,
                                                         float gain, float slope) {
    __shared__ __align__(16) float xs[FWD_RB][FWD_KC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n0 = (blockIdx.x * 8 + warp) * FPW;
    const int ks0 = blockIdx.y * FWD_KS, ks1 = min(K, ks0 + FWD_KS);
    const bool vec = (K & 3) == 0 && (((uintptr_t)x | (uintptr_t)w) & 15) == 0;
    // this warp's weight slices: [feature][256-tile][128-column half] float4 per lane
    float4 wv[FPW][FWD_KS / FWD_KC][2];
#pragma unroll
    for (int f = 0; f < FPW; ++f)
#pragma unroll
        for (int t = 0; t < FWD_KS / FWD_KC; ++t)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + f, k = ks0 + t * FWD_KC + h * 128 + lane * 4;
                float4 v = f4zero();
                if (n < N && k < ks1) {
                    const float* wr = w + (long long)n * K + k;
                    if (vec) v = ldg4(wr);
                    else { v.x = __ldg(wr); v.y = k + 1 < ks1 ? __ldg(wr + 1) : 0.f; v.z = k + 2 < ks1 ? __ldg(wr + 2) : 0.f; v.w = k + 3 < ks1 ? __ldg(wr + 3) : 0.f; }
                }
                wv[f][t][h] = v;
            }
    for (int b0 = 0; b0 < B; b0 += FWD_RB) {
        const int rows = min(FWD_RB, B - b0);
        float acc[FPW][FWD_RB];
#pragma unroll
        for (int f = 0; f < FPW; ++f)
#pragma unroll
            for (int r = 0; r < FWD_RB; ++r) acc[f][r] = 0.f;
#pragma unroll
        for (int t = 0; t < FWD_KS / FWD_KC; ++t) {
            const int k0 = ks0 + t * FWD_KC;
            if (k0 >= ks1) break;
            const int kc = min(FWD_KC, ks1 - k0);
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
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int r = 0; r < FWD_RB; ++r) {
                    const float4 xv = *reinterpret_cast<const float4*>(&xs[r][h * 128 + lane * 4]);
#pragma unroll
                    for (int f = 0; f < FPW; ++f) {
                        const float4 q = wv[f][t][h];
                        acc[f][r] = fmaf(xv.x, q.x, fmaf(xv.y, q.y, fmaf(xv.z, q.z, fmaf(xv.w, q.w, acc[f][r]))));
                    }
                }
        }
#pragma unroll
        for (int f = 0; f < FPW; ++f) {
            const int n = n0 + f;
            const float mine = row_sums_to_lanes(acc[f], lane);
            if (n < N && lane < rows) {
                if (partial) partial[((long long)blockIdx.y * B + b0 + lane) * N + n] = mine;
                else {
                    float t = (coef * mine + (bias ? __ldg(bias + n) : 0.f)) * gain;
                    y[(long long)(b0 + lane) * N + n] = t > 0.f ? t : t * slope;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) linear_fwd_finish_kernel(const float* __restrict__ partial, const float* __restrict__ bias, float* __restrict__ y,
                                                                int B, int N, int splits, float coef, float gain, float slope) {
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    if (i >= (long long)B * N) return;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += partial[(long long)sp * B * N + i];
    const float t = (coef * s + (bias ? __ldg(bias + (int)(i % N)) : 0.f)) * gain;
    y[i] = t > 0.f ? t : t * slope;
}

__device__ __forceinline__ float gu_of(const float* gy, const float* y, long long i, float gain, float slope) {
    const float g = __ldg(gy + i) * gain;
    return (y && !(__ldg(y + i) > 0.f)) ? g * slope : g;
}

// grid (k slices of 32 * KV, row groups of 8); NW warps split n, reduced through shared memory in warp order.  KV = 4 (128-bit weight loads)
// when K is large enough to fill the machine that way, KV = 1 (more, narrower CTAs) for the 512-wide mapping / affine layers.
template <int KV, int NW>
__global__ void __launch_bounds__(NW * 32) linear_bwd_data_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ w,
                                                              float* __restrict__ gx, int B, int K, int N, float coef, float gain, float slope) {
    constexpr int R = 8, NC = 512, KS = 32 * KV;
    __shared__ float gu[R][NC];
    __shared__ float red[NW][R][KS];
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
        for (int i = threadIdx.x; i < R * nn; i += NW * 32) {
            const int r = i / nn, c = i % nn;
            gu[r][c] = r < rows ? gu_of(gy, y, (long long)(b0 + r) * N + n0 + c, gain, slope) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = warp; c < nn; c += NW) {          // unrolled: 8 weight rows in flight per warp
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
    for (int i = threadIdx.x; i < R * KS; i += NW * 32) {
        const int r = i / KS, kk = i % KS;
        const int k = blockIdx.x * KS + kk;
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < NW; ++j) t += red[j][r][kk];
        if (r < rows && k < K) gx[(long long)(b0 + r) * K + k] = coef * t;
    }
}

// grid (k slices of 128, feature groups of 4 * FPW): a warp owns FPW features, a lane 4 consecutive k.  FPW = 4 for the wide layer
// (K = 8192: every x row slice is then fetched by N / 16 blocks instead of N / 4), 1 for the 512-wide layers (more blocks).
template <int FPW>
__global__ void __launch_bounds__(128) linear_bwd_weight_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ x,
                                                                float* __restrict__ gw, float* __restrict__ gb, int B, int K, int N,
                                                                float coef, float gain, float slope) {
    const int lane = threadIdx.x & 31;
    const int n0 = (blockIdx.y * 4 + (threadIdx.x >> 5)) * FPW;
    if (n0 >= N) return;
    const int k0 = blockIdx.x * 128 + lane * 4;
    const bool vec = (K & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)gw & 15) == 0;
    float4 acc[FPW];
    float sb[FPW];
#pragma unroll
    for (int f = 0; f < FPW; ++f) { acc[f] = f4zero(); sb[f] = 0.f; }
#pragma unroll 8
    for (int b = 0; b < B; ++b) {                     // unrolled: 8 rows of loads in flight (the loop is pure load latency otherwise)
        const float* xr = x + (long long)b * K;
        float4 xv = f4zero();
        if (vec) { if (k0 < K) xv = ldg4(xr + k0); }
        else {
            xv.x = k0 < K ? __ldg(xr + k0) : 0.f; xv.y = k0 + 1 < K ? __ldg(xr + k0 + 1) : 0.f;
            xv.z = k0 + 2 < K ? __ldg(xr + k0 + 2) : 0.f; xv.w = k0 + 3 < K ? __ldg(xr + k0 + 3) : 0.f;
        }
#pragma unroll
        for (int f = 0; f < FPW; ++f) {
            const float g = n0 + f < N ? gu_of(gy, y, (long long)b * N + n0 + f, gain, slope) : 0.f;
            sb[f] += g;
            fma4(acc[f], g, xv);
        }
    }
#pragma unroll
    for (int f = 0; f < FPW; ++f) {
        const int n = n0 + f;
        if (n >= N) break;
        float* dst = gw + (long long)n * K + k0;
        if (vec) { if (k0 < K) st4(dst, scale4(acc[f], coef)); }
        else {
            if (k0 < K) dst[0] = acc[f].x * coef;
            if (k0 + 1 < K) dst[1] = acc[f].y * coef;
            if (k0 + 2 < K) dst[2] = acc[f].z * coef;
            if (k0 + 3 < K) dst[3] = acc[f].w * coef;
        }
        if (gb && blockIdx.x == 0 && lane == 0) gb[n] = sb[f];
    }
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

// bytes of the partial-sum buffer sg2_linear_fwd needs (0: K fits one slice, no second pass)
extern "C" long long sg2_linear_fwd_workspace(int B, int K, int N) {
    const long long splits = ceil_div(K, lin::FWD_KS);
    return splits > 1 ? splits * B * N * (long long)sizeof(float) : 0;
}

extern "C" int sg2_linear_fwd(const float* x, const float* w, const float* bias, float* y, int B, int K, int N,
                              float coef, float gain, float slope, void* workspace, sg2_stream_t stream) {
    SG2_REQUIRE(x && w && y, "linear_fwd: null pointer");
    SG2_REQUIRE(B > 0 && K > 0 && N > 0, "linear_fwd: empty tensor");
    const int splits = (int)ceil_div(K, lin::FWD_KS);
    SG2_REQUIRE(splits == 1 || workspace, "linear_fwd: K spans several slices and needs the workspace of sg2_linear_fwd_workspace");
    float* partial = splits > 1 ? (float*)workspace : nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    // two features per warp when that still gives every SM a CTA (the wide discriminator layer), else one
    if (ceil_div(N, 16) * splits >= num_sms()) {
        dim3 grid((unsigned)ceil_div(N, 16), (unsigned)splits);
        lin::linear_fwd_kernel<2><<<grid, 256, 0, st>>>(x, w, bias, y, partial, B, K, N, coef, gain, slope);
    } else {
        dim3 grid((unsigned)ceil_div(N, 8), (unsigned)splits);
        lin::linear_fwd_kernel<1><<<grid, 256, 0, st>>>(x, w, bias, y, partial, B, K, N, coef, gain, slope);
    }
    int rc = launched("linear_fwd");
    if (rc || splits == 1) return rc;
    lin::linear_fwd_finish_kernel<<<(unsigned)ceil_div((long long)B * N, 256), 256, 0, st>>>(partial, bias, y, B, N, splits, coef, gain, slope);
    return launched("linear_fwd_finish");
}

extern "C" int sg2_linear_bwd_data(const float* gy, const float* y, const float* w, float* gx, int B, int K, int N,
                                   float coef, float gain, float slope, sg2_stream_t stream) {
    SG2_REQUIRE(gy && w && gx, "linear_bwd_data: null pointer");
    SG2_REQUIRE(B > 0 && K > 0 && N > 0, "linear_bwd_data: empty tensor");
    if (ceil_div(K, 128) * ceil_div(B, 8) >= num_sms()) {
        dim3 grid((unsigned)ceil_div(K, 128), (unsigned)ceil_div(B, 8));
        lin::linear_bwd_data_kernel<4, 4><<<grid, 128, 0, (cudaStream_t)stream>>>(gy, y, w, gx, B, K, N, coef, gain, slope);
    } else {
        dim3 grid((unsigned)ceil_div(K, 32), (unsigned)ceil_div(B, 8));
        lin::linear_bwd_data_kernel<1, 8><<<grid, 256, 0, (cudaStream_t)stream>>>(gy, y, w, gx, B, K, N, coef, gain, slope);
    }
    return launched("linear_bwd_data");
}

extern "C" int sg2_linear_bwd_weight(const float* gy, const float* y, const float* x, float* gw, float* gb, int B, int K, int N,
                                     float coef, float gain, float slope, sg2_stream_t stream) {
    SG2_REQUIRE(gy && x && gw, "linear_bwd_weight: null pointer");
    SG2_REQUIRE(B > 0 && K > 0 && N > 0, "linear_bwd_weight: empty tensor");
    if (ceil_div(K, 128) * ceil_div(N, 16) >= 2 * num_sms()) {
        dim3 grid((unsigned)ceil_div(K, 128), (unsigned)ceil_div(N, 16));
        lin::linear_bwd_weight_kernel<4><<<grid, 128, 0, (cudaStream_t)stream>>>(gy, y, x, gw, gb, B, K, N, coef, gain, slope);
    } else {
        dim3 grid((unsigned)ceil_div(K, 128), (unsigned)ceil_div(N, 4));
        SG2_REQUIRE(grid.y <= 65535, "linear_bwd_weight: too many output features (%d)", N);
        lin::linear_bwd_weight_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(gy, y, x, gw, gb, B, K, N, coef, gain, slope);
    }
    return launched("linear_bwd_weight");
}

extern "C" int sg2_pixelnorm(const float* x, float* y, int B, int K, float eps, sg2_stream_t stream) {
    SG2_REQUIRE(x && y && B > 0 && K > 0, "pixelnorm: bad arguments");
    lin::pixelnorm_kernel<<<(unsigned)ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(x, y, B, K, eps);
    return launched("pixelnorm");
}
