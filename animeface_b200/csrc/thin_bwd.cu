// Backward of a "thin-input" 1x1 convolution + bias + leaky ReLU in ONE pass over (gy, y): Discriminator.from_rgb
// (3 -> 32 @256^2, implementations/StyleGAN2/model.py:383-384; its backward is ATen's convolution_backward + the LeakyReLU
// backward + a bias reduction).  The separate form wrote gu = gy * lrelu'(y) (12 B per element) and read it back in the thin
// weight-gradient / data-gradient kernels; here gu lives in registers:
//     gu[pix, o]  = gy * gain * (y > 0 ? 1 : slope)
//     gw[o, c]    = coef * sum_pix gu[pix, o] * x[pix, c]          c < CIN <= 4
//     gb[o]       = sum_pix gu[pix, o]
//     gx[pix, c]  = coef * sum_o gu[pix, o] * w[o, c]               (optional: the G phase and R1 need it, the D phase does not)
// co / 4 threads share a pixel (a float4 of channels each); the per-thread sums meet in shared memory per block and a second
// kernel adds the blocks in a fixed order: deterministic.
#include "common.cuh"

namespace sg2 {
namespace thinb {

constexpr int kMaxCo = 64;

template <int CIN>
__global__ void __launch_bounds__(256) thin_in_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ x,
                                                          const float* __restrict__ w, float* __restrict__ gx, float* __restrict__ part,
                                                          long long P, int co, float slope, float gain, float coef) {
    __shared__ float sw[kMaxCo * 4];                       // [o][CIN] * coef
    __shared__ float red[(CIN + 1)][256 * 4 / 1];          // per quantity: [slot][co] = 256 threads x 4 channels
    for (int i = threadIdx.x; i < co * CIN; i += 256) sw[i] = __ldg(w + i) * coef;
    __syncthreads();
    const int cq = co >> 2, q = threadIdx.x % cq, slot = threadIdx.x / cq, slots = 256 / cq;
    float4 gw[CIN], gb = f4zero();
#pragma unroll
    for (int c = 0; c < CIN; ++c) gw[c] = f4zero();
    const long long step = (long long)gridDim.x * slots;
    const long long iters = (P + step - 1) / step;         // the same trip count for every thread: the shuffles below need full warps
    for (long long it = 0; it < iters; ++it) {
        const long long pix = it * step + (long long)blockIdx.x * slots + slot;
        const bool ok = pix < P;
        float4 gu = f4zero();
        float xv[CIN];
#pragma unroll
        for (int c = 0; c < CIN; ++c) xv[c] = 0.f;
        if (ok) {
            const long long o = pix * co + 4 * q;
            const float4 g = ldg4(gy + o), yv = ldg4(y + o);
            gu.x = g.x * gain * (yv.x > 0.f ? 1.f : slope); gu.y = g.y * gain * (yv.y > 0.f ? 1.f : slope);
            gu.z = g.z * gain * (yv.z > 0.f ? 1.f : slope); gu.w = g.w * gain * (yv.w > 0.f ? 1.f : slope);
#pragma unroll
            for (int c = 0; c < CIN; ++c) xv[c] = __ldg(x + pix * CIN + c);
        }
        gb = add4(gb, gu);
#pragma unroll
        for (int c = 0; c < CIN; ++c) fma4(gw[c], xv[c], gu);
        if (gx) {
            float s[CIN];
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                const int ob = 4 * q;
                s[c] = gu.x * sw[(ob + 0) * CIN + c] + gu.y * sw[(ob + 1) * CIN + c] + gu.z * sw[(ob + 2) * CIN + c] + gu.w * sw[(ob + 3) * CIN + c];
                for (int off = cq >> 1; off > 0; off >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], off);     // cq is a power of two <= 16
            }
            if (ok && q == 0) {
#pragma unroll
                for (int c = 0; c < CIN; ++c) gx[pix * CIN + c] = s[c];
            }
        }
    }
    // block reduction over the pixel slots, fixed order
    auto put = [&](int k, const float4& v) {
        float* r = &red[k][threadIdx.x * 4];
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    };
#pragma unroll
    for (int c = 0; c < CIN; ++c) put(c, gw[c]);
    put(CIN, gb);
    __syncthreads();
    for (int i = threadIdx.x; i < (CIN + 1) * co; i += 256) {
        const int k = i / co, ch = i - k * co;
        float s = 0.f;
        for (int sl = 0; sl < slots; ++sl) s += red[k][(sl * cq + (ch >> 2)) * 4 + (ch & 3)];
        part[(long long)blockIdx.x * (CIN + 1) * co + i] = s;
    }
}

// gw[o][c] = coef * sum_blocks part[b][c][o],  gb[o] = sum_blocks part[b][CIN][o].  A block owns 32 outputs; its 8 warps take every
// 8th block partial and meet in shared memory in warp order (fixed order).
__global__ void __launch_bounds__(256) thin_in_bwd_finish_kernel(const float* __restrict__ part, float* __restrict__ gw, float* __restrict__ gb,
                                                                 int blocks, int cin, int co, float coef) {
    __shared__ float sh[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane, total = (cin + 1) * co;
    float s = 0.f;
    if (i < total)
        for (int b = warp; b < blocks; b += 8) s += part[(long long)b * total + i];
    sh[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && i < total) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += sh[j][lane];
        const int k = i / co, o = i - k * co;
        if (k < cin) { if (gw) gw[o * cin + k] = t * coef; }
        else if (gb) gb[o] = t;
    }
}

static int blocks_for(long long P, int co) {
    const int slots = 256 / (co / 4);
    return (int)std::min<long long>(ceil_div(P, (long long)slots * 8), (long long)num_sms() * 4);
}

}  // namespace thinb
}  // namespace sg2

using namespace sg2;

extern "C" int64_t sg2_thin_in_bwd_workspace(int n, int hw, int cin, int co) {
    if (n <= 0 || hw <= 0 || cin <= 0 || co <= 0 || co % 4) return -1;
    return (int64_t)thinb::blocks_for((long long)n * hw, co) * (cin + 1) * co * (int64_t)sizeof(float);
}

extern "C" int sg2_thin_in_bwd(const float* gy, const float* y, const float* x, const float* w, float* gx, float* gw, float* gb,
                               void* workspace, int n, int hw, int cin, int co, float slope, float gain, float coef, sg2_stream_t stream) {
    SG2_REQUIRE(gy && y && x && w && workspace, "thin_in_bwd: null pointer");
    SG2_REQUIRE(n > 0 && hw > 0 && cin >= 1 && cin <= 4, "thin_in_bwd: 1..4 input channels");
    const int cq = co / 4;
    SG2_REQUIRE(co % 4 == 0 && co <= thinb::kMaxCo && cq >= 1 && (cq & (cq - 1)) == 0 && cq <= 16, "thin_in_bwd: co must be 4, 8, 16, 32 or 64 (co=%d)", co);
    SG2_REQUIRE((((uintptr_t)gy | (uintptr_t)y) & 15) == 0, "thin_in_bwd: gy / y must be 16-byte aligned");
    const long long P = (long long)n * hw;
    const int blocks = thinb::blocks_for(P, co);
    cudaStream_t st = (cudaStream_t)stream;
    float* part = (float*)workspace;
    switch (cin) {
        case 1: thinb::thin_in_bwd_kernel<1><<<blocks, 256, 0, st>>>(gy, y, x, w, gx, part, P, co, slope, gain, coef); break;
        case 2: thinb::thin_in_bwd_kernel<2><<<blocks, 256, 0, st>>>(gy, y, x, w, gx, part, P, co, slope, gain, coef); break;
        case 3: thinb::thin_in_bwd_kernel<3><<<blocks, 256, 0, st>>>(gy, y, x, w, gx, part, P, co, slope, gain, coef); break;
        default: thinb::thin_in_bwd_kernel<4><<<blocks, 256, 0, st>>>(gy, y, x, w, gx, part, P, co, slope, gain, coef); break;
    }
    int rc = launched("thin_in_bwd");
    if (rc) return rc;
    thinb::thin_in_bwd_finish_kernel<<<(unsigned)ceil_div((cin + 1) * co, 32), 256, 0, st>>>(part, gw, gb, blocks, cin, co, coef);
    return launched("thin_in_bwd_finish");
}
