// The demodulation coefficient of ModulatedConv2d and its gradients as three small kernels.
//
// Replaces, per modulated convolution and pass, the tensor expression of implementations/StyleGAN2/model.py:115-120 reduced to
// the [B,Co] coefficient (the per-sample weight tensor [B,Co,Ci,k,k] is never built):
//     wsq[o,i] = sum_k w[o,i,k]^2                                  (pow, reduce)
//     d[b,o]   = rsqrt(coef^2 * sum_i s[b,i]^2 wsq[o,i] + eps)     (pow, cuBLAS sgemm, mul, add, rsqrt)
// and their autograd backward (two more sgemms and ~8 elementwise launches): ~250 tiny launches per training step.
//   gt[b,o]  = -0.5 * gd[b,o] * d[b,o]^3 * coef^2
//   gs[b,i]  = 2 s[b,i] * sum_o gt[b,o] wsq[o,i]
//   gw[o,i,k] = 2 w[o,i,k] * sum_b gt[b,o] s[b,i]^2
// No atomics: every output element is summed by one thread in a fixed order.
#include "common.cuh"

namespace sg2 {
namespace demod {

constexpr int kMaxCi = 2048, kMaxB = 256;

// grid = Co, block = 256.  wsq row -> shared memory (and global, for the backward); then one warp per sample.
__global__ void __launch_bounds__(256) demod_fwd_kernel(const float* __restrict__ w, const float* __restrict__ s, float* __restrict__ wsq,
                                                        float* __restrict__ d, int B, int co, int ci, int kk, float coef2, float eps) {
    __shared__ float row[kMaxCi];
    const int o = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < ci; i += 256) {
        const float* src = w + ((long long)o * ci + i) * kk;
        float a = 0.f;
        for (int k = 0; k < kk; ++k) { const float v = __ldg(src + k); a = fmaf(v, v, a); }
        row[i] = a;
        wsq[(long long)o * ci + i] = a;
    }
    __syncthreads();
    // one warp per sample (8 samples in flight per block): no block-wide synchronisation inside the loop
    for (int b = warp; b < B; b += 8) {
        float a = 0.f;
        for (int i = lane; i < ci; i += 32) { const float v = __ldg(s + (long long)b * ci + i); a = fmaf(v * v, row[i], a); }
        a = warp_sum(a);
        if (lane == 0) d[(long long)b * co + o] = rsqrtf(coef2 * a + eps);
    }
}

// grid = Co, block = 256: gw[o,i,:] = 2 w[o,i,:] * sum_b gt[b,o] s[b,i]^2
__global__ void __launch_bounds__(256) demod_bwd_w_kernel(const float* __restrict__ w, const float* __restrict__ s, const float* __restrict__ d,
                                                          const float* __restrict__ gd, float* __restrict__ gw, int B, int co, int ci, int kk,
                                                          float coef2) {
    __shared__ float gt[kMaxB];
    const int o = blockIdx.x;
    for (int b = threadIdx.x; b < B; b += 256) {
        const float dv = __ldg(d + (long long)b * co + o);
        gt[b] = -0.5f * __ldg(gd + (long long)b * co + o) * dv * dv * dv * coef2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ci; i += 256) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) { const float v = __ldg(s + (long long)b * ci + i); a = fmaf(gt[b], v * v, a); }
        const long long base = ((long long)o * ci + i) * kk;
        for (int k = 0; k < kk; ++k) gw[base + k] = 2.f * __ldg(w + base + k) * a;
    }
}

// grid = (B, ci / 32), block = 256: gs[b,i] = 2 s[b,i] * sum_o gt[b,o] wsq[o,i].  Lane = input channel (coalesced rows of wsq), the 8
// warps split o and their partial sums meet in shared memory in warp order (fixed order: deterministic).
__global__ void __launch_bounds__(256) demod_bwd_s_kernel(const float* __restrict__ wsq, const float* __restrict__ s, const float* __restrict__ d,
                                                          const float* __restrict__ gd, float* __restrict__ gs, int B, int co, int ci, float coef2) {
    __shared__ float gt[kMaxCi];                          // indexed by o (co <= kMaxCi)
    __shared__ float part[8][32];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = threadIdx.x; o < co; o += 256) {
        const float dv = __ldg(d + (long long)b * co + o);
        gt[o] = -0.5f * __ldg(gd + (long long)b * co + o) * dv * dv * dv * coef2;
    }
    __syncthreads();
    const int i = blockIdx.y * 32 + lane;
    float a = 0.f;
    if (i < ci) {
#pragma unroll 8
        for (int o = warp; o < co; o += 8) a = fmaf(gt[o], __ldg(wsq + (long long)o * ci + i), a);
    }
    part[warp][lane] = a;
    __syncthreads();
    if (warp == 0 && i < ci) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += part[j][lane];
        gs[(long long)b * ci + i] = 2.f * __ldg(s + (long long)b * ci + i) * t;
    }
}

}  // namespace demod
}  // namespace sg2

using namespace sg2;

extern "C" int sg2_demod_fwd(const float* w, const float* s, float* wsq, float* d, int B, int co, int ci, int kk, float coef,
                             float eps, sg2_stream_t stream) {
    SG2_REQUIRE(w && s && wsq && d, "demod_fwd: null pointer");
    SG2_REQUIRE(B > 0 && co > 0 && ci > 0 && kk > 0 && ci <= demod::kMaxCi, "demod_fwd: need B, co, ci, k*k > 0 and ci <= %d", demod::kMaxCi);
    demod::demod_fwd_kernel<<<(unsigned)co, 256, 0, (cudaStream_t)stream>>>(w, s, wsq, d, B, co, ci, kk, coef * coef, eps);
    return launched("demod_fwd");
}

extern "C" int sg2_demod_bwd(const float* w, const float* s, const float* wsq, const float* d, const float* gd, float* gw, float* gs,
                             int B, int co, int ci, int kk, float coef, sg2_stream_t stream) {
    SG2_REQUIRE(w && s && wsq && d && gd, "demod_bwd: null pointer");
    SG2_REQUIRE(B > 0 && B <= demod::kMaxB && co > 0 && co <= demod::kMaxCi && ci > 0 && kk > 0, "demod_bwd: B <= %d, co <= %d", demod::kMaxB, demod::kMaxCi);
    cudaStream_t st = (cudaStream_t)stream;
    if (gw) {
        demod::demod_bwd_w_kernel<<<(unsigned)co, 256, 0, st>>>(w, s, d, gd, gw, B, co, ci, kk, coef * coef);
        int rc = launched("demod_bwd_w");
        if (rc) return rc;
    }
    if (gs) {
        demod::demod_bwd_s_kernel<<<dim3((unsigned)B, (unsigned)ceil_div(ci, 32)), 256, 0, st>>>(wsq, s, d, gd, gs, B, co, ci, coef * coef);
        return launched("demod_bwd_s");
    }
    return SG2_OK;
}
