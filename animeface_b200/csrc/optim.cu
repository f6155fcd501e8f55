// Fused Adam + EMA over one flat fp32 parameter buffer.
// Replaces torch.optim.Adam.step (implementations/StyleGAN2/utils.py:208-221, used at :85-86,:112-113;
// ~125 small per-tensor launches) and update_ema (nnutils/training.py:23-40; 81 lerps) by two launches.
// Math is torch.optim.Adam's (no amsgrad, no weight decay):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
//   ema = decay * ema + (1-decay) * p        (nnutils/training.py:33-34)
// HBM-bound: 4 reads + 3 writes (+ema read/write) of numel floats.
#include "common.cuh"

namespace sg2 {

__global__ void __launch_bounds__(256) ema_kernel(float* __restrict__ ema, const float* __restrict__ p, long long numel, float decay) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x)
        ema[i] = ema[i] * decay + p[i] * (1.f - decay);
}

__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v,
                                                       float* __restrict__ ema, long long numel,
                                                       const long long* __restrict__ step, float lr, float b1, float b2,
                                                       float eps, float gscale, float decay) {
    const double t = (double)*step;
    const float bc1 = (float)(1.0 - pow((double)b1, t));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
    const float step_size = lr / bc1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        const float pi = p[i] - step_size * (mi / denom);
        p[i] = pi;
        if (ema) ema[i] = ema[i] * decay + pi * (1.f - decay);
    }
}

// ---- multi-tensor Adam: ONE launch over the flat buffer with exact per-tensor step counts ---------------------------
// prep: one thread per tensor.  present[i] != 0 -> steps[i] += 1 and coef[i] = (lr / (1 - b1^t), 1 / sqrt(1 - b2^t));
// absent tensors get coef = (0, 0) and are skipped by the main kernel exactly like torch.optim.Adam skips grad=None.
__global__ void adam_multi_prep_kernel(long long* steps, const int* present, float2* coef, int nseg, float lr, float b1, float b2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg) return;
    if (!present[i]) { coef[i] = make_float2(0.f, 0.f); return; }
    const long long t = steps[i] + 1;
    steps[i] = t;
    coef[i] = make_float2((float)(lr / (1.0 - pow((double)b1, (double)t))), (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)t))));
}

__global__ void __launch_bounds__(256) adam_multi_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long numel, const long long* __restrict__ seg_off,
                                                         const float2* __restrict__ coef, int nseg, float b1, float b2, float eps, float gscale) {
    __shared__ int s_seg;
    const long long nq = numel >> 2;
    for (long long q0 = (long long)blockIdx.x * blockDim.x; q0 < nq; q0 += (long long)gridDim.x * blockDim.x) {
        if (threadIdx.x == 0) {                       // segment of the block's first element: binary search
            const long long e0 = q0 << 2;
            int lo = 0, hi = nseg - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (seg_off[mid] <= e0) lo = mid; else hi = mid - 1; }
            s_seg = lo;
        }
        __syncthreads();
        int seg = s_seg;
        const long long q = q0 + threadIdx.x;
        if (q < nq) {
            const long long e = q << 2;
            while (seg + 1 < nseg && seg_off[seg + 1] <= e) ++seg;
            const float2 c = coef[seg];
            if (c.y != 0.f) {
                const float4 gi = ldg4(g + e);
                float4 mi = *reinterpret_cast<const float4*>(m + e), vi = *reinterpret_cast<const float4*>(v + e), pi = *reinterpret_cast<const float4*>(p + e);
                const float gx[4] = {gi.x * gscale, gi.y * gscale, gi.z * gscale, gi.w * gscale};
                float mm[4] = {mi.x, mi.y, mi.z, mi.w}, vv[4] = {vi.x, vi.y, vi.z, vi.w}, pp[4] = {pi.x, pi.y, pi.z, pi.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mm[j] = b1 * mm[j] + (1.f - b1) * gx[j];
                    vv[j] = b2 * vv[j] + (1.f - b2) * gx[j] * gx[j];
                    pp[j] -= c.x * (mm[j] / (sqrtf(vv[j]) * c.y + eps));
                }
                st4(m + e, make_float4(mm[0], mm[1], mm[2], mm[3]));
                st4(v + e, make_float4(vv[0], vv[1], vv[2], vv[3]));
                st4(p + e, make_float4(pp[0], pp[1], pp[2], pp[3]));
            }
        }
        __syncthreads();
    }
}

template <class T>
__global__ void counter_add_kernel(T* c, int first, int count, int delta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) c[first + i] += (T)delta;
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_adam_multi(float* p, const float* g, float* m, float* v, int64_t numel, const int64_t* seg_off,
                              int64_t* steps, const int* present, void* coef_ws, int nseg,
                              float lr, float beta1, float beta2, float eps, float grad_scale, sg2_stream_t stream) {
    SG2_REQUIRE(p && g && m && v && seg_off && steps && present && coef_ws, "adam_multi: null pointer");
    SG2_REQUIRE(numel > 0 && numel % 4 == 0 && nseg > 0, "adam_multi: numel must be a positive multiple of 4");
    cudaStream_t st = (cudaStream_t)stream;
    adam_multi_prep_kernel<<<(nseg + 127) / 128, 128, 0, st>>>((long long*)steps, present, (float2*)coef_ws, nseg, lr, beta1, beta2);
    int rc = launched("adam_multi_prep");
    if (rc) return rc;
    const int blocks = (int)std::min<long long>(ceil_div(numel / 4, 256), (long long)num_sms() * 16);
    adam_multi_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, numel, (const long long*)seg_off, (const float2*)coef_ws, nseg, beta1, beta2, eps, grad_scale);
    return launched("adam_multi");
}

extern "C" int sg2_counter_add(void* counters, int first, int count, int delta, int is64, sg2_stream_t stream) {
    SG2_REQUIRE(counters && first >= 0 && count > 0, "counter_add: bad arguments");
    if (is64) counter_add_kernel<long long><<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>((long long*)counters, first, count, delta);
    else      counter_add_kernel<int><<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>((int*)counters, first, count, delta);
    return launched("counter_add");
}

extern "C" int sg2_adam_ema(float* p, const float* g, float* m, float* v, float* ema,
                            int64_t numel, const int64_t* step, float lr, float beta1, float beta2,
                            float eps, float grad_scale, float ema_decay, sg2_stream_t stream) {
    SG2_REQUIRE(p && g && m && v && step, "adam_ema: null pointer");
    SG2_REQUIRE(numel > 0, "adam_ema: empty buffer");
    const int blocks = (int)std::min<long long>(ceil_div(numel, 256), (long long)num_sms() * 16);
    adam_ema_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, ema, numel, (const long long*)step, lr, beta1, beta2, eps, grad_scale, ema_decay);
    return launched("adam_ema");
}

extern "C" int sg2_ema_update(float* ema, const float* p, int64_t numel, float decay, sg2_stream_t stream) {
    SG2_REQUIRE(ema && p, "ema_update: null pointer");
    SG2_REQUIRE(numel > 0, "ema_update: empty buffer");
    const int blocks = (int)std::min<long long>(ceil_div(numel, 256), (long long)num_sms() * 16);
    ema_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ema, p, numel, decay);
    return launched("ema_update");
}
