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

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_adam_ema(float* p, const float* g, float* m, float* v, float* ema,
                            int64_t numel, const int64_t* step, float lr, float beta1, float beta2,
                            float eps, float grad_scale, float ema_decay, sg2_stream_t stream) {
    SG2_REQUIRE(p && g && m && v && step, "adam_ema: null pointer");
    SG2_REQUIRE(numel > 0, "adam_ema: empty buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)std::min<long long>(ceil_div(numel, 256), (long long)num_sms() * 16);
    adam_ema_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, ema, numel, (const long long*)step, lr, beta1, beta2, eps, grad_scale, ema_decay);
    return launched("adam_ema");
}

extern "C" int sg2_ema_update(float* ema, const float* p, int64_t numel, float decay, sg2_stream_t stream) {
    SG2_REQUIRE(ema && p, "ema_update: null pointer");
    SG2_REQUIRE(numel > 0, "ema_update: empty buffer");
    const int blocks = (int)std::min<long long>(ceil_div(numel, 256), (long long)num_sms() * 16);
    ema_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ema, p, numel, decay);
    return launched("ema_update");
}
