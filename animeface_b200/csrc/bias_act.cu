// Fused bias + activation + gain + clamp, with first- and second-order gradient modes.
// Semantics: thirdparty/stylegan3_ops/ops/bias_act.py:86-115 (_bias_act_ref) for grad=0 and the
// gradient formulas of the plugin (thirdparty/stylegan3_ops/ops/bias_act.cu:17-141) for grad=1,2;
// host contract: thirdparty/stylegan3_ops/ops/bias_act.cpp:26-84.  Kernels are new: 128-bit
// vectorised over the flat dense buffer (HBM-bound: bytes = reads of present operands + one write).
#include "common.cuh"

namespace sg2 {

struct BiasActParams {
    const void *x, *b, *xref, *yref, *dy;
    void* y;
    long long numel, step_b;
    int size_b, grad;
    float alpha, gain, clamp;
};

template <class T> struct BAcc { typedef float type; };
template <> struct BAcc<double> { typedef double type; };

// One element.  A = activation index (1..9), G = p.grad at run time.
template <int A, int G, class S>
__device__ __forceinline__ S bias_act_eval(S x, S b, S xref, S yref, S dy, S alpha, S gain, S clamp) {
    const S one = (S)1, two = (S)2, exp_range = (S)80, half_exp_range = (S)40;
    const S selu_scale = (S)1.0507009873554804934193349852946, selu_alpha = (S)1.6732632423543772848170429916717;
    S yy = (G != 0 && gain != 0) ? yref / gain : (S)0;      // only the gradient modes look at the saved output
    S y = 0;
    if (G == 0) x += b; else xref += b;
    if (A == 1) { if (G == 0 || G == 1) y = x; }
    if (A == 2) { if (G == 0) y = (x > 0) ? x : (S)0; if (G == 1) y = (yy > 0) ? x : (S)0; }
    if (A == 3) { if (G == 0) y = (x > 0) ? x : x * alpha; if (G == 1) y = (yy > 0) ? x : x * alpha; }
    if (A == 4) {
        if (G == 0) { S c = exp(x); S d = one / c; y = (x < -exp_range) ? -one : (x > exp_range) ? one : (c - d) / (c + d); }
        if (G == 1) y = x * (one - yy * yy);
        if (G == 2) y = x * (one - yy * yy) * (-two * yy);
    }
    if (A == 5) {
        if (G == 0) y = (x < -exp_range) ? (S)0 : one / (exp(-x) + one);
        if (G == 1) y = x * yy * (one - yy);
        if (G == 2) y = x * yy * (one - yy) * (one - two * yy);
    }
    if (A == 6) {
        if (G == 0) y = (x >= 0) ? x : exp(x) - one;
        if (G == 1) y = (yy >= 0) ? x : x * (yy + one);
        if (G == 2) y = (yy >= 0) ? (S)0 : x * (yy + one);
    }
    if (A == 7) {
        if (G == 0) y = (x >= 0) ? selu_scale * x : (selu_scale * selu_alpha) * (exp(x) - one);
        if (G == 1) y = (yy >= 0) ? x * selu_scale : x * (yy + selu_scale * selu_alpha);
        if (G == 2) y = (yy >= 0) ? (S)0 : x * (yy + selu_scale * selu_alpha);
    }
    if (A == 8) {
        if (G == 0) y = (x > exp_range) ? x : log(exp(x) + one);
        if (G == 1) y = x * (one - exp(-yy));
        if (G == 2) { S c = exp(-yy); y = x * c * (one - c); }
    }
    if (A == 9) {
        if (G == 0) y = (x < -exp_range) ? (S)0 : x / (exp(-x) + one);
        else {
            S c = exp(xref), d = c + one;
            if (G == 1) y = (xref > half_exp_range) ? x : x * c * (xref + d) / (d * d);
            else        y = (xref > half_exp_range) ? (S)0 : x * c * (xref * (two - d) + two * d) / (d * d * d);
            yref = (xref < -exp_range) ? (S)0 : xref / (exp(-xref) + one) * gain;
        }
    }
    y *= gain * dy;
    if (clamp >= 0) {
        if (G == 0) y = (y > -clamp && y < clamp) ? y : (y >= 0) ? clamp : -clamp;
        else        y = (yref > -clamp && yref < clamp) ? y : (S)0;
    }
    return y;
}

template <class T, int A, int G>
__global__ void __launch_bounds__(256) bias_act_scalar(BiasActParams p) {
    typedef typename BAcc<T>::type S;
    const T* x = (const T*)p.x; const T* b = (const T*)p.b; const T* xr = (const T*)p.xref;
    const T* yr = (const T*)p.yref; const T* dy = (const T*)p.dy; T* y = (T*)p.y;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.numel; i += (long long)gridDim.x * blockDim.x) {
        S bv = b ? (S)b[(i / p.step_b) % p.size_b] : (S)0;
        S v = bias_act_eval<A, G, S>((S)x[i], bv, xr ? (S)xr[i] : (S)0, yr ? (S)yr[i] : (S)0, dy ? (S)dy[i] : (S)1,
                                     (S)p.alpha, (S)p.gain, (S)p.clamp);
        y[i] = (T)v;
    }
}

// fp32, 4 elements per thread.  BMODE 0: no bias, 1: channels_last (step_b == 1, size_b % 4 == 0),
// 2: step_b % 4 == 0 (the 4 elements share one bias value).
template <int A, int G, int BMODE>
__global__ void __launch_bounds__(256) bias_act_vec4(BiasActParams p) {
    const float* x = (const float*)p.x; const float* b = (const float*)p.b; const float* xr = (const float*)p.xref;
    const float* yr = (const float*)p.yref; const float* dy = (const float*)p.dy; float* y = (float*)p.y;
    const int nq = (int)(p.numel >> 2);               // numel <= INT_MAX (checked by the caller): 32-bit index math
    const int size_b = p.size_b, step_b = (int)p.step_b;
    constexpr int U = 4;                              // independent 128-bit accesses in flight per thread
    for (int q0 = blockIdx.x * blockDim.x * U + threadIdx.x; q0 < nq; q0 += gridDim.x * blockDim.x * U) {
        float4 xv[U], xrv[U], yrv[U], dyv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = q0 + u * blockDim.x;
            if (q < nq) {
                xv[u] = ldg4(x + 4 * (size_t)q);
                xrv[u] = xr ? ldg4(xr + 4 * (size_t)q) : f4zero();
                yrv[u] = yr ? ldg4(yr + 4 * (size_t)q) : f4zero();
                dyv[u] = dy ? ldg4(dy + 4 * (size_t)q) : make_float4(1.f, 1.f, 1.f, 1.f);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = q0 + u * blockDim.x;
            if (q >= nq) continue;
            const unsigned i = 4u * (unsigned)q;
            float4 bv = f4zero();
            if (BMODE == 1) bv = ldg4(b + (i % (unsigned)size_b));
            if (BMODE == 2) { float s = __ldg(b + (i / (unsigned)step_b) % (unsigned)size_b); bv = make_float4(s, s, s, s); }
            float4 o;
            o.x = bias_act_eval<A, G, float>(xv[u].x, bv.x, xrv[u].x, yrv[u].x, dyv[u].x, p.alpha, p.gain, p.clamp);
            o.y = bias_act_eval<A, G, float>(xv[u].y, bv.y, xrv[u].y, yrv[u].y, dyv[u].y, p.alpha, p.gain, p.clamp);
            o.z = bias_act_eval<A, G, float>(xv[u].z, bv.z, xrv[u].z, yrv[u].z, dyv[u].z, p.alpha, p.gain, p.clamp);
            o.w = bias_act_eval<A, G, float>(xv[u].w, bv.w, xrv[u].w, yrv[u].w, dyv[u].w, p.alpha, p.gain, p.clamp);
            st4_cs(y + 4 * (size_t)q, o);
        }
    }
}

template <int A, int G>
static int bias_act_dispatch_g(const BiasActParams& p, int dtype, cudaStream_t st) {
    const int threads = 256;
    auto aligned = [](const void* q) { return q == nullptr || ((uintptr_t)q % 16) == 0; };
    if (dtype == SG2_F32 && (p.numel % 4) == 0 && aligned(p.x) && aligned(p.b) && aligned(p.xref) &&
        aligned(p.yref) && aligned(p.dy) && aligned(p.y)) {
        int bmode = -1;
        if (!p.b) bmode = 0;
        else if (p.step_b == 1 && p.size_b % 4 == 0) bmode = 1;
        else if (p.step_b % 4 == 0) bmode = 2;
        if (bmode >= 0) {
            int blocks = (int)std::min<long long>(ceil_div(p.numel / 4, threads * 4), (long long)num_sms() * 16);
            if (bmode == 0)      bias_act_vec4<A, G, 0><<<blocks, threads, 0, st>>>(p);
            else if (bmode == 1) bias_act_vec4<A, G, 1><<<blocks, threads, 0, st>>>(p);
            else                 bias_act_vec4<A, G, 2><<<blocks, threads, 0, st>>>(p);
            return launched("bias_act_vec4");
        }
    }
    int blocks = (int)std::min<long long>(ceil_div(p.numel, threads), (long long)num_sms() * 32);
    if (dtype == SG2_F32)      bias_act_scalar<float, A, G><<<blocks, threads, 0, st>>>(p);
    else if (dtype == SG2_F16) bias_act_scalar<__half, A, G><<<blocks, threads, 0, st>>>(p);
    else                       bias_act_scalar<double, A, G><<<blocks, threads, 0, st>>>(p);
    return launched("bias_act_scalar");
}

template <int A>
static int bias_act_dispatch(const BiasActParams& p, int dtype, cudaStream_t st) {
    if (p.grad == 0) return bias_act_dispatch_g<A, 0>(p, dtype, st);
    if (p.grad == 1) return bias_act_dispatch_g<A, 1>(p, dtype, st);
    return bias_act_dispatch_g<A, 2>(p, dtype, st);
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_bias_act(const void* x, const void* b, const void* xref, const void* yref,
                            const void* dy, void* y, int dtype, int64_t numel,
                            int size_b, int64_t step_b, int grad, int act,
                            float alpha, float gain, float clamp, sg2_stream_t stream) {
    // checks mirror thirdparty/stylegan3_ops/ops/bias_act.cpp:29-41
    SG2_REQUIRE(x && y, "bias_act: null pointer");
    SG2_REQUIRE(numel > 0 && numel <= 2147483647LL, "bias_act: x is empty or too large");
    SG2_REQUIRE(grad >= 0 && grad <= 2, "bias_act: grad must be 0, 1 or 2");
    SG2_REQUIRE(dtype == SG2_F32 || dtype == SG2_F16 || dtype == SG2_F64, "bias_act: unsupported dtype %d", dtype);
    SG2_REQUIRE(!b || (size_b > 0 && step_b > 0), "bias_act: b has wrong number of elements");
    SG2_REQUIRE(act >= 1 && act <= 9, "bias_act: no CUDA kernel found for the specified activation func");
    BiasActParams p;
    p.x = x; p.b = b; p.xref = xref; p.yref = yref; p.dy = dy; p.y = y;
    p.numel = numel; p.step_b = b ? step_b : 1; p.size_b = b ? size_b : 1; p.grad = grad;
    p.alpha = alpha; p.gain = gain; p.clamp = clamp;
    cudaStream_t st = (cudaStream_t)stream;
    switch (act) {
        case 1: return bias_act_dispatch<1>(p, dtype, st);
        case 2: return bias_act_dispatch<2>(p, dtype, st);
        case 3: return bias_act_dispatch<3>(p, dtype, st);
        case 4: return bias_act_dispatch<4>(p, dtype, st);
        case 5: return bias_act_dispatch<5>(p, dtype, st);
        case 6: return bias_act_dispatch<6>(p, dtype, st);
        case 7: return bias_act_dispatch<7>(p, dtype, st);
        case 8: return bias_act_dispatch<8>(p, dtype, st);
        default: return bias_act_dispatch<9>(p, dtype, st);
    }
}
