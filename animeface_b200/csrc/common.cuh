// Shared helpers for libsg2b200 (host error plumbing + small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/sg2b200.h"

namespace sg2 {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;
extern long long* g_trace;          // sg2_debug_trace buffer (device pointer) or null

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// Record a launch and turn a launch-time error into SG2_ELAUNCH.
inline int launched(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SG2_ELAUNCH, "%s: %s", what, cudaGetErrorString(e));
    return SG2_OK;
}

#define SG2_REQUIRE(cond, ...) do { if (!(cond)) return sg2::fail(SG2_EINVAL, __VA_ARGS__); } while (0)
#define SG2_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return sg2::fail(SG2_ELAUNCH, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

inline int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming store: written once, not re-read by this kernel
__device__ __forceinline__ void st4_cs(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void fma4(float4& a, float s, const float4& v) {
    a.x = fmaf(s, v.x, a.x); a.y = fmaf(s, v.y, a.y); a.z = fmaf(s, v.z, a.z); a.w = fmaf(s, v.w, a.w);
}
// the same as two packed f32x2 instructions (sm_100 FFMA2: one issue slot per two lanes, same IEEE result per component).
// Pays where a kernel is short of issue slots and its operands sit in aligned register pairs anyway; measured per kernel.
__device__ __forceinline__ void fma4_x2(float4& a, float s, const float4& v) {
    const float2 ss = make_float2(s, s);
    const float2 lo = __ffma2_rn(ss, make_float2(v.x, v.y), make_float2(a.x, a.y));
    const float2 hi = __ffma2_rn(ss, make_float2(v.z, v.w), make_float2(a.z, a.w));
    a = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 mul4(const float4& a, const float4& b) {
    return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 scale4(const float4& a, float s) {
    return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace sg2
