// Internal parameter blocks shared by the convolution kernels (SIMT and tcgen05) and the C-ABI shim.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sg2 {

struct ConvParams {
    const float* x;        // [n,h,w,ci] NHWC dense
    const void* wp;        // packed weight (layout depends on impl)
    float* y;              // output, strides ys (n,c,h,w) in elements
    long long ys[4];
    int n, h, w, ci, co, k;
    const float* in_scale;   // [n,ci] or null
    const float* out_scale;  // [n,co] or null
    const float* bias;       // [co] or null
    const float* noise;      // [n,h,w] or null
    int act;                 // 1 linear, 3 lrelu
    float alpha, gain;
};

struct WgradParams {
    const float* x;        // [n,h,w,ci]
    const float* gy;       // [n,h,w,co]
    float* dw;             // [co,ci,k,k]
    int n, h, w, ci, co, k;
    float coef;
    const float* in_scale;   // [n,ci] or null
    const float* out_scale;  // [n,co] or null
    long long chunk;         // pixels per split (set by the launcher)
    // deterministic reduction: when `ws` is set every pixel split stores its partial dw into ws[part][co*ci*k*k] (plain stores, each
    // element exactly once per part) and wgrad_sum_parts adds the parts in a fixed order; null -> fp32 atomics into dw
    float* ws;
};

int conv_fwd_simt(const ConvParams& p, cudaStream_t st);
int conv_wgrad_simt(WgradParams p, int accumulate, cudaStream_t st);
// number of partial buffers the kernel that takes this shape would write (same dispatch as the launchers)
int wgrad_parts_simt(const WgradParams& p);
int wgrad_parts_tc(const WgradParams& p);
int wgrad_sum_parts(const float* ws, float* dw, long long size, int parts, int accumulate, cudaStream_t st);
int conv_pack_simt(const float* w, float* wp, int co, int ci, int k, float coef, int transpose, cudaStream_t st);

// thin 1x1 layers, <= 4 channels on one side (conv_thin.cu); SG2_ENOTSUP when the shape / layout is not theirs
int conv_fwd_thin(const ConvParams& p, cudaStream_t st);
int conv_wgrad_thin(const WgradParams& p, cudaStream_t st);
int wgrad_parts_thin(const WgradParams& p);     // 0 when the thin kernel does not take the shape

// tcgen05 weight gradient on fp32 operands (wgrad_tc.cu)
bool wgrad_tc_supported(int n, int h, int w, int ci, int co, int k);
int conv_wgrad_tc(WgradParams p, int accumulate, cudaStream_t st);

// tcgen05 path, halo variant (conv_halo.cu): patch loaded/converted once per tile, taps = shifted descriptor windows
bool conv_halo_supported(int n, int h, int w, int ci, int co, int k);
int conv_fwd_halo(const ConvParams& p, int precise, cudaStream_t st);
int conv_pack_halo(const float* w, void* wp, int co, int ci, int k, float coef, int transpose, int precise, cudaStream_t st);
long long conv_packed_bytes_halo(int co, int ci, int k);

// bf16 pair-planes operands (planes.cu): TMA -> tcgen05 without a transform pass
bool conv_halo_pl_supported(int n, int h, int w, int ci, int co, int k);
int conv_fwd_halo_pl(const void* x_planes, const ConvParams& p, int accumulate, cudaStream_t st);   // p.x unused; weight packed as impl 4
bool wgrad_pl_supported(int n, int h, int w, int ci, int co, int k);
long long wgrad_pl_workspace_bytes(int n, int h, int w, int ci, int co, int k);
int conv_wgrad_pl(const void* x_planes, const void* gy_planes, float* dw, void* workspace, int n, int h, int w, int ci, int co, int k,
                  float coef, int accumulate, cudaStream_t st);

}  // namespace sg2
