// "Thin" 1x1 convolutions: one side of the layer has at most 4 channels (RGB).
//
// On the StyleGAN2 path these are Discriminator.from_rgb (3 -> 32 @256^2, implementations/StyleGAN2/model.py:383-384) and
// the seven ToImage modulated convolutions (ci -> 3, model.py:239-250), plus their data and weight gradients.  They hold
// ~0.1 % of the step's flops but touch the largest tensors of the model, i.e. they are pure HBM streams: a GEMM-tiled
// kernel (conv_simt.cu pads K = 3 to 16 or N = 3 to 32) spends its time on zeros.  Here every kernel makes exactly one
// coalesced pass over the wide tensor:
//   thin_in_kernel    y[pix, co]   = act(os * sum_{c<CIN} x[pix,c] is[c] W[c,co] + bias + noise)     CIN <= 4, co % 4 == 0
//   thin_out_kernel   y[pix, o<=4] = act(os * sum_{tap,c} x[pix+tap,c] is[c] W[tap,c,o] + bias + noise)  ci % 4 == 0, k = 1 | 3
//                     (k = 3: the data gradient of the minibatch-stddev channel, 512 -> 1 @4^2, model.py:388-389)
//   thin_wgrad_kernel R[C, T<=4]   = sum_pix wide[pix,C] * thin[pix,T]  (per-sample scales applied per block)
// Weights arrive in the SIMT pack layout [cin][cout] (k = 1) of conv_pack_simt_kernel.  Algorithmic bytes = the wide
// tensor once (+ the thin one), the HBM roofline of DESIGN.md section 3.
#include "common.cuh"
#include "conv.h"

namespace sg2 {
namespace thin {

constexpr int kMaxW = 9 * 512;            // floats of weight kept in shared memory (k*k * thin * wide channels)

// ---------------------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256) thin_in_kernel(ConvParams p) {
    __shared__ __align__(16) float sw[kMaxW];                      // [CIN][co]
    for (int i = threadIdx.x; i < CIN * p.co; i += blockDim.x) sw[i] = ((const float*)p.wp)[i];
    __syncthreads();
    const int cq = p.co >> 2, hw = p.h * p.w;
    const long long P = (long long)p.n * hw;
    const bool per_sample = p.in_scale || p.out_scale;
    auto pixel = [&](long long pix, int q) {
        const int b = per_sample ? (int)(pix / hw) : 0;
        float4 acc = f4zero();
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            float xv = __ldg(p.x + pix * CIN + c);
            if (p.in_scale) xv *= __ldg(p.in_scale + (long long)b * CIN + c);
            fma4(acc, xv, *reinterpret_cast<const float4*>(&sw[c * p.co + 4 * q]));
        }
        if (p.out_scale) acc = mul4(acc, ldg4(p.out_scale + (long long)b * p.co + 4 * q));
        if (p.bias) acc = add4(acc, ldg4(p.bias + 4 * q));
        if (p.noise) { const float nz = __ldg(p.noise + pix); acc = add4(acc, make_float4(nz, nz, nz, nz)); }
        if (p.act == 3) {
            acc.x = acc.x > 0.f ? acc.x : acc.x * p.alpha; acc.y = acc.y > 0.f ? acc.y : acc.y * p.alpha;
            acc.z = acc.z > 0.f ? acc.z : acc.z * p.alpha; acc.w = acc.w > 0.f ? acc.w : acc.w * p.alpha;
        }
        st4_cs(p.y + pix * p.co + 4 * q, scale4(acc, p.gain));     // y is dense NHWC (checked by the launcher)
    };
    if (256 % cq == 0) {
        // the channel quad of a thread is fixed and its pixel advances by a constant: no division in the loop, two pixels in flight
        const int q = threadIdx.x % cq, ppb = 256 / cq;
        const long long step = (long long)gridDim.x * ppb;
        long long pix = (long long)blockIdx.x * ppb + threadIdx.x / cq;
        for (; pix + step < P; pix += 2 * step) { pixel(pix, q); pixel(pix + step, q); }
        if (pix < P) pixel(pix, q);
        return;
    }
    const long long total = P * cq;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        pixel(idx / cq, (int)(idx % cq));
}

// ---------------------------------------------------------------------------------------------------------------
// LANES lanes share one pixel: each takes the channel quads l, l + LANES, ...; partial dot products meet in a shuffle tree.
template <int COUT, int LANES>
__global__ void __launch_bounds__(256) thin_out_kernel(ConvParams p) {
    __shared__ __align__(16) float sw[kMaxW];                      // transposed to [tap][COUT][ci] for float4 reads
    const int taps = p.k * p.k, pad = p.k >> 1;
    for (int i = threadIdx.x; i < taps * COUT * p.ci; i += blockDim.x) {
        const int c = i % p.ci, o = (i / p.ci) % COUT, t = i / (p.ci * COUT);
        sw[i] = ((const float*)p.wp)[((long long)t * p.ci + c) * COUT + o];
    }
    __syncthreads();
    const int hw = p.h * p.w, cq = p.ci >> 2;
    const long long P = (long long)p.n * hw;
    const int lane = threadIdx.x & (LANES - 1);
    const long long slot = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / LANES;
    const long long nslots = (long long)gridDim.x * blockDim.x / LANES;
    for (long long pix0 = slot; pix0 < (P + nslots - 1) / nslots * nslots; pix0 += nslots) {   // uniform trip count per warp
        const bool ok = pix0 < P;
        const long long pix = ok ? pix0 : P - 1;
        const int b = (int)(pix / hw), r = (int)(pix % hw), oy = r / p.w, ox = r % p.w;
        float acc[COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
        for (int t = 0; t < taps; ++t) {
            const int iy = oy + t / p.k - pad, ix = ox + t % p.k - pad;
            if (iy < 0 || iy >= p.h || ix < 0 || ix >= p.w) continue;           // zero padding
            const float* xp = p.x + (((long long)b * p.h + iy) * p.w + ix) * p.ci;
            const float* wt = sw + t * COUT * p.ci;
            for (int q = lane; q < cq; q += LANES) {
                float4 xv = ldg4(xp + 4 * q);
                if (p.in_scale) xv = mul4(xv, ldg4(p.in_scale + (long long)b * p.ci + 4 * q));
#pragma unroll
                for (int o = 0; o < COUT; ++o) {
                    const float4 w4 = *reinterpret_cast<const float4*>(&wt[o * p.ci + 4 * q]);
                    acc[o] = fmaf(xv.x, w4.x, fmaf(xv.y, w4.y, fmaf(xv.z, w4.z, fmaf(xv.w, w4.w, acc[o]))));
                }
            }
        }
#pragma unroll
        for (int s = LANES >> 1; s > 0; s >>= 1)
#pragma unroll
            for (int o = 0; o < COUT; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
        if (ok && lane == 0) {
            const float nz = p.noise ? __ldg(p.noise + pix) : 0.f;
            float* yp = p.y + b * p.ys[0] + oy * p.ys[2] + ox * p.ys[3];
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                float v = acc[o];
                if (p.out_scale) v *= __ldg(p.out_scale + (long long)b * COUT + o);
                if (p.bias) v += __ldg(p.bias + o);
                v += nz;
                if (p.act == 3) v = v > 0.f ? v : v * p.alpha;
                yp[(long long)o * p.ys[1]] = v * p.gain;
            }
        }
    }
}

// Few input channels (ci <= 64, k = 1: the ToImage layers at 128^2 / 256^2): one thread per pixel.  A warp's 32 pixels are
// contiguous in x, so its float4 loads sweep one contiguous span and its stores are coalesced per output plane -- no
// shuffle tree, no lanes idling in the store.
template <int COUT>
__global__ void __launch_bounds__(256) thin_out_pix_kernel(ConvParams p) {
    __shared__ __align__(16) float sw[COUT * 64];                  // [COUT][ci]
    for (int i = threadIdx.x; i < COUT * p.ci; i += blockDim.x) {
        const int o = i / p.ci, c = i % p.ci;
        sw[i] = ((const float*)p.wp)[c * COUT + o];
    }
    __syncthreads();
    const int hw = p.h * p.w, cq = p.ci >> 2;
    const long long P = (long long)p.n * hw;
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < P; pix += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(pix / hw), r = (int)(pix % hw), oy = r / p.w, ox = r % p.w;
        const float* xp = p.x + pix * p.ci;
        const float* sp = p.in_scale ? p.in_scale + (long long)b * p.ci : nullptr;
        float acc[COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
        for (int q = 0; q < cq; ++q) {
            float4 xv = ldg4(xp + 4 * q);
            if (sp) xv = mul4(xv, ldg4(sp + 4 * q));
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                const float4 w4 = *reinterpret_cast<const float4*>(&sw[o * p.ci + 4 * q]);
                acc[o] = fmaf(xv.x, w4.x, fmaf(xv.y, w4.y, fmaf(xv.z, w4.z, fmaf(xv.w, w4.w, acc[o]))));
            }
        }
        const float nz = p.noise ? __ldg(p.noise + pix) : 0.f;
        float* yp = p.y + b * p.ys[0] + oy * p.ys[2] + ox * p.ys[3];
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            float v = acc[o];
            if (p.out_scale) v *= __ldg(p.out_scale + (long long)b * COUT + o);
            if (p.bias) v += __ldg(p.bias + o);
            v += nz;
            if (p.act == 3) v = v > 0.f ? v : v * p.alpha;
            yp[(long long)o * p.ys[1]] = v * p.gain;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// R[c, t] = sum_pix wide[pix, c] * thin[pix, t]; grid = (splits, n): a block stays inside one sample so the per-sample
// scales factor out of its partial sum.  dw[co][ci] gets R (thin side = ci) or R^T (thin side = co) through fp32 atomics.
struct WgParams {
    const float* wide; const float* thin;      // [P, C], [P, T] dense
    const float* wide_scale; const float* thin_scale;   // [n, C], [n, T] or null
    float* dw;
    float* ws;                                  // per-block partial buffers (deterministic reduction) or null (atomics)
    int n, hw, C, T;
    int thin_is_ci;                             // 1: dw[c_wide][t]  (from_rgb), 0: dw[t][c_wide] (ToImage)
    float coef;
    int pix_per_block;
};

template <int T>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(WgParams p) {
    __shared__ float red[256 * 4 * T];
    const int cq = p.C >> 2;
    const int b = blockIdx.y;
    const int pbeg = blockIdx.x * p.pix_per_block, pend = min(p.hw, pbeg + p.pix_per_block);
    float acc[4][T];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int t = 0; t < T; ++t) acc[j][t] = 0.f;
    // thread = (pixel slot, channel quad): consecutive threads read consecutive 16 B of one pixel
    const int q = threadIdx.x % cq, slot = threadIdx.x / cq, nslot = blockDim.x / cq;
    if (slot < nslot) {
        const float* wp = p.wide + ((long long)b * p.hw) * p.C + 4 * q;
        const float* tp = p.thin + ((long long)b * p.hw) * T;
        for (int pix = pbeg + slot; pix < pend; pix += nslot) {
            const float4 w4 = ldg4(wp + (long long)pix * p.C);
            float th[T];
#pragma unroll
            for (int t = 0; t < T; ++t) th[t] = __ldg(tp + (long long)pix * T + t);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                acc[0][t] = fmaf(w4.x, th[t], acc[0][t]); acc[1][t] = fmaf(w4.y, th[t], acc[1][t]);
                acc[2][t] = fmaf(w4.z, th[t], acc[2][t]); acc[3][t] = fmaf(w4.w, th[t], acc[3][t]);
            }
        }
    }
    // reduce over the pixel slots of the block (threads with the same q), then one atomic per (channel, t)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int t = 0; t < T; ++t) red[(j * T + t) * 256 + threadIdx.x] = acc[j][t];
    __syncthreads();
    for (int o = threadIdx.x; o < cq * 4 * T; o += blockDim.x) {
        const int qq = o % cq, jt = o / cq;                       // jt = j * T + t
        float s = 0.f;
        for (int sl = 0; sl < nslot; ++sl) s += red[jt * 256 + sl * cq + qq];
        const int j = jt / T, t = jt % T, c = 4 * qq + j;
        float sc = p.coef;
        if (p.wide_scale) sc *= __ldg(p.wide_scale + (long long)b * p.C + c);
        if (p.thin_scale) sc *= __ldg(p.thin_scale + (long long)b * T + t);
        const long long el = p.thin_is_ci ? (long long)c * T + t : (long long)t * p.C + c;
        if (p.ws) p.ws[((long long)blockIdx.y * gridDim.x + blockIdx.x) * ((long long)p.C * T) + el] = s * sc;
        else atomicAdd(p.dw + el, s * sc);
    }
}

}  // namespace thin

// The launchers below are tried first by conv_fwd_simt / conv_wgrad_simt; they return SG2_ENOTSUP when the shape or
// layout is not the thin one, and the generic fp32 kernels run instead.
int conv_fwd_thin(const ConvParams& p, cudaStream_t st) {
    const long long P = (long long)p.n * p.h * p.w;
    if (P > 2147483647LL / 4) return SG2_ENOTSUP;
    const bool y_dense_nhwc = p.ys[1] == 1 && p.ys[3] == p.co && p.ys[2] == (long long)p.w * p.co && p.ys[0] == (long long)p.h * p.w * p.co;
    if (p.k == 1 && p.ci <= 4 && (p.co % 4) == 0 && p.ci * p.co <= thin::kMaxW && y_dense_nhwc && ((uintptr_t)p.y % 16) == 0) {
        const long long work = P * (p.co / 4);
        const int blocks = (int)std::min<long long>(ceil_div(work, 256), (long long)num_sms() * 16);
        switch (p.ci) {
            case 1: thin::thin_in_kernel<1><<<blocks, 256, 0, st>>>(p); break;
            case 2: thin::thin_in_kernel<2><<<blocks, 256, 0, st>>>(p); break;
            case 3: thin::thin_in_kernel<3><<<blocks, 256, 0, st>>>(p); break;
            default: thin::thin_in_kernel<4><<<blocks, 256, 0, st>>>(p); break;
        }
        return launched("conv_thin_in");
    }
    if (p.co <= 4 && (p.ci % 4) == 0 && p.k * p.k * p.ci * p.co <= thin::kMaxW && ((uintptr_t)p.x % 16) == 0) {
        const int cq = p.ci / 4;
        if (p.k == 1 && p.ci <= 64 && P >= 65536) {
            const int blocks = (int)std::min<long long>(ceil_div(P, 256), (long long)num_sms() * 16);
            switch (p.co) {
                case 1: thin::thin_out_pix_kernel<1><<<blocks, 256, 0, st>>>(p); break;
                case 2: thin::thin_out_pix_kernel<2><<<blocks, 256, 0, st>>>(p); break;
                case 3: thin::thin_out_pix_kernel<3><<<blocks, 256, 0, st>>>(p); break;
                default: thin::thin_out_pix_kernel<4><<<blocks, 256, 0, st>>>(p); break;
            }
            return launched("conv_thin_out_pix");
        }
        const int lanes = cq >= 32 ? 32 : (cq >= 16 ? 16 : (cq >= 8 ? 8 : 4));
        const int blocks = (int)std::min<long long>(ceil_div(P * lanes, 256), (long long)num_sms() * 16);
#define SG2_THIN_OUT(CO, L) thin::thin_out_kernel<CO, L><<<blocks, 256, 0, st>>>(p)
#define SG2_THIN_OUT_L(CO) do { if (lanes == 32) SG2_THIN_OUT(CO, 32); else if (lanes == 16) SG2_THIN_OUT(CO, 16); \
                                else if (lanes == 8) SG2_THIN_OUT(CO, 8); else SG2_THIN_OUT(CO, 4); } while (0)
        switch (p.co) {
            case 1: SG2_THIN_OUT_L(1); break;
            case 2: SG2_THIN_OUT_L(2); break;
            case 3: SG2_THIN_OUT_L(3); break;
            default: SG2_THIN_OUT_L(4); break;
        }
#undef SG2_THIN_OUT_L
#undef SG2_THIN_OUT
        return launched("conv_thin_out");
    }
    return SG2_ENOTSUP;
}

static bool thin_wgrad_plan(const WgradParams& wp, thin::WgParams& p, int& splits) {
    if (wp.k != 1) return false;
    if (wp.ci <= 4 && (wp.co % 4) == 0 && wp.co <= 1024) {
        p.wide = wp.gy; p.thin = wp.x; p.wide_scale = wp.out_scale; p.thin_scale = wp.in_scale;
        p.C = wp.co; p.T = wp.ci; p.thin_is_ci = 1;
    } else if (wp.co <= 4 && (wp.ci % 4) == 0 && wp.ci <= 1024) {
        p.wide = wp.x; p.thin = wp.gy; p.wide_scale = wp.in_scale; p.thin_scale = wp.out_scale;
        p.C = wp.ci; p.T = wp.co; p.thin_is_ci = 0;
    } else {
        return false;
    }
    if (((uintptr_t)p.wide % 16) != 0) return false;
    p.dw = wp.dw; p.ws = wp.ws; p.n = wp.n; p.hw = wp.h * wp.w; p.coef = wp.coef;
    // enough blocks to fill the machine, at least 64 pixels per pixel slot
    const int nslot = 256 / (p.C / 4) > 0 ? 256 / (p.C / 4) : 1;
    splits = std::max(1, std::min((4 * num_sms() + p.n - 1) / p.n, p.hw / (nslot * 16) + 1));
    p.pix_per_block = (p.hw + splits - 1) / splits;
    splits = (p.hw + p.pix_per_block - 1) / p.pix_per_block;
    return true;
}

int wgrad_parts_thin(const WgradParams& wp) {
    thin::WgParams p;
    int splits;
    return thin_wgrad_plan(wp, p, splits) ? splits * wp.n : 0;
}

int conv_wgrad_thin(const WgradParams& wp, cudaStream_t st) {
    thin::WgParams p;
    int splits;
    if (!thin_wgrad_plan(wp, p, splits)) return SG2_ENOTSUP;
    dim3 grid((unsigned)splits, (unsigned)p.n);
    switch (p.T) {
        case 1: thin::thin_wgrad_kernel<1><<<grid, 256, 0, st>>>(p); break;
        case 2: thin::thin_wgrad_kernel<2><<<grid, 256, 0, st>>>(p); break;
        case 3: thin::thin_wgrad_kernel<3><<<grid, 256, 0, st>>>(p); break;
        default: thin::thin_wgrad_kernel<4><<<grid, 256, 0, st>>>(p); break;
    }
    return launched("conv_wgrad_thin");
}

}  // namespace sg2
