// fp32 SIMT implicit-GEMM convolution (stride 1, "same" zero padding, k in {1,3}), NHWC.
// This is the exact-fp32 kernel family: it backs layers the tensor-core kernel does not take
// (channel counts not multiple of 8: RGB in/out, the 513-channel minibatch-stddev conv) and is the
// on-device cross-check for the tcgen05 kernels.  Replaces the ATen calls listed in sg2b200.h
// (implementations/StyleGAN2/model.py:106-132, :29-53) -- no per-sample weight tensor is ever built:
// the style scale is applied to the activation tile on load, demodulation in the epilogue.
//
// GEMM view (fwd):  M = n*h*w pixels, N = co, K = k*k*ci.   Tile 128 x BN x 16, 256 threads,
// 8 x (BN/16) outputs per thread, register-prefetch double buffering.
#include "common.cuh"
#include "conv.h"

namespace sg2 {

constexpr int kBM = 128, kBK = 16, kAPad = 4;

template <int BN>
__global__ void __launch_bounds__(256) conv_fwd_simt_kernel(ConvParams p) {
    constexpr int TN = BN / 16;
    constexpr int BLOADS = (4 * BN + 255) / 256;
    __shared__ __align__(16) float As[kBK][kBM + kAPad];
    __shared__ __align__(16) float Bs[kBK][BN + 4];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long P = (long long)p.n * p.h * p.w;
    const long long m0 = (long long)blockIdx.x * kBM;
    const int n0 = blockIdx.y * BN;
    const int pad = p.k >> 1;
    const int hw = p.h * p.w;
    const bool a_vec = (p.ci % 4) == 0, b_vec = (p.co % 4) == 0;

    // the two A slots this thread fetches: float4 index i = tid + 256*j -> pixel i/4, k-quad i%4
    int a_oy[2], a_ox[2], a_b[2];
    bool a_ok[2];
    const float* a_base[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = tid + 256 * j, m = i >> 2;
        const long long pix = m0 + m;
        a_ok[j] = pix < P;
        const long long pp = a_ok[j] ? pix : 0;
        a_b[j] = (int)(pp / hw);
        const int r = (int)(pp % hw);
        a_oy[j] = r / p.w; a_ox[j] = r % p.w;
        a_base[j] = p.x + pp * p.ci;
    }

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int kchunks = (p.ci + kBK - 1) / kBK;
    const int iters = p.k * p.k * kchunks;
    float4 ra[2], rb[BLOADS];

    auto fetch = [&](int it) {
        const int t = it / kchunks, c0 = (it % kchunks) * kBK;
        const int dy = t / p.k - pad, dx = t % p.k - pad;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int kq = (tid + 256 * j) & 3;
            const int cc = c0 + kq * 4;
            const int iy = a_oy[j] + dy, ix = a_ox[j] + dx;
            float4 v = f4zero();
            if (a_ok[j] && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w && cc < p.ci) {
                const float* src = a_base[j] + ((long long)dy * p.w + dx) * p.ci + cc;
                if (a_vec) {
                    v = ldg4(src);
                    if (p.in_scale) v = mul4(v, ldg4(p.in_scale + (long long)a_b[j] * p.ci + cc));
                } else {
                    float e[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        e[q] = (cc + q < p.ci) ? __ldg(src + q) : 0.f;
                        if (p.in_scale && cc + q < p.ci) e[q] *= __ldg(p.in_scale + (long long)a_b[j] * p.ci + cc + q);
                    }
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
            ra[j] = v;
        }
#pragma unroll
        for (int j = 0; j < BLOADS; ++j) {
            const int i = tid + 256 * j;
            float4 v = f4zero();
            if (i < 4 * BN) {
                const int kk = i / (BN / 4), nq = i % (BN / 4);
                const int cc = c0 + kk, nn = n0 + nq * 4;
                if (cc < p.ci && nn < p.co) {
                    const float* src = (const float*)p.wp + ((long long)t * p.ci + cc) * p.co + nn;
                    if (b_vec) v = ldg4(src);
                    else {
                        float e[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) e[q] = (nn + q < p.co) ? __ldg(src + q) : 0.f;
                        v = make_float4(e[0], e[1], e[2], e[3]);
                    }
                }
            }
            rb[j] = v;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = tid + 256 * j, m = i >> 2, kq = i & 3;
            As[kq * 4 + 0][m] = ra[j].x; As[kq * 4 + 1][m] = ra[j].y;
            As[kq * 4 + 2][m] = ra[j].z; As[kq * 4 + 3][m] = ra[j].w;
        }
#pragma unroll
        for (int j = 0; j < BLOADS; ++j) {
            const int i = tid + 256 * j;
            if (i < 4 * BN) { const int kk = i / (BN / 4), nq = i % (BN / 4); st4(&Bs[kk][nq * 4], rb[j]); }
        }
    };

    fetch(0);
    for (int it = 0; it < iters; ++it) {
        stash();
        __syncthreads();
        if (it + 1 < iters) fetch(it + 1);
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            float a[8], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if (TN == 2) { const float2 b0 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]); b[0] = b0.x; b[1] = b0.y; }
            if (TN >= 4) { const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]); b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; }
            if (TN == 8) { const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]); b[4 % TN] = b1.x; b[5 % TN] = b1.y; b[6 % TN] = b1.z; b[7 % TN] = b1.w; }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue: y = gain * act(out_scale * acc + bias + noise)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long pix = m0 + ty * 8 + i;
        if (pix >= P) continue;
        const int b = (int)(pix / hw), r = (int)(pix % hw), oy = r / p.w, ox = r % p.w;
        const float nz = p.noise ? __ldg(p.noise + pix) : 0.f;
        float* yp = p.y + b * p.ys[0] + oy * p.ys[2] + ox * p.ys[3];
#pragma unroll
        for (int g4 = 0; g4 < (TN + 3) / 4; ++g4) {
            constexpr int GW = TN < 4 ? TN : 4;
            const int cbase = n0 + (TN == 2 ? tx * 2 : (g4 == 0 ? tx * 4 : 64 + tx * 4));
            float o[4];
#pragma unroll
            for (int j = 0; j < GW; ++j) {
                const int co = cbase + j;
                float v = acc[i][g4 * 4 + j];
                if (co < p.co) {
                    if (p.out_scale) v *= __ldg(p.out_scale + (long long)b * p.co + co);
                    if (p.bias) v += __ldg(p.bias + co);
                    v += nz;
                    if (p.act == 3) v = v > 0.f ? v : v * p.alpha;
                    v *= p.gain;
                }
                o[j] = v;
            }
            if (GW == 4 && p.ys[1] == 1 && cbase + 3 < p.co && ((reinterpret_cast<uintptr_t>(yp + cbase) & 15) == 0)) {
                st4(yp + cbase, make_float4(o[0], o[1], o[2], o[3]));
            } else {
#pragma unroll
                for (int j = 0; j < GW; ++j) if (cbase + j < p.co) yp[(long long)(cbase + j) * p.ys[1]] = o[j];
            }
        }
    }
}

// Weight gradient.  Per tap t: dW_t[ci, co] = sum_pix x[pix + off(t), ci] * gy[pix, co].
// Tile 64(ci) x 64(co) x 16(pixels), split over pixel ranges (blockIdx.z), fp32 atomics into dw.
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(WgradParams p) {
    __shared__ __align__(16) float As[kBK][64 + 4];
    __shared__ __align__(16) float Bs[kBK][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int mtiles = (p.ci + 63) / 64;
    const int t = blockIdx.x / mtiles, m0 = (blockIdx.x % mtiles) * 64, n0 = blockIdx.y * 64;
    const int pad = p.k >> 1, dy = t / p.k - pad, dx = t % p.k - pad;
    const int hw = p.h * p.w;
    const long long P = (long long)p.n * hw;
    const long long pbeg = (long long)blockIdx.z * p.chunk, pend = min(P, pbeg + p.chunk);
    const int lk = tid >> 4, lq = tid & 15;     // loader: pixel lk of the step, channel quad lq
    const bool a_vec = (p.ci % 4) == 0, b_vec = (p.co % 4) == 0;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra, rb;
    auto fetch = [&](long long pstep) {
        const long long pix = pstep + lk;
        ra = f4zero(); rb = f4zero();
        if (pix < pend) {
            const int b = (int)(pix / hw), r = (int)(pix % hw), oy = r / p.w, ox = r % p.w;
            const int iy = oy + dy, ix = ox + dx, ca = m0 + lq * 4, cb = n0 + lq * 4;
            if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w && ca < p.ci) {
                const float* src = p.x + (pix + (long long)dy * p.w + dx) * p.ci + ca;
                if (a_vec) { ra = ldg4(src); if (p.in_scale) ra = mul4(ra, ldg4(p.in_scale + (long long)b * p.ci + ca)); }
                else {
                    float e[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        e[q] = (ca + q < p.ci) ? __ldg(src + q) : 0.f;
                        if (p.in_scale && ca + q < p.ci) e[q] *= __ldg(p.in_scale + (long long)b * p.ci + ca + q);
                    }
                    ra = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
            if (cb < p.co) {
                const float* src = p.gy + pix * p.co + cb;
                if (b_vec) { rb = ldg4(src); if (p.out_scale) rb = mul4(rb, ldg4(p.out_scale + (long long)b * p.co + cb)); }
                else {
                    float e[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        e[q] = (cb + q < p.co) ? __ldg(src + q) : 0.f;
                        if (p.out_scale && cb + q < p.co) e[q] *= __ldg(p.out_scale + (long long)b * p.co + cb + q);
                    }
                    rb = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
        }
    };

    if (pbeg < pend) fetch(pbeg);
    for (long long ps = pbeg; ps < pend; ps += kBK) {
        st4(&As[lk][lq * 4], ra);
        st4(&Bs[lk][lq * 4], rb);
        __syncthreads();
        if (ps + kBK < pend) fetch(ps + kBK);
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int kk2 = p.k * p.k;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = m0 + ty * 4 + i;
        if (ci >= p.ci) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tx * 4 + j;
            if (co < p.co) {
                const long long o = ((long long)co * p.ci + ci) * kk2 + t;
                if (p.ws) p.ws[(long long)blockIdx.z * ((long long)p.co * p.ci * kk2) + o] = acc[i][j] * p.coef;
                else atomicAdd(p.dw + o, acc[i][j] * p.coef);
            }
        }
    }
}

// w[co][ci][k][k] -> wp[t][kin][nout] * coef   (transpose: kin=co, nout=ci, taps flipped)
__global__ void conv_pack_simt_kernel(const float* __restrict__ w, float* __restrict__ wp, int co, int ci, int k,
                                      float coef, int transpose) {
    const int kk2 = k * k;
    const long long total = (long long)co * ci * kk2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int kin_n = transpose ? co : ci, nout_n = transpose ? ci : co;
        const int nout = (int)(idx % nout_n);
        long long r = idx / nout_n;
        const int kin = (int)(r % kin_n);
        const int t = (int)(r / kin_n);
        const int o = transpose ? kin : nout, i = transpose ? nout : kin, ts = transpose ? kk2 - 1 - t : t;
        wp[idx] = w[((long long)o * ci + i) * kk2 + ts] * coef;
    }
}

int conv_fwd_simt(const ConvParams& p, cudaStream_t st) {
    {   // RGB-side 1x1 layers: one coalesced pass instead of a padded GEMM tile (conv_thin.cu)
        const int rc = conv_fwd_thin(p, st);
        if (rc != SG2_ENOTSUP) return rc;
    }
    const long long P = (long long)p.n * p.h * p.w;
    const int mt = (int)ceil_div(P, kBM);
    if (p.co <= 32) {
        dim3 grid(mt, (p.co + 31) / 32);
        conv_fwd_simt_kernel<32><<<grid, 256, 0, st>>>(p);
    } else if (p.co <= 64) {
        dim3 grid(mt, 1);
        conv_fwd_simt_kernel<64><<<grid, 256, 0, st>>>(p);
    } else {
        dim3 grid(mt, (p.co + 127) / 128);
        conv_fwd_simt_kernel<128><<<grid, 256, 0, st>>>(p);
    }
    return launched("conv_fwd_simt");
}

static long long simt_wgrad_splits(const WgradParams& p, long long& chunk) {
    const long long P = (long long)p.n * p.h * p.w;
    const int kk2 = p.k * p.k;
    const int mtiles = (p.ci + 63) / 64, ntiles = (p.co + 63) / 64;
    const long long tiles = (long long)mtiles * ntiles * kk2;
    long long want = std::max<long long>(1, (4LL * num_sms()) / tiles);
    long long splits = std::min<long long>(want, ceil_div(P, 256));
    splits = std::max<long long>(1, std::min<long long>(splits, 65535));
    chunk = ceil_div(ceil_div(P, splits), kBK) * kBK;
    return ceil_div(P, chunk);
}

int wgrad_parts_simt(const WgradParams& p) {
    const int thin = wgrad_parts_thin(p);
    if (thin > 0) return thin;
    long long chunk;
    return (int)simt_wgrad_splits(p, chunk);
}

// dw[i] (+)= sum_parts ws[part][i], parts added in index order
__global__ void __launch_bounds__(256) wgrad_sum_parts_kernel(const float* __restrict__ ws, float* __restrict__ dw, long long size, int parts, int accumulate) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= size) return;
    float s = 0.f;
    for (int k = 0; k < parts; ++k) s += ws[(long long)k * size + i];
    dw[i] = accumulate ? dw[i] + s : s;
}

int wgrad_sum_parts(const float* ws, float* dw, long long size, int parts, int accumulate, cudaStream_t st) {
    wgrad_sum_parts_kernel<<<(unsigned)ceil_div(size, 256), 256, 0, st>>>(ws, dw, size, parts, accumulate);
    return launched("wgrad_sum_parts");
}

int conv_wgrad_simt(WgradParams p, int accumulate, cudaStream_t st) {
    const int kk2 = p.k * p.k;
    if (!accumulate && !p.ws) {
        cudaError_t e = cudaMemsetAsync(p.dw, 0, sizeof(float) * (size_t)p.co * p.ci * kk2, st);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "conv_wgrad: memset: %s", cudaGetErrorString(e));
    }
    {
        const int rc = conv_wgrad_thin(p, st);
        if (rc != SG2_ENOTSUP) return rc;
    }
    const int mtiles = (p.ci + 63) / 64, ntiles = (p.co + 63) / 64;
    long long chunk;
    const long long splits = simt_wgrad_splits(p, chunk);
    p.chunk = chunk;
    dim3 grid((unsigned)(mtiles * kk2), (unsigned)ntiles, (unsigned)splits);
    conv_wgrad_simt_kernel<<<grid, 256, 0, st>>>(p);
    return launched("conv_wgrad_simt");
}

int conv_pack_simt(const float* w, float* wp, int co, int ci, int k, float coef, int transpose, cudaStream_t st) {
    const long long total = (long long)co * ci * k * k;
    const int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 8);
    conv_pack_simt_kernel<<<blocks, 256, 0, st>>>(w, wp, co, ci, k, coef, transpose);
    return launched("conv_pack_simt");
}

}  // namespace sg2
