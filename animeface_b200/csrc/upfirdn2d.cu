// upfirdn2d for sm_100a: pad -> zero-insert upsample -> FIR -> decimate in one launch.
// Semantics follow thirdparty/stylegan3_ops/ops/upfirdn2d.py:161-207 (_upfirdn2d_ref) and the
// plugin entry thirdparty/stylegan3_ops/ops/upfirdn2d.cpp:10; the kernels are new.
//
//   y[n,c,jy,jx] = gain * sum_{ty,tx} U[jy*downy + ty - pady0, jx*downx + tx - padx0] * wt[ty][tx]
//   U[ky,kx]     = x[ky/upy, kx/upx] when ky%upy==0, kx%upx==0 and inside the image, else 0
//   wt[ty][tx]   = flip ? f[ty][tx] : f[fh-1-ty][fw-1-tx]
//
// HBM-bound (algorithmic bytes = |x| + |y|).  Two kernels:
//   * upfirdn2d_vec4_nhwc : fp32 channels_last, C % 4 == 0 -- one thread = one output pixel x 4
//     channels, 128-bit coalesced loads/stores, taps served from L1, only non-zero phases visited.
//   * upfirdn2d_generic<T>: any dtype / strides / factors, thread order follows y's memory order.
#include <stdlib.h>
#include "common.cuh"

namespace sg2 {

struct UpfirdnParams {
    const void* x; const float* f; void* y;
    int n, c, in_h, in_w, out_h, out_w;
    long long xs[4], ys[4];          // element strides (n,c,h,w)
    int fh, fw, upx, upy, downx, downy, padx0, pady0, flip;
    float gain;
    int order[4];                    // dims of y sorted by decreasing stride (memory order)
    long long total;
    int sepok;                       // ring walk: separable evaluation allowed (SG2_UPF_SEP=0 disables it: the A/B switch)
};

__device__ __forceinline__ int pos_mod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

template <class T> struct Acc { typedef float type; };
template <> struct Acc<double> { typedef double type; };

template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_generic(UpfirdnParams p) {
    typedef typename Acc<T>::type acc_t;
    const T* __restrict__ x = (const T*)p.x;
    T* __restrict__ y = (T*)p.y;
    const int size[4] = {p.n, p.c, p.out_h, p.out_w};
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        int coord[4];
        long long r = idx;
#pragma unroll
        for (int k = 3; k >= 0; --k) { int d = p.order[k]; coord[d] = (int)(r % size[d]); r /= size[d]; }
        const int in_ = coord[0], ic = coord[1], jy = coord[2], jx = coord[3];
        const int basey = jy * p.downy - p.pady0, basex = jx * p.downx - p.padx0;
        const int ty0 = pos_mod(-basey, p.upy), tx0 = pos_mod(-basex, p.upx);
        const T* xp = x + in_ * p.xs[0] + ic * p.xs[1];
        acc_t acc = 0;
        for (int ty = ty0; ty < p.fh; ty += p.upy) {
            int ky = basey + ty;
            if (ky < 0) continue;
            int iy = ky / p.upy;
            if (iy >= p.in_h) break;
            int fy = p.flip ? ty : p.fh - 1 - ty;
            for (int tx = tx0; tx < p.fw; tx += p.upx) {
                int kx = basex + tx;
                if (kx < 0) continue;
                int ix = kx / p.upx;
                if (ix >= p.in_w) break;
                int fx = p.flip ? tx : p.fw - 1 - tx;
                acc += (acc_t)xp[iy * p.xs[2] + ix * p.xs[3]] * (acc_t)p.f[fy * p.fw + fx];
            }
        }
        y[in_ * p.ys[0] + ic * p.ys[1] + jy * p.ys[2] + jx * p.ys[3]] = (T)(acc * (acc_t)p.gain);
    }
}

// fp32 channels_last fast path.  Grid: x = pixel tiles of one image row strip, threads cover
// (pixels-in-tile x channel quads) with the channel quad fastest => 128-bit coalesced.
constexpr int kMaxTaps = 32 * 32;

__global__ void __launch_bounds__(256) upfirdn2d_vec4_nhwc(UpfirdnParams p) {
    __shared__ float sf[kMaxTaps];
    for (int i = threadIdx.x; i < p.fh * p.fw; i += blockDim.x) {
        int ty = i / p.fw, tx = i % p.fw;          // stored already oriented: sf[ty][tx] = wt * gain
        int fy = p.flip ? ty : p.fh - 1 - ty, fx = p.flip ? tx : p.fw - 1 - tx;
        sf[i] = p.f[fy * p.fw + fx] * p.gain;
    }
    __syncthreads();
    const float* __restrict__ x = (const float*)p.x;
    float* __restrict__ y = (float*)p.y;
    const int cq = p.c >> 2;
    const long long total = (long long)p.n * p.out_h * p.out_w * cq;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int q = (int)(idx % cq);
        long long pix = idx / cq;
        int jx = (int)(pix % p.out_w); pix /= p.out_w;
        int jy = (int)(pix % p.out_h);
        int in_ = (int)(pix / p.out_h);
        const int basey = jy * p.downy - p.pady0, basex = jx * p.downx - p.padx0;
        const int ty0 = pos_mod(-basey, p.upy), tx0 = pos_mod(-basex, p.upx);
        const float* xp = x + in_ * p.xs[0] + 4 * q;
        float4 acc = f4zero();
        for (int ty = ty0; ty < p.fh; ty += p.upy) {
            int ky = basey + ty;
            if (ky < 0) continue;
            int iy = ky / p.upy;
            if (iy >= p.in_h) break;
            const float* xr = xp + iy * p.xs[2];
            const float* fr = sf + ty * p.fw;
#pragma unroll 4
            for (int tx = tx0; tx < p.fw; tx += p.upx) {
                int kx = basex + tx;
                if (kx < 0) continue;
                int ix = kx / p.upx;
                if (ix >= p.in_w) break;
                fma4(acc, fr[tx], ldg4(xr + ix * p.xs[3]));
            }
        }
        st4_cs(y + in_ * p.ys[0] + jy * p.ys[2] + jx * p.ys[3] + 4 * q, acc);
    }
}


// Strip version of the channels_last fast path: block = (x range, strip of output rows, image); a thread owns one output
// column x 4 channels and walks down the strip.  All index math is 32-bit; the x-direction tap range is hoisted.
constexpr int kUpfStrip = 8;

__global__ void __launch_bounds__(256) upfirdn2d_strip_nhwc(UpfirdnParams p) {
    __shared__ float sf[kMaxTaps];
    for (int i = threadIdx.x; i < p.fh * p.fw; i += blockDim.x) {
        int ty = i / p.fw, tx = i % p.fw;
        int fy = p.flip ? ty : p.fh - 1 - ty, fx = p.flip ? tx : p.fw - 1 - tx;
        sf[i] = p.f[fy * p.fw + fx] * p.gain;
    }
    __syncthreads();
    const int cq = p.c >> 2, xs = 256 / cq;
    const int q = threadIdx.x % cq, jx = blockIdx.x * xs + threadIdx.x / cq;
    if (jx >= p.out_w) return;
    const int img = blockIdx.z, jy0 = blockIdx.y * kUpfStrip, jy1 = min(p.out_h, jy0 + kUpfStrip);
    const float* __restrict__ x = (const float*)p.x + (size_t)img * p.xs[0] + 4 * q;
    float* __restrict__ y = (float*)p.y + (size_t)img * p.ys[0] + 4 * q;
    const int xrow = (int)p.xs[2], xpix = (int)p.xs[3], yrow = (int)p.ys[2], ypix = (int)p.ys[3];
    // x-direction taps of this column: tx = txa + k*upx, input column ixa + k, k in [0, nx)
    const int basex = jx * p.downx - p.padx0;
    int txa = pos_mod(-basex, p.upx);
    int ixa = (basex + txa) / p.upx;              // exact division (may be negative)
    if (ixa < 0) { txa += -ixa * p.upx; ixa = 0; }
    int nx = txa < p.fw ? (p.fw - 1 - txa) / p.upx + 1 : 0;
    nx = min(nx, p.in_w - ixa);
    for (int jy = jy0; jy < jy1; ++jy) {
        const int basey = jy * p.downy - p.pady0;
        int tya = pos_mod(-basey, p.upy);
        int iya = (basey + tya) / p.upy;
        if (iya < 0) { tya += -iya * p.upy; iya = 0; }
        int ny = tya < p.fh ? (p.fh - 1 - tya) / p.upy + 1 : 0;
        ny = min(ny, p.in_h - iya);
        float4 acc = f4zero();
        for (int a = 0; a < ny; ++a) {
            const float* xr = x + (size_t)(iya + a) * xrow + (size_t)ixa * xpix;
            const float* fr = sf + (tya + a * p.upy) * p.fw + txa;
#pragma unroll 4
            for (int b = 0; b < nx; ++b) fma4(acc, fr[b * p.upx], ldg4(xr + (size_t)b * xpix));
        }
        st4_cs(y + (size_t)jy * yrow + (size_t)jx * ypix, acc);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Register-ring fast path for up = 1 (FIR, optionally decimating) -- the blur in front of every discriminator
// convolution of the StyleGAN3-style networks (conv2d_resample.py:101-108: [1,3,3,1] x [1,3,3,1], down 1 or 2).
//   y[jy][jx] = sum_{a<FH} sum_{b<FW} wt[a][b] * x[jy*DOWN - pady0 + a][jx*DOWN - padx0 + b]      (zero outside)
// A thread owns one output column (x 4 channels for NHWC, V = float4; one plane for NCHW, V = float) and walks down a
// strip of STRIP output rows keeping the FH x FW input window in registers: going one row down costs DOWN x FW loads
// instead of FH x FW -- each input element is fetched FW / DOWN times per thread-column instead of FH*FW / DOWN^2, and
// neighbouring columns' fetches hit L1.  The strip loop is fully unrolled so the ring is addressed with constants.
template <bool B> struct BoolTag { static constexpr bool value = B; };
template <class V> struct RingVec;
template <> struct RingVec<float4> {
    static __device__ __forceinline__ float4 zero() { return f4zero(); }
    static __device__ __forceinline__ float4 ld(const float* p) { return ldg4(p); }
    static __device__ __forceinline__ void st(float* p, const float4& v) { st4_cs(p, v); }
    static __device__ __forceinline__ void fma(float4& a, float s, const float4& v) { fma4(a, s, v); }
    static __device__ __forceinline__ void fma_x2(float4& a, float s, const float4& v) { fma4_x2(a, s, v); }
};
template <> struct RingVec<float> {
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void st(float* p, float v) { __stcs(p, v); }
    static __device__ __forceinline__ void fma(float& a, float s, float v) { a = fmaf(s, v, a); }
    static __device__ __forceinline__ void fma_x2(float& a, float s, float v) { a = fmaf(s, v, a); }
};

template <class V, int FH, int FW, int DOWN, int COLS, int STRIP>
__global__ void __launch_bounds__(256) upfirdn2d_ring_kernel(UpfirdnParams p) {
    typedef RingVec<V> R;
    constexpr bool NHWC = sizeof(V) == 16;
    constexpr int WW = (COLS - 1) * DOWN + FW;     // input window width of the COLS adjacent outputs of a thread
    // weights oriented like the generic kernels: wt[a][b] = f[flip ? a : FH-1-a][flip ? b : FW-1-b] * gain
    float wt[FH][FW];
#pragma unroll
    for (int a = 0; a < FH; ++a)
#pragma unroll
        for (int b = 0; b < FW; ++b)
            wt[a][b] = __ldg(p.f + (p.flip ? a : FH - 1 - a) * FW + (p.flip ? b : FW - 1 - b)) * p.gain;
    int jx, img, strip = blockIdx.y;
    const float* x; float* y;
    int xrow, xcol, yrow, ycol;                    // element strides of one row / one column step
    if (NHWC) {
        const int cq = p.c >> 2, xs = 256 / cq;
        const int q = threadIdx.x % cq;
        jx = (blockIdx.x * xs + threadIdx.x / cq) * COLS;
        img = blockIdx.z;
        x = (const float*)p.x + (size_t)img * p.xs[0] + 4 * q;
        y = (float*)p.y + (size_t)img * p.ys[0] + 4 * q;
        xcol = p.c; ycol = p.c; xrow = p.in_w * p.c; yrow = p.out_w * p.c;
    } else {
        // one plane: work items (strip, column group) flattened so that a block is full whatever out_w is
        const int ncols = (p.out_w + COLS - 1) / COLS;
        const int item = blockIdx.x * 256 + threadIdx.x;
        strip = item / ncols;
        jx = (item - strip * ncols) * COLS;
        img = blockIdx.z;                          // plane index n * C + c
        x = (const float*)p.x + (size_t)img * p.in_h * p.in_w;
        y = (float*)p.y + (size_t)img * p.out_h * p.out_w;
        xcol = 1; ycol = 1; xrow = p.in_w; yrow = p.out_w;
    }
    if (jx >= p.out_w || strip * STRIP >= p.out_h) return;
    const int jy0 = strip * STRIP;
    const int ix0 = jx * DOWN - p.padx0, iy0 = jy0 * DOWN - p.pady0;
    bool colok[WW];
#pragma unroll
    for (int b = 0; b < WW; ++b) colok[b] = (unsigned)(ix0 + b) < (unsigned)p.in_w;
    const float* xc = x + (ptrdiff_t)ix0 * xcol;
    if constexpr (DOWN > 1 && NHWC) {
        // Input-stationary ring: the FH x WW input window lives in registers, going one output row down loads DOWN new
        // rows.  Measured on B200 (U4, fraction of HBM peak, this ring | the output-stationary walk below): NHWC down 2
        // 0.86 | 0.81, NHWC filter 0.59 | 0.64, NCHW down 2 0.65 | 0.74, NCHW filter 0.49 | 0.56 -- hence this condition.
        V ring[FH][WW];
        auto load_row = [&](int k, V (&dst)[WW]) {     // input row iy0 + k
            const int iy = iy0 + k;
            const bool rowok = (unsigned)iy < (unsigned)p.in_h;
            const float* xr = xc + (ptrdiff_t)iy * xrow;
#pragma unroll
            for (int b = 0; b < WW; ++b) dst[b] = (rowok && colok[b]) ? R::ld(xr + (ptrdiff_t)b * xcol) : R::zero();
        };
#pragma unroll
        for (int k = 0; k < FH - DOWN; ++k) load_row(k, ring[k % FH]);
#pragma unroll
        for (int s = 0; s < STRIP; ++s) {
#pragma unroll
            for (int k = FH - DOWN; k < FH; ++k) load_row(s * DOWN + k, ring[(s * DOWN + k) % FH]);
#pragma unroll
            for (int e = 0; e < COLS; ++e) {
                V acc = R::zero();
#pragma unroll
                for (int a = 0; a < FH; ++a)
#pragma unroll
                    for (int b = 0; b < FW; ++b) R::fma(acc, wt[a][b], ring[(s * DOWN + a) % FH][e * DOWN + b]);
                if (jy0 + s < p.out_h && jx + e < p.out_w) R::st(y + (size_t)(jy0 + s) * yrow + (size_t)(jx + e) * ycol, acc);
            }
        }
        return;
    }
    // Output-stationary walk: an input row is loaded once (WW values, transient) and scattered into the NA = ceil(FH/DOWN)
    // output rows it contributes to; an output row is stored when its last input row has passed.  The ring holds
    // accumulators (NA x COLS values) instead of the FH x WW input window, so a thread needs ~50 registers instead of
    // ~80: more threads per SM and room for the compiler to issue the next rows' loads early -- the kernel is bound by
    // memory-level parallelism, not by the FMAs.  Fully unrolled: every ring index is a constant.
    constexpr int NA = (FH + DOWN - 1) / DOWN;
    constexpr int NROWS = (STRIP - 1) * DOWN + FH;           // input rows a strip touches
    // A rank-1 filter (setup_filter's outer product of a 1-D kernel: every blur of the path) is applied separably: the
    // horizontal taps once per input row, then one FMA per output row the input row feeds -- COLS * (FW + NA) multiply-adds per
    // input row instead of COLS * NA * FW.  Detected from the weights themselves (exact for the dyadic [1,3,3,1] family; a
    // filter that is not recognised takes the general walk, same results to rounding).
    bool sep = wt[0][0] != 0.f;
#pragma unroll
    for (int a = 0; a < FH; ++a)
#pragma unroll
        for (int b = 0; b < FW; ++b) {
            const float lhs = wt[a][b] * wt[0][0], rhs = wt[a][0] * wt[0][b];
            sep = sep && fabsf(lhs - rhs) <= 2.4e-7f * fabsf(rhs);
        }
    float fx[FW], fy[FH];
#pragma unroll
    for (int b = 0; b < FW; ++b) fx[b] = sep ? wt[0][b] / wt[0][0] : 0.f;
#pragma unroll
    for (int a = 0; a < FH; ++a) fy[a] = wt[a][0];
    V acc[NA][COLS];
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int e = 0; e < COLS; ++e) acc[i][e] = R::zero();
    // packed f32x2 FMAs in this walk: measured on B200 (U4 filter NHWC, fraction of HBM peak) general 0.54 -> 0.56,
    // separable 0.59 -> 0.68; the input-stationary ring above keeps scalar FMAs (packed: 0.86 -> 0.70, its rotating window
    // does not sit in aligned register pairs)
    // (NCHW rows from three aligned 128-bit loads per 4 outputs instead of 7 scalar ones were tried: 0.57 -> 0.50 of HBM peak --
    // 12 values fetched for 7 used, plus the selects that pick the window out of them)
    auto walk = [&](auto separable) {
        constexpr bool SEP = decltype(separable)::value;
        auto fma = [](V& a, float s, const V& v) { R::fma_x2(a, s, v); };
#pragma unroll
        for (int k = 0; k < NROWS; ++k) {
            V row[WW];
            {
                const int iy = iy0 + k;
                const bool rowok = (unsigned)iy < (unsigned)p.in_h;
                const float* xr = xc + (ptrdiff_t)iy * xrow;
#pragma unroll
                for (int b = 0; b < WW; ++b) row[b] = (rowok && colok[b]) ? R::ld(xr + (ptrdiff_t)b * xcol) : R::zero();
            }
            V hrow[COLS];
            if (SEP) {
#pragma unroll
                for (int e = 0; e < COLS; ++e) {
                    hrow[e] = R::zero();
#pragma unroll
                    for (int b = 0; b < FW; ++b) fma(hrow[e], fx[b], row[e * DOWN + b]);
                }
            }
#pragma unroll
            for (int a = 0; a < FH; ++a) {
                // input row k is tap row a of output row s = (k - a) / DOWN
                if ((k - a) >= 0 && (k - a) % DOWN == 0 && (k - a) / DOWN < STRIP) {
                    const int s = (k - a) / DOWN;
#pragma unroll
                    for (int e = 0; e < COLS; ++e) {
                        if (SEP) fma(acc[s % NA][e], fy[a], hrow[e]);
                        else {
#pragma unroll
                            for (int b = 0; b < FW; ++b) fma(acc[s % NA][e], wt[a][b], row[e * DOWN + b]);
                        }
                    }
                    if (a == FH - 1) {                           // last tap row of output s: store and recycle the accumulator
#pragma unroll
                        for (int e = 0; e < COLS; ++e) {
                            if (jy0 + s < p.out_h && jx + e < p.out_w) R::st(y + (size_t)(jy0 + s) * yrow + (size_t)(jx + e) * ycol, acc[s % NA][e]);
                            acc[s % NA][e] = R::zero();
                        }
                    }
                }
            }
        }
    };
    if (sep && p.sepok) walk(BoolTag<true>{}); else walk(BoolTag<false>{});
}

template <class V, int FH, int FW, int DOWN, int COLS, int STRIP>
static int launch_ring_strip(const UpfirdnParams& p_in, cudaStream_t st) {
    constexpr bool NHWC = sizeof(V) == 16;
    static int sepok = -1;
    if (sepok < 0) { const char* e = getenv("SG2_UPF_SEP"); sepok = e ? atoi(e) != 0 : 1; }
    UpfirdnParams p = p_in;
    p.sepok = sepok;
    dim3 grid;
    if (NHWC) grid = dim3((unsigned)ceil_div(ceil_div(p.out_w, COLS), 256 / (p.c / 4)), (unsigned)ceil_div(p.out_h, STRIP), (unsigned)p.n);
    else grid = dim3((unsigned)ceil_div(ceil_div(p.out_w, COLS) * ceil_div(p.out_h, STRIP), 256), 1u, (unsigned)(p.n * p.c));
    upfirdn2d_ring_kernel<V, FH, FW, DOWN, COLS, STRIP><<<grid, 256, 0, st>>>(p);
    return launched("upfirdn2d_ring");
}

template <class V, int FH, int FW, int DOWN, int COLS>
static int launch_ring_cols(const UpfirdnParams& p, cudaStream_t st) {
    // Rows per strip: a strip re-reads FH - DOWN halo rows, so 16 rows cost 19 % extra reads for the 4x4 blur and 32 rows 9 %;
    // but the strip body is fully unrolled, and an image of 257 rows (U4: 256 + the blur's padding) would leave a 9th strip of 32
    // with one row.  33 rows cut it into 8 strips.  Small images keep 16 (more blocks).  SG2_UPF_STRIP overrides (sweep).
    static int forced = 0;
    if (!forced) { const char* e = getenv("SG2_UPF_STRIP"); forced = e ? atoi(e) : -1; }
    int strip = forced > 0 ? forced : (p.out_h >= 128 ? (p.out_h % 32 == 0 ? 32 : (p.out_h % 33 <= 1 || p.out_h % 32 <= 1 ? 33 : 32)) : 16);
    if (DOWN != 1 && forced <= 0) strip = 16;
    if (strip == 33) return launch_ring_strip<V, FH, FW, DOWN, COLS, 33>(p, st);
    if (strip == 32) return launch_ring_strip<V, FH, FW, DOWN, COLS, 32>(p, st);
    return launch_ring_strip<V, FH, FW, DOWN, COLS, 16>(p, st);
}

template <class V, int FH, int FW, int DOWN>
static int launch_ring(const UpfirdnParams& p, cudaStream_t st) {
    // NHWC: a thread = one pixel column x 4 channels.  NCHW: COLS adjacent columns of one plane.
    if constexpr (sizeof(V) == 16) {
        // adjacent pixels per thread: with 1 every input pixel is fetched FW times through L1 (4 x the data for the 4x4 blur);
        // 4 pixels share a 7-wide window (1.75 x).  Measured on B200 with the separable walk (U4 filter, fraction of HBM peak):
        // 1 -> 0.68, 2 -> 0.69, 4 -> 0.73; the decimating form keeps 1 (0.86; 2 -> 0.71, 4 -> 0.31).  SG2_UPF_COLS_NHWC overrides.
        static int cols = 0;
        if (!cols) { const char* e = getenv("SG2_UPF_COLS_NHWC"); cols = e ? atoi(e) : -1; }
        const int use = cols > 0 ? cols : (DOWN == 1 ? 4 : 1);
        if (use == 4) return launch_ring_cols<V, FH, FW, DOWN, 4>(p, st);
        if (use == 2) return launch_ring_cols<V, FH, FW, DOWN, 2>(p, st);
        return launch_ring_cols<V, FH, FW, DOWN, 1>(p, st);
    } else {
        // columns per thread, measured on B200 (U4, NCHW, fraction of HBM peak): filter 1 -> 0.35, 2 -> 0.48, 4 -> 0.56;
        // down 2: 1 -> 0.54, 2 -> 0.74, 4 -> 0.52.  SG2_UPF_COLS overrides the choice for that sweep.
        static int cols = 0;
        if (!cols) { const char* e = getenv("SG2_UPF_COLS"); cols = e ? atoi(e) : -1; }
        if (cols < 0) return DOWN == 1 ? launch_ring_cols<V, FH, FW, DOWN, 4>(p, st) : launch_ring_cols<V, FH, FW, DOWN, 2>(p, st);
        if (cols == 4) return launch_ring_cols<V, FH, FW, DOWN, 4>(p, st);
        if (cols == 2) return launch_ring_cols<V, FH, FW, DOWN, 2>(p, st);
        return launch_ring_cols<V, FH, FW, DOWN, 1>(p, st);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// NCHW filter form (up = down = 1) with coalesced 128-bit row loads and warp shuffles: the ring kernel's thread fetches the 7-wide
// window of its 4 outputs with 7 scalar loads that overlap 7-fold between neighbours (L1-bound: 0.57 of HBM peak on U4; this
// kernel: 0.61).  Here a
// lane loads ONE aligned float4 of the input row (a warp reads 512 contiguous bytes) and takes the 3 values it lacks from the
// next lane with shuffles; lane 31 only feeds lane 30, so consecutive warps overlap by one group (31 / 32 of the lanes produce
// outputs).  Output columns of a thread: jx = 4 g + padx0 + e (e < 4), whose windows start inside its own group g.
// One warp = one work item (plane, strip of STRIP output rows, chunk of 31 groups); rows walk output-stationary like the ring.
template <int FH, int FW, int STRIP>
__global__ void __launch_bounds__(256) upfirdn2d_nchw_shfl_kernel(UpfirdnParams p, int gmin, int chunks, int strips) {
    static_assert(FW <= 4, "a window spans the thread's own group and the next one");
    float wt[FH][FW];
#pragma unroll
    for (int a = 0; a < FH; ++a)
#pragma unroll
        for (int b = 0; b < FW; ++b)
            wt[a][b] = __ldg(p.f + (p.flip ? a : FH - 1 - a) * FW + (p.flip ? b : FW - 1 - b)) * p.gain;
    bool sep = wt[0][0] != 0.f && p.sepok;
#pragma unroll
    for (int a = 0; a < FH; ++a)
#pragma unroll
        for (int b = 0; b < FW; ++b) {
            const float lhs = wt[a][b] * wt[0][0], rhs = wt[a][0] * wt[0][b];
            sep = sep && fabsf(lhs - rhs) <= 2.4e-7f * fabsf(rhs);
        }
    float fx[FW], fy[FH];
#pragma unroll
    for (int b = 0; b < FW; ++b) fx[b] = sep ? wt[0][b] / wt[0][0] : 0.f;
#pragma unroll
    for (int a = 0; a < FH; ++a) fy[a] = wt[a][0];

    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nitems = (long long)p.n * p.c * strips * chunks;
    if (item >= nitems) return;
    const int chunk = (int)(item % chunks);
    const int strip = (int)((item / chunks) % strips);
    const long long plane = item / ((long long)chunks * strips);
    const int g = gmin + chunk * 31 + lane;                 // aligned input group of this lane: columns 4g .. 4g+3
    const bool gok = g >= 0 && 4 * g < p.in_w;              // in_w % 4 == 0: a group is inside the row or outside, never astride
    const int jx0 = 4 * g + p.padx0;                        // first output column of this lane
    const int jy0 = strip * STRIP, iy0 = jy0 - p.pady0;
    const float* xp = (const float*)p.x + plane * p.in_h * p.in_w + 4 * g;
    float* yp = (float*)p.y + plane * p.out_h * p.out_w;
    constexpr int NROWS = STRIP - 1 + FH;
    float acc[FH][4];
#pragma unroll
    for (int i = 0; i < FH; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
    auto walk = [&](auto separable) {
        constexpr bool SEP = decltype(separable)::value;
#pragma unroll
        for (int k = 0; k < NROWS; ++k) {
            const int iy = iy0 + k;
            float4 v = f4zero();
            if (gok && (unsigned)iy < (unsigned)p.in_h) v = ldg4(xp + (size_t)iy * p.in_w);
            float w[7];
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            w[4] = __shfl_down_sync(0xffffffffu, v.x, 1); w[5] = __shfl_down_sync(0xffffffffu, v.y, 1); w[6] = __shfl_down_sync(0xffffffffu, v.z, 1);
            float h[4];
            if (SEP) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    h[e] = 0.f;
#pragma unroll
                    for (int b = 0; b < FW; ++b) h[e] = fmaf(fx[b], w[e + b], h[e]);
                }
            }
#pragma unroll
            for (int a = 0; a < FH; ++a) {
                if (k - a >= 0 && k - a < STRIP) {
                    const int s = k - a;                        // input row k is tap row a of output row s
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (SEP) acc[s % FH][e] = fmaf(fy[a], h[e], acc[s % FH][e]);
                        else {
#pragma unroll
                            for (int b = 0; b < FW; ++b) acc[s % FH][e] = fmaf(wt[a][b], w[e + b], acc[s % FH][e]);
                        }
                    }
                    if (a == FH - 1) {                          // last tap row of output row s: store, recycle the accumulator
                        if (lane < 31 && jy0 + s < p.out_h) {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if ((unsigned)(jx0 + e) < (unsigned)p.out_w) __stcs(yp + (size_t)(jy0 + s) * p.out_w + jx0 + e, acc[s % FH][e]);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[s % FH][e] = 0.f;
                    }
                }
            }
        }
    };
    if (sep) walk(BoolTag<true>{}); else walk(BoolTag<false>{});
}

template <int FH, int FW>
static int launch_nchw_shfl(const UpfirdnParams& p_in, cudaStream_t st) {
    UpfirdnParams p = p_in;
    static int sepok = -1;
    if (sepok < 0) { const char* e = getenv("SG2_UPF_SEP"); sepok = e ? atoi(e) != 0 : 1; }
    p.sepok = sepok;
    constexpr int STRIP = 32;
    // groups whose windows reach an output column: jx = 4g + padx0 + e in [0, out_w)  ->  g in [gmin, gmax]
    auto fdiv = [](int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; };
    const int gmin = -fdiv(p.padx0 + 3, 4), gmax = fdiv(p.out_w - 1 - p.padx0, 4);
    const int chunks = (int)ceil_div(gmax - gmin + 1, 31), strips = (int)ceil_div(p.out_h, STRIP);
    const long long items = (long long)p.n * p.c * strips * chunks;
    upfirdn2d_nchw_shfl_kernel<FH, FW, STRIP><<<(unsigned)ceil_div(items, 8), 256, 0, st>>>(p, gmin, chunks, strips);
    return launched("upfirdn2d_nchw_shfl");
}

// returns SG2_ENOTSUP when the call is not one of the ring shapes
template <class V>
static int try_ring(const UpfirdnParams& p, cudaStream_t st) {
    if (p.upx != 1 || p.upy != 1 || p.downx != p.downy || p.fh != p.fw) return SG2_ENOTSUP;
    if constexpr (sizeof(V) == 4) {
        // NCHW filter form with 16-byte aligned rows: the shuffle kernel (SG2_UPF_NCHW_SHFL=0 keeps the ring, for the A/B)
        static int shfl = -1;
        if (shfl < 0) { const char* e = getenv("SG2_UPF_NCHW_SHFL"); shfl = e ? atoi(e) != 0 : 1; }
        // ... when its lanes are reasonably full: a row of G groups takes ceil(G / 31) warps (U4: 65 groups -> 3 warps, 70 % of the lanes)
        const int groups = (p.out_w - 1 - p.padx0) / 4 + (p.padx0 + 3) / 4 + 1;
        const bool full = groups * 100 >= 65 * 31 * (int)ceil_div(groups, 31);
        if (shfl && full && p.downx == 1 && (p.in_w % 4) == 0 && (((uintptr_t)p.x) & 15) == 0 && p.padx0 >= 0) {
            if (p.fh == 4) return launch_nchw_shfl<4, 4>(p, st);
            if (p.fh == 3) return launch_nchw_shfl<3, 3>(p, st);
        }
    }
    if (p.fh == 4 && p.downx == 1) return launch_ring<V, 4, 4, 1>(p, st);
    if (p.fh == 4 && p.downx == 2) return launch_ring<V, 4, 4, 2>(p, st);
    if (p.fh == 3 && p.downx == 1) return launch_ring<V, 3, 3, 1>(p, st);
    if (p.fh == 2 && p.downx == 2) return launch_ring<V, 2, 2, 2>(p, st);
    return SG2_ENOTSUP;
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_upfirdn2d(const void* x, const float* f, void* y, int dtype,
                             int n, int c, int in_h, int in_w, const int64_t x_strides[4],
                             int out_h, int out_w, const int64_t y_strides[4],
                             int fh, int fw, int upx, int upy, int downx, int downy,
                             int padx0, int padx1, int pady0, int pady1,
                             int flip, float gain, sg2_stream_t stream) {
    // argument checks mirror thirdparty/stylegan3_ops/ops/upfirdn2d.cpp:13-34
    SG2_REQUIRE(x && f && y, "upfirdn2d: null pointer");
    SG2_REQUIRE(n > 0 && c > 0 && in_h > 0 && in_w > 0, "upfirdn2d: x has zero size");
    SG2_REQUIRE(fh >= 1 && fw >= 1, "upfirdn2d: f must be at least 1x1");
    SG2_REQUIRE(upx >= 1 && upy >= 1, "upfirdn2d: upsampling factor must be at least 1");
    SG2_REQUIRE(downx >= 1 && downy >= 1, "upfirdn2d: downsampling factor must be at least 1");
    SG2_REQUIRE(dtype == SG2_F32 || dtype == SG2_F16 || dtype == SG2_F64, "upfirdn2d: unsupported dtype %d", dtype);
    const int ow = (in_w * upx + padx0 + padx1 - fw + downx) / downx;
    const int oh = (in_h * upy + pady0 + pady1 - fh + downy) / downy;
    SG2_REQUIRE(ow >= 1 && oh >= 1, "upfirdn2d: output must be at least 1x1");
    SG2_REQUIRE(ow == out_w && oh == out_h, "upfirdn2d: output size mismatch (expected %dx%d, got %dx%d)", oh, ow, out_h, out_w);
    const long long total = (long long)n * c * out_h * out_w;
    SG2_REQUIRE(total <= 2147483647LL && (long long)n * c * in_h * in_w <= 2147483647LL, "upfirdn2d: tensor is too large");

    UpfirdnParams p;
    p.x = x; p.f = f; p.y = y;
    p.n = n; p.c = c; p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
    for (int i = 0; i < 4; ++i) { p.xs[i] = x_strides[i]; p.ys[i] = y_strides[i]; }
    p.fh = fh; p.fw = fw; p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy;
    p.padx0 = padx0; p.pady0 = pady0; p.flip = flip ? 1 : 0; p.gain = gain; p.total = total;
    // memory order of y: sort dims by decreasing stride (size-1 dims go first, they do not matter)
    int ord[4] = {0, 1, 2, 3};
    const int size[4] = {n, c, out_h, out_w};
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j) {
            long long si = size[ord[i]] == 1 ? (1LL << 62) : p.ys[ord[i]];
            long long sj = size[ord[j]] == 1 ? (1LL << 62) : p.ys[ord[j]];
            if (sj > si) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
        }
    for (int i = 0; i < 4; ++i) p.order[i] = ord[i];

    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    const bool nhwc_dense = p.xs[1] == 1 && p.ys[1] == 1 && (c % 4) == 0 &&
                            p.xs[3] == c && p.ys[3] == c && p.xs[2] == (long long)in_w * c &&
                            p.ys[2] == (long long)out_w * c && (p.xs[0] % 4) == 0 && (p.ys[0] % 4) == 0 &&
                            ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0;
    const int cq = c / 4;
    const bool nchw_dense = p.xs[3] == 1 && p.ys[3] == 1 && p.xs[2] == in_w && p.ys[2] == out_w &&
                            p.xs[1] == (long long)in_h * in_w && p.ys[1] == (long long)out_h * out_w &&
                            p.xs[0] == (long long)c * in_h * in_w && p.ys[0] == (long long)c * out_h * out_w;
    if (dtype == SG2_F32 && nhwc_dense && cq <= 256 && (256 % cq) == 0 && n <= 65535) {
        const int rc = try_ring<float4>(p, st);
        if (rc != SG2_ENOTSUP) return rc;
    }
    if (dtype == SG2_F32 && nchw_dense && (long long)n * c <= 65535) {
        const int rc = try_ring<float>(p, st);
        if (rc != SG2_ENOTSUP) return rc;
    }
    if (dtype == SG2_F32 && nhwc_dense && fh * fw <= kMaxTaps && cq <= 256 && (256 % cq) == 0 && n <= 65535) {
        const int xs = 256 / cq;
        dim3 grid((unsigned)ceil_div(out_w, xs), (unsigned)ceil_div(out_h, kUpfStrip), (unsigned)n);
        upfirdn2d_strip_nhwc<<<grid, threads, 0, st>>>(p);
        return launched("upfirdn2d_strip_nhwc");
    }
    if (dtype == SG2_F32 && nhwc_dense && fh * fw <= kMaxTaps) {
        long long work = total / 4;
        int blocks = (int)std::min<long long>(ceil_div(work, threads), (long long)num_sms() * 32);
        upfirdn2d_vec4_nhwc<<<blocks, threads, 0, st>>>(p);
        return launched("upfirdn2d_vec4_nhwc");
    }
    int blocks = (int)std::min<long long>(ceil_div(total, threads), (long long)num_sms() * 32);
    if (dtype == SG2_F32)      upfirdn2d_generic<float><<<blocks, threads, 0, st>>>(p);
    else if (dtype == SG2_F16) upfirdn2d_generic<__half><<<blocks, threads, 0, st>>>(p);
    else                       upfirdn2d_generic<double><<<blocks, threads, 0, st>>>(p);
    return launched("upfirdn2d_generic");
}
