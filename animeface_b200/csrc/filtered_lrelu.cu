// filtered_lrelu: the alias-suppressed non-linearity of the StyleGAN3 generator as ONE kernel.
//
// Replaces thirdparty/stylegan3_ops/ops/filtered_lrelu.py:50-268 / filtered_lrelu.cu:133-1093 (called from
// implementations/StyleGAN3/model.py:186-190), whose arithmetic is that of _filtered_lrelu_ref (:121-147):
//     xb = x + b[c]
//     z  = up^2 * FIR_fu( zero-insert(xb, up), padded )                                  (the "z grid", zh x zw)
//     a  = clamp( lrelu_slope(z) * gain )
//     y  = decimate_down( FIR_fd(a) )
// Composed from bias_act / upfirdn2d this writes and re-reads the up^2-times larger z and a (three tensors of 16 B per z
// element).  Here a CTA owns an output tile of one (sample, channel) plane and keeps every intermediate in shared memory:
//   sX  input tile (+ bias, zero outside the image)        -> horizontal up-FIR (polyphase: only the taps that hit a sample)
//   sH  [input rows][z columns]                             -> vertical up-FIR, activation (or the stored sign mask)
//   sZ  [z rows][z columns]                                 -> down FIR: separable (sD = horizontal pass) or a full 2-D filter
// HBM traffic: x once (+ tile halos), y once, and one byte per z element for the sign mask when a backward will follow.
//
// The backward pass is the same kernel (mode 2): dy takes the place of x, the filters swap roles (flipped), up <-> down, and the
// activation is replaced by the stored mask -- a = z * gain * {slope, 1, 0}[mask] -- on the SAME z grid; the down stage reads
// its window at an offset (doff) so that no re-padding of the z grid is needed.
// Filters arrive oriented for correlation (the host flips them): z[u] = sum_t fu[t] * xup[u + t], y[o] = sum_s fd[s] * a[o*down + s + doff].
#include "common.cuh"

namespace sg2 {
namespace flr {

struct Params {
    const float* x; const float* b; float* y; unsigned char* mask;
    const float* fu; const float* fd;
    int fu_2d, fd_2d, planes, channels;
    int in_h, in_w, up, pad0x, pad0y, zh, zw, fu_n, fd_n, down, doffx, doffy, out_h, out_w;
    int otw, oth;                 // output tile
    int ztw, zth, itw, ith;       // z tile / input tile extents (upper bounds, fixed per launch)
    float up_gain, gain, slope, clamp;
    int mode;                     // 0 activation, 1 activation + write mask, 2 multiply by mask
};

__device__ __forceinline__ int floor_div(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
__device__ __forceinline__ int ceil_div_i(int a, int b) { return -floor_div(-a, b); }

// 2-D down filter: a thread owns one output column and 4 consecutive output rows, with a sliding register window down the z
// column -- per 4 multiply-adds one broadcast tap load and one data load instead of 8 loads (the first version of this stage
// was bound by the shared-memory pipe).  Rows past the tile are read as zero.
template <int DOWN>
__device__ __forceinline__ void down2d_block(const Params& p, const float* sZ, const float* sFd, float* yp, int tid, int tw, int th,
                                             int zh_t, int ox0, int oy0) {
    constexpr int W = 3 * DOWN + 1;
    const int nyb = (th + 3) >> 2;
    for (int i = tid; i < nyb * tw; i += 256) {
        const int yb = i / tw, ox = i - yb * tw;
        const int oy = yb * 4, zr0 = oy * DOWN;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int sx = 0; sx < p.fd_n; ++sx) {
            const float* col = sZ + (size_t)zr0 * p.ztw + ox * DOWN + sx;
            float w[W];
#pragma unroll
            for (int k = 0; k < W - 1; ++k) w[k] = (zr0 + k < zh_t) ? col[(size_t)k * p.ztw] : 0.f;
            for (int sy = 0; sy < p.fd_n; ++sy) {
                w[W - 1] = (zr0 + sy + W - 1 < zh_t) ? col[(size_t)(sy + W - 1) * p.ztw] : 0.f;
                const float tap = sFd[sy * p.fd_n + sx];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = fmaf(tap, w[j * DOWN], acc[j]);
#pragma unroll
                for (int k = 0; k < W - 1; ++k) w[k] = w[k + 1];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (oy + j < th) yp[(size_t)(oy0 + oy + j) * p.out_w + ox0 + ox] = acc[j];
    }
}

// UP: the up-sampling factor as a compile-time constant (1, 2, 4: the polyphase index arithmetic becomes shifts), 0 = run time.
// Loops are (row by warp, column by lane): no integer division per element.
template <int UP>
__global__ void __launch_bounds__(256) filtered_lrelu_kernel(const Params p) {
    const int up = UP ? UP : p.up;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    extern __shared__ float smem[];
    float* sX = smem;                                  // [ith][itw]
    float* sH = sX + p.ith * p.itw;                    // [ith][ztw]      (later reused as sD [zth][otw])
    const int sh_elems = max(p.ith * p.ztw, p.zth * p.otw);
    float* sZ = sH + sh_elems;                         // [zth][ztw]
    float* sFu = sZ + p.zth * p.ztw;                   // fu_n
    float* sFd = sFu + (p.fu_2d ? p.fu_n * p.fu_n : p.fu_n);      // fd_n or fd_n^2
    const int tid = threadIdx.x;
    const int plane = blockIdx.z;
    const int ox0 = blockIdx.x * p.otw, oy0 = blockIdx.y * p.oth;
    const int tw = min(p.otw, p.out_w - ox0), th = min(p.oth, p.out_h - oy0);
    // z tile: what the output tile reads
    const int zx0 = ox0 * p.down + p.doffx, zy0 = oy0 * p.down + p.doffy;
    const int zw_t = (tw - 1) * p.down + p.fd_n, zh_t = (th - 1) * p.down + p.fd_n;
    // input tile: the samples the z tile's up-FIR windows hit
    const int ix0 = ceil_div_i(zx0 - p.pad0x, up), iy0 = ceil_div_i(zy0 - p.pad0y, up);
    const int ix1 = floor_div(zx0 + zw_t + p.fu_n - 2 - p.pad0x, up), iy1 = floor_div(zy0 + zh_t + p.fu_n - 2 - p.pad0y, up);
    const int iw_t = max(0, ix1 - ix0 + 1), ih_t = max(0, iy1 - iy0 + 1);

    const int nfu = p.fu_2d ? p.fu_n * p.fu_n : p.fu_n;
    for (int i = tid; i < nfu; i += 256) sFu[i] = __ldg(p.fu + i);
    const int nfd = p.fd_2d ? p.fd_n * p.fd_n : p.fd_n;
    for (int i = tid; i < nfd; i += 256) sFd[i] = __ldg(p.fd + i);
    // ---- input tile (+ bias)
    const float bias = p.b ? __ldg(p.b + plane % p.channels) : 0.f;
    const float* xp = p.x + (size_t)plane * p.in_h * p.in_w;
    for (int r = warp; r < ih_t; r += 8)
        for (int c = lane; c < iw_t; c += 32) {
            const int gy = iy0 + r, gx = ix0 + c;
            float v = 0.f;
            if ((unsigned)gy < (unsigned)p.in_h && (unsigned)gx < (unsigned)p.in_w) v = __ldg(xp + (size_t)gy * p.in_w + gx) + bias;
            sX[r * p.itw + c] = v;
        }
    __syncthreads();
    const size_t mplane = (size_t)plane * p.zh * p.zw;
    // mask ownership (mode 1: doff = 0, the z tile starts at this tile's first output): up to where the next tile's z tile
    // starts; the last tile also owns the tail of the grid
    const int own_x1 = (ox0 + tw >= p.out_w) ? p.zw : (ox0 + tw) * p.down, own_y1 = (oy0 + th >= p.out_h) ? p.zh : (oy0 + th) * p.down;
    // activation (or stored mask) of one z element -> sZ; elements outside the z grid are zero
    auto finish_z = [&](int v, int u, float acc) {
        const int gzy = zy0 + v, gzx = zx0 + u;
        float a = 0.f;
        if ((unsigned)gzy < (unsigned)p.zh && (unsigned)gzx < (unsigned)p.zw) {
            const float z = acc * p.up_gain;
            if (p.mode == 2) {
                const unsigned char code = p.mask[mplane + (size_t)gzy * p.zw + gzx];
                a = z * p.gain * (code == 1 ? 1.f : (code == 0 ? p.slope : 0.f));
            } else {
                unsigned char code = z > 0.f ? 1 : 0;
                a = (z > 0.f ? z : z * p.slope) * p.gain;
                if (p.clamp >= 0.f && fabsf(a) > p.clamp) { a = a > 0.f ? p.clamp : -p.clamp; code = 2; }
                if (p.mode == 1 && gzx < own_x1 && gzy < own_y1) p.mask[mplane + (size_t)gzy * p.zw + gzx] = code;
            }
        }
        sZ[v * p.ztw + u] = a;
    };
    if (!p.fu_2d) {
        // ---- horizontal up-FIR: sH[r][u] = sum_t fu[t] * xup[zx0 + u + t], xup[q] = x[(q - pad0x) / up] when divisible
        for (int r = warp; r < ih_t; r += 8)
            for (int u = lane; u < zw_t; u += 32) {
                const int q0 = zx0 + u - p.pad0x;                               // xup coordinate (minus padding) of tap 0
                int t = ((-q0) % up + up) % up;                                 // first tap that lands on a sample
                float acc = 0.f;
                const float* row = sX + r * p.itw;
                for (int c = (q0 + t) / up - ix0; t < p.fu_n; t += up, ++c) acc = fmaf(sFu[t], row[c], acc);
                sH[r * p.ztw + u] = acc;
            }
        __syncthreads();
        // ---- vertical up-FIR + activation / mask -> sZ
        for (int v = warp; v < zh_t; v += 8) {
            const int q0 = zy0 + v - p.pad0y;
            const int t0 = ((-q0) % up + up) % up, r0 = (q0 + t0) / up - iy0;
            const bool rowok = (unsigned)(zy0 + v) < (unsigned)p.zh;
            for (int u = lane; u < zw_t; u += 32) {
                float acc = 0.f;
                if (rowok)
                    for (int t = t0, r = r0; t < p.fu_n; t += up, ++r) acc = fmaf(sFu[t], sH[r * p.ztw + u], acc);
                finish_z(v, u, acc);
            }
        }
    } else {
        // ---- 2-D up filter (the adjoint of a radial down filter), polyphase: the z elements of one residue class (v mod up, u mod up)
        // are a dense correlation of the input tile with the sub-filter fu[t0y + ky*up][t0x + kx*up].  A thread owns one column of the
        // class and 4 consecutive members along y, with a sliding register window down the input column: per 4 multiply-adds one
        // broadcast tap load and one data load (consecutive threads -> consecutive columns: no bank conflicts).
        const int ny = (zh_t + up - 1) / up, nx = (zw_t + up - 1) / up, nyb = (ny + 3) >> 2;
        const int kn = (p.fu_n + up - 1) / up;
        const int items = up * up * nyb * nx;
        for (int i = tid; i < items; i += 256) {
            const int mx = i % nx;
            int rest = i / nx;
            const int yb = rest % nyb; rest /= nyb;
            const int rx = rest % up, ry = rest / up;
            const int u = rx + mx * up, v0 = ry + (yb * 4) * up;
            if (u >= zw_t || v0 >= zh_t) continue;
            const int q0x = zx0 + u - p.pad0x, q0y = zy0 + v0 - p.pad0y;
            const int t0x = ((-q0x) % up + up) % up, t0y = ((-q0y) % up + up) % up;
            const int c0 = (q0x + t0x) / up - ix0, r0 = (q0y + t0y) / up - iy0;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int kx = 0; kx < kn; ++kx) {
                const int tx = t0x + kx * up;
                if (tx >= p.fu_n) break;
                const float* col = sX + (size_t)r0 * p.itw + c0 + kx;
                // rows r0 .. r0+2 preloaded; input rows past the tile's last are never read with a non-zero tap, but stay in bounds
                float w0 = col[0], w1 = (r0 + 1 < p.ith) ? col[p.itw] : 0.f, w2 = (r0 + 2 < p.ith) ? col[2 * p.itw] : 0.f;
                for (int ky = 0; ky < kn; ++ky) {
                    const int ty = t0y + ky * up;
                    if (ty >= p.fu_n) break;
                    const float w3 = (r0 + ky + 3 < p.ith) ? col[(ky + 3) * p.itw] : 0.f;
                    const float tap = sFu[ty * p.fu_n + tx];
                    acc[0] = fmaf(tap, w0, acc[0]); acc[1] = fmaf(tap, w1, acc[1]); acc[2] = fmaf(tap, w2, acc[2]); acc[3] = fmaf(tap, w3, acc[3]);
                    w0 = w1; w1 = w2; w2 = w3;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (v0 + j * up < zh_t) finish_z(v0 + j * up, u, acc[j]);
        }
    }
    __syncthreads();
    // ---- down FIR + decimation
    float* yp = p.y + (size_t)plane * p.out_h * p.out_w;
    if (p.fd_2d) {
        // a thread owns one output column and 4 consecutive output rows, with a sliding register window down the z column:
        // per 4 multiply-adds one broadcast tap load and `down` data loads instead of 8 loads
        if (p.down == 1) down2d_block<1>(p, sZ, sFd, yp, tid, tw, th, zh_t, ox0, oy0);
        else if (p.down == 2) down2d_block<2>(p, sZ, sFd, yp, tid, tw, th, zh_t, ox0, oy0);
        else if (p.down == 4) down2d_block<4>(p, sZ, sFd, yp, tid, tw, th, zh_t, ox0, oy0);
        else {
            for (int i = tid; i < th * tw; i += 256) {
                const int oy = i / tw, ox = i - oy * tw;
                const float* zr = sZ + (oy * p.down) * p.ztw + ox * p.down;
                float acc = 0.f;
                for (int sy = 0; sy < p.fd_n; ++sy)
                    for (int sx = 0; sx < p.fd_n; ++sx) acc = fmaf(sFd[sy * p.fd_n + sx], zr[sy * p.ztw + sx], acc);
                yp[(size_t)(oy0 + oy) * p.out_w + ox0 + ox] = acc;
            }
        }
    } else {
        float* sD = sH;                                                      // [zh_t][otw]
        for (int v = warp; v < zh_t; v += 8)
            for (int ox = lane; ox < tw; ox += 32) {
                const float* z = sZ + v * p.ztw + ox * p.down;
                float acc = 0.f;
                for (int sx = 0; sx < p.fd_n; ++sx) acc = fmaf(sFd[sx], z[sx], acc);
                sD[v * p.otw + ox] = acc;
            }
        __syncthreads();
        for (int oy = warp; oy < th; oy += 8)
            for (int ox = lane; ox < tw; ox += 32) {
                float acc = 0.f;
                for (int sy = 0; sy < p.fd_n; ++sy) acc = fmaf(sFd[sy], sD[(oy * p.down + sy) * p.otw + ox], acc);
                yp[(size_t)(oy0 + oy) * p.out_w + ox0 + ox] = acc;
            }
    }
}

}  // namespace flr
}  // namespace sg2

using namespace sg2;

// x [planes = N*C][in_h][in_w], y [planes][out_h][out_w] dense NCHW fp32; mask [planes][zh][zw] bytes (mode 1 writes, mode 2 reads).
extern "C" int sg2_filtered_lrelu(const float* x, const float* b, float* y, void* mask, const float* fu, const float* fd, int fu_2d, int fd_2d,
                                  int planes, int channels, int in_h, int in_w, int up, int pad0x, int pad0y, int zh, int zw,
                                  int fu_n, int fd_n, int down, int doffx, int doffy, int out_h, int out_w,
                                  float up_gain, float gain, float slope, float clamp, int mode, sg2_stream_t stream) {
    SG2_REQUIRE(x && y && fu && fd, "filtered_lrelu: null pointer");
    SG2_REQUIRE(planes > 0 && channels > 0 && planes % channels == 0, "filtered_lrelu: planes must be a multiple of channels");
    SG2_REQUIRE(in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && zh > 0 && zw > 0, "filtered_lrelu: empty tensor");
    SG2_REQUIRE(up >= 1 && down >= 1 && fu_n >= 1 && fd_n >= 1 && fu_n <= (fu_2d ? 32 : 64) && fd_n <= 32,
                "filtered_lrelu: up/down >= 1, fu <= 64 taps (32 x 32 when 2-D), fd <= 32 taps");
    SG2_REQUIRE(mode >= 0 && mode <= 2 && (mode == 0 || mask), "filtered_lrelu: mode 1/2 need the mask buffer");
    SG2_REQUIRE(planes <= 65535 * 1, "filtered_lrelu: too many planes for one launch (%d)", planes);
    flr::Params p;
    p.x = x; p.b = b; p.y = y; p.mask = (unsigned char*)mask; p.fu = fu; p.fd = fd;
    p.fu_2d = fu_2d; p.fd_2d = fd_2d; p.planes = planes; p.channels = channels;
    p.in_h = in_h; p.in_w = in_w; p.up = up; p.pad0x = pad0x; p.pad0y = pad0y; p.zh = zh; p.zw = zw; p.fu_n = fu_n; p.fd_n = fd_n;
    p.down = down; p.doffx = doffx; p.doffy = doffy; p.out_h = out_h; p.out_w = out_w;
    p.up_gain = up_gain; p.gain = gain; p.slope = slope; p.clamp = clamp; p.mode = mode;
    // output tile: the largest of 32x32, 32x16, 16x16, 16x8, 8x8 whose shared-memory footprint fits 96 KB
    static const int tiles[5][2] = {{32, 32}, {32, 16}, {16, 16}, {16, 8}, {8, 8}};
    size_t bytes = 0;
    bool ok = false;
    for (int i = 0; i < 5 && !ok; ++i) {
        p.otw = std::min(tiles[i][0], out_w); p.oth = std::min(tiles[i][1], out_h);
        p.ztw = (p.otw - 1) * down + fd_n; p.zth = (p.oth - 1) * down + fd_n;
        p.itw = (p.ztw + fu_n - 2) / up + 3; p.ith = (p.zth + fu_n - 2) / up + 3;
        const size_t fl = (size_t)p.ith * p.itw + std::max((size_t)p.ith * p.ztw, (size_t)p.zth * p.otw) + (size_t)p.zth * p.ztw +
                          (size_t)(fu_2d ? fu_n * fu_n : fu_n) + (size_t)(fd_2d ? fd_n * fd_n : fd_n);
        bytes = fl * sizeof(float);
        ok = bytes <= 96 * 1024;
    }
    if (!ok) return fail(SG2_ENOTSUP, "filtered_lrelu: no tile fits shared memory (up=%d down=%d fu=%d fd=%d)", up, down, fu_n, fd_n);
    static size_t configured = 0;
    if (bytes > 48 * 1024 && bytes > configured) {
        cudaError_t e = cudaSuccess;
        for (auto k : {flr::filtered_lrelu_kernel<0>, flr::filtered_lrelu_kernel<1>, flr::filtered_lrelu_kernel<2>, flr::filtered_lrelu_kernel<4>})
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) return fail(SG2_ELAUNCH, "filtered_lrelu: cannot opt in to 96 KB of shared memory: %s", cudaGetErrorString(e));
        configured = 96 * 1024;
    }
    dim3 grid((unsigned)ceil_div(out_w, p.otw), (unsigned)ceil_div(out_h, p.oth), (unsigned)planes);
    cudaStream_t st = (cudaStream_t)stream;
    if (up == 1) flr::filtered_lrelu_kernel<1><<<grid, 256, bytes, st>>>(p);
    else if (up == 2) flr::filtered_lrelu_kernel<2><<<grid, 256, bytes, st>>>(p);
    else if (up == 4) flr::filtered_lrelu_kernel<4><<<grid, 256, bytes, st>>>(p);
    else flr::filtered_lrelu_kernel<0><<<grid, 256, bytes, st>>>(p);
    return launched("filtered_lrelu");
}
