// Hardware experiment (sm_100a): do tcgen05.mma shared-memory descriptors accept
//   (1) a K-major SWIZZLE_128B operand whose start address is shifted by an arbitrary number of 128-byte rows and whose
//       8-row groups are an arbitrary number of rows apart (SBO not a multiple of 1024 B), and
//   (2) an MN-major SWIZZLE_128B operand whose start is shifted by K rows, with overlapping 64-wide MN blocks (LBO = 128 B)?
// If the swizzle is a function of the absolute shared-memory address (tiles written by TMA at 1024-aligned bases), both
// give exact results and the convolution kernels can feed TMA-written bf16 tiles to the MMA without a re-layout pass:
// taps become descriptor offsets.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o exp_umma_shift exp_umma_shift.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include "../animeface_b200/csrc/tc_common.cuh"

namespace sg2 { thread_local char g_err[512] = ""; std::atomic<long long> g_launches{0}; long long* g_trace = nullptr; }
using namespace sg2::tc;

constexpr int ROWS = 320;                 // rows of 128 B in the activation buffer
__host__ __device__ inline float pval(int r, int c) { return (float)(((r * 7 + c * 3) % 13) - 6); }
__host__ __device__ inline float wval(int n, int k) { return (float)(((n * 5 + k * 11) % 7) - 3); }
__host__ __device__ inline float gval(int k, int m) { return (float)(((k * 3 + m * 5) % 11) - 5); }

struct Variant { int mode, r0, sbo_rows, base_off, lbo_bytes, n; };

__device__ __forceinline__ uint64_t kdesc(uint32_t addr, uint32_t sbo, uint32_t boff) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(boff & 7) << 49) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t mndesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t boff) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(boff & 7) << 49) | ((uint64_t)2 << 61);
}

// out: [128][256] fp32
__global__ void __launch_bounds__(128, 1) exp_kernel(Variant v, float* out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t P = base;                               // ROWS x 128 B, swizzled by absolute address
    const uint32_t Wt = base + ROWS * 128;                 // second operand (40960 = 40 * 1024: aligned)
    const uint32_t bar = Wt + 256 * 128 + 64 * 128;
    const uint32_t tslot = bar + 16;
    const int tid = threadIdx.x, warp = tid >> 5;
    // activation rows: P[r][c], chunk (c/8) ^ (r&7)
    for (int i = tid; i < ROWS * 64; i += 128) {
        const int r = i / 64, c = i % 64;
        const uint32_t off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
        if (v.mode == 2) *reinterpret_cast<__half*>(gen + off) = __float2half(pval(r, c));
        else *reinterpret_cast<__nv_bfloat16*>(gen + off) = __float2bfloat16(pval(r, c));
    }
    if (v.mode == 0 || v.mode == 2) {
        // B: K-major SW128 weights [64 n rows][64 k]
        for (int i = tid; i < 64 * 64; i += 128) {
            const int n = i / 64, k = i % 64;
            const uint32_t off = ROWS * 128 + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(gen + off) = __float2bfloat16(wval(n, k));
        }
    } else {
        // A: MN-major SW128 gy tile: rows = k (16 pixels), 2 MN blocks of 64 (block stride 16 rows * 128 = 2048... use 4096)
        for (int i = tid; i < 16 * 128; i += 128) {
            const int k = i / 128, m = i % 128;
            const int blk = m >> 6, c = m & 63;
            const uint32_t off = ROWS * 128 + blk * 4096 + k * 128 + (((c >> 3) ^ (k & 7)) << 4) + (c & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(gen + off) = __float2bfloat16(gval(k, m));
        }
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tslot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_d;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tslot));
    if (warp == 0 && elect_one()) {
        if (v.mode == 0 || v.mode == 2) {
            // mode 2: A = fp16, B = bf16 in one kind::f16 instruction (a_format 0, b_format 1)
            const uint32_t idesc = v.mode == 2 ? (idesc_f16(128, 64) | (1u << 10)) : idesc_bf16(128, 64);
            const uint32_t a0 = P + v.r0 * 128;
            for (int kq = 0; kq < 4; ++kq)
                mma_bf16(tmem_d, kdesc(a0 + kq * 32, v.sbo_rows * 128, v.base_off ? ((a0 >> 7) & 7) : 0), kmajor_desc(Wt + kq * 32), idesc, kq != 0);
        } else {
            const uint32_t idesc = idesc_bf16_mn(128, v.n);
            const uint32_t b0 = P + v.r0 * 128;
            for (int kq = 0; kq < 1; ++kq)
                mma_bf16(tmem_d, mnmajor_desc(Wt, 4096, 1024), mndesc(b0, v.lbo_bytes, 1024, v.base_off ? ((b0 >> 7) & 7) : 0), idesc, 0);
        }
        mma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c = 0; c < 256 / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + c * 16, r);
        for (int j = 0; j < 16; ++j) out[(size_t)tid * 256 + c * 16 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_d, 256); }
}

int main() {
    const int smem = 1024 + ROWS * 128 + 256 * 128 + 64 * 128 + 256;
    cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* d_out;
    cudaMalloc(&d_out, 128 * 256 * 4);
    std::vector<float> h(128 * 256);
    std::vector<Variant> vs;
    for (int boff = 0; boff < 2; ++boff)
        for (int sbo : {8, 10, 18})
            for (int r0 : {0, 1, 3, 8, 11}) vs.push_back({0, r0, sbo, boff, 0, 64});
    for (int boff = 0; boff < 2; ++boff) {
        for (int r0 : {0, 1, 2, 9}) vs.push_back({1, r0, 8, boff, 4096, 64});      // one shifted 64-wide block
        for (int r0 : {0, 1, 5}) vs.push_back({1, r0, 8, boff, 128, 192});          // three overlapping blocks, LBO = one row
        for (int r0 : {0, 3}) vs.push_back({1, r0, 8, boff, 34 * 128, 192});        // blocks 34 rows apart (kernel rows)
    }
    vs.push_back({2, 0, 8, 0, 0, 64});
    vs.push_back({2, 3, 10, 0, 0, 64});
    for (const Variant& v : vs) {
        cudaMemset(d_out, 0xff, 128 * 256 * 4);
        exp_kernel<<<1, 128, smem>>>(v, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d r0 %d sbo %d boff %d lbo %d: CUDA error %s\n", v.mode, v.r0, v.sbo_rows, v.base_off, v.lbo_bytes, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d_out, 128 * 256 * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < v.n; ++n) {
                double ref = 0;
                if (v.mode == 0 || v.mode == 2) {
                    const int row = v.r0 + (m / 8) * v.sbo_rows + (m % 8);
                    for (int k = 0; k < 64; ++k) ref += (double)pval(row, k) * wval(n, k);
                } else {
                    const int j = n / 64, c = n % 64;
                    const int shift = j * (v.lbo_bytes / 128);
                    for (int k = 0; k < 16; ++k) ref += (double)gval(k, m) * pval(v.r0 + shift + k, c);
                }
                const double err = fabs(ref - h[(size_t)m * 256 + n]);
                if (err > maxerr) maxerr = err;
                if (err > 1e-3) ++bad;
            }
        printf("mode %s r0 %2d sbo_rows %2d base_off %d lbo %5d N %3d : max err %.3g  bad %d/%d %s\n", v.mode == 1 ? "MN" : (v.mode == 2 ? "Kx" : "K "), v.r0, v.sbo_rows,
               v.base_off, v.lbo_bytes, v.n, maxerr, bad, 128 * v.n, bad ? "WRONG" : "exact");
    }
    return 0;
}
