// Shared PTX wrappers for the tcgen05 / TMA / mbarrier kernels of libsg2b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace sg2 {
namespace tc {

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (a suspend-time hint on try_wait -- the hardware parks the thread instead of polling -- was measured: the polling goes away but
// every hand-over wakes later; forward convs 3-5 % slower, step 50.3 -> 52.3 ms.  Plain polling stays.)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// mbar_wait that adds the cycles spent waiting to `acc` when tracing is on (sg2_debug_trace)
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool trace, long long& acc) {
    if (trace) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        acc += clock64() - t0;
    } else {
        mbar_wait(bar, parity);
    }
}
// one elected lane of a converged warp (elect.sync): the form ptxas recognises as "exactly one thread", so the
// tcgen05 / TMA instructions under it are emitted without a per-thread ELECT loop around each of them
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same instruction with its shared-memory descriptors given as (low, high) 32-bit words: an issuing loop that advances
// only the low words (start address) keeps its arithmetic in 32 bits
__device__ __forceinline__ void mma_f16_words(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: issue several, then ONE tmem_ld_wait(), then reg_fence() each destination array so the
// compiler cannot move a consumer of the registers above the wait
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void reg_fence(uint32_t (&r)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) asm volatile("" : "+r"(r[j]));
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row atoms of 1024 B, dense):
// start>>4 | LBO(=1, ignored for swizzled K-major)<<16 | SBO(1024>>4)<<32 | version 1<<46 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: D=f32 (bit4), A=B=bf16 (bits 7,10), both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// MN-major, SWIZZLE_128B descriptor: rows of 128 B hold 64 consecutive M/N elements of one k; 8 k-rows form a
// 1024 B atom.  LBO = byte stride between 64-element MN blocks, SBO = byte stride between 8-row k groups.
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 with both operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t idesc_bf16_mn(int m, int n) {
    return idesc_bf16(m, n) | (1u << 15) | (1u << 16);
}
// kind::tf32, K-major operands: a_format = b_format = 2
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t rna_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return u;
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 q;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(addr));
    return q;
}
__device__ __forceinline__ void sts4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// kind::f16 with fp16 operands (a_format = b_format = 0), K-major
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// fp16 big/small split of two floats: big = rn_f16(v) (saturating), small = rn_f16((v - big) * 2^11).
// big carries 11 mantissa bits, the scaled residual the next 11: together 22 bits like the tf32 pair, at the fp16/bf16
// MMA rate.  The 2^11 keeps the residual in the normal fp16 range; the consumer multiplies the cross-term sum by 2^-11.
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& big, uint32_t& small) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(big) : "f"(b), "f"(a));        // {b (high), a (low)}
    const float2 bf = __half22float2(*reinterpret_cast<const __half2*>(&big));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(small) : "f"((b - bf.y) * 2048.f), "f"((a - bf.x) * 2048.f));
}

// split two floats into packed bf16x2 hi and lo words (element 0 in the low half)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    float2 hf = __bfloat1622float2(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}


// 4-D fp32 NHWC tensor map: dims (c, w, h, n), box (32 channels, bw, bh, bb), SWIZZLE_128B, OOB -> 0.
static inline int make_nhwc_map(CUtensorMap* map, const float* x, int n, int h, int w, int c, int bw, int bh, int bb, const char* who) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(SG2_ELAUNCH, "%s: cuTensorMapEncodeTiled is not available from the driver", who);
    if (((uintptr_t)x & 15) != 0 || (c % 4) != 0) return fail(SG2_EINVAL, "%s: tensor must be 16-byte aligned with C %% 4 == 0", who);
    const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t strides[3] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, (cuuint64_t)h * w * c * 4};
    const cuuint32_t box[4] = {32u, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SG2_ELAUNCH, "%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
    return SG2_OK;
}

// 5-D bf16 map over a "pair planes" tensor [2 planes][n][h][w][c] (hi = bf16(v), lo = bf16(v - hi)): dims (c, w, h, n, plane),
// box (64 channels = one 128-byte row, bw, bh, bb, 1), SWIZZLE_128B, OOB -> 0.  A box lands in shared memory as rows of 128 B in
// (b, y, x) order with the 16-byte chunks XOR-swizzled by the row's ABSOLUTE shared-memory address bits [7:9] -- the layout a
// tcgen05 SWIZZLE_128B descriptor reads, for any row-shifted start address (scripts/exp_umma_shift.cu, profiles/r2a_umma_shift.txt).
static inline int make_planes_map(CUtensorMap* map, const void* planes, int n, int h, int w, int c, int bw, int bh, int bb, const char* who) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(SG2_ELAUNCH, "%s: cuTensorMapEncodeTiled is not available from the driver", who);
    if (((uintptr_t)planes & 15) != 0 || (c % 8) != 0) return fail(SG2_EINVAL, "%s: planes must be 16-byte aligned with C %% 8 == 0", who);
    const cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n, 2};
    const cuuint64_t strides[4] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2, (cuuint64_t)n * h * w * c * 2};
    const cuuint32_t box[5] = {64u, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb, 1u};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)planes, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SG2_ELAUNCH, "%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
    return SG2_OK;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// TMA stores from a SWIZZLE_128B shared-memory tile (rows of 128 B = 32 fp32 channels of one pixel, (b, y, x) order) into an NHWC
// tensor; elements of the box that fall outside the tensor are not written.  `_add` reduces into global memory instead.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// K-major SWIZZLE_128B descriptor with an explicit stride between 8-row groups (rows of 128 B; the start may sit on any row)
__device__ __forceinline__ uint64_t kmajor_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// Split `total` (128 or 32) pixels into a box of the [n, h, w] grid: as wide as possible first.
static inline bool pixel_box(int total, int h, int w, int& bw, int& bh, int& bb) {
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (!pow2(w) || !pow2(h) || w > 256 || h > 256 || w < 4 || h < 4) return false;
    bw = w < total ? w : total;
    bh = (total / bw) < h ? (total / bw) : h;
    bb = total / (bw * bh);
    return bw * bh * bb == total && bb <= 256;
}
// The same for kernels that accept ragged images: any image at least `total` pixels wide is cut into row pieces of
// `total` pixels; the last piece of a row hangs over the edge (TMA fills the overhang with zeros on load).
static inline bool pixel_box_ragged(int total, int h, int w, int& bw, int& bh, int& bb) {
    if (pixel_box(total, h, w, bw, bh, bb)) return true;
    if (w < total || w > 8192 || h < 1 || h > 8192) return false;
    bw = total; bh = 1; bb = 1;
    return true;
}

}  // namespace tc
}  // namespace sg2
