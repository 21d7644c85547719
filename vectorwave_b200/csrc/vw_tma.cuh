// vw_tma.cuh -- device helpers shared by the tile kernels (vw_fused.cu, vw_lean.cu): mbarrier + 1-D bulk async copies
// (TMA), the boundary extension of one position, and the staging of one tile with its halo into shared memory.
#pragma once
#include <stdint.h>

#include "vw_internal.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copies (TMA)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    // try_wait with a suspend-time hint: the hardware parks the warp until the phase completes (or the hint runs out)
    // instead of handing issue slots to a polling loop while the other CTAs of the SM compute
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
            : "memory");
    } while (!done);
}
// global -> shared, completes on the mbarrier; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, bulk async-group completion
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// boundary extension of one position (used only for the few out-of-range halo samples)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t wrap_mod(int64_t i, int64_t n) { i %= n; return i < 0 ? i + n : i; }
__device__ __forceinline__ double ext_load(const double *__restrict__ row, int64_t pos, int64_t n, int mode) {
    if (pos >= 0 && pos < n) return __ldg(row + pos);
    if (mode == VW_PERIODIC) return __ldg(row + wrap_mod(pos, n));
    if (mode == VW_SYMMETRIC) { int64_t m = wrap_mod(pos, 2 * n); return __ldg(row + (m < n ? m : 2 * n - 1 - m)); }
    return 0.0;  // zero padding / linear span
}

// Stage positions [pos0, pos0+count) of `row` (length n, boundary `mode`) into dst[0..count).
// TMA path: contiguous in-range pieces as bulk copies on `bar` (thread 0), everything else by hand.
// Returns nothing; caller waits on `bar` (when use_tma) and __syncthreads().
__device__ __forceinline__ void stage_tile(double *dst, const double *__restrict__ row, int64_t pos0, int count, int64_t n,
                                           int mode, bool use_tma, uint64_t *bar, bool row_is_null) {
    const int tid = threadIdx.x;
    if (row_is_null) {
        for (int i = tid; i < count; i += (int)blockDim.x) dst[i] = 0.0;
        if (use_tma && tid == 0) mbar_expect_tx(bar, 0);
        return;
    }
    if (!use_tma) {
        for (int i = tid; i < count; i += (int)blockDim.x) dst[i] = ext_load(row, pos0 + i, n, mode);
        return;
    }
    if (mode == VW_PERIODIC) {
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)count * 8u);
            int done = 0;
            int64_t p = wrap_mod(pos0, n);
            while (done < count) {
                int64_t piece = n - p;
                if (piece > count - done) piece = count - done;
                bulk_g2s(dst + done, row + p, (uint32_t)piece * 8u, bar);
                done += (int)piece;
                p = 0;
            }
        }
        return;
    }
    // non-periodic: one in-range piece [lo, hi), the rest (zeros or mirror) by hand
    int64_t lo = pos0 < 0 ? 0 : pos0, hi = pos0 + count > n ? n : pos0 + count;
    if (hi < lo) hi = lo;
    int a = (int)(lo - pos0), b = (int)(hi - pos0);  // dst[a..b) in range
    if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)(b - a) * 8u);
        if (b > a) bulk_g2s(dst + a, row + lo, (uint32_t)(b - a) * 8u, bar);
    }
    for (int i = tid; i < a; i += (int)blockDim.x) dst[i] = ext_load(row, pos0 + i, n, mode);
    for (int i = b + tid; i < count; i += (int)blockDim.x) dst[i] = ext_load(row, pos0 + i, n, mode);
}

// L2 prefetch of [p, p + count) doubles clipped to the row [0, n): the future CTA's bulk copy then hits L2 instead of
// queueing behind the stores in HBM (16-byte units; the rows are 16-byte aligned whenever the bulk path is on)
__device__ __forceinline__ void prefetch_l2_span(const double *row, long long p, int count, long long n) {
    long long lo = p < 0 ? 0 : p, hi = p + count > n ? n : p + count;
    lo = (lo + 1) & ~1ll; hi &= ~1ll;
    if (hi > lo)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row + lo), "r"((uint32_t)((hi - lo) * 8)) : "memory");
}


}  // namespace
