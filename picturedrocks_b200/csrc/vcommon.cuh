// vcommon.cuh — device helpers shared by the V-layout streaming kernels (vstream.cu, vell.cu):
// mbarrier + 1-D TMA bulk copies for the per-warp rings, and the shared-memory probe of the
// selected gene's per-cell shift code.
#pragma once
#include "common.cuh"

namespace prb {

__device__ __forceinline__ uint32_t v_smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void v_mbar_init(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void v_bulk_load(uint32_t dst, const void *src, uint32_t bytes,
                                            uint32_t bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// the same copy, its lines marked "evict first" in L2: a stream that is read exactly once must
// not push the tables the launch keeps updating (hits) out of the cache
__device__ __forceinline__ uint64_t v_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t v_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// 8-byte store whose line should stay in L2 (a table that the next kernel reads back)
__device__ __forceinline__ void v_store2_keep(uint32_t *p, uint32_t a, uint32_t b, uint64_t policy)
{
    asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(p), "r"(a), "r"(b),
                 "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void v_bulk_load_hint(uint32_t dst, const void *src, uint32_t bytes,
                                                 uint32_t bar, uint64_t policy)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void v_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

// Shared-window address of the first byte of dynamic shared memory in a kernel without
// static shared memory (sm_100: 1 KB of the window is reserved by the system).  The kernel
// traps if the assumption does not hold.
constexpr int V_SMEM_BASE = 1024;

// shift code of the cell whose index (< SPAN) is idx: the tile is the first SPAN bytes of
// dynamic shared memory, so the address is idx + an immediate (no add instruction)
__device__ __forceinline__ uint32_t v_probe(uint32_t idx)
{
    uint32_t r;
    asm volatile("ld.shared.u8 %0, [%1 + %2];" : "=r"(r) : "r"(idx), "n"(V_SMEM_BASE));
    return r;
}

// PTX shl clamps: a shift amount >= 32 gives 0 (the shift code of a cell with x_j = 0).
__device__ __forceinline__ uint32_t v_shl1(uint32_t s)
{
    uint32_t r;
    asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(s));
    return r;
}

}  // namespace prb
