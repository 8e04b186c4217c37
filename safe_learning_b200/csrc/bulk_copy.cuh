// bulk_copy.cuh -- TMA bulk copies (cp.async.bulk, SASS UBLKCP) of contiguous GP tables into shared
// memory, completion tracked by an mbarrier transaction count (SASS SYNCS.ARRIVE.TRANS64 / SYNCS
// try_wait).  Used where a table slice (training inputs X, gamma = L^-T alpha, alpha) is a
// block-wide shared-memory landing: one elected thread issues the copies, nobody spends LDG/STS
// issue slots on them and the data for the NEXT chunk lands while the current one is consumed.
//
// Contract of cp.async.bulk: source, destination and size are multiples of 16 bytes.  The GP
// tables are allocated with even row counts of padding by the host (functions.py: Xs / gamma /
// alpha are zero-padded), so a slice of `rows` rows is copied as round_up(rows * row_bytes, 16).
#pragma once
#include <stdint.h>

namespace slb_bulk {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals)
                 : "memory");
}

// make the initialised barrier visible to the async proxy (the copy engine)
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// generic-proxy accesses to shared memory ordered before subsequent async-proxy accesses
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// one arrival + the number of bytes the bulk copies of this phase will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void copy_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}

// spin until the phase with the given parity has completed
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const uint32_t addr = smem_addr(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ unsigned round16(unsigned bytes) { return (bytes + 15u) & ~15u; }

}  // namespace slb_bulk
