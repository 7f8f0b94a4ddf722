// stage.cuh -- staging the row tables into shared memory with the TMA bulk-copy engine.
//
// One elected thread posts `cp.async.bulk` copies (global -> shared, completion counted in bytes on
// an mbarrier); every thread of the block then waits on the barrier's phase.  Compared with a
// cooperative LDG/STS loop this removes ~27 load/store trips per thread from the start of every
// block and lets the copy run while the threads set up their state.  (SASS: UBLKCP + SYNCS.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace g2048 {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Call from ALL threads of the block.  `dst*` are 16-byte aligned shared addresses, `src*` 16-byte
// aligned global addresses, sizes multiples of 16 (second copy optional: bytes1 == 0).
// Returns once both tables are visible to every thread.
__device__ __forceinline__ void stage_bulk(void *dst0, const void *src0, uint32_t bytes0,
                                           void *dst1, const void *src1, uint32_t bytes1)
{
    __shared__ __align__(8) uint64_t mbar;
    const uint32_t bar = smem_addr(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes0 + bytes1) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_addr(dst0)), "l"(src0), "r"(bytes0), "r"(bar) : "memory");
        if (bytes1)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr(dst1)), "l"(src1), "r"(bytes1), "r"(bar) : "memory");
    }
    // phase 0 completes when the arrive above has happened and all bytes have landed
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar) : "memory");
}

}  // namespace g2048
