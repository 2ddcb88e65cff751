// tma.cuh — bulk asynchronous global -> shared copies (the 1-D form of the Blackwell/Hopper Tensor Memory Accelerator path:
// PTX `cp.async.bulk`, SASS `UBLKCP`) completing on a shared-memory mbarrier.  Used by the tile render kernels to stage
// their per-tile instance records: one elected thread issues the copies for batch i+1 while the CTA composites batch i.
//
// Rules the callers follow (PTX ISA, cp.async.bulk): source, destination and size are multiples of 16 bytes; the issuing
// thread first arms the barrier with the total byte count (`arrive.expect_tx`); every copy must have completed (its phase
// observed) before the CTA exits, because the shared memory it targets is released with the CTA.
#pragma once
#include <stdint.h>

namespace envgs {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// make the initialised barriers visible to the async proxy before the first bulk copy targets them
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// spin until the phase with the given parity has completed (try_wait suspends the thread for a hardware-chosen time slice)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(a), "r"(parity)
                     : "memory");
    } while (!ok);
}

}  // namespace envgs
