// b200z_tma.cuh -- bulk asynchronous copies global -> shared memory through the TMA engine (cp.async.bulk with an mbarrier
// that counts the bytes as they land; SASS: UBLKCP + SYNCS).  One thread arms the barrier and issues the copies, the CTA
// waits on the barrier's phase; no registers are staged and the copy engine runs beside the threads' own loads.
// Sizes are multiples of 16 bytes, both addresses 16-byte aligned.  (tests/cuda_emu builds the plain loop instead.)
#pragma once
#include <stdint.h>

namespace b200z {

#if defined(B200Z_EMU)
struct BulkBarrier {
	uint64_t word;
};
__device__ inline void bulk_barrier_init(BulkBarrier *) {}
__device__ inline void bulk_copy_start(BulkBarrier *, void *dst, const void *src, uint32_t bytes, uint32_t piece = 16384) {
	(void)piece;
	memcpy(dst, src, bytes);
}
__device__ inline void bulk_expect(BulkBarrier *, uint32_t) {}
__device__ inline void bulk_wait(BulkBarrier *, uint32_t) {}
#else
struct __align__(8) BulkBarrier {
	uint64_t word;
};
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one thread, before anybody uses the barrier (followed by a __syncthreads())
__device__ __forceinline__ void bulk_barrier_init(BulkBarrier *b) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(b)) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// the issuing thread: the bytes the barrier's current phase waits for (one arrival: this one)
__device__ __forceinline__ void bulk_expect(BulkBarrier *b, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// the issuing thread: `bytes` from global to shared memory, in pieces (each a cp.async.bulk that reports to the barrier)
__device__ __forceinline__ void bulk_copy_start(BulkBarrier *b, void *dst, const void *src, uint32_t bytes, uint32_t piece = 16384) {
	const uint32_t bar = smem_u32(b);
	for (uint32_t o = 0; o < bytes; o += piece) {
		const uint32_t n = bytes - o < piece ? bytes - o : piece;
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
		                 smem_u32(reinterpret_cast<uint8_t *>(dst) + o)),
		             "l"(reinterpret_cast<const uint8_t *>(src) + o), "r"(n), "r"(bar)
		             : "memory");
	}
}
// everybody: until the phase with parity `phase` has completed (all expected bytes have landed)
__device__ __forceinline__ void bulk_wait(BulkBarrier *b, uint32_t phase) {
	const uint32_t bar = smem_u32(b);
	uint32_t done = 0;
	while (!done) {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done)
		             : "r"(bar), "r"(phase)
		             : "memory");
	}
}
#endif

} // namespace b200z
