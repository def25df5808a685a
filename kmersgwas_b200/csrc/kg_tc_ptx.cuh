// kg_tc_ptx.cuh -- thin inline-PTX wrappers for the sm_100a features the tensor-core engines use:
// mbarrier, 1-D bulk async copy (TMA engine, UBLKCP), tcgen05 (TMEM alloc, UMMA, commit, TMEM loads)
// and the shared-memory matrix / instruction descriptors of tcgen05.mma.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t kg_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (converged) warp: the same lane every time, so that tcgen05.commit sees the MMAs it issued
__device__ __forceinline__ bool kg_elect_one() {
	uint32_t pred;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "elect.sync _|p, 0xffffffff;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(pred));
	return pred != 0;
}

// byte permute (PRMT): result byte n = byte (sel nibble n & 7) of {b, a}; nibble bit 3 replicates that byte's sign
__device__ __forceinline__ uint32_t kg_prmt(uint32_t a, uint32_t b, uint32_t sel) {
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
	return d;
}
// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void kg_mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kg_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kg_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void kg_mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(kg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void kg_mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kg_mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "KG_WAIT_%=:\n\t"
#ifdef KG_MBAR_TEST_WAIT
	    "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#else
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#endif
	    "@p bra KG_DONE_%=;\n\t"
	    "bra KG_WAIT_%=;\n\t"
	    "KG_DONE_%=:\n\t}" ::"r"(kg_smem_u32(bar)), "r"(parity) : "memory");
}

// for waiters that are far from the critical path: poll with a back-off so that they do not steal issue slots
__device__ __forceinline__ void kg_mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
	uint32_t done = 0;
	for (;;) {
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(done) : "r"(kg_smem_u32(bar)), "r"(parity) : "memory");
		if (done) return;
		__nanosleep(200);
	}
}

// ---- bulk async copy global -> shared (1-D TMA), completion on an mbarrier -----------------------
__device__ __forceinline__ void kg_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 kg_smem_u32(smem_dst)),
	             "l"(gmem_src), "r"(bytes), "r"(kg_smem_u32(bar))
	             : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void kg_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05: TMEM allocation ---------------------------------------------------------------------
__device__ __forceinline__ void kg_tmem_alloc(uint32_t *smem_result, uint32_t n_cols) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(kg_smem_u32(smem_result)), "r"(n_cols)
	             : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void kg_tmem_dealloc(uint32_t taddr, uint32_t n_cols) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(n_cols) : "memory");
}
__device__ __forceinline__ void kg_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void kg_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tcgen05.mma descriptors ----------------------------------------------------------------------
// Shared-memory matrix descriptor, no swizzle ("interleave"), both majors:
//   bits [0,14) start address >> 4, [16,30) leading-dimension byte offset >> 4, [32,46) stride-dimension
//   byte offset >> 4, [46,48) version = 1 (Blackwell), [61,64) layout type = 0 (SWIZZLE_NONE).
// K-major operand (element (mn, k) of 1 byte): 16 B of K are contiguous; 8 MN rows are 16 B apart (one
//   128-byte core matrix); core matrices step by LBO along K and by SBO along MN.
// MN-major operand: 16 B of MN are contiguous; 8 K rows are 16 B apart; core matrices step by SBO along MN
//   ... and by LBO along K (see kg_kinship_tc.cuh for the layout it builds).
__device__ __forceinline__ uint64_t kg_umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
	uint64_t d = 0;
	d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
	d |= 1ull << 46;
	return d;
}
// Instruction descriptor for kind::i8: D = S32 (c_format 2 at bits [4,6)), A/B format at [7,10)/[10,13)
// (0 = u8, 1 = s8), a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N >> 3 at [17,23),
// M >> 4 at [24,29).
__host__ __device__ __forceinline__ uint32_t kg_umma_idesc_i8(uint32_t m, uint32_t n, bool a_signed, bool b_signed,
                                                             bool a_mn_major, bool b_mn_major) {
	return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((a_mn_major ? 1u : 0u) << 15) |
	       ((b_mn_major ? 1u : 0u) << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void kg_umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
// Same with the A operand in tensor memory (lane = row, 4 one-byte K elements per 32-bit column).
__device__ __forceinline__ void kg_umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void kg_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
	asm volatile(
	    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
	    "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
	    "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
	    : "memory");
}
__device__ __forceinline__ void kg_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mbarrier arrive once all tcgen05.mma issued so far by this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void kg_umma_commit(uint64_t *bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(kg_smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns -------------------------
__device__ __forceinline__ void kg_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
	      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
	    : "r"(taddr)
	    : "memory");
}
__device__ __forceinline__ void kg_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
