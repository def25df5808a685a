// kg_probe.cuh -- live measurement of the int8 tensor-pipe peak (roofline denominator of the tcgen05 kernels).
// MEASURED_PEAKS.json carries an HBM and a bf16 figure only; the scan filter and the kinship Gram run kind::i8 MMAs, so
// bench.py asks the library to measure what back-to-back tcgen05.mma.kind::i8 (M = 128, N = 256, K = 32, A from
// tensor memory, B from shared memory; operands are garbage, one issuing warp per SM, one commit at the end) reach on
// this very GPU: profiles/probes/umma_rate.cu is the stand-alone ancestor of this kernel.
#pragma once
#include "kg_tc_ptx.cuh"

#define KG_PROBE_SMEM (40 * 1024)

__global__ void __launch_bounds__(128, 1) kg_probe_umma_i8_kernel(int n_mma) {
	extern __shared__ uint8_t kg_probe_smem[];
	uint8_t *base = reinterpret_cast<uint8_t *>(((uintptr_t)kg_probe_smem + 1023) & ~(uintptr_t)1023);
	__shared__ uint64_t bar;
	__shared__ uint32_t slot;
	const uint32_t warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) { kg_mbar_init(&bar, 1); kg_fence_mbar_init(); }
	for (uint32_t i = threadIdx.x; i < 32 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0x01010101u;
	if (warp == 0) kg_tmem_alloc(&slot, 512);
	kg_fence_proxy_async();
	kg_tc_fence_before();
	__syncthreads();
	kg_tc_fence_after();
	const uint32_t tmem = slot;
	if (warp == 0) {
		const uint32_t idesc = kg_umma_idesc_i8(128, 256, false, true, false, false);
		// B: 256 columns x 32 bytes of K, K-major core matrices: 8 columns x 16 B = 128 B, 2 along K (LBO 128), SBO 256
		const uint64_t bd = kg_umma_smem_desc(kg_smem_u32(base), 128, 256);
		if (kg_elect_one()) {
			for (int i = 0; i < n_mma; i++) kg_umma_i8_ts(tmem, tmem + 256 + (uint32_t)(i & 15) * 8, bd, idesc, 1);
			kg_umma_commit(&bar);
		}
		__syncwarp();
		kg_mbar_wait(&bar, 0);
	}
	kg_tc_fence_before();
	__syncthreads();
	if (warp == 0) kg_tmem_dealloc(tmem, 512);
}
