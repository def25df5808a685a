// tmem_contention.cu -- micro-benchmark (perf experiment, not product code): does tcgen05.ld traffic (the filter's
// epilogue reading accumulators back) slow the MMA stream down, and does it matter whether the MMA's A operand comes
// from tensor memory (TS) or shared memory (SS)?  One CTA per SM; one warp issues 4096 back-to-back tcgen05.mma
// kind::i8 (M 128, N 112, K 32), `n_ld_warps` x 4 other warps loop over tcgen05.ld.32x32b.x16 of 112 columns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I kmersgwas_b200/csrc -o scratch/tmem_contention profiles/probes/tmem_contention.cu
#include <cstdio>
#include "kg_tc_ptx.cuh"

__global__ void __launch_bounds__(32 * 9, 1) probe(int ss_mode, int ld_sets, int n_mma, long long *out, int writers) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t *base = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	__shared__ uint64_t bar;
	__shared__ uint32_t slot;
	__shared__ volatile int stop_flag;
	__shared__ unsigned long long ld_count;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) { kg_mbar_init(&bar, 1); kg_fence_mbar_init(); stop_flag = 0; ld_count = 0; }
	for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0x01010101u;
	if (warp == 0) kg_tmem_alloc(&slot, 512);
	kg_fence_proxy_async();
	kg_tc_fence_before();
	__syncthreads();
	kg_tc_fence_after();
	const uint32_t tmem = slot;
	if (warp == 0) {
		const uint32_t idesc = kg_umma_idesc_i8(128, 112, false, true, false, false);
		const uint64_t bd = kg_umma_smem_desc(kg_smem_u32(base), 128, 9216);
		const uint64_t ad = kg_umma_smem_desc(kg_smem_u32(base) + 147456 - 8192, 128, 256);
		long long t0 = 0, t1 = 0;
		uint32_t phase = 0;
		for (int rep = 0; rep < 2; rep++) {
			if (rep == 1 && lane == 0) stop_flag = 2;   // readers count from here
			__syncwarp();
			t0 = clock64();
			if (kg_elect_one()) {
				for (int k = 0; k < n_mma; k++) {
					const uint32_t kk = (uint32_t)k & 15u;
					if (ss_mode) kg_umma_i8(tmem, ad + (uint64_t)(kk & 1) * 256, bd + kk * 16, idesc, 1);
					else kg_umma_i8_ts(tmem, tmem + 256 + kk * 8, bd + kk * 16, idesc, 1);
				}
				kg_umma_commit(&bar);
			}
			__syncwarp();
			kg_mbar_wait(&bar, phase);
			phase ^= 1;
			t1 = clock64();
		}
		if (lane == 0) {
			stop_flag = 1;
			if (blockIdx.x == 0) { out[0] = t1 - t0; }
		}
	} else if ((int)(warp - 1) < 4 * ld_sets) {
		// readers: accumulator-sized sweeps over columns 384..495 (not touched by the MMAs) of this warp's lane quarter
		const uint32_t taddr = tmem + 384 + (((warp & 3u) * 32u) << 16);   // a warp may only touch the lane quarter warp % 4
		unsigned long long n = 0;
		uint32_t sink = 0;
		while (stop_flag != 1) {
			if (writers) {   // tcgen05.st instead: the expanders' side (7 x 16 columns per sweep, then wait::st)
				uint32_t v[16];
#pragma unroll
				for (int j = 0; j < 16; j++) v[j] = lane + j;
				for (uint32_t c0 = 0; c0 < 112; c0 += 16) kg_tmem_st16(taddr + c0, v);
				kg_tmem_st_wait();
			} else
			for (uint32_t c0 = 0; c0 < 112; c0 += 32) {
				uint32_t v[16], u[16];
				kg_tmem_ld16(taddr + c0, v);
				if (c0 + 16 < 112) kg_tmem_ld16(taddr + c0 + 16, u);
				kg_tmem_ld_wait();
				sink ^= v[0] ^ u[3];
			}
			if (stop_flag == 2) n++;
		}
		if (lane == 0) atomicAdd(&ld_count, n);
		if (sink == 0x12345678u) out[3] = sink;
	}
	kg_tc_fence_before();
	__syncthreads();
	if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (long long)ld_count;
	if (warp == 0) kg_tmem_dealloc(tmem, 512);
}

int main() {
	long long *d, h[4];
	cudaMalloc(&d, 32);
	const int smem = 162 * 1024;
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	const int n_mma = 4096;
	for (int writers = 0; writers < 2; writers++)
	for (int ss = 0; ss < 2; ss++)
		for (int sets = (writers ? 1 : 0); sets <= 2; sets++) {
			cudaMemset(d, 0, 32);
			probe<<<148, 32 * 9, smem>>>(ss, sets, n_mma, d, writers);
			cudaError_t e = cudaDeviceSynchronize();
			if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
			cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
			// one sweep of one warp = 32 lanes x 112 columns x 4 B = 14 336 B
			const double bytes_per_clk = h[0] > 0 ? (double)h[1] * 14336.0 / (double)h[0] : 0.0;
			printf("A from %s, %d %s warps: %6.1f clk/MMA, they moved %6.1f B/clk (%.0f KB per 36 MMAs)\n", ss ? "smem" : "TMEM", 4 * sets, writers ? "tcgen05.st" : "tcgen05.ld",
			       (double)h[0] / n_mma, bytes_per_clk, bytes_per_clk * ((double)h[0] / n_mma) * 36 / 1024.0);
		}
	return 0;
}
