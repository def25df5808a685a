// umma_rate.cu -- micro-benchmark (perf experiment, not product code): issue rate of back-to-back tcgen05.mma
// instructions, M = 128, K = 32 bytes, for UMMA N in {32..256}; A from tensor memory (TS) or shared memory (SS),
// kind::i8 vs kind::f8f6f4 (e4m3).  Operands are garbage; only the cycle count matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I kmersgwas_b200/csrc -o scratch/umma_rate profiles/probes/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "kg_tc_ptx.cuh"

__device__ __forceinline__ void umma_f8_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f8_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 = i8 TS, 1 = i8 SS, 2 = f8 TS, 3 = f8 SS;  dsplit: alternate between two accumulator buffers every `dsplit` MMAs
__global__ void __launch_bounds__(704, 1) probe(int mode, uint32_t n, int n_mma, int batch, long long *cycles, uint32_t mma_warp, int proto) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t *base = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	__shared__ uint64_t bar, stop, a_full[2], a_empty[2], turn[2], done2[2];
	__shared__ uint32_t slot;
	const uint32_t warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) { kg_mbar_init(&bar, proto == 1 ? (uint32_t)((n_mma + batch - 1) / batch) : (proto == 3 ? 2u * (uint32_t)((n_mma / 2 + batch - 1) / batch) : 1u)); kg_mbar_init(&stop, 1);
		for (int i = 0; i < 2; i++) { kg_mbar_init(&a_full[i], 1); kg_mbar_init(&a_empty[i], 1); kg_mbar_init(&turn[i], 1); kg_mbar_init(&done2[i], (uint32_t)(n_mma / batch / 2)); }
		kg_fence_mbar_init(); }
	for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0x01010101u;
	if (warp == mma_warp) kg_tmem_alloc(&slot, 512);
	kg_fence_proxy_async();
	kg_tc_fence_before();
	__syncthreads();
	kg_tc_fence_after();
	const uint32_t tmem = slot;
	if (proto == 5 && (warp == mma_warp || warp + 1 == mma_warp)) {
		const uint32_t k = mma_warp - warp;   // issuer 0 / 1
		const uint32_t idesc = kg_umma_idesc_i8(128, n, false, true, false, false);
		const uint64_t bd = kg_umma_smem_desc(kg_smem_u32(base), 128, 9216);
		const int n_batches = n_mma / batch / 2;   // per issuer and repetition
		long long t0 = 0, t1 = 0;
		uint32_t waits = 0;
		for (int rep = 0; rep < 2; rep++) {
			t0 = clock64();
			for (int i = 0; i < n_batches; i++) {
				if (k == 1 || i > 0 || rep > 0) { kg_mbar_wait(&turn[k], waits & 1); waits++; }
				if (kg_elect_one()) {
					for (int q = 0; q < batch; q++) {
						const uint32_t kk = (uint32_t)q & 15u;
						kg_umma_i8_ts(tmem, tmem + 256 + kk * 8, bd + kk * 16, idesc, (q != 0 || ((2 * i + k) & 3) != 0) ? 1u : 0u);
					}
				}
				__syncwarp();
				if ((threadIdx.x & 31) == 0) kg_mbar_arrive(&turn[k ^ 1]);
				if (kg_elect_one()) kg_umma_commit(&done2[k]);
				__syncwarp();
			}
			kg_mbar_wait(&done2[k], rep & 1);
			t1 = clock64();
		}
		if (k == 0) {
			// the other issuer is waiting for one more turn: consume nothing, just release the pollers
			if ((threadIdx.x & 31) == 0) kg_mbar_arrive(&stop);
			if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) *cycles = t1 - t0;
		}
	} else if (proto == 2 && warp == (mma_warp ? 0u : 1u)) {
		// helper = the expanders of the filter kernel without the work: wait until the stage is free, hand it back as full
		uint32_t st = 0, par = 1;
		for (int rep = 0; rep < 2; rep++)
			for (int i = 0; i < n_mma; i += batch) {
				kg_mbar_wait(&a_empty[st], par);
				kg_tc_fence_after();
				kg_tc_fence_before();
				__syncwarp();
				if ((threadIdx.x & 31) == 0) kg_mbar_arrive(&a_full[st]);
				if (++st == 2) { st = 0; par ^= 1; }
			}
	} else if (warp != mma_warp && !(proto == 3 && warp + 1 == mma_warp)) {
		kg_mbar_wait(&stop, 0);   // spinner warps: poll like the waiting roles of the filter kernel
	} else {
		// idesc: i8 -> kg_umma_idesc_i8; f8f6f4: D = F32 (c_format 1 at [4,6)), A/B format e4m3 = 0
		const uint32_t idesc = (mode < 2) ? kg_umma_idesc_i8(128, n, false, true, false, false) : ((1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24));
		const uint64_t bd = kg_umma_smem_desc(kg_smem_u32(base), 128, n <= 128 ? 9216 : 4608);   // B: K-major core matrices (all inside the 147 KB image)
		const uint64_t ad = kg_umma_smem_desc(kg_smem_u32(base) + 147456 - 8192, 128, 256);   // A (SS): 128 rows x 32 B
		const bool second = proto == 3 && warp != mma_warp;
		const uint32_t d_t = tmem + (second ? 128u : 0u);
		if (proto == 3) n_mma /= 2;
		const uint32_t a_t = tmem + 256 + (second ? 128u : 0u);                                                 // A (TS): columns 256..
		uint32_t phase = 0, st = 0, par = 0;
		long long t0 = 0, t1 = 0;
		for (int rep = 0; rep < 2; rep++) {   // rep 0 = warm-up
			t0 = clock64();
			for (int i = 0; i < n_mma; i += batch) {
				if (proto == 2) { kg_mbar_wait(&a_full[st], par); kg_tc_fence_after(); }
				if (kg_elect_one()) {
					for (int k = 0; k < batch; k++) {
						const uint32_t kk = (uint32_t)k & 15u;
						if (mode == 0) kg_umma_i8_ts(d_t, a_t + kk * 8, bd + kk * 16, idesc, 1);
						else if (mode == 1) kg_umma_i8(d_t, ad + (uint64_t)(kk & 1) * 256, bd + kk * 16, idesc, 1);
						else if (mode == 2) umma_f8_ts(d_t, a_t + kk * 8, bd + kk * 16, idesc, 1);
						else umma_f8_ss(d_t, ad + (uint64_t)(kk & 1) * 256, bd + kk * 16, idesc, 1);
					}
					kg_umma_commit(proto == 2 ? &a_empty[st] : &bar);
				}
				__syncwarp();
				if (proto == 0) { kg_mbar_wait(&bar, phase); phase ^= 1; }
				if (proto == 2 && ++st == 2) { st = 0; par ^= 1; }
			}
			if (proto == 2) { if (kg_elect_one()) kg_umma_commit(&bar); __syncwarp(); }   // drain
			if (proto != 0) { kg_mbar_wait(&bar, phase); phase ^= 1; }   // proto 3: both issuers' commits complete the phase
			t1 = clock64();
		}
		if ((threadIdx.x & 31) == 0 && !second) kg_mbar_arrive(&stop);
		if ((threadIdx.x & 31) == 0 && blockIdx.x == 0 && !second) *cycles = t1 - t0;
	}
	kg_tc_fence_before();
	__syncthreads();
	if (warp == mma_warp) kg_tmem_dealloc(tmem, 512);
}

int main() {
	long long *d_c, h_c;
	cudaMalloc(&d_c, 8);
	const int smem = 162 * 1024;
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	const int n_mma = 4096;
	for (int proto : {1, 5}) {
		for (uint32_t n : {112u}) {
			for (int batch : {4096, 16, 8, 4}) {
				if (proto == 5 && batch == 4096) continue;
				probe<<<148, 704, smem>>>(0, n, n_mma, batch, d_c, 21u, proto);
				cudaError_t e = cudaDeviceSynchronize();
				if (e != cudaSuccess) { printf("proto %d N=%u: %s\n", proto, n, cudaGetErrorString(e)); return 1; }
				cudaMemcpy(&h_c, d_c, 8, cudaMemcpyDeviceToHost);
				printf("proto %d (1 = one issuer, 5 = two issuers alternating batches of one stream, same accumulator), i8 TS N=%3u batch %4d : %7.1f clk/MMA\n", proto, n, batch, (double)h_c / n_mma);
			}
		}
	}
	return 0;
}
