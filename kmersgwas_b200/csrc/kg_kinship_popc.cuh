// kg_kinship_popc.cuh -- kinship Gram accumulation with AND + POPC on transposed bit columns.
// Exact integer cross-check engine (KG_OPT_KINSHIP_ENGINE = 1); the tensor-core engine is in
// kg_kinship_tc.cuh.
//
// Reference: update_emma_kinshhip_calculation (/root/reference/src/kmers_multiple_databases.cpp:418-438)
//   K[i][j] += 1 ^ g_i ^ g_j  (j < i) over kept rows.  With G = B^T B (co-presence counts),
//   c_i = G[i][i] and M kept rows:  K[i][j] = M - c_i - c_j + 2 G[i][j]   (exact integers).
#pragma once
#include "kg_common.cuh"

// keep bit per row (MAC filter of load_kmers :117-121) + kept-row count.
// One thread per row; also usable for memory-order views (mask = valid bits).
__global__ void __launch_bounds__(256) kg_prefilter_kernel(KgRowView view, const uint64_t *__restrict__ mask,
                                                          uint32_t n_used, uint32_t min_count,
                                                          uint32_t *__restrict__ keep_bits,
                                                          unsigned long long *__restrict__ kept_count) {
	const uint64_t n32 = (view.n_rows + 31) / 32 * 32;
	for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n32;
	     r += (uint64_t)gridDim.x * blockDim.x) {
		bool keep = false;
		if (r < view.n_rows) {
			const uint64_t *row = view.base + r * view.stride;
			uint32_t c = 0;
			for (uint32_t k = 0; k < view.w_in; k++) c += __popcll(__ldg(row + 1 + k) & __ldg(mask + k));
			keep = c >= min_count && c + min_count <= n_used;
		}
		const uint32_t bits = __ballot_sync(0xffffffffu, keep);
		if ((threadIdx.x & 31) == 0) {
			keep_bits[r >> 5] = bits;
			if (bits) atomicAdd(kept_count, (unsigned long long)__popc(bits));
		}
	}
}

// grid = (pair tiles of 64x64 samples over the lower triangle, row splits); block = 256.
// The view must be in MEMORY (used-sample) order: raw tile when the column map is the identity,
// squeezed tile otherwise.  Sample block I = presence word I of the row.
#define KG_KIN_CHUNK_ROWS 1024
__global__ void __launch_bounds__(256) kg_kinship_popc_kernel(KgRowView view, const uint32_t *__restrict__ keep_bits,
                                                             uint32_t n_used, uint32_t n_tiles64,
                                                             unsigned long long *__restrict__ G) {
	// col[w][s]: bits of sample s (0..63 of the block) for rows 32w..32w+31 of the chunk
	__shared__ __align__(16) uint32_t colI[KG_KIN_CHUNK_ROWS / 32][64];
	__shared__ __align__(16) uint32_t colJ[KG_KIN_CHUNK_ROWS / 32][64];

	// decode pair tile index -> (I, J), J <= I
	uint32_t I = 0, rem = blockIdx.x;
	while (rem > I) { rem -= I + 1; I++; }
	const uint32_t J = rem;
	const bool diag = (I == J);

	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
	int acc[4][4];
#pragma unroll
	for (int a = 0; a < 4; a++)
#pragma unroll
		for (int b = 0; b < 4; b++) acc[a][b] = 0;

	const uint64_t n_chunks = (view.n_rows + KG_KIN_CHUNK_ROWS - 1) / KG_KIN_CHUNK_ROWS;
	for (uint64_t ch = blockIdx.y; ch < n_chunks; ch += gridDim.y) {
		const uint64_t r0 = ch * KG_KIN_CHUNK_ROWS;
		// transpose: each warp takes row groups warp, warp+8, ...
		for (uint32_t rg = warp; rg < KG_KIN_CHUNK_ROWS / 32; rg += 8) {
			const uint64_t r = r0 + (uint64_t)rg * 32 + lane;
			uint64_t wi = 0, wj = 0;
			if (r < view.n_rows && ((keep_bits[r >> 5] >> (r & 31)) & 1u)) {
				const uint64_t *row = view.base + r * view.stride;
				if (I < view.w_in) wi = __ldg(row + 1 + I);
				if (J < view.w_in) wj = __ldg(row + 1 + J);
			}
			uint32_t mi_lo = 0, mi_hi = 0, mj_lo = 0, mj_hi = 0;
#pragma unroll
			for (int k = 0; k < 32; k++) {
				const uint32_t b0 = __ballot_sync(0xffffffffu, (wi >> k) & 1ull);
				const uint32_t b1 = __ballot_sync(0xffffffffu, (wi >> (k + 32)) & 1ull);
				if (lane == (uint32_t)k) { mi_lo = b0; mi_hi = b1; }
			}
			colI[rg][lane] = mi_lo;
			colI[rg][32 + lane] = mi_hi;
			if (!diag) {
#pragma unroll
				for (int k = 0; k < 32; k++) {
					const uint32_t b0 = __ballot_sync(0xffffffffu, (wj >> k) & 1ull);
					const uint32_t b1 = __ballot_sync(0xffffffffu, (wj >> (k + 32)) & 1ull);
					if (lane == (uint32_t)k) { mj_lo = b0; mj_hi = b1; }
				}
				colJ[rg][lane] = mj_lo;
				colJ[rg][32 + lane] = mj_hi;
			}
		}
		__syncthreads();
		const uint32_t(*cj)[64] = diag ? colI : colJ;
#pragma unroll 4
		for (uint32_t w = 0; w < KG_KIN_CHUNK_ROWS / 32; w++) {
			const uint4 a4 = *reinterpret_cast<const uint4 *>(&colI[w][ti * 4]);
			const uint4 b4 = *reinterpret_cast<const uint4 *>(&cj[w][tj * 4]);
			const uint32_t av[4] = {a4.x, a4.y, a4.z, a4.w};
			const uint32_t bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
			for (int a = 0; a < 4; a++)
#pragma unroll
				for (int b = 0; b < 4; b++) acc[a][b] += __popc(av[a] & bv[b]);
		}
		__syncthreads();
	}
#pragma unroll
	for (int a = 0; a < 4; a++)
#pragma unroll
		for (int b = 0; b < 4; b++) {
			const uint32_t i = I * 64 + ti * 4 + a, j = J * 64 + tj * 4 + b;
			if (i < n_used && j <= i && acc[a][b] != 0)
				atomicAdd(&G[(size_t)i * n_used + j], (unsigned long long)acc[a][b]);
		}
	(void)n_tiles64;
}

// ibs[i][j] = M - c_i - c_j + 2 G[i][j]  (j < i), c_i = G[i][i]
__global__ void kg_kinship_finalize_kernel(const unsigned long long *__restrict__ G, uint32_t n,
                                           unsigned long long M, unsigned long long *__restrict__ ibs) {
	const uint64_t total = (uint64_t)n * n;
	for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
	     idx += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t i = (uint32_t)(idx / n), j = (uint32_t)(idx % n);
		unsigned long long v = 0;
		if (j < i) v = M - G[(size_t)i * n + i] - G[(size_t)j * n + j] + 2ull * G[idx];
		ibs[idx] = v;
	}
}
