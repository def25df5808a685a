// kg_synth.cuh -- device twin of the counter-based synthetic table generator documented in
// oracle/oracle.c (kgo_synth_rows).  One thread per (row, word); word 0 is the k-mer id.
#pragma once
#include "kg_common.cuh"

__global__ void kg_synth_rows_kernel(uint64_t seed, uint64_t first_row, uint64_t n_rows, uint32_t w_file,
                                     uint64_t last_mask, uint64_t *__restrict__ out) {
	const uint64_t stride = (uint64_t)w_file + 1;
	const uint64_t total = n_rows * stride;
	for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
	     idx += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t i = idx / stride;
		const uint32_t k = (uint32_t)(idx - i * stride);
		const uint64_t r = first_row + i;
		const uint64_t base_r = kg_mix64(seed ^ kg_mix64(r));
		if (k == 0) {
			out[idx] = r * 1024ull + (kg_mix64(base_r ^ 0x4B3Aull) & 1023ull);
			continue;
		}
		const uint32_t w = k - 1;
		uint64_t s = r;
		if (r > 0 && (kg_mix64(base_r ^ 0xD00Dull) & 63ull) == 0) s = r - 1;
		const uint64_t b = kg_mix64(seed ^ kg_mix64(s));
		const unsigned level = (unsigned)((kg_mix64(b ^ 0x1E7E1ull) >> 7) & 15ull);
		uint64_t q[5];
#pragma unroll
		for (int j = 0; j < 5; j++) q[j] = kg_mix64(b + 8ull * (uint64_t)w + (uint64_t)j + 1ull);
		uint64_t v;
		switch (level) {
		case 0: v = q[0] & q[1] & q[2] & q[3] & q[4]; break;
		case 1: v = q[0] & q[1] & q[2]; break;
		case 2: case 11: v = q[0] & q[1]; break;
		case 3: case 12: v = q[0] & (q[1] | q[2]); break;
		case 7: case 14: v = q[0] | (q[1] & q[2]); break;
		case 8: case 15: v = q[0] | q[1]; break;
		case 9: v = q[0] | q[1] | q[2]; break;
		case 10: v = q[0] | q[1] | q[2] | q[3] | q[4]; break;
		default: v = q[0]; break;
		}
		if (w == w_file - 1) v &= last_mask;
		out[idx] = v;
	}
}
