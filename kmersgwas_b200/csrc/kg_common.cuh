// kg_common.cuh -- shared device/host helpers for the sm_100a kernels of the kmersGWAS hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kmersgwas_b200.h"

#define KG_SM_COUNT_DEFAULT 148

// A "row view": how a kernel finds row r's presence words.
//   word k of row r = base[r * stride + 1 + k]  for k < w_in, zero beyond;  k-mer = base[r * stride]
// Raw .table tiles: stride = 1 + W_file, w_in = W_file (file column order).
// Squeezed tiles  : stride = 1 + W_mem,  w_in = W_mem  (memory = phenotype order).
struct KgRowView {
	const uint64_t *base;
	uint64_t n_rows;
	uint32_t stride;
	uint32_t w_in;
};

__host__ __device__ __forceinline__ uint64_t kg_mix64(uint64_t x) {
	x += 0x9e3779b97f4a7c15ull;
	x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
	x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
	return x ^ (x >> 31);
}

// 128-bit streaming load that does not allocate in L1 (tiles are read once).
__device__ __forceinline__ uint4 kg_ldg_stream(const uint4 *p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
