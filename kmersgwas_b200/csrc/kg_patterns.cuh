// kg_patterns.cuh -- distinct presence/absence patterns on the device (--pattern_counter).
//
// Reference: MultipleKmersDataBases::update_presence_absence_pattern_counter
// (/root/reference/src/kmers_multiple_databases.cpp:367-380): every row load_kmers kept is hashed -- over its
// W_mem memory-order words, seed ^= Hash64(word) + 0x9e3779b97f4a7c15 + (seed << 6) + (seed >> 2)
// (src/kmer_general.h:32-41 for Hash64) -- and the hash goes into a set; the CLI reports the set's size.  Here the set
// is an open-addressing table of the same 64-bit hashes in HBM (linear probing, atomicCAS), so the count equals the
// reference's even where two patterns collide in the hash.
#pragma once
#include "kg_common.cuh"

#define KG_PAT_EMPTY 0xFFFFFFFFFFFFFFFFull

__host__ __device__ __forceinline__ uint64_t kg_hash64(uint64_t key) {
	key = (key ^ (key >> 33)) * 0xff51afd7ed558ccdull;
	key = (key ^ (key >> 33)) * 0xc4ceb9fe1a85ec53ull;
	return key ^ (key >> 33);
}

// insert `key`; returns true if it was not in the set.  slots is a power of two; the caller keeps the load below 1/2.
__device__ __forceinline__ bool kg_pat_insert(unsigned long long *table, uint64_t slots, uint64_t key, unsigned long long *has_empty_key) {
	if (key == KG_PAT_EMPTY) return atomicExch(has_empty_key, 1ull) == 0ull;   // the one value the table cannot hold
	uint64_t at = kg_mix64(key) & (slots - 1);
	for (;;) {
		const unsigned long long seen = atomicCAS(table + at, KG_PAT_EMPTY, (unsigned long long)key);
		if (seen == KG_PAT_EMPTY) return true;
		if (seen == key) return false;
		at = (at + 1) & (slots - 1);
	}
}

// view: memory-order rows (the raw tile when the column map is the identity); mask = valid-column mask per word of the view
__global__ void kg_patterns_rows_kernel(KgRowView view, const uint64_t *__restrict__ mask, uint32_t w_mem, uint32_t n_used, uint32_t min_count,
                                        unsigned long long *table, uint64_t slots, unsigned long long *counters /*[0] distinct [1] has empty key [2] kept*/) {
	unsigned long long added = 0, kept = 0;
	for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < view.n_rows; r += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t *row = view.base + r * view.stride + 1;
		uint32_t cnt = 0;
		uint64_t seed = 0;
		for (uint32_t w = 0; w < w_mem; w++) {
			const uint64_t v = w < view.w_in ? (row[w] & mask[w]) : 0ull;
			cnt += __popcll(v);
			seed ^= kg_hash64(v) + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2);
		}
		if (!(cnt >= min_count && cnt + min_count <= n_used)) continue;   // load_kmers :121
		kept++;
		if (kg_pat_insert(table, slots, seed, counters + 1)) added++;
	}
	if (added) atomicAdd(counters + 0, added);
	if (kept) atomicAdd(counters + 2, kept);
}

__global__ void kg_patterns_keys_kernel(const unsigned long long *__restrict__ keys, uint64_t n, unsigned long long *table, uint64_t slots,
                                        unsigned long long *counters) {
	unsigned long long added = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		if (kg_pat_insert(table, slots, keys[i], counters + 1)) added++;
	if (added) atomicAdd(counters + 0, added);
}

// old table -> new (larger) table; counters[0] is rebuilt by the inserts
__global__ void kg_patterns_rehash_kernel(const unsigned long long *__restrict__ old_table, uint64_t old_slots, unsigned long long *table, uint64_t slots,
                                          unsigned long long *counters) {
	unsigned long long added = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < old_slots; i += (uint64_t)gridDim.x * blockDim.x) {
		const unsigned long long k = old_table[i];
		if (k != KG_PAT_EMPTY && kg_pat_insert(table, slots, k, counters + 1)) added++;
	}
	if (added) atomicAdd(counters + 0, added);
}

// compact the keys of the table into out (any order); *n_out counts them
__global__ void kg_patterns_export_kernel(const unsigned long long *__restrict__ table, uint64_t slots, unsigned long long *out, uint64_t cap,
                                          unsigned long long *n_out) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (uint64_t)gridDim.x * blockDim.x) {
		const unsigned long long k = table[i];
		if (k == KG_PAT_EMPTY) continue;
		const unsigned long long at = atomicAdd(n_out, 1ull);
		if (at < cap) out[at] = k;
	}
}
