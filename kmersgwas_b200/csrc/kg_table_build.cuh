// kg_table_build.cuh -- construction of the presence/absence table.
//
// Reference: MultipleKmersDataBasesMerger::load_kmers (/root/reference/src/kmers_merge_multiple_databaes.cpp:85-121):
// for a range of the sorted list of all k-mers, every accession's sorted k-mer list is looked up in a hash map
// k-mer -> row, and bit `accession` of the row is set when the k-mer is one of the listed ones.  On the GPU the list of
// all k-mers of the range is the sorted array itself: one thread per (accession, k-mer) binary-searches it and ORs the
// bit into the row (k-mers that are not in the list are ignored, as in the reference).
#pragma once
#include "kg_common.cuh"

// rows[i] = {all[i], 0, ..., 0}
__global__ void kg_table_init_kernel(const uint64_t *__restrict__ all, uint64_t n_all, uint32_t w, uint64_t *__restrict__ table) {
	const uint64_t total = n_all * (uint64_t)(w + 1);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t r = i / (w + 1);
		table[i] = (i - r * (w + 1)) == 0 ? all[r] : 0ull;
	}
}

// packed: the accessions' k-mers of this range back to back, accession a = [off[a], off[a + 1])
__global__ void kg_table_mark_kernel(const uint64_t *__restrict__ all, uint64_t n_all, const uint64_t *__restrict__ packed,
                                     const uint64_t *__restrict__ off, uint32_t n_acc, uint32_t w, unsigned long long *__restrict__ table) {
	const uint64_t total = off[n_acc];
	for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
		// accession of element e: last a with off[a] <= e
		uint32_t lo = 0, hi = n_acc;
		while (hi - lo > 1) {
			const uint32_t mid = (lo + hi) >> 1;
			if (off[mid] <= e) lo = mid; else hi = mid;
		}
		const uint32_t a = lo;
		const uint64_t k = packed[e];
		uint64_t l = 0, h = n_all;
		while (l < h) {
			const uint64_t m = (l + h) >> 1;
			if (all[m] < k) l = m + 1; else h = m;
		}
		if (l < n_all && all[l] == k) atomicOr(table + l * (uint64_t)(w + 1) + 1 + (a >> 6), 1ull << (a & 63));
	}
}
