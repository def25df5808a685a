// host check of the device heap algorithms against oracle.c's restatement of std::priority_queue
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../kmersgwas_b200/csrc/kg_select.cuh"
#include "../oracle/oracle.h"
int main() {
	for (int trial = 0; trial < 200; trial++) {
		const uint32_t K = 1 + rand() % (trial % 3 == 0 ? 2000 : 50);
		const int n = rand() % 2000;
		const uint32_t kpad = kg_select_kmax_pad(K);
		void *raw = aligned_alloc(16, kpad * 8 + 32); double *hs = (double *)((char *)raw + 8);
		void *raw2 = aligned_alloc(16, kpad * 4 + 32); uint32_t *hlp = (uint32_t *)((char *)raw2 + 4);
		struct { uint32_t *p; uint32_t *data() { return p; } uint32_t &operator[](size_t i) { return p[i]; } } hl{hlp};
		std::vector<uint64_t> pk(K), pr(K);
		uint32_t size = 0;
		// second copy driven by the split form (scores by the sequential half, slots from the records)
		void *raw3 = aligned_alloc(16, kpad * 8 + 32); double *hs2 = (double *)((char *)raw3 + 8);
		std::vector<uint32_t> hl2(kpad + 8);
		kgo_heap *o = kgo_heap_new(K);
		for (int i = 0; i < n; i++) {
			const double s = (double)(rand() % 97);   // many ties
			const uint64_t kmer = 1000 + i, row = i;
			kgo_heap_add(o, kmer, s, row);
			uint32_t slot;
			{
				const KgHeapPtr m2{hs2, hl2.data()};
				KgHeapRec rec; rec.cand = 0; bool admit = true;
				if (size < K) { rec.leaf = 0; rec.pos = size; rec.info = (kg_heap_push_up_scores(m2, (int32_t)size, s) << 8) | (KG_REC_PUSH << 16); }
				else if (s > hs2[0]) kg_heap_replace_top_scores(m2, (int32_t)size, (int32_t)kpad, s, rec);
				else admit = false;
				if (admit) kg_heap_apply_slots_seq(hl2.data(), rec);
			}
			if (size < K) { slot = size; kg_heap_push_up(hs, hl.data(), (int32_t)size, s, slot); size++; }
			else { if (!(s > hs[0])) continue; slot = hl[0]; kg_heap_replace_top(hs, hl.data(), (int32_t)size, (int32_t)kpad, s, slot); }
			pk[slot] = kmer; pr[slot] = row;
		}
		for (uint32_t i = 0; i < size; i++) if (hs2[i] != hs[i] || hl2[i] != hl[i]) { printf("trial %d: split form differs at %u\n", trial, i); return 1; }
		free(raw3);
		// compare by popping copies: push layout into a fresh oracle heap
		kgo_heap *d = kgo_heap_new(K ? K : 1);
		for (uint32_t i = 0; i < size; i++) kgo_heap_add(d, pk[hl[i]], hs[i], pr[hl[i]]);
		const uint64_t so = kgo_heap_size(o), sd = kgo_heap_size(d);
		if (so != sd) { printf("trial %d: size %llu vs %llu\n", trial, (unsigned long long)so, (unsigned long long)sd); return 1; }
		std::vector<uint64_t> k1(so), r1(so), k2(so), r2(so); std::vector<double> s1(so), s2(so);
		kgo_heap_dump(o, k1.data(), s1.data(), r1.data());
		kgo_heap_dump(d, k2.data(), s2.data(), r2.data());
		for (uint64_t i = 0; i < so; i++) if (k1[i] != k2[i] || s1[i] != s2[i] || r1[i] != r2[i]) { printf("trial %d K %u n %d: mismatch at %llu\n", trial, K, n, (unsigned long long)i); return 1; }
		kgo_heap_free(o); kgo_heap_free(d); free(raw); free(raw2);
	}
	printf("ok\n");
	return 0;
}
