// host_capi.cpp -- C entry points of the HOST library (libkmersgwas_host.so) so that bench.py and the
// tests can drive the product's own host path (association_driver + BestAssociationsHeap) through
// ctypes on tiles that live in memory instead of in a .table file.
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <vector>

#include "association_driver.h"
#include "best_associations_heap.h"
#include "kmersgwas_b200.h"

namespace {
thread_local std::string g_err;
}

struct kgh_session {
	kg_ctx *ctx = nullptr;
	std::vector<BestAssociationsHeap> heaps;
	std::vector<BestAssociationsHeap *> hp;
	AssociationDriverState state;
	std::size_t stride = 0;
};

extern "C" {

const char *kgh_last_error(void) { return g_err.c_str(); }

void kgh_session_destroy(kgh_session *s) {
	if (!s) return;
	if (s->ctx) kg_ctx_destroy(s->ctx);
	delete s;
}

kgh_session *kgh_session_create(int device, uint64_t n_file, uint64_t n_used, const uint32_t *map_word,
                                const uint32_t *map_bit, const float *y, uint32_t n_pheno, uint64_t min_count,
                                const uint64_t *kbest, void *stream, int scan_engine, int log_hits) {
	std::unique_ptr<kgh_session> s(new kgh_session());
	kg_shape shape;
	shape.n_file = n_file; shape.n_used = n_used; shape.map_word = map_word; shape.map_bit = map_bit;
	if (kg_ctx_create(device, &shape, stream, &s->ctx) != KG_OK) { g_err = kg_last_error(nullptr); return nullptr; }
	if (kg_set_option(s->ctx, KG_OPT_SCAN_ENGINE, scan_engine) != KG_OK ||
	    kg_scan_set_phenotypes(s->ctx, y, n_pheno, min_count) != KG_OK) {
		g_err = kg_last_error(s->ctx);
		kgh_session_destroy(s.release());
		return nullptr;
	}
	for (uint32_t p = 0; p < n_pheno; p++) s->heaps.emplace_back((std::size_t)kbest[p]);
	for (auto &h : s->heaps) s->hp.push_back(&h);
	s->stride = 1 + (std::size_t)((n_file + 63) / 64);
	s->state.log_hits = log_hits != 0;
	return s.release();
}

// The last round of the call may stay in flight on the device (its hits are replayed by the next call or by
// kgh_session_finish); host rows must stay valid until then.  Every accessor below finishes first.
int kgh_session_associate(kgh_session *s, const uint64_t *rows, uint64_t n_rows, uint64_t first_row_id) {
	try {
		kgh_associate_rows(s->ctx, s->hp.data(), s->hp.size(), rows, n_rows, first_row_id, s->stride, s->state);
		return 0;
	} catch (const std::exception &e) {
		g_err = e.what();
		return 1;
	}
}

int kgh_session_finish(kgh_session *s) {
	try {
		kgh_associate_finish(s->ctx, s->hp.data(), s->hp.size(), s->state);
		return 0;
	} catch (const std::exception &e) {
		g_err = e.what();
		return 1;
	}
}

void *kgh_session_ctx(kgh_session *s) { return s->ctx; }
uint64_t kgh_session_heap_size(kgh_session *s, uint32_t p) { kgh_session_finish(s); return s->heaps[p].size(); }
uint64_t kgh_session_tested(kgh_session *s, uint32_t p) { kgh_session_finish(s); return s->heaps[p].number_of_insertion(); }
double kgh_session_threshold(kgh_session *s, uint32_t p) { kgh_session_finish(s); return s->heaps[p].device_threshold(); }

void kgh_session_heap_dump(kgh_session *s, uint32_t p, uint64_t *kmers, double *scores, uint64_t *rows) {
	kgh_session_finish(s);
	const std::vector<AssociationScoreHeap> e = s->heaps[p].entries_in_pop_order();
	for (std::size_t i = 0; i < e.size(); i++) {
		kmers[i] = std::get<0>(e[i]);
		scores[i] = std::get<1>(e[i]);
		rows[i] = std::get<2>(e[i]);
	}
}

void kgh_session_stats(kgh_session *s, uint64_t *rounds, uint64_t *hits_replayed, uint64_t *rows_scored, uint64_t *rows_kept) {
	kgh_session_finish(s);
	if (rounds) *rounds = s->state.rounds;
	if (hits_replayed) *hits_replayed = s->state.hits_replayed;
	if (rows_scored) *rows_scored = s->state.rows_scored;
	if (rows_kept) *rows_kept = s->state.rows_kept;
}

// host wall time per driver phase so far, ns: wait for device, copy hits, group, replay, thresholds + submit
void kgh_session_host_ns(kgh_session *s, uint64_t *out5) {
	out5[0] = s->state.ns_wait; out5[1] = s->state.ns_copy; out5[2] = s->state.ns_group;
	out5[3] = s->state.ns_replay; out5[4] = s->state.ns_submit;
}

void kgh_session_io_bytes(kgh_session *s, uint64_t *h2d_small, uint64_t *d2h) {
	if (h2d_small) *h2d_small = s->state.h2d_small_bytes;
	if (d2h) *d2h = s->state.d2h_bytes;
}

// Multi-shard merge: replay the shards' hit logs, in global row order, into shard 0's heaps
// (which are reset first).  Used by the world_size > 1 tests and the multi-GPU bench.
uint64_t kgh_session_log_size(kgh_session *s) { kgh_session_finish(s); return s->state.hit_log_size(); }
void kgh_session_log_copy(kgh_session *s, kg_hit *out) {
	for (const auto &chunk : s->state.hit_log) {
		memcpy(out, chunk.p, chunk.n * sizeof(kg_hit));
		out += chunk.n;
	}
}

// Standalone heap object for merging gathered logs on rank 0.
struct kgh_heapset {
	std::vector<BestAssociationsHeap> heaps;
};
kgh_heapset *kgh_heapset_create(const uint64_t *kbest, uint32_t n_pheno) {
	kgh_heapset *h = new kgh_heapset();
	for (uint32_t p = 0; p < n_pheno; p++) h->heaps.emplace_back((std::size_t)kbest[p]);
	return h;
}
void kgh_heapset_destroy(kgh_heapset *h) { delete h; }
// hits: concatenated shard logs (any order); rows_kept: total kept rows over all shards
void kgh_heapset_merge(kgh_heapset *h, const kg_hit *hits, uint64_t n_hits, uint64_t rows_kept) {
	std::vector<kg_hit> all(hits, hits + n_hits);
	std::vector<BestAssociationsHeap *> hp;
	for (auto &x : h->heaps) hp.push_back(&x);
	kgh_merge_hit_log(all, rows_kept, hp.data(), hp.size());
}
uint64_t kgh_heapset_size(kgh_heapset *h, uint32_t p) { return h->heaps[p].size(); }
uint64_t kgh_heapset_tested(kgh_heapset *h, uint32_t p) { return h->heaps[p].number_of_insertion(); }
void kgh_heapset_dump(kgh_heapset *h, uint32_t p, uint64_t *kmers, double *scores, uint64_t *rows) {
	const std::vector<AssociationScoreHeap> e = h->heaps[p].entries_in_pop_order();
	for (std::size_t i = 0; i < e.size(); i++) {
		kmers[i] = std::get<0>(e[i]);
		scores[i] = std::get<1>(e[i]);
		rows[i] = std::get<2>(e[i]);
	}
}
// Direct heap access for CPU-only tests of the host heap against the oracle.
void kgh_heapset_add(kgh_heapset *h, uint32_t p, uint64_t kmer, double score, uint64_t row) {
	h->heaps[p].add_association(kmer, score, row);
}

}  // extern "C"
