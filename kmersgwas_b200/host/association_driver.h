// association_driver.h -- the batch loop of associate_kmers (reference associate_kmers.cpp:123-148)
// over the C ABI: device candidates -> exact replay through BestAssociationsHeap.
// Used by MultipleKmersDataBases::add_kmers_to_heaps, the CLI and the C API for bench / tests.
#ifndef KGH_ASSOCIATION_DRIVER_H
#define KGH_ASSOCIATION_DRIVER_H

#include <vector>

#include "best_associations_heap.h"
#include "kmersgwas_b200.h"

struct AssociationDriverState {
	uint64_t rows_scored = 0;        // rows scored since the phenotypes were set on the context
	uint64_t rounds = 0;             // threshold refreshes so far
	uint64_t hits_replayed = 0;      // candidates replayed through the heaps
	std::vector<kg_hit> hit_buf;
	std::vector<double> thr;
	// Multi-GPU shards: keep every replayed candidate so that the shards' logs can be merged and
	// replayed once more, in global row order, through the final heaps (kgh_merge_shards).
	bool log_hits = false;
	std::vector<kg_hit> hit_log;
	uint64_t rows_kept = 0;          // rows that passed the MAC filter (all rounds)
	uint64_t d2h_bytes = 0;          // hits + counters copied back from the device
	uint64_t h2d_small_bytes = 0;    // thresholds sent to the device (the tiles themselves are counted by the caller)
};

// Score rows [0, n_rows) (raw .table rows, host or device memory) against the phenotypes already set
// on ctx, feeding heaps[p].  Row ids = first_row_id + index.  Throws std::runtime_error on ABI errors.
void kgh_associate_rows(kg_ctx *ctx, BestAssociationsHeap *const *heaps, std::size_t n_heaps, const uint64_t *rows,
                        uint64_t n_rows, uint64_t first_row_id, std::size_t stride_words, AssociationDriverState &state);

// Exact merge of row-sharded scans.  Each shard ran kgh_associate_rows on its own
// contiguous row block with log_hits = true and its own (local) heaps.  A shard's local threshold is
// never above the sequential reference heap's threshold at the same row (the reference heap has seen a
// superset of rows), so the union of the logs contains every row the reference heap would accept.
// Replaying the logs in global row order through fresh heaps reproduces the reference state, ties
// included; rows_kept sums to the reference's number_of_insertion().
void kgh_merge_shards(std::vector<AssociationDriverState *> &shards, BestAssociationsHeap *const *final_heaps,
                      std::size_t n_heaps);

#endif
