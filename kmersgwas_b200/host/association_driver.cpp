// association_driver.cpp -- see association_driver.h.
//
// The batch is cut into rounds.  Each round the device reports every (row, p) whose score is
// > heap_p.lowest_score as of the START of the round (every kept row while heap_p is not full); the
// hits come back sorted by (phenotype, row) and are replayed through the real heaps; thresholds are
// then refreshed.  lowest_score never decreases (best_associations_heap.cpp:43-59), so the candidate
// set is a superset of what the sequential reference heap accepts, and a dropped row could not have
// changed the heap (strict '>'): the heap state, ties included, is identical to the reference's.
#include "association_driver.h"

#include <algorithm>
#include <stdexcept>
#include <string>

namespace {
const uint64_t kSubTileRows = 1ull << 20;   // rows per device submit: the H2D copy of sub-tile i+1 overlaps the kernels of i
const uint64_t kMaxRoundRows = 1ull << 24;  // rows between two threshold refreshes once the heaps are warm
const uint64_t kHitBudget = 1ull << 21;     // expected hits per round (the device buffer holds 1 << 22 by default)

void check(kg_ctx *ctx, kg_status st, const char *what) {
	if (st != KG_OK) throw std::runtime_error(std::string(what) + ": " + kg_last_error(ctx));
}
}  // namespace

void kgh_associate_rows(kg_ctx *ctx, BestAssociationsHeap *const *heaps, std::size_t P, const uint64_t *rows,
                        uint64_t n_rows, uint64_t first_row_id, std::size_t stride, AssociationDriverState &S) {
	S.thr.resize(P);
	uint64_t kept_before = 0;
	{
		std::size_t n_hits = 0;
		uint64_t seen = 0;
		check(ctx, kg_scan_fetch(ctx, nullptr, 0, &n_hits, &seen, &kept_before), "kg_scan_fetch");
	}
	uint64_t done = 0;
	while (done < n_rows) {
		std::size_t cold = 0, need = 0;
		uint64_t kmax = 1;
		for (std::size_t j = 0; j < P; j++) {
			S.thr[j] = heaps[j]->device_threshold();
			kmax = std::max<uint64_t>(kmax, heaps[j]->capacity());
			if (S.thr[j] < 0) {
				cold++;
				need = std::max(need, heaps[j]->capacity() - heaps[j]->size());
			}
		}
		uint64_t round;
		if (cold) {
			// every kept row of a cold phenotype is a hit: bound rows x cold phenotypes
			round = std::min<uint64_t>(need + need / 4 + 64, std::max<uint64_t>(kHitBudget / cold, 1));
		} else {
			// warm: expected hits per phenotype ~ K * round / rows_so_far -> the round grows with the scan
			const double per_row = (double)(P * kmax) / (double)std::max<uint64_t>(S.rows_scored, 1);
			round = std::min<uint64_t>(std::max<uint64_t>((uint64_t)((double)kHitBudget / per_row), 4096), kMaxRoundRows);
		}
		round = std::min<uint64_t>(std::max<uint64_t>(round, 1), n_rows - done);
		std::size_t n_hits = 0;
		uint64_t kept_now = 0;
		for (;;) {
			check(ctx, kg_scan_set_thresholds(ctx, S.thr.data(), (uint32_t)P), "kg_scan_set_thresholds");
			for (uint64_t off = 0; off < round; off += kSubTileRows) {
				const uint64_t n = std::min<uint64_t>(kSubTileRows, round - off);
				check(ctx, kg_scan_submit(ctx, rows + (done + off) * stride, n, first_row_id + done + off), "kg_scan_submit");
			}
			uint64_t seen = 0;
			const kg_status st = kg_scan_fetch(ctx, nullptr, 0, &n_hits, &seen, &kept_now);
			if (st == KG_ERR_HITS_OVERFLOW && round > 1) {
				round = std::max<uint64_t>(round / 4, 1);  // the library rolled its counters back
				continue;
			}
			check(ctx, st, "kg_scan_fetch");
			break;
		}
		S.hit_buf.resize(n_hits);
		if (n_hits) check(ctx, kg_scan_fetch(ctx, S.hit_buf.data(), n_hits, &n_hits, nullptr, nullptr), "kg_scan_fetch");
		check(ctx, kg_scan_clear_hits(ctx), "kg_scan_clear_hits");
		const uint64_t kept_round = kept_now - kept_before;
		kept_before = kept_now;
		std::size_t i = 0;
		for (std::size_t j = 0; j < P; j++) {  // hits are sorted by (phenotype, row)
			std::size_t e = i;
			while (e < n_hits && S.hit_buf[e].pheno == j) e++;
			heaps[j]->add_hits(S.hit_buf.data() + i, e - i);
			heaps[j]->note_tested_rows((std::size_t)(kept_round - (e - i)));
			i = e;
		}
		if (S.log_hits) S.hit_log.insert(S.hit_log.end(), S.hit_buf.begin(), S.hit_buf.end());
		S.rows_kept += kept_round;
		S.d2h_bytes += (uint64_t)n_hits * sizeof(kg_hit) + 4 * sizeof(uint64_t);
		S.h2d_small_bytes += (uint64_t)P * sizeof(double);
		done += round;
		S.rows_scored += round;
		S.rounds++;
		S.hits_replayed += n_hits;
	}
}

void kgh_merge_shards(std::vector<AssociationDriverState *> &shards, BestAssociationsHeap *const *final_heaps,
                      std::size_t P) {
	std::vector<kg_hit> all;
	uint64_t kept = 0;
	for (AssociationDriverState *s : shards) {
		all.insert(all.end(), s->hit_log.begin(), s->hit_log.end());
		kept += s->rows_kept;
	}
	std::sort(all.begin(), all.end(), [](const kg_hit &a, const kg_hit &b) {
		return a.pheno != b.pheno ? a.pheno < b.pheno : a.row < b.row;
	});
	std::size_t i = 0;
	for (std::size_t j = 0; j < P; j++) {
		std::size_t e = i;
		while (e < all.size() && all[e].pheno == j) e++;
		final_heaps[j]->add_hits(all.data() + i, e - i);
		final_heaps[j]->note_tested_rows((std::size_t)(kept - (e - i)));
		i = e;
	}
}
