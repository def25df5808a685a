// association_driver.h -- the batch loop of associate_kmers (reference associate_kmers.cpp:123-148)
// over the C ABI: device candidates -> exact replay through BestAssociationsHeap.
// Used by MultipleKmersDataBases::add_kmers_to_heaps, the CLI and the C API for bench / tests.
//
// The reference fans one CTPL task per phenotype out over a loaded batch (associate_kmers.cpp:134-141) and joins
// them before loading the next batch.  Here the device scores all phenotypes of a round of rows at once, and
// the host keeps the same per-phenotype task structure for the part that stays on the CPU -- replaying the
// device's candidate hits through the P independent heaps on a small thread pool -- while the device already
// scans the next round (one hit interval in flight, kg_scan_mark / kg_scan_fetch).
#ifndef KGH_ASSOCIATION_DRIVER_H
#define KGH_ASSOCIATION_DRIVER_H

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "best_associations_heap.h"
#include "kmersgwas_b200.h"

unsigned kgh_host_threads();

// Minimal fork-join pool: run(n, fn) calls fn(i) for i in [0, n) on the workers and the calling thread.
class KghTaskPool {
	public:
		explicit KghTaskPool(unsigned n_threads);
		~KghTaskPool();
		KghTaskPool(const KghTaskPool &) = delete;
		KghTaskPool &operator=(const KghTaskPool &) = delete;
		void run(std::size_t n_tasks, const std::function<void(std::size_t)> &fn);
		// Same, but task i always runs on thread i mod threads(): per-phenotype heaps (240 KB each at K = 10001) then
		// stay in the L2 of the core that replayed them last round instead of migrating with a dynamic schedule.
		void run_static(std::size_t n_tasks, const std::function<void(std::size_t)> &fn);
		unsigned threads() const { return (unsigned)m_workers.size() + 1; }
	private:
		void worker(unsigned index);
		void drain(unsigned index);
		std::vector<std::thread> m_workers;
		std::mutex m_mu;
		std::condition_variable m_cv_work, m_cv_done;
		const std::function<void(std::size_t)> *m_fn = nullptr;
		std::size_t m_n = 0;
		std::atomic<std::size_t> m_next{0};
		std::size_t m_active = 0;
		uint64_t m_generation = 0;
		bool m_stop = false;
		bool m_static = false;
};

struct AssociationDriverState {
	AssociationDriverState();
	~AssociationDriverState();
	AssociationDriverState(const AssociationDriverState &) = delete;
	AssociationDriverState &operator=(const AssociationDriverState &) = delete;

	uint64_t rows_scored = 0;        // rows whose hits have been replayed
	uint64_t rows_submitted = 0;     // rows handed to the device since the phenotypes were set
	uint64_t rounds = 0;             // hit intervals so far
	uint64_t hits_replayed = 0;      // candidates replayed through the heaps
	uint64_t rows_kept = 0;          // rows that passed the MAC filter (replayed intervals)
	uint64_t d2h_bytes = 0;          // hits + counters copied back from the device
	uint64_t h2d_small_bytes = 0;    // thresholds sent to the device (the tiles themselves are counted by the caller)
	// host wall time per phase (ns): waiting for the device, copying hits, grouping, replaying, thresholds + submit
	uint64_t ns_wait = 0, ns_copy = 0, ns_group = 0, ns_replay = 0, ns_submit = 0;
	std::vector<double> thr;
	// Multi-GPU shards: keep every replayed candidate so that the shards' logs can be merged and
	// replayed once more, in global row order, through the final heaps (kgh_merge_shards).
	bool log_hits = false;
	// Append-only chunks of 128 MB on transparent huge pages: appending a round's hits must not re-copy a 100+ MB
	// vector nor take a 4 KB page fault per 128 hits.
	struct LogChunk { kg_hit *p; std::size_t n, cap; };
	std::vector<LogChunk> hit_log;
	void log_append(const kg_hit *hits, std::size_t n);
	std::size_t hit_log_size() const {
		std::size_t n = 0;
		for (const auto &c : hit_log) n += c.n;
		return n;
	}

	// ---- pipeline state
	// the round in flight, so that a hit-buffer overflow found later (next call, or kgh_associate_finish) can be redone in
	// smaller pieces: the caller keeps host rows valid until kgh_associate_finish anyway
	const uint64_t *flight_rows = nullptr;
	uint64_t flight_n = 0, flight_first_id = 0;
	std::size_t flight_stride = 0;
	uint64_t shrink = 1;             // divides the round length after an overflow (sticky for the rest of the scan)
	bool in_flight = false;          // the open device interval holds submitted rows whose hits are not replayed yet
	uint64_t in_flight_rows = 0;
	uint64_t kept_seen = 0;          // rows_kept total reported by the last fetch
	kg_ctx *pinned_owner = nullptr;
	kg_hit *hit_buf = nullptr;       // pinned (kg_host_alloc)
	std::size_t hit_cap = 0;
	std::vector<kg_hit> bucketed;    // hits of the current interval grouped by phenotype
	std::vector<std::size_t> bucket_off;
	std::vector<std::size_t> slice_count;   // [threads][P] scratch of the parallel counting sort
	KghTaskPool *pool = nullptr;
};

// Score rows [0, n_rows) (raw .table rows, host or device memory) against the phenotypes already set
// on ctx, feeding heaps[p].  Row ids = first_row_id + index.  Throws std::runtime_error on ABI errors.
// On return the LAST round may still be in flight on the device (its hits not yet in the heaps): call
// kgh_associate_finish before reading the heaps.  Host rows must stay valid until then.
void kgh_associate_rows(kg_ctx *ctx, BestAssociationsHeap *const *heaps, std::size_t n_heaps, const uint64_t *rows,
                        uint64_t n_rows, uint64_t first_row_id, std::size_t stride_words, AssociationDriverState &state);
void kgh_associate_finish(kg_ctx *ctx, BestAssociationsHeap *const *heaps, std::size_t n_heaps, AssociationDriverState &state);

// Exact merge of row-sharded scans.  Each shard ran kgh_associate_rows on its own
// contiguous row block with log_hits = true and its own (local) heaps.  A shard's local threshold is
// never above the sequential reference heap's threshold at the same row (the reference heap has seen a
// superset of rows), so the union of the logs contains every row the reference heap would accept.
// Replaying the logs in global row order through fresh heaps reproduces the reference state, ties
// included; rows_kept sums to the reference's number_of_insertion().
void kgh_merge_shards(std::vector<AssociationDriverState *> &shards, BestAssociationsHeap *const *final_heaps,
                      std::size_t n_heaps);
// Same for already concatenated logs (any order).
void kgh_merge_hit_log(std::vector<kg_hit> &all, uint64_t rows_kept, BestAssociationsHeap *const *final_heaps,
                       std::size_t n_heaps);

#endif
