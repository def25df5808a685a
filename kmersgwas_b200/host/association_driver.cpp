// association_driver.cpp -- see association_driver.h.
//
// The rows are cut into rounds = device hit intervals.  For each round the device reports every (row, p) whose
// score is > heap_p.lowest_score as the host knew it when the round was submitted (every kept row while heap_p
// is not full).  lowest_score never decreases (best_associations_heap.cpp:43-59), so a stale threshold only adds
// candidates: the candidate set is a superset of what the sequential reference heap accepts, and a dropped row
// could not have changed the heap (strict '>').  The candidates are grouped by phenotype, sorted by row and
// replayed through the real heaps -- one task per phenotype, like the reference's CTPL fan-out -- which
// reproduces the reference heap state, ties included.
#include "association_driver.h"

#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>

namespace {
const uint64_t kMaxRoundRows = 1ull << 23;  // rows per round once the heaps are warm
const uint64_t kMinWarmRound = 1ull << 10;   // floor of a warm round (the hit budget decides above it)
const uint64_t kHitBudget = 1ull << 21;     // expected hits per round (the device buffer holds 1 << 22 by default)

uint64_t now_ns() {
	return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void check(kg_ctx *ctx, kg_status st, const char *what) {
	if (st != KG_OK) throw std::runtime_error(std::string(what) + ": " + kg_last_error(ctx));
}
}  // namespace

// Host threads for the per-phenotype replay tasks: KMERSGWAS_HOST_THREADS if set, else the cores of the box
// divided by the ranks sharing it (torchrun's LOCAL_WORLD_SIZE), at most 32.
unsigned kgh_host_threads() {
	if (const char *e = getenv("KMERSGWAS_HOST_THREADS")) {
		const int v = atoi(e);
		if (v > 0) return (unsigned)v;
	}
	unsigned hw = std::thread::hardware_concurrency();
	if (hw == 0) hw = 4;
	unsigned ranks = 1;
	if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
		const int v = atoi(e);
		if (v > 0) ranks = (unsigned)v;
	}
	return std::max(1u, std::min(hw / ranks, 32u));
}

// ------------------------------------------------------------------------------------------ pool
KghTaskPool::KghTaskPool(unsigned n_threads) {
	for (unsigned i = 1; i < n_threads; i++) m_workers.emplace_back([this, i] { worker(i); });
}

KghTaskPool::~KghTaskPool() {
	{
		std::lock_guard<std::mutex> lk(m_mu);
		m_stop = true;
	}
	m_cv_work.notify_all();
	for (std::thread &t : m_workers) t.join();
}

void KghTaskPool::drain(unsigned index) {
	if (m_static) {
		const std::size_t T = m_workers.size() + 1;
		for (std::size_t i = index; i < m_n; i += T) (*m_fn)(i);
		return;
	}
	for (;;) {
		const std::size_t i = m_next.fetch_add(1, std::memory_order_relaxed);
		if (i >= m_n) return;
		(*m_fn)(i);
	}
}

void KghTaskPool::worker(unsigned index) {
	uint64_t seen = 0;
	for (;;) {
		{
			std::unique_lock<std::mutex> lk(m_mu);
			m_cv_work.wait(lk, [&] { return m_stop || m_generation != seen; });
			if (m_stop) return;
			seen = m_generation;
		}
		drain(index);
		{
			std::lock_guard<std::mutex> lk(m_mu);
			if (--m_active == 0) m_cv_done.notify_one();
		}
	}
}

void KghTaskPool::run_static(std::size_t n_tasks, const std::function<void(std::size_t)> &fn) {
	m_static = true;    // read by the workers only after the generation bump under the mutex in run()
	run(n_tasks, fn);
	m_static = false;
}

void KghTaskPool::run(std::size_t n_tasks, const std::function<void(std::size_t)> &fn) {
	if (n_tasks == 0) return;
	if (m_workers.empty() || n_tasks == 1) {
		for (std::size_t i = 0; i < n_tasks; i++) fn(i);
		return;
	}
	{
		std::lock_guard<std::mutex> lk(m_mu);
		m_fn = &fn;
		m_n = n_tasks;
		m_next.store(0, std::memory_order_relaxed);
		m_active = m_workers.size();
		m_generation++;
	}
	m_cv_work.notify_all();
	drain(0);
	std::unique_lock<std::mutex> lk(m_mu);
	m_cv_done.wait(lk, [&] { return m_active == 0; });
	m_fn = nullptr;
}

// ------------------------------------------------------------------------------------------ state
AssociationDriverState::AssociationDriverState() {}

AssociationDriverState::~AssociationDriverState() {
	if (hit_buf && pinned_owner) kg_host_free(pinned_owner, hit_buf);
	for (LogChunk &c : hit_log) free(c.p);
	delete pool;
}

void AssociationDriverState::log_append(const kg_hit *hits, std::size_t n) {
	while (n) {
		if (hit_log.empty() || hit_log.back().n == hit_log.back().cap) {
			const std::size_t cap = 1u << 22;   // 4 M hits = 128 MB
			void *p = nullptr;
			if (posix_memalign(&p, 2u << 20, cap * sizeof(kg_hit)) != 0) throw std::bad_alloc();
			madvise(p, cap * sizeof(kg_hit), MADV_HUGEPAGE);
			hit_log.push_back(LogChunk{static_cast<kg_hit *>(p), 0, cap});
		}
		LogChunk &c = hit_log.back();
		const std::size_t m = std::min(n, c.cap - c.n);
		memcpy(c.p + c.n, hits, m * sizeof(kg_hit));
		c.n += m;
		hits += m;
		n -= m;
	}
}

// ------------------------------------------------------------------------------------------ replay
// Fetch the interval that is in flight and replay it.  Returns false if the device reported a hit overflow
// (the interval was dropped; nothing was replayed).
static bool replay_in_flight(kg_ctx *ctx, BestAssociationsHeap *const *heaps, std::size_t P, AssociationDriverState &S) {
	if (!S.in_flight) return true;
	std::size_t n_hits = 0;
	uint64_t seen = 0, kept_now = 0;
	uint64_t t0 = now_ns();
	kg_status st = kg_scan_fetch(ctx, nullptr, 0, &n_hits, &seen, &kept_now);
	S.ns_wait += now_ns() - t0;
	t0 = now_ns();
	if (st == KG_ERR_HITS_OVERFLOW) {
		S.in_flight = false;
		return false;
	}
	check(ctx, st, "kg_scan_fetch");
	if (n_hits) {
		if (S.hit_cap < n_hits) {
			if (S.hit_buf) kg_host_free(S.pinned_owner, S.hit_buf);
			S.hit_buf = nullptr;
			S.hit_cap = 0;
			const std::size_t cap = std::max<std::size_t>(n_hits + n_hits / 2, 1u << 16);
			void *p = nullptr;
			check(ctx, kg_host_alloc(ctx, cap * sizeof(kg_hit), &p), "kg_host_alloc");
			S.hit_buf = static_cast<kg_hit *>(p);
			S.hit_cap = cap;
			S.pinned_owner = ctx;
		}
		check(ctx, kg_scan_fetch(ctx, S.hit_buf, S.hit_cap, &n_hits, &seen, &kept_now), "kg_scan_fetch");
	}
	const uint64_t kept_round = kept_now - S.kept_seen;
	S.kept_seen = kept_now;
	S.ns_copy += now_ns() - t0;
	t0 = now_ns();

	// group by phenotype (counting sort, parallel over slices of the hit buffer), then one task per phenotype:
	// sort by row + replay through its heap
	if (!S.pool) S.pool = new KghTaskPool(std::min<unsigned>(kgh_host_threads(), (unsigned)P));
	const std::size_t T = n_hits >= 8192 ? S.pool->threads() : 1;
	const std::size_t slice = (n_hits + T - 1) / std::max<std::size_t>(T, 1);
	S.slice_count.assign(T * P, 0);
	const kg_hit *const raw = S.hit_buf;
	S.pool->run(T, [&](std::size_t t) {
		std::size_t *cnt = S.slice_count.data() + t * P;
		const std::size_t b = t * slice, e = std::min(n_hits, b + slice);
		for (std::size_t i = b; i < e; i++) cnt[raw[i].pheno]++;
	});
	S.bucket_off.assign(P + 1, 0);
	{
		std::size_t run = 0;   // slice_count[t][j] becomes the write position of slice t inside bucket j
		for (std::size_t j = 0; j < P; j++) {
			S.bucket_off[j] = run;
			for (std::size_t t = 0; t < T; t++) {
				const std::size_t c = S.slice_count[t * P + j];
				S.slice_count[t * P + j] = run;
				run += c;
			}
		}
		S.bucket_off[P] = run;
	}
	S.bucketed.resize(n_hits);
	kg_hit *const base = S.bucketed.data();
	S.pool->run(T, [&](std::size_t t) {
		std::size_t *at = S.slice_count.data() + t * P;
		const std::size_t b = t * slice, e = std::min(n_hits, b + slice);
		for (std::size_t i = b; i < e; i++) base[at[raw[i].pheno]++] = raw[i];
	});
	S.ns_group += now_ns() - t0;
	t0 = now_ns();
	const std::vector<std::size_t> &off = S.bucket_off;
	// one task per phenotype; in sharded runs task 0 appends the round's hits to the shard log meanwhile
	const std::size_t log_task = (S.log_hits && n_hits) ? 1 : 0;
	S.pool->run_static(P + log_task, [&](std::size_t task) {
		if (task >= P) { S.log_append(raw, n_hits); return; }
		const std::size_t j = task;
		kg_hit *b = base + off[j], *e = base + off[j + 1];
		std::sort(b, e, [](const kg_hit &x, const kg_hit &y) { return x.row < y.row; });
		heaps[j]->add_hits(b, (std::size_t)(e - b));
		heaps[j]->note_tested_rows((std::size_t)(kept_round - (uint64_t)(e - b)));
	});
	S.ns_replay += now_ns() - t0;
	S.rows_kept += kept_round;
	S.rows_scored += S.in_flight_rows;
	S.hits_replayed += n_hits;
	S.d2h_bytes += (uint64_t)n_hits * sizeof(kg_hit) + 8 * sizeof(uint64_t);
	S.rounds++;
	S.in_flight = false;
	S.in_flight_rows = 0;
	return true;
}

// Redo the round that overflowed the device hit buffer (it was dropped as a whole) in rounds a quarter as long.
static void redo_flight(kg_ctx *ctx, BestAssociationsHeap *const *heaps, std::size_t P, AssociationDriverState &S) {
	check(ctx, kg_scan_discard(ctx), "kg_scan_discard");
	if (!S.flight_rows || S.flight_n == 0) throw std::runtime_error("hit buffer overflow and the round's rows are gone (raise KG_OPT_HIT_CAPACITY)");
	if (S.shrink > (1ull << 40)) throw std::runtime_error("hit buffer overflow even with one-row rounds (raise KG_OPT_HIT_CAPACITY)");
	const uint64_t *rows = S.flight_rows;
	const uint64_t n = S.flight_n, first = S.flight_first_id;
	const std::size_t stride = S.flight_stride;
	S.rows_submitted -= n;
	S.shrink *= 4;
	S.flight_rows = nullptr;
	S.flight_n = 0;
	kgh_associate_rows(ctx, heaps, P, rows, n, first, stride, S);
}

void kgh_associate_finish(kg_ctx *ctx, BestAssociationsHeap *const *heaps, std::size_t P, AssociationDriverState &S) {
	while (!replay_in_flight(ctx, heaps, P, S)) redo_flight(ctx, heaps, P, S);
}

void kgh_associate_rows(kg_ctx *ctx, BestAssociationsHeap *const *heaps, std::size_t P, const uint64_t *rows,
                        uint64_t n_rows, uint64_t first_row_id, std::size_t stride, AssociationDriverState &S) {
	S.thr.resize(P);
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
		if (cold && S.in_flight) {
			// cold heaps take every kept row: their thresholds must be exact, so no round stays in flight
			if (!replay_in_flight(ctx, heaps, P, S)) redo_flight(ctx, heaps, P, S);
			continue;
		}
		uint64_t round;
		if (cold) {
			// every kept row of a cold phenotype is a hit: bound rows x cold phenotypes
			round = std::min<uint64_t>(need + need / 4 + 64, std::max<uint64_t>(kHitBudget / cold, 1));
		} else {
			// warm: expected hits per phenotype ~ K * round / rows_so_far (x2: thresholds are one round stale)
			const double per_row = 2.0 * (double)(P * kmax) / (double)std::max<uint64_t>(S.rows_submitted, 1);
			round = std::min<uint64_t>(std::max<uint64_t>((uint64_t)((double)kHitBudget / per_row), kMinWarmRound), kMaxRoundRows);
		}
		round = std::max<uint64_t>(round / S.shrink, 1);
		round = std::min<uint64_t>(round, n_rows - done);

		const uint64_t t_sub = now_ns();
		check(ctx, kg_scan_set_thresholds(ctx, S.thr.data(), (uint32_t)P), "kg_scan_set_thresholds");
		S.h2d_small_bytes += (uint64_t)P * sizeof(double);
		// one submit per round: the library scans device rows in place and cuts host rows into sub-tiles itself
		check(ctx, kg_scan_submit(ctx, rows + done * stride, round, first_row_id + done), "kg_scan_submit");
		S.ns_submit += now_ns() - t_sub;
		// the device now works on this round; meanwhile replay the previous one
		if (!replay_in_flight(ctx, heaps, P, S)) {
			// overflow in the PREVIOUS round: everything not replayed is dropped (this round too); redo the previous
			// round in smaller pieces, then come back to this one with the shorter rounds
			redo_flight(ctx, heaps, P, S);
			continue;
		}
		check(ctx, kg_scan_mark(ctx), "kg_scan_mark");
		S.in_flight = true;
		S.in_flight_rows = round;
		S.flight_rows = rows + done * stride;
		S.flight_n = round;
		S.flight_first_id = first_row_id + done;
		S.flight_stride = stride;
		S.rows_submitted += round;
		done += round;
		if (cold && !replay_in_flight(ctx, heaps, P, S)) redo_flight(ctx, heaps, P, S);
	}
}

void kgh_merge_hit_log(std::vector<kg_hit> &all, uint64_t kept, BestAssociationsHeap *const *final_heaps, std::size_t P) {
	// group by phenotype (stable counting sort), then one task per phenotype: order by row (shard logs are already
	// row-sorted, so this is usually a no-op check) and replay through the final heap
	std::vector<std::size_t> off(P + 1, 0);
	for (const kg_hit &h : all)
		if (h.pheno < P) off[h.pheno + 1]++;
	for (std::size_t j = 0; j < P; j++) off[j + 1] += off[j];
	std::vector<kg_hit> grouped(off[P]);
	{
		std::vector<std::size_t> at(off.begin(), off.end() - 1);
		for (const kg_hit &h : all)
			if (h.pheno < P) grouped[at[h.pheno]++] = h;
	}
	std::vector<kg_hit>().swap(all);
	KghTaskPool pool(std::min<unsigned>(kgh_host_threads(), (unsigned)std::max<std::size_t>(P, 1)));
	pool.run(P, [&](std::size_t j) {
		kg_hit *b = grouped.data() + off[j], *e = grouped.data() + off[j + 1];
		auto by_row = [](const kg_hit &x, const kg_hit &y) { return x.row < y.row; };
		if (!std::is_sorted(b, e, by_row)) std::sort(b, e, by_row);
		final_heaps[j]->add_hits(b, (std::size_t)(e - b));
		final_heaps[j]->note_tested_rows((std::size_t)(kept - (uint64_t)(e - b)));
	});
}

void kgh_merge_shards(std::vector<AssociationDriverState *> &shards, BestAssociationsHeap *const *final_heaps,
                      std::size_t P) {
	std::vector<kg_hit> all;
	uint64_t kept = 0;
	std::size_t total = 0;
	for (AssociationDriverState *s : shards) total += s->hit_log_size();
	all.reserve(total);
	for (AssociationDriverState *s : shards) {
		for (const auto &chunk : s->hit_log) all.insert(all.end(), chunk.p, chunk.p + chunk.n);
		kept += s->rows_kept;
	}
	kgh_merge_hit_log(all, kept, final_heaps, P);
}
