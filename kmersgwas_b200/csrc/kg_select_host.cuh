// kg_select_host.cuh -- host side of the device-resident BestAssociationsHeap set (kg_select.cuh) and the C ABI's
// kg_select_* entry points.  Included by kg_abi.cu after kg_tc.cuh.
#pragma once
#include "kg_filter_retune.cuh"
#include "kg_select.cuh"

static void kg_sel_free(KgSelState *s) {
	cudaFree(s->d_kbest); cudaFree(s->d_hsize); cudaFree(s->d_hslot); cudaFree(s->d_cand_count); cudaFree(s->d_order);
	cudaFree(s->d_log_count); cudaFree(s->d_hscore); cudaFree(s->d_floor); cudaFree(s->d_pay_kmer); cudaFree(s->d_pay_row);
	cudaFree(s->d_hstat); cudaFree(s->d_sort_buf); cudaFree(s->d_status); cudaFree(s->d_digest); cudaFree(s->d_cand); cudaFree(s->d_log);
	if (s->h_status) cudaFreeHost(s->h_status);
	const double growth = s->growth;
	const uint64_t max_round = s->max_round, cc = s->cand_cap_opt, lc = s->log_cap_opt;
	*s = KgSelState();
	s->growth = growth; s->max_round = max_round; s->cand_cap_opt = cc; s->log_cap_opt = lc;   // options survive
}

static KgSelectParams kg_sel_params(kg_ctx *c) {
	KgSelState &s = c->sel;
	KgSelectParams p;
	memset(&p, 0, sizeof p);
	p.n_pheno = s.n_pheno;
	p.kmax = s.kmax;
	p.kbest = s.d_kbest;
	p.h_score = s.d_hscore;
	p.h_slot = s.d_hslot;
	p.pay_kmer = s.d_pay_kmer;
	p.pay_row = s.d_pay_row;
	p.h_size = s.d_hsize;
	p.h_stat = s.d_hstat;
	p.cand = s.d_cand;
	p.cand_count = s.d_cand_count;
	p.cand_cap = s.cand_cap;
	p.order = s.d_order;
	p.sort_buf = s.d_sort_buf;
	p.sort_stride = s.sort_stride;
	p.sort_smem = s.sort_smem;
	p.status = s.d_status;
	p.thr = c->d_thr;
	p.floor_thr = s.floor_set ? s.d_floor : nullptr;
	if (s.log_enabled) { p.log = s.d_log; p.log_count = s.d_log_count; p.log_cap = s.log_cap; }
	return p;
}

template <typename T>
static kg_status kg_sel_alloc(kg_ctx *c, T **p, size_t n, bool zero, const char *what) {
	cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc %s (%zu bytes): %s", what, n * sizeof(T), cudaGetErrorString(e));
	if (zero) KG_CUDA(c, cudaMemsetAsync(*p, 0, std::max<size_t>(n, 1) * sizeof(T), c->stream));
	return KG_OK;
}
#define KG_SEL_ALLOC(ptr, n, zero, what)                                     \
	do {                                                                     \
		kg_status st_ = kg_sel_alloc(c, &(ptr), (size_t)(n), (zero), (what)); \
		if (st_ != KG_OK) return st_;                                        \
	} while (0)

extern "C" kg_status kg_select_end(kg_ctx *c) {
	if (!c) return KG_ERR_INVALID;
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	kg_sel_free(&c->sel);
	return KG_OK;
}

extern "C" kg_status kg_select_begin(kg_ctx *c, const uint64_t *k_best, uint32_t n_pheno, uint32_t flags) {
	if (!c || !k_best) return KG_ERR_INVALID;
	if (!c->d_y_lane || n_pheno != c->n_pheno) KG_FAIL(c, KG_ERR_STATE, "kg_select_begin: call kg_scan_set_phenotypes with the same number of phenotypes first");
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	kg_sel_free(&c->sel);
	KgSelState &s = c->sel;
	s.n_pheno = n_pheno;
	s.kbest.resize(n_pheno);
	uint64_t kmax = 1;
	for (uint32_t p = 0; p < n_pheno; p++) {
		if (k_best[p] < 1 || k_best[p] > 0x7FFFFFFFull) KG_FAIL(c, KG_ERR_INVALID, "kg_select_begin: k_best[%u] = %llu out of range", p, (unsigned long long)k_best[p]);
		s.kbest[p] = (uint32_t)k_best[p];
		kmax = std::max<uint64_t>(kmax, k_best[p]);
	}
	s.kmax = (uint32_t)kmax;
	// shared memory: the heap (12 bytes per entry) next to the sort / staging scratch
	int max_smem = 0;
	KG_CUDA(c, cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
	s.sort_smem = 8192;
	while (s.sort_smem > 1024 && kg_select_smem_bytes(s.kmax, s.sort_smem) > (size_t)max_smem) s.sort_smem >>= 1;
	s.smem = kg_select_smem_bytes(s.kmax, s.sort_smem);
	if (s.smem > (size_t)max_smem) {
		kg_sel_free(&s);
		KG_FAIL(c, KG_ERR_INVALID, "kg_select_begin: a heap of %llu entries does not fit shared memory (use the host replay path)", (unsigned long long)kmax);
	}
	s.cand_cap = (uint32_t)std::min<uint64_t>(s.cand_cap_opt ? s.cand_cap_opt : std::max<uint64_t>(4 * kmax, 65536), 1u << 24);
	s.sort_stride = 2;
	while (s.sort_stride < s.cand_cap) s.sort_stride <<= 1;
	s.log_enabled = (flags & KG_SELECT_LOG) != 0;
	s.log_cap = s.log_enabled ? (uint32_t)std::min<uint64_t>(s.log_cap_opt ? s.log_cap_opt : 32 * kmax + 65536, 1u << 26) : 0;
	s.fill_rows = kmax + kmax / 4 + 64;
	s.rows_submitted = 0;
	const size_t P = n_pheno;
	KG_SEL_ALLOC(s.d_kbest, P, false, "heap capacities");
	KG_CUDA(c, cudaMemcpyAsync(s.d_kbest, s.kbest.data(), P * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
	KG_SEL_ALLOC(s.d_hsize, P, true, "heap sizes");
	KG_SEL_ALLOC(s.d_hslot, P * s.kmax, false, "heap slots");
	KG_SEL_ALLOC(s.d_hscore, P * s.kmax, true, "heap scores");
	KG_SEL_ALLOC(s.d_pay_kmer, P * s.kmax, false, "heap k-mers");
	KG_SEL_ALLOC(s.d_pay_row, P * s.kmax, false, "heap rows");
	KG_SEL_ALLOC(s.d_hstat, 2 * P, true, "heap counters");
	KG_SEL_ALLOC(s.d_cand_count, P, true, "candidate counters");
	KG_SEL_ALLOC(s.d_cand, P * s.cand_cap, false, "candidate segments");
	KG_SEL_ALLOC(s.d_order, P * s.cand_cap, false, "candidate order");
	KG_SEL_ALLOC(s.d_sort_buf, P * s.sort_stride, false, "sort scratch");
	KG_SEL_ALLOC(s.d_status, KG_SEL_ST_WORDS, true, "selection status");
	KG_SEL_ALLOC(s.d_digest, 1, true, "digest");
	KG_SEL_ALLOC(s.d_floor, P, false, "threshold floors");
	if (s.log_enabled) {
		KG_SEL_ALLOC(s.d_log, P * s.log_cap, false, "admission log");
		KG_SEL_ALLOC(s.d_log_count, P, true, "admission log counters");
	}
	KG_CUDA(c, cudaMallocHost((void **)&s.h_status, KG_SEL_ST_WORDS * sizeof(unsigned long long)));
	// thresholds: every heap is empty
	c->h_thr.assign(c->p_alloc, -1.0);
	KG_CUDA(c, cudaMemcpyAsync(c->d_thr, c->h_thr.data(), (size_t)c->p_alloc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	KG_CUDA(c, cudaFuncSetAttribute(kg_select_replay_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem));
	KG_CUDA(c, cudaFuncSetAttribute(kg_select_replay_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	s.active = true;
	return KG_OK;
}

// after a round's scan kernels: bookkeeping, heap replay, filter constants for the next round
static kg_status kg_sel_finish_round(kg_ctx *c, uint64_t n_rows, uint64_t first_row_id, bool filter_counters) {
	KgSelState &s = c->sel;
	timing_begin(c, KG_KERNEL_SCAN_SELECT, n_rows);
	kg_select_round_end_kernel<<<1, 128, 0, c->stream>>>(s.d_status, s.d_cand_count, s.n_pheno, s.cand_cap, n_rows, first_row_id,
	                                                     c->d_counters, filter_counters ? c->tc.d_group_count : nullptr, c->tc.p_pad / 16);
	KG_LAUNCH_CHECK(c);
	KgSelectParams prm = kg_sel_params(c);
	prm.first_row = first_row_id;
	kg_select_replay_kernel<false><<<s.n_pheno, KG_SEL_THREADS, s.smem, c->stream>>>(prm);
	KG_LAUNCH_CHECK(c);
	kg_status st = kg_tc_retune(c, false);
	timing_end(c);
	return st;
}

// One round of rows (device memory) through the scan kernels and the heaps.
static kg_status kg_sel_round(kg_ctx *c, const uint64_t *dev, uint64_t n_rows, uint64_t first_row_id) {
	KgSelState &s = c->sel;
	const bool cold = s.rows_submitted < s.fill_rows;
	bool use_tc = !cold && c->scan_engine != 1 && kg_tc_scan_available(c);
	if (!cold && c->scan_engine == 2 && !kg_tc_scan_available(c))
		KG_FAIL(c, KG_ERR_INVALID, "tensor filter engine unavailable for this shape: %s", c->tc.why_unavailable.c_str());
	kg_status st;
	if (use_tc) {
		st = kg_tc_scan_tile(c, dev, n_rows, first_row_id);   // ends with kg_sel_finish_round
		if (st != KG_OK) return st;
	} else {
		KgRowView view;
		st = memory_view(c, dev, n_rows, &view);
		if (st != KG_OK) return st;
		KgScanParams prm = scan_params(c, view, first_row_id);
		st = launch_exact_pt<0>(c, prm);
		if (st != KG_OK) return st;
		st = kg_sel_finish_round(c, n_rows, first_row_id, false);
		if (st != KG_OK) return st;
	}
	s.rows_submitted += n_rows;
	return KG_OK;
}

// cut a tile into rounds: [fill] then geometric growth, so that the expected candidates per phenotype and round stay
// near growth x K whatever the scan position (a row at position m is admitted with probability ~ K / m)
static uint64_t kg_sel_next_round(const KgSelState &s, uint64_t left) {
	uint64_t len;
	if (s.rows_submitted < s.fill_rows) len = std::min<uint64_t>(s.fill_rows - s.rows_submitted, s.cand_cap);   // every kept row is a candidate
	else len = std::min<uint64_t>(std::max<uint64_t>((uint64_t)(s.growth * (double)s.rows_submitted), s.min_round), s.max_round);
	// Whole 128-row blocks (unless a cap forbids it): a round that starts at an odd row of a table with an odd number of
	// words per row is not 16-byte aligned, and the filter then scans a device-to-device COPY of the round
	// (kg_tc_aligned_tile) -- half of the cold phase's rounds did
	const uint64_t up = (len + 127) & ~(uint64_t)127;
	if (up <= s.max_round && (s.rows_submitted >= s.fill_rows || up <= s.cand_cap)) len = up;
	return std::max<uint64_t>(1, std::min(len, left));
}

static kg_status kg_sel_submit(kg_ctx *c, const uint64_t *rows, uint64_t n_rows, uint64_t first_row_id) {
	const size_t stride = (size_t)c->w_file + 1;
	if (is_device_pointer(rows)) {
		c->cur_slot = -1;
		if (c->pat_attached) {
			kg_status st = patterns_attached_tile(c, rows, n_rows);
			if (st != KG_OK) return st;
		}
		for (uint64_t off = 0; off < n_rows;) {
			const uint64_t n = kg_sel_next_round(c->sel, n_rows - off);
			kg_status st = kg_sel_round(c, rows + off * stride, n, first_row_id + off);
			if (st != KG_OK) return st;
			off += n;
		}
		return KG_OK;
	}
	// host rows: sub-tiles through the two device slots (H2D of sub-tile i + 1 under the kernels of sub-tile i)
	for (uint64_t off = 0; off < n_rows; off += kHostSubTileRows) {
		const uint64_t nt = std::min<uint64_t>(kHostSubTileRows, n_rows - off);
		const uint64_t *dev = nullptr;
		kg_status st = acquire_tile(c, rows + off * stride, nt, &dev);
		if (st != KG_OK) return st;
		if (c->pat_attached) {
			st = patterns_attached_tile(c, dev, nt);
			if (st != KG_OK) return st;
		}
		for (uint64_t o2 = 0; o2 < nt;) {
			const uint64_t n = kg_sel_next_round(c->sel, nt - o2);
			st = kg_sel_round(c, dev + o2 * stride, n, first_row_id + off + o2);
			if (st != KG_OK) return st;
			o2 += n;
		}
		st = release_tile(c);
		if (st != KG_OK) return st;
	}
	return KG_OK;
}

static kg_status kg_sel_read_status(kg_ctx *c) {
	KgSelState &s = c->sel;
	KG_CUDA(c, cudaStreamSynchronize(c->copy_stream));
	KG_CUDA(c, cudaMemcpyAsync(s.h_status, s.d_status, KG_SEL_ST_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	timing_resolve(c);
	return KG_OK;
}

extern "C" kg_status kg_select_sync(kg_ctx *c, uint64_t *rows_applied, uint64_t *rows_kept) {
	if (!c) return KG_ERR_INVALID;
	if (!c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_select_sync: call kg_select_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	kg_status st = kg_sel_read_status(c);
	if (st != KG_OK) return st;
	KgSelState &s = c->sel;
	if (rows_applied) *rows_applied = s.h_status[KG_SEL_ST_ROWS_APPLIED];
	if (rows_kept) *rows_kept = s.h_status[KG_SEL_ST_KEPT];
	if (s.h_status[KG_SEL_ST_LOG_OVERFLOW])
		KG_FAIL(c, KG_ERR_NOMEM, "kg_select: the admission log of a phenotype overflowed (%u entries): raise KG_OPT_SELECT_LOG_CAP", s.log_cap);
	if (s.h_status[KG_SEL_ST_POISON]) {
		// recovery: forget the rounds that were not applied, shorten the rounds, clear the poison
		const unsigned long long fail_row = s.h_status[KG_SEL_ST_FAIL_ROW];
		s.rows_submitted = s.h_status[KG_SEL_ST_ROWS_APPLIED];
		s.growth = std::max(s.growth / 4.0, 1e-4);
		s.min_round = std::max<uint64_t>(s.min_round / 4, std::min<uint64_t>(64, s.cand_cap));   // a round of <= cand_cap rows cannot overflow
		KG_CUDA(c, cudaMemsetAsync(s.d_status + KG_SEL_ST_POISON, 0, sizeof(unsigned long long), c->stream));
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		KG_FAIL(c, KG_ERR_HITS_OVERFLOW, "kg_select: the round starting at row id %llu overflowed a candidate segment (%u per phenotype); "
		        "%llu rows are applied, resubmit the rest", fail_row, s.cand_cap, (unsigned long long)s.rows_submitted);
	}
	return KG_OK;
}

extern "C" kg_status kg_select_stats(kg_ctx *c, uint64_t *rounds, uint64_t *candidates, uint64_t *admitted, uint64_t *reorders) {
	if (!c) return KG_ERR_INVALID;
	if (!c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_select_stats: call kg_select_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	kg_status st = kg_sel_read_status(c);
	if (st != KG_OK) return st;
	KgSelState &s = c->sel;
	if (rounds) *rounds = s.h_status[KG_SEL_ST_ROUNDS];
	if (candidates) *candidates = s.h_status[KG_SEL_ST_CANDS];
	if (reorders) *reorders = s.h_status[KG_SEL_ST_REORDERS];
	if (admitted) {
		std::vector<unsigned long long> hs(2 * (size_t)s.n_pheno);
		KG_CUDA(c, cudaMemcpy(hs.data(), s.d_hstat, hs.size() * 8, cudaMemcpyDeviceToHost));
		unsigned long long t = 0;
		for (uint32_t p = 0; p < s.n_pheno; p++) t += hs[2 * p];
		*admitted = t;
	}
	return KG_OK;
}

extern "C" size_t kg_select_state_len(const kg_ctx *c) { return c && c->sel.active ? kg_select_state_words(c->sel.n_pheno, c->sel.kmax) : 0; }
extern "C" uint32_t kg_select_kmax(const kg_ctx *c) { return c && c->sel.active ? c->sel.kmax : 0; }

// device scratch big enough for `bytes`, reusing the context's squeeze scratch (idle between rounds of a synchronised stream)
static kg_status kg_sel_scratch(kg_ctx *c, size_t bytes, void **out) {
	if (c->squeezed_cap < bytes) {
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		if (c->d_squeezed) KG_CUDA(c, cudaFree(c->d_squeezed));
		c->d_squeezed = nullptr;
		c->squeezed_cap = 0;
		cudaError_t me = cudaMalloc((void **)&c->d_squeezed, bytes);
		if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc(%zu) scratch: %s", bytes, cudaGetErrorString(me));
		c->squeezed_cap = bytes;
	}
	*out = c->d_squeezed;
	return KG_OK;
}

extern "C" kg_status kg_select_export(kg_ctx *c, uint64_t *state) {
	if (!c || !state) return KG_ERR_INVALID;
	if (!c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_select_export: call kg_select_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	KgSelState &s = c->sel;
	const size_t words = kg_select_state_words(s.n_pheno, s.kmax);
	unsigned long long *dst = reinterpret_cast<unsigned long long *>(state);
	const bool dev = is_device_pointer(state);
	if (!dev) {
		KG_CUDA(c, cudaStreamSynchronize(c->stream));   // the scratch below doubles as the squeeze buffer of queued rounds
		void *tmp = nullptr;
		kg_status st = kg_sel_scratch(c, words * 8, &tmp);
		if (st != KG_OK) return st;
		dst = static_cast<unsigned long long *>(tmp);
	}
	const KgSelectParams prm = kg_sel_params(c);
	dim3 grid(std::max(1u, std::min((s.kmax + 255) / 256, 64u)), s.n_pheno);
	kg_select_export_kernel<<<grid, 256, 0, c->stream>>>(prm, dst);
	KG_LAUNCH_CHECK(c);
	if (!dev) {
		KG_CUDA(c, cudaMemcpyAsync(state, dst, words * 8, cudaMemcpyDeviceToHost, c->stream));
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	return KG_OK;
}

extern "C" kg_status kg_select_import(kg_ctx *c, const uint64_t *state, uint64_t rows_applied, uint64_t rows_kept) {
	if (!c || !state) return KG_ERR_INVALID;
	if (!c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_select_import: call kg_select_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	KgSelState &s = c->sel;
	const size_t words = kg_select_state_words(s.n_pheno, s.kmax);
	const unsigned long long *src = reinterpret_cast<const unsigned long long *>(state);
	if (!is_device_pointer(state)) {
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		void *tmp = nullptr;
		kg_status st = kg_sel_scratch(c, words * 8, &tmp);
		if (st != KG_OK) return st;
		KG_CUDA(c, cudaMemcpyAsync(tmp, state, words * 8, cudaMemcpyHostToDevice, c->stream));
		src = static_cast<const unsigned long long *>(tmp);
	}
	const KgSelectParams prm = kg_sel_params(c);
	dim3 grid(std::max(1u, std::min((s.kmax + 255) / 256, 64u)), s.n_pheno);
	kg_select_import_kernel<<<grid, 256, 0, c->stream>>>(prm, src);
	KG_LAUNCH_CHECK(c);
	unsigned long long st_words[KG_SEL_ST_WORDS] = {0};
	st_words[KG_SEL_ST_ROWS_APPLIED] = rows_applied;
	st_words[KG_SEL_ST_KEPT] = rows_kept;
	KG_CUDA(c, cudaMemcpyAsync(s.d_status, st_words, sizeof st_words, cudaMemcpyHostToDevice, c->stream));
	kg_status st = kg_tc_retune(c, true);
	if (st != KG_OK) return st;
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	s.rows_submitted = rows_applied;
	return KG_OK;
}

extern "C" kg_status kg_select_digest(kg_ctx *c, uint64_t *digest) {
	if (!c || !digest) return KG_ERR_INVALID;
	if (!c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_select_digest: call kg_select_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	KgSelState &s = c->sel;
	KG_CUDA(c, cudaMemsetAsync(s.d_digest, 0, 8, c->stream));
	const KgSelectParams prm = kg_sel_params(c);
	dim3 grid(std::max(1u, std::min((s.kmax + 255) / 256, 16u)), s.n_pheno);
	kg_select_digest_kernel<<<grid, 256, 0, c->stream>>>(prm, s.d_digest);
	KG_LAUNCH_CHECK(c);
	unsigned long long h = 0;
	KG_CUDA(c, cudaMemcpyAsync(&h, s.d_digest, 8, cudaMemcpyDeviceToHost, c->stream));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	*digest = h;
	return KG_OK;
}

extern "C" kg_status kg_select_thresholds(kg_ctx *c, double *thr) {
	if (!c || !thr) return KG_ERR_INVALID;
	if (!c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_select_thresholds: call kg_select_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaMemcpyAsync(thr, c->d_thr, (size_t)c->n_pheno * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	return KG_OK;
}

extern "C" kg_status kg_select_log_reset(kg_ctx *c) {
	if (!c) return KG_ERR_INVALID;
	if (!c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_select_log_reset: call kg_select_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	KgSelState &s = c->sel;
	if (s.log_enabled) KG_CUDA(c, cudaMemsetAsync(s.d_log_count, 0, (size_t)s.n_pheno * sizeof(uint32_t), c->stream));
	KG_CUDA(c, cudaMemsetAsync(s.d_status + KG_SEL_ST_KEPT, 0, sizeof(unsigned long long), c->stream));
	return KG_OK;
}

extern "C" kg_status kg_select_log_counts(kg_ctx *c, uint64_t *counts) {
	if (!c || !counts) return KG_ERR_INVALID;
	if (!c->sel.active || !c->sel.log_enabled) KG_FAIL(c, KG_ERR_STATE, "kg_select_log_counts: selection without KG_SELECT_LOG");
	KG_CUDA(c, cudaSetDevice(c->device));
	KgSelState &s = c->sel;
	std::vector<uint32_t> h(s.n_pheno);
	KG_CUDA(c, cudaMemcpyAsync(h.data(), s.d_log_count, (size_t)s.n_pheno * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	for (uint32_t p = 0; p < s.n_pheno; p++) counts[p] = std::min(h[p], s.log_cap);
	return KG_OK;
}

extern "C" kg_status kg_select_log_export(kg_ctx *c, uint64_t *entries, const uint64_t *offsets) {
	if (!c || !offsets) return KG_ERR_INVALID;
	if (!c->sel.active || !c->sel.log_enabled) KG_FAIL(c, KG_ERR_STATE, "kg_select_log_export: selection without KG_SELECT_LOG");
	KG_CUDA(c, cudaSetDevice(c->device));
	KgSelState &s = c->sel;
	const uint64_t total = offsets[s.n_pheno];
	if (total == 0) return KG_OK;
	if (!entries) KG_FAIL(c, KG_ERR_INVALID, "kg_select_log_export: null output");
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	uint64_t *d_off = nullptr;
	KG_CUDA(c, cudaMalloc((void **)&d_off, (size_t)(s.n_pheno + 1) * 8));
	cudaMemcpyAsync(d_off, offsets, (size_t)(s.n_pheno + 1) * 8, cudaMemcpyHostToDevice, c->stream);
	KgCand *dst = reinterpret_cast<KgCand *>(entries);
	const bool dev = is_device_pointer(entries);
	if (!dev) {
		void *tmp = nullptr;
		kg_status st = kg_sel_scratch(c, (size_t)total * sizeof(KgCand), &tmp);
		if (st != KG_OK) { cudaFree(d_off); return st; }
		dst = static_cast<KgCand *>(tmp);
	}
	dim3 grid(32, s.n_pheno);
	kg_select_log_pack_kernel<<<grid, 256, 0, c->stream>>>(s.d_log, s.d_log_count, s.log_cap, d_off, dst);
	c->launches++;
	cudaError_t e0 = cudaGetLastError();
	cudaError_t e1 = cudaSuccess;
	if (!dev) e1 = cudaMemcpyAsync(entries, dst, (size_t)total * sizeof(KgCand), cudaMemcpyDeviceToHost, c->stream);
	cudaError_t e2 = cudaStreamSynchronize(c->stream);
	cudaFree(d_off);
	KG_CUDA(c, e0); KG_CUDA(c, e1); KG_CUDA(c, e2);
	return KG_OK;
}

extern "C" kg_status kg_select_replay(kg_ctx *c, const uint64_t *entries, const uint64_t *offsets, uint64_t rows, uint64_t kept) {
	if (!c || !offsets) return KG_ERR_INVALID;
	if (!c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_select_replay: call kg_select_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	KgSelState &s = c->sel;
	const uint64_t total = offsets[s.n_pheno];
	if (total && !entries) KG_FAIL(c, KG_ERR_INVALID, "kg_select_replay: null entries");
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	uint64_t *d_off = nullptr;
	KG_CUDA(c, cudaMalloc((void **)&d_off, (size_t)(s.n_pheno + 1) * 8));
	cudaMemcpyAsync(d_off, offsets, (size_t)(s.n_pheno + 1) * 8, cudaMemcpyHostToDevice, c->stream);
	const KgCand *src = reinterpret_cast<const KgCand *>(entries);
	if (total && !is_device_pointer(entries)) {
		void *tmp = nullptr;
		kg_status st = kg_sel_scratch(c, (size_t)total * sizeof(KgCand), &tmp);
		if (st != KG_OK) { cudaFree(d_off); return st; }
		cudaMemcpyAsync(tmp, entries, (size_t)total * sizeof(KgCand), cudaMemcpyHostToDevice, c->stream);
		src = static_cast<const KgCand *>(tmp);
	}
	KgSelectParams prm = kg_sel_params(c);
	prm.cand = src;
	prm.cand_off = d_off;
	timing_begin(c, KG_KERNEL_SCAN_SELECT, rows);
	if (total) kg_select_replay_kernel<true><<<s.n_pheno, KG_SEL_THREADS, s.smem, c->stream>>>(prm);
	c->launches++;
	cudaError_t e0 = cudaGetLastError();
	unsigned long long add[2] = {rows, kept};
	// totals: rows_applied / kept live next to each other in the status block
	static_assert(KG_SEL_ST_KEPT == KG_SEL_ST_ROWS_APPLIED + 1, "status layout");
	kg_status st = kg_sel_read_status(c);
	cudaFree(d_off);
	KG_CUDA(c, e0);
	if (st != KG_OK) return st;
	add[0] += s.h_status[KG_SEL_ST_ROWS_APPLIED];
	add[1] += s.h_status[KG_SEL_ST_KEPT];
	KG_CUDA(c, cudaMemcpyAsync(s.d_status + KG_SEL_ST_ROWS_APPLIED, add, sizeof add, cudaMemcpyHostToDevice, c->stream));
	st = kg_tc_retune(c, false);
	timing_end(c);
	if (st != KG_OK) return st;
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	s.rows_submitted += rows;
	return KG_OK;
}

extern "C" kg_status kg_select_export_scores(kg_ctx *c, uint64_t min_row, double *scores_dev) {
	if (!c || !scores_dev) return KG_ERR_INVALID;
	if (!c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_select_export_scores: call kg_select_begin first");
	if (!is_device_pointer(scores_dev)) KG_FAIL(c, KG_ERR_INVALID, "kg_select_export_scores: the output must be device memory");
	KG_CUDA(c, cudaSetDevice(c->device));
	KgSelState &s = c->sel;
	const KgSelectParams prm = kg_sel_params(c);
	dim3 grid(std::max(1u, std::min((s.kmax + 255) / 256, 64u)), s.n_pheno);
	kg_select_export_scores_kernel<<<grid, 256, 0, c->stream>>>(prm, min_row, scores_dev);
	KG_LAUNCH_CHECK(c);
	return KG_OK;
}

extern "C" kg_status kg_select_set_floor(kg_ctx *c, const double *scores, uint32_t n_heaps) {
	if (!c || !scores || n_heaps == 0) return KG_ERR_INVALID;
	if (!c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_select_set_floor: call kg_select_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	KgSelState &s = c->sel;
	if (!s.floor_set) {
		std::vector<double> m1(s.n_pheno, -1.0);
		KG_CUDA(c, cudaMemcpyAsync(s.d_floor, m1.data(), (size_t)s.n_pheno * 8, cudaMemcpyHostToDevice, c->stream));
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		s.floor_set = true;
	}
	kg_select_floor_kernel<<<s.n_pheno, 256, 0, c->stream>>>(scores, n_heaps, s.n_pheno, s.kmax, s.d_kbest, s.d_floor);
	KG_LAUNCH_CHECK(c);
	return KG_OK;
}
