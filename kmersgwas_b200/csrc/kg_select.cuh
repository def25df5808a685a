// kg_select.cuh -- BestAssociationsHeap on the device (SURVEY.md section 8, rows a7 / a8 and the north star's
// "device-side top-K that replaces the CPU heap").
//
// The reference keeps, per phenotype, a std::priority_queue<tuple<kmer, score, row>, vector, cmp_second> capped at K
// (/root/reference/src/best_associations_heap.cpp:43-59, src/kmer_general.h:113-128): push while not full, afterwards
// replace the minimum only on a STRICTLY larger score.  Which of several equal-score minima is evicted, and the pop
// order of equal scores (the rank in the .bim names), is decided by libstdc++'s heap layout, so an
// exact replacement has to perform the same sequence of push_heap / pop_heap element moves.  This file does that on
// the GPU: one CTA per phenotype holds the heap (score + payload slot per position, 12 bytes) in SHARED memory and one
// thread replays the round's candidates in row order with libstdc++'s __push_heap / __adjust_heap (bits/stl_heap.h,
// restated like oracle/oracle.c does for the CPU); the other threads sort the candidates by row (bitonic, shared or
// global memory), stage them, and move the heap between shared and global memory.  The scan kernels feed it through
// per-phenotype candidate segments; thresholds, the tensor filter's bound constants and its column order are
// recomputed on the device after every round (kg_filter_retune_kernel), so the host is not in the scan loop at all.
#pragma once
#include <string.h>

#include "kg_common.cuh"

struct KgCand {      // one candidate association: what add_kmers_to_heap hands to add_association (:281-283)
	uint64_t row;
	uint64_t kmer;
	double score;
};

#define KG_SEL_THREADS 256
#define KG_SEL_STAGE 1024          // candidates staged in shared memory per replay block (24 KB)

// device-resident status words of a selection (kg_ctx::sel.d_status)
enum {
	KG_SEL_ST_POISON = 0,        // != 0: a round overflowed a candidate segment; every later round is ignored
	KG_SEL_ST_ROWS_APPLIED = 1,  // rows of the rounds applied to the heaps so far
	KG_SEL_ST_KEPT = 2,          // rows that passed the MAC filter in those rounds
	KG_SEL_ST_FAIL_ROW = 3,      // first row id of the round that overflowed
	KG_SEL_ST_ROUND_KEPT = 4,    // scratch: kept rows of the round being scanned
	KG_SEL_ST_ROUND_OK = 5,      // scratch: 1 while the replay of the current round may run
	KG_SEL_ST_LOG_OVERFLOW = 6,  // != 0: the admission log of a phenotype ran out of room
	KG_SEL_ST_ROUNDS = 7,        // rounds applied
	KG_SEL_ST_CANDS = 8,         // candidates replayed
	KG_SEL_ST_REORDERS = 9,      // times the filter's column order was rebuilt
	KG_SEL_ST_WORDS = 16
};

struct KgSelectParams {
	uint32_t n_pheno;
	uint32_t kmax;                 // stride of the per-phenotype heap arrays (>= every kbest[p])
	const uint32_t *kbest;         // [P] capacity of heap p (BestAssociationsHeap::m_n_res)
	// the heaps, in libstdc++ layout order (position 0 = top = lowest score)
	double *h_score;               // [P][kmax]
	uint32_t *h_slot;              // [P][kmax] payload slot of the entry at that position
	uint64_t *pay_kmer, *pay_row;  // [P][kmax] by slot
	uint32_t *h_size;              // [P]
	unsigned long long *h_stat;    // [P][2] cnt_push, cnt_pops (plot_stat)
	// candidates of the round
	const KgCand *cand;            // unsorted: [P][cand_cap]; presorted: packed, segment p = [cand_off[p], cand_off[p+1])
	uint32_t *cand_count;          // [P] (unsorted mode; zeroed by the replay)
	const uint64_t *cand_off;      // [P + 1] (presorted mode)
	uint32_t cand_cap;
	uint32_t *order;               // [P][cand_cap] scratch: candidate indices in row order
	unsigned long long *sort_buf;  // [P][sort_stride] scratch for sorts that do not fit shared memory
	uint32_t sort_stride;          // power of two >= cand_cap
	uint32_t sort_smem;            // keys the shared-memory scratch holds (power of two)
	uint64_t first_row;            // id of the round's first row: sort keys are (row - first_row) << 32 | index
	unsigned long long *status;    // KG_SEL_ST_*
	double *thr;                   // [P] out: lowest kept score once the heap is full, else -1
	const double *floor_thr;       // [P] or NULL: candidates with score <= floor are dropped (multi-GPU threshold exchange)
	// admission log (row shards other than the first): every candidate the heap admitted, in row order
	KgCand *log;                   // [P][log_cap] or NULL
	uint32_t *log_count;           // [P]
	uint32_t log_cap;
};

__host__ __device__ __forceinline__ long long kg_dbits(double x) {
#ifdef __CUDA_ARCH__
	return __double_as_longlong(x);
#else
	long long r;
	memcpy(&r, &x, 8);
	return r;
#endif
}
#ifdef KG_SEL_INTCMP
#define KG_GT(a, b) (kg_dbits(a) > kg_dbits(b))
#else
#define KG_GT(a, b) ((a) > (b))
#endif
// ---- libstdc++ heap algorithms on (score, slot) pairs; cmp_second(l, r) = l.score > r.score (min-heap) ------------
// hs has its element 1 on a 16-byte boundary, so the two children 2h+1, 2h+2 of a node are one 128-bit load.
__host__ __device__ __forceinline__ void kg_heap_push_up(double *hs, uint32_t *hl, int32_t hole, double v, uint32_t vs) {
	// __push_heap(first, hole, top = 0, value): while (hole > top && comp(first[parent], value)) move the parent down.
	// A newly admitted score is just above the heap's minimum, so it climbs almost to the root: the ancestors of a
	// position are known in advance, and four levels of them are loaded per shared-memory round trip (nothing written
	// in between touches an ancestor).
	while (hole > 0) {
		const int32_t p1 = (hole - 1) >> 1;
		const int32_t p2 = p1 > 0 ? (p1 - 1) >> 1 : 0;
		const int32_t p3 = p2 > 0 ? (p2 - 1) >> 1 : 0;
		const int32_t p4 = p3 > 0 ? (p3 - 1) >> 1 : 0;
		const double s1 = hs[p1], s2 = hs[p2], s3 = hs[p3], s4 = hs[p4];
		const uint32_t l1 = hl[p1], l2 = hl[p2], l3 = hl[p3], l4 = hl[p4];
		if (!KG_GT(s1, v)) break;
		hs[hole] = s1; hl[hole] = l1; hole = p1;
		if (hole == 0 || !KG_GT(s2, v)) break;
		hs[hole] = s2; hl[hole] = l2; hole = p2;
		if (hole == 0 || !KG_GT(s3, v)) break;
		hs[hole] = s3; hl[hole] = l3; hole = p3;
		if (hole == 0 || !KG_GT(s4, v)) break;
		hs[hole] = s4; hl[hole] = l4; hole = p4;
	}
	hs[hole] = v;
	hl[hole] = vs;
}

// pop_heap + pop_back on a heap of `len` entries, then push_back + push_heap of (v_new, slot_new): the reference's
// m_best_kmers.pop(); m_best_kmers.push(new_res) (:53-54).  len is unchanged.
//
// One thread runs this, so the cost is the chain of dependent shared-memory round trips.  __adjust_heap always walks the
// hole down to a leaf choosing the smaller child (ties: the right one), independent of the value that is re-inserted, so
// TWO levels are resolved per round trip: the children (2h+1, 2h+2) and the grandchildren (4h+3 .. 4h+6) of the hole
// are contiguous, and with &hs[1] / &hs[3] on 16-byte and &hl[1] / &hl[3] on 8- / 16-byte boundaries they are five
// vector loads issued back to back (scores and slots together, so the slot moves never wait on their own loads).
// `cap` = entries that may be READ (>= len; reads past the heap's end are clamped to it and their values unused).
__host__ __device__ __forceinline__ void kg_heap_replace_top(double *hs, uint32_t *hl, int32_t len, int32_t cap, double v_new, uint32_t slot_new) {
	if (len > 1) {
		// __pop_heap: value = last element, *last = *first (discarded by pop_back), __adjust_heap(first, 0, len - 1, value)
		const double v = hs[len - 1];
		const uint32_t vs = hl[len - 1];
		const int32_t n = len - 1;
		const int32_t lim = (n - 1) / 2;      // nodes below lim have both children inside [0, n)
		const int32_t g_max = cap - 4;        // last index a 4-entry grandchild load may start at (cap >= 8, odd start kept below)
		int32_t hole = 0;
		while (hole < lim) {
			const int32_t l = 2 * hole + 1;                           // children l, l + 1
			int32_t g = 2 * l + 1;                                    // grandchildren g .. g + 3 (g = 3 mod 4)
			g = g > g_max ? 3 : g;                                    // out of range: any valid aligned index (values unused)
			const double2 s12 = *reinterpret_cast<const double2 *>(hs + l);
			const uint2 l12 = *reinterpret_cast<const uint2 *>(hl + l);
			const double2 s34 = *reinterpret_cast<const double2 *>(hs + g);
			const double2 s56 = *reinterpret_cast<const double2 *>(hs + g + 2);
			const uint4 l36 = *reinterpret_cast<const uint4 *>(hl + g);
			// comp(first[right], first[left]) = right.score > left.score: take the left child, else (ties too) the right one
			const bool left = KG_GT(s12.y, s12.x);
			const int32_t c = left ? l : l + 1;
			hs[hole] = left ? s12.x : s12.y;
			hl[hole] = left ? l12.x : l12.y;
			hole = c;
			if (!(c < lim)) break;
			const double a = left ? s34.x : s56.x, b = left ? s34.y : s56.y;
			const uint32_t la = left ? l36.x : l36.z, lb = left ? l36.y : l36.w;
			const bool left2 = KG_GT(b, a);
			hs[c] = left2 ? a : b;
			hl[c] = left2 ? la : lb;
			hole = 2 * c + (left2 ? 1 : 2);
		}
		if ((n & 1) == 0 && hole == (n - 2) / 2) {   // the last inner node has a left child only
			const int32_t ch = 2 * hole + 1;
			hs[hole] = hs[ch];
			hl[hole] = hl[ch];
			hole = ch;
		}
		kg_heap_push_up(hs, hl, hole, v, vs);
	}
	// push_back at position len - 1, push_heap
	kg_heap_push_up(hs, hl, len - 1, v_new, slot_new);
}

__device__ __forceinline__ void kg_bitonic_sort_u64(unsigned long long *buf, uint32_t n2) {
	for (uint32_t k = 2; k <= n2; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
				const uint32_t ixj = i ^ j;
				if (ixj > i) {
					const unsigned long long a = buf[i], b = buf[ixj];
					const bool up = (i & k) == 0;
					if ((a > b) == up) { buf[i] = b; buf[ixj] = a; }
				}
			}
			__syncthreads();
		}
}

// shared memory: scores at byte 8 (kmax_pad doubles: &hs[1] and &hs[3] are 16-byte aligned), slots at byte
// 20 + 8 kmax_pad (= 4 mod 16: &hl[1] is 8-byte and &hl[3] 16-byte aligned), then the scratch =
// max(sort_smem * 8, KG_SEL_STAGE * 24).  kmax_pad = kmax rounded up to a multiple of 4, + 8 entries of read slack.
__host__ __device__ inline uint32_t kg_select_kmax_pad(uint32_t kmax) { return ((kmax + 3u) & ~3u) + 8u; }
__host__ __device__ inline size_t kg_select_hl_offset(uint32_t kmax) { return 20 + (size_t)kg_select_kmax_pad(kmax) * 8; }
__host__ __device__ inline size_t kg_select_scratch_offset(uint32_t kmax) { return (kg_select_hl_offset(kmax) + (size_t)kg_select_kmax_pad(kmax) * 4 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t kg_select_smem_bytes(uint32_t kmax, uint32_t sort_smem) {
	const size_t scratch = (size_t)sort_smem * 8 > (size_t)KG_SEL_STAGE * sizeof(KgCand) ? (size_t)sort_smem * 8 : (size_t)KG_SEL_STAGE * sizeof(KgCand);
	return kg_select_scratch_offset(kmax) + scratch;
}

// grid = P, block = KG_SEL_THREADS.  PRESORTED: the candidates are already in row order (merge of shard logs).
template <bool PRESORTED>
__global__ void __launch_bounds__(KG_SEL_THREADS) kg_select_replay_kernel(const KgSelectParams prm) {
	extern __shared__ __align__(16) unsigned char kg_sel_smem[];
	const uint32_t p = blockIdx.x;
	if (!PRESORTED && prm.status[KG_SEL_ST_ROUND_OK] == 0ull) return;
	const int32_t kpad = (int32_t)kg_select_kmax_pad(prm.kmax);
	double *hs = reinterpret_cast<double *>(kg_sel_smem + 8);                       // &hs[1] is 16-byte aligned
	uint32_t *hl = reinterpret_cast<uint32_t *>(kg_sel_smem + kg_select_hl_offset(prm.kmax));
	unsigned char *scratch = kg_sel_smem + kg_select_scratch_offset(prm.kmax);

	uint32_t n;
	const KgCand *cand;
	if (PRESORTED) {
		n = (uint32_t)(prm.cand_off[p + 1] - prm.cand_off[p]);
		cand = prm.cand + prm.cand_off[p];
	} else {
		n = prm.cand_count[p];
		cand = prm.cand + (size_t)p * prm.cand_cap;
	}
	if (n == 0) {   // nothing to replay: only a raised floor can move the threshold
		if (threadIdx.x == 0 && prm.floor_thr && prm.floor_thr[p] > prm.thr[p]) prm.thr[p] = prm.floor_thr[p];
		return;
	}
	const uint32_t K = prm.kbest[p];

	// ---- candidates in row order (rows of one phenotype are distinct, so the order is total)
	uint32_t *order = PRESORTED ? nullptr : prm.order + (size_t)p * prm.cand_cap;
	if (!PRESORTED && n > 1) {
		uint32_t n2 = 2;
		while (n2 < n) n2 <<= 1;
		unsigned long long *buf = n2 <= prm.sort_smem ? reinterpret_cast<unsigned long long *>(scratch)
		                                             : prm.sort_buf + (size_t)p * prm.sort_stride;
		for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x)
			buf[i] = i < n ? (((unsigned long long)(cand[i].row - prm.first_row) << 32) | i) : ~0ull;
		__syncthreads();
		kg_bitonic_sort_u64(buf, n2);
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) order[i] = (uint32_t)buf[i];
		__syncthreads();
	}

	// ---- heap: global -> shared
	uint32_t size = prm.h_size[p];
	{
		const double *gs = prm.h_score + (size_t)p * prm.kmax;
		const uint32_t *gl = prm.h_slot + (size_t)p * prm.kmax;
		for (uint32_t i = threadIdx.x; i < size; i += blockDim.x) { hs[i] = gs[i]; hl[i] = gl[i]; }
	}
	__syncthreads();

	// ---- replay: thread 0 runs add_association over the candidates, block by block
	KgCand *stage = reinterpret_cast<KgCand *>(scratch);
	uint64_t *pk = prm.pay_kmer + (size_t)p * prm.kmax, *pr = prm.pay_row + (size_t)p * prm.kmax;
	KgCand *log = prm.log ? prm.log + (size_t)p * prm.log_cap : nullptr;
	uint32_t n_log = log ? prm.log_count[p] : 0;
	const double floor_thr = prm.floor_thr ? prm.floor_thr[p] : -1.0;
	unsigned long long pushes = 0, pops = 0;
#ifdef KG_SEL_PROFILE
	long long t_loop = 0, t_rep = 0, t_all0 = clock64();
#endif
	__shared__ uint32_t s_warp_cnt[KG_SEL_THREADS / 32];
	__shared__ uint32_t s_m;
	__shared__ uint32_t s_size_now;
	for (uint32_t b0 = 0; b0 < n; b0 += KG_SEL_STAGE) {
		const uint32_t m_in = min((uint32_t)KG_SEL_STAGE, n - b0);
		// Stage the block, dropping IN PARALLEL (order kept) every candidate the sequential loop would reject on its
		// first comparison anyway: the heap's minimum only rises, so a score that is not above the minimum at the
		// start of the block (or above the floor) cannot be admitted later in the block either.
		if (threadIdx.x == 0) { s_m = 0; s_size_now = size; }
		__syncthreads();
		const bool full = s_size_now >= K;
		const double top = full ? hs[0] : 0.0;
		for (uint32_t c0 = 0; c0 < m_in; c0 += KG_SEL_THREADS) {
			const uint32_t i = c0 + threadIdx.x;
			KgCand c;
			bool keep = false;
			if (i < m_in) {
				c = cand[(PRESORTED || n == 1) ? b0 + i : order[b0 + i]];
				keep = (!full || c.score > top) && !(c.score <= floor_thr);
			}
			const uint32_t bal = __ballot_sync(0xffffffffu, keep);
			if ((threadIdx.x & 31) == 0) s_warp_cnt[threadIdx.x >> 5] = __popc(bal);
			__syncthreads();
			uint32_t base = s_m;
			for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) base += s_warp_cnt[w];
			if (keep) stage[base + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u))] = c;
			__syncthreads();
			if (threadIdx.x == 0) {
				uint32_t t = 0;
				for (uint32_t w = 0; w < KG_SEL_THREADS / 32; w++) t += s_warp_cnt[w];
				s_m += t;
			}
			__syncthreads();
		}
		const uint32_t m = s_m;
		if (threadIdx.x == 0) {
#ifdef KG_SEL_PROFILE
			const long long tl0 = clock64();
#endif
			for (uint32_t i = 0; i < m; i++) {
				const double s = stage[i].score;
				uint32_t slot;
				if (size < K) {                                  // :45-48 heap not full: push
					if (s <= floor_thr) continue;                // (only in multi-GPU shards > 0; never while exactness matters)
					slot = size;
					kg_heap_push_up(hs, hl, (int32_t)size, s, slot);
					size++;
				} else {
					if (!(s > hs[0]) || s <= floor_thr) continue;   // :50 strict '>' against lowest_score = top
					slot = hl[0];
#ifdef KG_SEL_PROFILE
					const long long tr0 = clock64();
#endif
					kg_heap_replace_top(hs, hl, (int32_t)size, kpad, s, slot);
#ifdef KG_SEL_PROFILE
					t_rep += clock64() - tr0;
#endif
					pops++;
				}
				pushes++;
				pk[slot] = stage[i].kmer;
				pr[slot] = stage[i].row;
				if (log) {
					if (n_log < prm.log_cap) log[n_log] = stage[i];
					n_log++;
				}
			}
#ifdef KG_SEL_PROFILE
			t_loop += clock64() - tl0;
#endif
		}
		__syncthreads();
	}
#ifdef KG_SEL_PROFILE
	if (threadIdx.x == 0 && p == 0) {
		atomicAdd(prm.status + 10, (unsigned long long)t_loop);
		atomicAdd(prm.status + 11, (unsigned long long)t_rep);
		atomicAdd(prm.status + 12, (unsigned long long)(clock64() - t_all0));
		atomicAdd(prm.status + 13, pops);
	}
#endif

	// ---- heap: shared -> global; threshold for the scan kernels (only thread 0 knows the new size)
	__shared__ uint32_t s_size;
	if (threadIdx.x == 0) s_size = size;
	__syncthreads();
	size = s_size;
	{
		double *gs = prm.h_score + (size_t)p * prm.kmax;
		uint32_t *gl = prm.h_slot + (size_t)p * prm.kmax;
		for (uint32_t i = threadIdx.x; i < size; i += blockDim.x) { gs[i] = hs[i]; gl[i] = hl[i]; }
	}
	if (threadIdx.x == 0) {
		prm.h_size[p] = size;
		double t = size < K ? -1.0 : hs[0];                      // BestAssociationsHeap::device_threshold
		if (floor_thr > t) t = floor_thr;
		prm.thr[p] = t;
		prm.h_stat[2 * p] += pushes;
		prm.h_stat[2 * p + 1] += pops;
		if (!PRESORTED) prm.cand_count[p] = 0;
		if (log) {
			if (n_log > prm.log_cap) { prm.status[KG_SEL_ST_LOG_OVERFLOW] = 1ull; n_log = prm.log_cap; }
			prm.log_count[p] = n_log;
		}
		atomicAdd(prm.status + KG_SEL_ST_CANDS, (unsigned long long)n);
	}
}

// End of a round's scan kernels (one CTA): decides whether the round may be applied.  A candidate segment that
// overflowed poisons the selection: this and every later round are ignored (no heap sees a partial round), the host
// learns about it at the next kg_select_sync and resubmits from KG_SEL_ST_ROWS_APPLIED in smaller rounds.  Also does the
// per-tile bookkeeping of kg_tile_end_kernel for the filter's counters.
__global__ void kg_select_round_end_kernel(unsigned long long *status, uint32_t *cand_count, uint32_t n_pheno, uint32_t cand_cap,
                                           uint64_t round_rows, uint64_t first_row, unsigned long long *interval_cnt,
                                           unsigned long long *tile_cnt, uint32_t n_groups) {
	__shared__ int s_over;
	if (threadIdx.x == 0) s_over = 0;
	__syncthreads();
	for (uint32_t p = threadIdx.x; p < n_pheno; p += blockDim.x)
		if (cand_count[p] > cand_cap) s_over = 1;
	__syncthreads();
	const bool poisoned = status[KG_SEL_ST_POISON] != 0ull;
	const bool ok = !poisoned && !s_over;
	if (!ok)
		for (uint32_t p = threadIdx.x; p < n_pheno; p += blockDim.x) cand_count[p] = 0;
	__syncthreads();
	if (threadIdx.x == 0) {
		if (ok) {
			status[KG_SEL_ST_ROWS_APPLIED] += round_rows;
			status[KG_SEL_ST_KEPT] += status[KG_SEL_ST_ROUND_KEPT];
			status[KG_SEL_ST_ROUNDS] += 1;
		} else if (!poisoned) {
			status[KG_SEL_ST_POISON] = 1ull;
			status[KG_SEL_ST_FAIL_ROW] = first_row;
		}
		status[KG_SEL_ST_ROUND_KEPT] = 0ull;
		status[KG_SEL_ST_ROUND_OK] = ok ? 1ull : 0ull;
		if (tile_cnt) {
			unsigned long long t = 0;
			for (uint32_t i = 0; i < n_groups; i++) t += tile_cnt[i];
			interval_cnt[5] += interval_cnt[2];
			interval_cnt[6] += t;
			interval_cnt[7] += tile_cnt[24];
			interval_cnt[2] = 0;
		}
	}
	__syncthreads();
	if (tile_cnt && threadIdx.x < 32) tile_cnt[threadIdx.x] = 0;
}

// ---- heap state export / import / digest ------------------------------------------------------------------------
// State image (u64 words): [P][4] header {size, cnt_push, cnt_pops, reserved}, then [P][kmax][3] entries
// {kmer, score bits, row} in libstdc++ layout order (position 0 = top).  Pushing the entries of a phenotype in this
// order into an empty std::priority_queue reproduces the layout (every parent <= its child: no element moves).
__host__ __device__ inline size_t kg_select_state_words(uint32_t n_pheno, uint32_t kmax) { return (size_t)n_pheno * 4 + (size_t)n_pheno * kmax * 3; }

__global__ void kg_select_export_kernel(const KgSelectParams prm, unsigned long long *out) {
	const uint32_t p = blockIdx.y;
	const uint32_t size = prm.h_size[p];
	unsigned long long *hdr = out + (size_t)p * 4;
	unsigned long long *ent = out + (size_t)prm.n_pheno * 4 + (size_t)p * prm.kmax * 3;
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		hdr[0] = size; hdr[1] = prm.h_stat[2 * p]; hdr[2] = prm.h_stat[2 * p + 1]; hdr[3] = 0;
	}
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < prm.kmax; i += gridDim.x * blockDim.x) {
		unsigned long long k = 0, s = 0, r = 0;
		if (i < size) {
			const uint32_t slot = prm.h_slot[(size_t)p * prm.kmax + i];
			k = prm.pay_kmer[(size_t)p * prm.kmax + slot];
			r = prm.pay_row[(size_t)p * prm.kmax + slot];
			s = (unsigned long long)__double_as_longlong(prm.h_score[(size_t)p * prm.kmax + i]);
		}
		ent[(size_t)i * 3] = k; ent[(size_t)i * 3 + 1] = s; ent[(size_t)i * 3 + 2] = r;
	}
}

__global__ void kg_select_import_kernel(const KgSelectParams prm, const unsigned long long *in) {
	const uint32_t p = blockIdx.y;
	const unsigned long long *hdr = in + (size_t)p * 4;
	const unsigned long long *ent = in + (size_t)prm.n_pheno * 4 + (size_t)p * prm.kmax * 3;
	const uint32_t size = (uint32_t)min((unsigned long long)prm.kbest[p], hdr[0]);
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		prm.h_size[p] = size;
		prm.h_stat[2 * p] = hdr[1];
		prm.h_stat[2 * p + 1] = hdr[2];
		prm.thr[p] = size < prm.kbest[p] || size == 0 ? -1.0 : __longlong_as_double((long long)ent[1]);
	}
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < size; i += gridDim.x * blockDim.x) {
		prm.h_slot[(size_t)p * prm.kmax + i] = i;
		prm.pay_kmer[(size_t)p * prm.kmax + i] = ent[(size_t)i * 3];
		prm.h_score[(size_t)p * prm.kmax + i] = __longlong_as_double((long long)ent[(size_t)i * 3 + 1]);
		prm.pay_row[(size_t)p * prm.kmax + i] = ent[(size_t)i * 3 + 2];
	}
}

// Order-sensitive 64-bit digest of the heaps (layout order): sum over (p, position, field) of a position-keyed mix.
__global__ void kg_select_digest_kernel(const KgSelectParams prm, unsigned long long *out) {
	const uint32_t p = blockIdx.y;
	const uint32_t size = prm.h_size[p];
	unsigned long long acc = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < size; i += gridDim.x * blockDim.x) {
		const uint32_t slot = prm.h_slot[(size_t)p * prm.kmax + i];
		const unsigned long long key = ((unsigned long long)p << 40) ^ ((unsigned long long)i << 2);
		acc += kg_mix64(prm.pay_kmer[(size_t)p * prm.kmax + slot] ^ kg_mix64(key));
		acc += kg_mix64((unsigned long long)__double_as_longlong(prm.h_score[(size_t)p * prm.kmax + i]) ^ kg_mix64(key + 1));
		acc += kg_mix64(prm.pay_row[(size_t)p * prm.kmax + slot] ^ kg_mix64(key + 2));
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) acc += kg_mix64(((unsigned long long)p << 40) ^ size ^ 0x5151ull);
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// pack the admission logs: out[off[p] + i] = log[p][i]
__global__ void kg_select_log_pack_kernel(const KgCand *log, const uint32_t *log_count, uint32_t log_cap, const uint64_t *off, KgCand *out) {
	const uint32_t p = blockIdx.y;
	const uint32_t n = min(log_count[p], log_cap);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		out[off[p] + i] = log[(size_t)p * log_cap + i];
}

// ---- multi-GPU threshold exchange ---------------------------------------------------------------------------------
// Scores of the heap entries whose row id is >= min_row, position by position ([P][kmax]; -1 elsewhere): what a row
// shard contributes to the exchange -- its OWN rows only, so that the shards' contributions are disjoint row sets.
__global__ void kg_select_export_scores_kernel(const KgSelectParams prm, uint64_t min_row, double *out) {
	const uint32_t p = blockIdx.y;
	const uint32_t size = prm.h_size[p];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < prm.kmax; i += gridDim.x * blockDim.x) {
		double v = -1.0;
		if (i < size) {
			const uint32_t slot = prm.h_slot[(size_t)p * prm.kmax + i];
			if (prm.pay_row[(size_t)p * prm.kmax + slot] >= min_row) v = prm.h_score[(size_t)p * prm.kmax + i];
		}
		out[(size_t)p * prm.kmax + i] = v;
	}
}

// floor[p] = the kbest[p]-th largest score among the union of n_heaps score sets (scores[g][p][kmax], negative = no
// entry), or unchanged when the union holds fewer than kbest[p] scores.  When the sets come from DISJOINT rows that all
// precede the rows a context still has to scan, this is a lower bound of the sequential heap's threshold at every such
// row.  Radix select over the 64-bit patterns of the (non-negative) scores, 8 bits per pass; NaN is ignored.
// grid = P, block = 256.
__global__ void __launch_bounds__(256) kg_select_floor_kernel(const double *scores, uint32_t n_heaps, uint32_t n_pheno,
                                                              uint32_t kmax, const uint32_t *kbest, double *floor_out) {
	__shared__ uint32_t hist[256];
	__shared__ unsigned long long s_prefix;
	__shared__ uint32_t s_want;
	const uint32_t p = blockIdx.x;
	uint32_t want = kbest[p];          // rank (1 = largest) of the score looked for among the elements matching the prefix
	unsigned long long prefix = 0;
	for (int shift = 56; shift >= 0; shift -= 8) {
		for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
		__syncthreads();
		const unsigned long long mask_hi = shift == 56 ? 0ull : (~0ull << (shift + 8));
		for (uint32_t g = 0; g < n_heaps; g++) {
			const double *s = scores + ((size_t)g * n_pheno + p) * kmax;
			for (uint32_t i = threadIdx.x; i < kmax; i += blockDim.x) {
				const double v = s[i];
				if (!(v >= 0.0)) continue;
				const unsigned long long b = (unsigned long long)__double_as_longlong(v);
				if ((b & mask_hi) == prefix) atomicAdd(&hist[(b >> shift) & 255], 1u);
			}
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t run = 0;
			int d = 255;
			for (; d >= 0; d--) {
				if (run + hist[d] >= want) break;
				run += hist[d];
			}
			if (d < 0) { s_want = 0; }   // fewer than kbest scores in total
			else { s_want = want - run; s_prefix = prefix | ((unsigned long long)d << shift); }
		}
		__syncthreads();
		if (s_want == 0) return;
		want = s_want;
		prefix = s_prefix;
		__syncthreads();
	}
	if (threadIdx.x == 0 && __longlong_as_double((long long)prefix) > floor_out[p]) floor_out[p] = __longlong_as_double((long long)prefix);
}
