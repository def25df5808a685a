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
// restated like oracle/oracle.c does for the CPU) -- on the SCORES; a second warp applies the same moves to the slot
// array from the records that thread writes (kg_heap_apply_slots_warp) and stores the payload; the other threads sort
// the candidates by row (bitonic, shared or global memory), stage them, and move the heap between shared and global memory.  The scan kernels feed it through
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
#define KG_SEL_RING 64            // admission records in flight between the sequential thread and the slot warp
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
#define KG_GT(a, b) ((a) > (b))
// ---- libstdc++ heap algorithms on (score, slot) pairs; cmp_second(l, r) = l.score > r.score (min-heap) ------------
// The heap is two arrays, scores hs[] and slots hl[], with hs[1] on a 16-byte boundary and hl[1] / hl[3] on 8- / 16-byte
// ones, so the two children 2h+1, 2h+2 of a node (and its four grandchildren 4h+3 .. 4h+6) are single vector loads.
// The algorithms are written once over a memory policy:
//   KgHeapPtr     plain pointers (tests/heap_host_check.cu runs them on the host against the oracle's priority_queue)
//   KgHeapShared  32-bit shared-window addresses + ld/st.shared: what the replay kernel uses.  With generic pointers into
//                 dynamic shared memory the compiler re-derives the window base (S2R SR_CgaCtaId, ~30 cycles each) several
//                 times per admission, inside the one dependent chain that bounds the kernel.
struct KgHeapPtr {
	double *hs;
	uint32_t *hl;
	__host__ __device__ __forceinline__ double lds(int32_t i) const { return hs[i]; }
	__host__ __device__ __forceinline__ uint32_t ldl(int32_t i) const { return hl[i]; }
	__host__ __device__ __forceinline__ double2 lds2(int32_t i) const { return *reinterpret_cast<const double2 *>(hs + i); }
	__host__ __device__ __forceinline__ uint2 ldl2(int32_t i) const { return *reinterpret_cast<const uint2 *>(hl + i); }
	__host__ __device__ __forceinline__ uint4 ldl4(int32_t i) const { return *reinterpret_cast<const uint4 *>(hl + i); }
	__host__ __device__ __forceinline__ void sts(int32_t i, double v) const { hs[i] = v; }
	__host__ __device__ __forceinline__ void stl(int32_t i, uint32_t v) const { hl[i] = v; }
};
#ifdef __CUDACC__
struct KgHeapShared {
	uint32_t hs_a, hl_a;   // shared-window byte addresses of hs[0], hl[0]
	__device__ __forceinline__ KgHeapShared(const double *hs, const uint32_t *hl) {
		hs_a = (uint32_t)__cvta_generic_to_shared(hs);
		hl_a = (uint32_t)__cvta_generic_to_shared(hl);
		asm volatile("" : "+r"(hs_a), "+r"(hl_a));   // opaque: kept in registers, never re-derived
	}
	__device__ __forceinline__ double lds(int32_t i) const {
		double v;
		asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(hs_a + 8u * (uint32_t)i));
		return v;
	}
	__device__ __forceinline__ uint32_t ldl(int32_t i) const {
		uint32_t v;
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(hl_a + 4u * (uint32_t)i));
		return v;
	}
	__device__ __forceinline__ double2 lds2(int32_t i) const {
		double2 v;
		asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(hs_a + 8u * (uint32_t)i));
		return v;
	}
	__device__ __forceinline__ uint2 ldl2(int32_t i) const {
		uint2 v;
		asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(hl_a + 4u * (uint32_t)i));
		return v;
	}
	__device__ __forceinline__ uint4 ldl4(int32_t i) const {
		uint4 v;
		asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(hl_a + 4u * (uint32_t)i));
		return v;
	}
	__device__ __forceinline__ void sts(int32_t i, double v) const {
		asm volatile("st.shared.f64 [%0], %1;" ::"r"(hs_a + 8u * (uint32_t)i), "d"(v));
	}
	__device__ __forceinline__ void stl(int32_t i, uint32_t v) const {
		asm volatile("st.shared.u32 [%0], %1;" ::"r"(hl_a + 4u * (uint32_t)i), "r"(v));
	}
};
#endif

#ifdef __CUDA_ARCH__
#define KG_HEAP_UNROLL _Pragma("unroll")
#else
#define KG_HEAP_UNROLL
#endif
// __push_heap(first, hole, top = 0, value): while (hole > top && comp(first[parent], value)) move the parent down.
// The ancestors of a position are known in advance (j-th ancestor of 1-based position i: i >> j), so CLIMB levels of them
// are loaded back to back in one shared-memory round trip (nothing written in between touches an ancestor), compared
// independently, and moved down while every comparison below them said "above v".  A newly admitted score is just above
// the heap's minimum and climbs almost to the root (~log2 K levels): CLIMB = 16 covers that in one trip.  Pushes that
// rarely climb (the fill phase, the re-inserted leaf of a pop) use CLIMB = 4.
template <int CLIMB, class M>
__host__ __device__ __forceinline__ void kg_heap_push_up(const M m, int32_t hole, double v, uint32_t vs) {
	while (hole > 0) {
		const uint32_t i1 = (uint32_t)hole + 1u;
		int32_t q[CLIMB];
		double s[CLIMB];
		uint32_t l[CLIMB];
		KG_HEAP_UNROLL
		for (int j = 0; j < CLIMB; j++) {
			const uint32_t a = i1 >> (j + 1);
			q[j] = a ? (int32_t)a - 1 : 0;
		}
		KG_HEAP_UNROLL
		for (int j = 0; j < CLIMB; j++) s[j] = m.lds(q[j]);
		KG_HEAP_UNROLL
		for (int j = 0; j < CLIMB; j++) l[j] = m.ldl(q[j]);
		bool go = true;
		KG_HEAP_UNROLL
		for (int j = 0; j < CLIMB; j++) {
			go = go && (i1 >> (j + 1)) != 0u && KG_GT(s[j], v);
			if (go) {
				m.sts(hole, s[j]);
				m.stl(hole, l[j]);
				hole = q[j];
			}
		}
		if (!go) break;
	}
	m.sts(hole, v);
	m.stl(hole, vs);
}

// pop_heap + pop_back on a heap of `len` entries, then push_back + push_heap of (v_new, slot_new): the reference's
// m_best_kmers.pop(); m_best_kmers.push(new_res) (:53-54).  len is unchanged.
//
// One thread runs this, so the cost is the chain of dependent shared-memory round trips.  __adjust_heap always walks the
// hole down to a leaf choosing the smaller child (ties: the right one), independent of the value that is re-inserted, so
// TWO levels are resolved per round trip: the children (2h+1, 2h+2) and the grandchildren (4h+3 .. 4h+6) of the hole
// are contiguous: five vector loads issued back to back (scores and slots together, so the slot moves never wait on
// their own loads).  The re-inserted value (the old last leaf) then climbs from the leaf the hole reached: its first
// comparison is against the entry just moved into the hole's parent, which is still in a register -- usually the climb
// ends there without another round trip.
// `cap` = entries that may be READ (>= len; reads past the heap's end are clamped to it and their values unused).
template <class M>
__host__ __device__ __forceinline__ void kg_heap_replace_top(const M m, int32_t len, int32_t cap, double v_new, uint32_t slot_new) {
	if (len > 1) {
		// __pop_heap: value = last element, *last = *first (discarded by pop_back), __adjust_heap(first, 0, len - 1, value)
		const double v = m.lds(len - 1);
		const uint32_t vs = m.ldl(len - 1);
		const int32_t n = len - 1;
		const int32_t lim = (n - 1) / 2;      // nodes below lim have both children inside [0, n)
		const int32_t g_max = cap - 4;        // last index a 4-entry grandchild load may start at (cap >= 8, odd start kept below)
		int32_t hole = 0;
		double moved = 0.0;                   // score now at the hole's parent (valid once hole > 0)
		while (hole < lim) {
			const int32_t l = 2 * hole + 1;                           // children l, l + 1
			int32_t g = 2 * l + 1;                                    // grandchildren g .. g + 3 (g = 3 mod 4)
			g = g > g_max ? 3 : g;                                    // out of range: any valid aligned index (values unused)
			const double2 s12 = m.lds2(l);
			const uint2 l12 = m.ldl2(l);
			const double2 s34 = m.lds2(g);
			const double2 s56 = m.lds2(g + 2);
			const uint4 l36 = m.ldl4(g);
			// comp(first[right], first[left]) = right.score > left.score: take the left child, else (ties too) the right one
			const bool left = KG_GT(s12.y, s12.x);
			const int32_t c = left ? l : l + 1;
			moved = left ? s12.x : s12.y;
			m.sts(hole, moved);
			m.stl(hole, left ? l12.x : l12.y);
			hole = c;
			if (!(c < lim)) break;
			const double a = left ? s34.x : s56.x, b = left ? s34.y : s56.y;
			const uint32_t la = left ? l36.x : l36.z, lb = left ? l36.y : l36.w;
			const bool left2 = KG_GT(b, a);
			moved = left2 ? a : b;
			m.sts(c, moved);
			m.stl(c, left2 ? la : lb);
			hole = 2 * c + (left2 ? 1 : 2);
		}
		if ((n & 1) == 0 && hole == (n - 2) / 2) {   // the last inner node has a left child only
			const int32_t ch = 2 * hole + 1;
			moved = m.lds(ch);
			m.sts(hole, moved);
			m.stl(hole, m.ldl(ch));
			hole = ch;
		}
		if (hole > 0 && !KG_GT(moved, v)) {          // __push_heap stops at once: the parent is not above v
			m.sts(hole, v);
			m.stl(hole, vs);
		} else {
			kg_heap_push_up<2>(m, hole, v, vs);
		}
	}
	// push_back at position len - 1, push_heap
	kg_heap_push_up<2>(m, len - 1, v_new, slot_new);
}

// ---- the same algorithms split in two, for the replay kernel --------------------------------------------------------
// Only the SCORES decide where entries move; the slots (and the payload behind them) just follow.  The kernel's one
// sequential thread therefore runs the score half and writes a 16-byte record of what it did; a second warp applies the
// records to the slot array, all levels of a path at once, and stores the payload.  The sequential thread is bound by
// the number of instructions on its dependent chain (~3.3 cycles each, profiles/r02_select_replay_ncu.md), and the slot
// half was ~40 % of them.
struct KgHeapRec {
	uint32_t cand;   // index of the admitted candidate in the staged block
	uint32_t leaf;   // pop: the leaf the hole reached (its ancestors are the sift path)
	uint32_t info;   // bits 0-7 c1: levels the re-inserted last entry climbed from the leaf; 8-15 c2: levels the new entry
	                 // climbed from `pos`; bits 16-17 KG_REC_*; bit 18: the pop sifted (heap had more than one entry)
	uint32_t pos;    // position the new entry was pushed at (len - 1 after a pop, the old size in the fill phase)
};
#define KG_REC_POP 0u
#define KG_REC_PUSH 1u
#define KG_REC_END 2u
#define KG_REC_SIFTED (1u << 18)

// __push_heap on the scores alone; returns the number of levels climbed
template <class M>
__host__ __device__ __forceinline__ uint32_t kg_heap_push_up_scores(const M m, int32_t hole, double v) {
	uint32_t climbed = 0;
	while (hole > 0) {
		const int32_t p1 = (hole - 1) >> 1;
		const int32_t p2 = p1 > 0 ? (p1 - 1) >> 1 : 0;
		const double s1 = m.lds(p1), s2 = m.lds(p2);
		if (!KG_GT(s1, v)) break;
		m.sts(hole, s1); hole = p1; climbed++;
		if (hole == 0 || !KG_GT(s2, v)) break;
		m.sts(hole, s2); hole = p2; climbed++;
	}
	m.sts(hole, v);
	return climbed;
}

// pop + push of kg_heap_replace_top on the scores alone; fills leaf / info of the record
template <class M>
__host__ __device__ __forceinline__ void kg_heap_replace_top_scores(const M m, int32_t len, int32_t cap, double v_new, KgHeapRec &rec) {
	uint32_t c1 = 0, sifted = 0;
	rec.leaf = 0;
	if (len > 1) {
		const double v = m.lds(len - 1);
		const int32_t n = len - 1;
		const int32_t lim = (n - 1) / 2;
		const int32_t g_max = cap - 4;
		int32_t hole = 0;
		double moved = 0.0;
		while (hole < lim) {
			const int32_t l = 2 * hole + 1;
			int32_t g = 2 * l + 1;
			g = g > g_max ? 3 : g;
			const double2 s12 = m.lds2(l);
			const double2 s34 = m.lds2(g);
			const double2 s56 = m.lds2(g + 2);
			const bool left = KG_GT(s12.y, s12.x);
			const int32_t c = left ? l : l + 1;
			moved = left ? s12.x : s12.y;
			m.sts(hole, moved);
			hole = c;
			if (!(c < lim)) break;
			const double a = left ? s34.x : s56.x, b = left ? s34.y : s56.y;
			const bool left2 = KG_GT(b, a);
			moved = left2 ? a : b;
			m.sts(c, moved);
			hole = 2 * c + (left2 ? 1 : 2);
		}
		if ((n & 1) == 0 && hole == (n - 2) / 2) {
			const int32_t ch = 2 * hole + 1;
			moved = m.lds(ch);
			m.sts(hole, moved);
			hole = ch;
		}
		rec.leaf = (uint32_t)hole;
		sifted = KG_REC_SIFTED;
		if (hole > 0 && !KG_GT(moved, v)) m.sts(hole, v);
		else c1 = kg_heap_push_up_scores(m, hole, v);
	}
	const uint32_t c2 = kg_heap_push_up_scores(m, len - 1, v_new);
	rec.pos = (uint32_t)(len - 1);
	rec.info = c1 | (c2 << 8) | (KG_REC_POP << 16) | sifted;
}

#ifdef __CUDACC__
// kg_heap_replace_top_scores for the replay kernel.  Same moves; the sift's dependent chain is cut to
//     ld.shared (children) -> compare -> select -> compare -> select -> ld.shared
// per TWO levels (profiles/r02_select_replay_ncu.md: the generic loop spent half its time on the two data-dependent
// bounds branches and on index -> address arithmetic between the last compare and the next load):
//  - every entry above the last two levels has both children and all four grandchildren, so the first
//    (depth of the last entry - 1) / 2 double steps run in a counted loop with no bounds test at all;
//  - the loop carries the shared-memory ADDRESS of the hole's children pair, A = hs + 8 + 16 h; both candidates for the
//    next A are formed while the second compare is in flight, and the loads take A as it is.
struct KgHeapShape {   // what the sift needs to know about a heap of `len` entries (the same for every admission once it is full)
	int32_t len, lim, iters, edge;
};
__device__ __forceinline__ KgHeapShape kg_heap_shape(int32_t len) {
	KgHeapShape sh;
	const int32_t n = len - 1;                                   // entries during the sift of a pop
	sh.len = len;
	sh.lim = (n - 1) / 2;                                        // entries below lim have both children inside [0, n)
	sh.iters = n >= 1 ? (31 - __clz(n) - 1) >> 1 : 0;            // depth of entry n - 1 is floor(log2 n)
	sh.edge = (n >= 2 && (n & 1) == 0) ? (n - 2) / 2 : -1;       // the last inner entry when it has a left child only
	return sh;
}
__device__ __forceinline__ void kg_heap_replace_top_scores(const KgHeapShared m, const KgHeapShape sh, int32_t cap, double v_new, KgHeapRec &rec) {
	uint32_t c1 = 0, sifted = 0;
	rec.leaf = 0;
	const int32_t len = sh.len;
	if (len > 1) {
		const double v = m.lds(len - 1);
		const int32_t lim = sh.lim;
		const int32_t g_max = cap - 4;
		double moved = 0.0;
		const int32_t iters = sh.iters;
		uint32_t A = m.hs_a + 8u;
		const uint32_t C = 24u - 3u * m.hs_a;
		for (int32_t it = 0; it < iters; it++) {
			const uint32_t G = 2u * A - m.hs_a + 8u;      // hs + 24 + 32 h: grandchildren 4h+3 .. 4h+6
			double2 s12, s34, s56;
			asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(s12.x), "=d"(s12.y) : "r"(A));
			asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(s34.x), "=d"(s34.y) : "r"(G));
			asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(s56.x), "=d"(s56.y) : "r"(G + 16u));
			const bool left = KG_GT(s12.y, s12.x);
			const double m1 = left ? s12.x : s12.y;
			const double a = left ? s34.x : s56.x, b = left ? s34.y : s56.y;
			const bool left2 = KG_GT(b, a);
			moved = left2 ? a : b;
			const uint32_t pre = 4u * A + C + (left ? 0u : 32u);
			asm volatile("st.shared.f64 [%0], %1;" ::"r"((A + m.hs_a - 8u) >> 1), "d"(m1));      // hs[h]
			asm volatile("st.shared.f64 [%0], %1;" ::"r"(A + (left ? 0u : 8u)), "d"(moved));      // hs[2h + 1 + !left]
			A = left2 ? pre : pre + 16u;
		}
		int32_t hole = (int32_t)((A - m.hs_a - 8u) >> 4);
		while (hole < lim) {                                  // the last level or two, with the bounds tests
			const int32_t l = 2 * hole + 1;
			int32_t g = 2 * l + 1;
			g = g > g_max ? 3 : g;
			const double2 s12 = m.lds2(l);
			const double2 s34 = m.lds2(g);
			const double2 s56 = m.lds2(g + 2);
			const bool left = KG_GT(s12.y, s12.x);
			const int32_t c = left ? l : l + 1;
			moved = left ? s12.x : s12.y;
			m.sts(hole, moved);
			hole = c;
			if (!(c < lim)) break;
			const double a = left ? s34.x : s56.x, b = left ? s34.y : s56.y;
			const bool left2 = KG_GT(b, a);
			moved = left2 ? a : b;
			m.sts(c, moved);
			hole = 2 * c + (left2 ? 1 : 2);
		}
		if (hole == sh.edge) {                                // the last inner entry has a left child only
			const int32_t ch = 2 * hole + 1;
			moved = m.lds(ch);
			m.sts(hole, moved);
			hole = ch;
		}
		rec.leaf = (uint32_t)hole;
		sifted = KG_REC_SIFTED;
		if (hole > 0 && !KG_GT(moved, v)) m.sts(hole, v);
		else c1 = kg_heap_push_up_scores(m, hole, v);
	}
	const uint32_t c2 = kg_heap_push_up_scores(m, len - 1, v_new);
	rec.pos = (uint32_t)(len - 1);
	rec.info = c1 | (c2 << 8) | (KG_REC_POP << 16) | sifted;
}
#endif

__host__ __device__ __forceinline__ uint32_t kg_heap_depth(uint32_t i) {   // level of position i (root = 0)
#ifdef __CUDA_ARCH__
	return 31u - (uint32_t)__clz((int)(i + 1u));
#else
	uint32_t d = 0;
	while (((i + 1u) >> (d + 1)) != 0u) d++;
	return d;
#endif
}

// What a record does to the slot array, one move after the other (the definition; tests/heap_host_check.cu runs it
// against the oracle).  Net effect of the sift + the re-inserted entry's climb of c1 levels on the path node(0) = root ..
// node(d) = leaf: entries above level d - c1 move up one level, level d - c1 takes the old last entry, the levels below
// it end up unchanged.  Then the new entry's climb of c2 levels along the ancestors of `pos`.  Returns the new entry's slot.
__host__ __device__ inline uint32_t kg_heap_apply_slots_seq(uint32_t *hl, const KgHeapRec &rec) {
	const uint32_t c1 = rec.info & 0xffu, c2 = (rec.info >> 8) & 0xffu, type = (rec.info >> 16) & 3u;
	uint32_t slot = rec.pos;                              // fill phase: slot = position at admission
	if (type == KG_REC_POP) {
		slot = hl[0];                                     // the evicted entry's slot is reused
		if (rec.info & KG_REC_SIFTED) {
			const uint32_t last = hl[rec.pos];
			const uint32_t d = kg_heap_depth(rec.leaf), i1 = rec.leaf + 1u;
			for (uint32_t j = 0; j + c1 < d; j++) hl[(i1 >> (d - j)) - 1u] = hl[(i1 >> (d - j - 1)) - 1u];
			hl[(i1 >> c1) - 1u] = last;
		}
	}
	const uint32_t q1 = rec.pos + 1u;
	for (uint32_t t = 0; t < c2; t++) hl[(q1 >> t) - 1u] = hl[(q1 >> (t + 1)) - 1u];
	hl[(q1 >> c2) - 1u] = slot;
	return slot;
}

#ifdef __CUDACC__
// The same, by one warp: every level of the path (then of the climb) is one lane.
__device__ __forceinline__ uint32_t kg_heap_apply_slots_warp(uint32_t *hl, const KgHeapRec &rec, uint32_t lane) {
	const uint32_t c1 = rec.info & 0xffu, c2 = (rec.info >> 8) & 0xffu, type = (rec.info >> 16) & 3u;
	uint32_t slot = rec.pos, val = 0;
	if (type == KG_REC_POP) {
		slot = hl[0];
		if (rec.info & KG_REC_SIFTED) {
			const uint32_t last = hl[rec.pos];
			const uint32_t d = kg_heap_depth(rec.leaf), i1 = rec.leaf + 1u, nmove = d - c1;
			if (lane < nmove) val = hl[(i1 >> (d - lane - 1)) - 1u];
			__syncwarp();
			if (lane < nmove) hl[(i1 >> (d - lane)) - 1u] = val;
			else if (lane == nmove) hl[(i1 >> c1) - 1u] = last;
		}
		__syncwarp();
	}
	const uint32_t q1 = rec.pos + 1u;
	if (lane < c2) val = hl[(q1 >> (lane + 1)) - 1u];
	__syncwarp();
	if (lane < c2) hl[(q1 >> lane) - 1u] = val;
	else if (lane == c2) hl[(q1 >> c2) - 1u] = slot;
	__syncwarp();
	return slot;
}
#endif

// pointer forms (host check)
__host__ __device__ __forceinline__ void kg_heap_push_up(double *hs, uint32_t *hl, int32_t hole, double v, uint32_t vs) {
	kg_heap_push_up<4>(KgHeapPtr{hs, hl}, hole, v, vs);
}
__host__ __device__ __forceinline__ void kg_heap_replace_top(double *hs, uint32_t *hl, int32_t len, int32_t cap, double v_new, uint32_t slot_new) {
	kg_heap_replace_top(KgHeapPtr{hs, hl}, len, cap, v_new, slot_new);
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

	// ---- replay: add_association over the candidates, block by block (thread 0: scores; warp 1: slots + payload)
	KgCand *stage = reinterpret_cast<KgCand *>(scratch);
	const KgHeapShared heap(hs, hl);
	uint64_t *pk = prm.pay_kmer + (size_t)p * prm.kmax, *pr = prm.pay_row + (size_t)p * prm.kmax;
	KgCand *log = prm.log ? prm.log + (size_t)p * prm.log_cap : nullptr;
	uint32_t n_log = log ? prm.log_count[p] : 0;
	const double floor_thr = prm.floor_thr ? prm.floor_thr[p] : -1.0;
	unsigned long long pushes = 0, pops = 0;
	__shared__ uint32_t s_warp_cnt[KG_SEL_THREADS / 32];
	__shared__ uint32_t s_m;
	__shared__ uint32_t s_size_now;
	// record ring between the sequential thread (warp 0, lane 0) and the slot warp (warp 1); positions only grow
	__shared__ __align__(16) KgHeapRec s_ring[KG_SEL_RING];
	__shared__ uint32_t s_head, s_tail, s_nlog;
	uint32_t a_ring = (uint32_t)__cvta_generic_to_shared(s_ring), a_head = (uint32_t)__cvta_generic_to_shared(&s_head),
	         a_tail = (uint32_t)__cvta_generic_to_shared(&s_tail);
	asm volatile("" : "+r"(a_ring), "+r"(a_head), "+r"(a_tail));   // kept in registers (see KgHeapShared)
	if (threadIdx.x == 0) { s_head = 0; s_tail = 0; s_nlog = n_log; }
	uint32_t ring_pos = 0;                                 // producer: records written; consumer: records applied
	uint32_t tail_seen = 0;                                // producer: the slot warp's position when last read
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
			// ---- the sequential half: add_association on the scores, one record per admission
			const KgHeapShape shape = kg_heap_shape((int32_t)K);   // used once size == K
			double s_next = m ? stage[0].score : 0.0;            // the next candidate's score is fetched one admission ahead
			for (uint32_t i = 0; i <= m; i++) {
				KgHeapRec rec;
				if (i < m) {
					const double s = s_next;
					s_next = stage[i + 1 < m ? i + 1 : i].score;
					if (size < K) {                                  // :45-48 heap not full: push
						if (s <= floor_thr) continue;                // (only in multi-GPU shards > 0; never while exactness matters)
						rec.leaf = 0;
						rec.pos = size;
						rec.info = (kg_heap_push_up_scores(heap, (int32_t)size, s) << 8) | (KG_REC_PUSH << 16);
						size++;
					} else {
						if (!(s > heap.lds(0)) || s <= floor_thr) continue;   // :50 strict '>' against lowest_score = top
						kg_heap_replace_top_scores(heap, shape, kpad, s, rec);
						pops++;
					}
					pushes++;
					rec.cand = i;
				} else {
					rec.cand = rec.leaf = rec.pos = 0;
					rec.info = KG_REC_END << 16;                     // the slot warp leaves its loop for this block
				}
				while (ring_pos - tail_seen >= KG_SEL_RING)          // ring full as far as known: look at the slot warp's position
					asm volatile("ld.relaxed.cta.shared.u32 %0, [%1];" : "=r"(tail_seen) : "r"(a_tail));
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a_ring + 16u * (ring_pos % KG_SEL_RING)), "r"(rec.cand),
				             "r"(rec.leaf), "r"(rec.info), "r"(rec.pos));
				ring_pos++;
				asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a_head), "r"(ring_pos) : "memory");
			}
		} else if ((threadIdx.x >> 5) == 1) {
			// ---- the slot half: apply the records to hl[], store the payload and the admission log
			const uint32_t lane = threadIdx.x & 31;
			uint32_t nl = s_nlog;
			for (bool open = true; open;) {
				uint32_t h;
				do {   // lane 0 polls, the warp stays converged on one value
					h = 0;
					if (lane == 0) asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(h) : "r"(a_head) : "memory");
					h = __shfl_sync(0xffffffffu, h, 0);
					if (h == ring_pos) __nanosleep(32);
				} while (h == ring_pos);
				__syncwarp();   // lane 0's acquire orders the other lanes' reads of the records as well
				for (; ring_pos != h; ring_pos++) {
					const uint4 r4 = *reinterpret_cast<const uint4 *>(&s_ring[ring_pos % KG_SEL_RING]);
					KgHeapRec rec;
					rec.cand = r4.x; rec.leaf = r4.y; rec.info = r4.z; rec.pos = r4.w;
					if (((rec.info >> 16) & 3u) == KG_REC_END) { open = false; ring_pos++; break; }
					const uint32_t slot = kg_heap_apply_slots_warp(hl, rec, lane);
					if (lane == 0) {
						pk[slot] = stage[rec.cand].kmer;
						pr[slot] = stage[rec.cand].row;
					} else if (lane == 1 && log) {
						if (nl < prm.log_cap) log[nl] = stage[rec.cand];
					}
					nl++;
				}
				__syncwarp();
				if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a_tail), "r"(ring_pos) : "memory");
			}
			if (lane == 0) s_nlog = nl;
		}
		__syncthreads();
	}
	n_log = s_nlog;

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
