// kg_scan_filter.cuh -- int8 tensor-core FILTER in front of the exact score kernel (scan engine 2).
//
// With P = 101 phenotype columns the scan is a (rows x N) . (N x P) contraction of presence bits with
// phenotype values (hard part 2).  The reference's result is defined by a float32
// summation order, so the tensor core cannot produce it -- but it can prove, for almost every
// (row, phenotype) pair, that the reference score cannot exceed the heap threshold.  Only the pairs it
// cannot rule out are re-scored in the reference's fp32 order (kg_scan_pair_kernel per (row, phenotype) pair;
// kg_scan_exact_kernel in row-list mode for column groups with long lists), so the reported hits stay bit-identical.
//
// Bound (DESIGN.md section 4 has the derivation).  Per phenotype p, host side (kg_tc.cuh):
//   ybar = sum_ref / N,  c_i = y_i - ybar,  s = max|c_i| / 127,  q_i = rint(c_i / s) in [-127, 127],
//   e_i = c_i - s q_i (|e_i| <= s/2),  e_tot = sum e_i,  A = sum |y_i|,
//   gamma = float32 error of the reference's lane sums.
// Per row (bits S, N1 = |S|, m = min(N1, N - N1), den = N1 (N - N1)) the tensor core gives the EXACT integer
//   Q = sum_{i in S} q_i, and
//   |r_ref| <= N s ( |Q| + F(m) + kappa ),      kappa = (|e_tot| + gamma A + |N ybar - sum_ref|) / s + 1
//   F(m) = largest possible |sum of e_i / s| over m samples = max(sum of the m largest positive, m largest negative
//   errors) <= m / 2;  score_ref = r_ref^2 / den > thr  ==>  |Q| >= alpha sqrt(den) - F(m) - kappa, alpha = sqrt(thr)/(N s)
// with alpha rounded down and sqrt(den) rounded down, so no qualifying pair is ever dropped.
// The phenotype columns are sorted by alpha and tested 16 at a time: max |Q| of the group against the group's
// smallest alpha and largest kappa.  A row with no surviving group is ruled out for every phenotype.
//
// Kernel: persistent, one CTA per SM, 128-row blocks, warp-specialised (23 warps; the numbers below are the split for
// tables wider than 4 presence words, narrow tables trade expander warps for epilogue sets: kg_filter_split_*):
//   warp 0        bulk-async-copies raw 128-row blocks (contiguous 128 * 8(1+W) bytes) into a 4-stage ring
//   warps 1-12    expanders: presence bits -> s8 {0, -1} with 8 PRMTs per 32 bits (byte-permute against constant
//                 tables, ALU pipe only), stored with tcgen05.st straight into TENSOR MEMORY (A operand from TMEM: lane =
//                 row, 4 bytes of K per column).  TMEM holds the two accumulator buffers (2 x P_pad columns) and the A
//                 stages: at N = 1135, P = 101 that is 2 x 112 + 2 x 144 columns = two stages of half a row block each.
//                 A thread expands its words of a stage BEFORE it waits for the stage, so the stage is held only for
//                 the stores.
//   warps 21-22   MMA issuers: tcgen05.mma kind::i8 (M = 128, N = P_pad, K = 32 per instruction, A from TMEM, B from
//                 shared memory) into a double-buffered TMEM accumulator; B (negated quantised phenotypes, P_pad x K_pad
//                 s8, K-major core matrices, no swizzle) stays resident in shared memory.  TWO issuers take turns with
//                 the A-stage batches because a tcgen05.commit stalls its issuing thread for ~670 cycles (measured:
//                 profiles/probes/umma_rate.cu, profiles/r01_umma_probe.md) -- see the role's comment.
//   warps 13-20   epilogue (two sets of 4 warps alternate blocks, one accumulator buffer each): tcgen05.ld of the
//                 128 x P_pad accumulators; column 0 of B is -1 over the used columns, so its accumulator is the row
//                 popcount (MAC filter) for free; per 16-column group max |Q| (3-input min/max) against the group's
//                 loosest bound; (row, group) pairs that survive (rare) are appended to the group's list together with
//                 the row popcount and the group's 16 accumulators, for the per-column re-test (kg_pair_select_kernel)
#pragma once
#include "kg_common.cuh"
#include "kg_tc_ptx.cuh"

#define KG_F_ROWS 128            // rows per block = UMMA M
#define KG_F_MAX_A_STAGES 8      // A stages live in tensor memory next to the two accumulator buffers:
#define KG_F_TMEM_COLS 512       //   2 x p_pad accumulator columns + a_stages x 16 a_words columns <= 512
#define KG_F_RAW_STAGES 16        // most raw row-block stages (as many as fit next to the B image are used).  Narrow rows need MANY:
                                 // a stage is 128 x 8(1+W) bytes (2 KB at W = 1) and ~40 KB per SM must be in flight to cover the
                                 // HBM latency at full bandwidth (4 stages: N = 64 ran at 0.11 of the HBM peak)
#define KG_F_EXPAND_WARP0 1
// Role split = template parameters <NEXP expander warps, NACC accumulator buffers = epilogue sets of 4 warps>, always
// 1 + NEXP + 4 NACC + KG_F_MMA_WARPS = 23 warps.  The epilogue's work per row block does not shrink with the table's width
// (P_pad accumulators per row whatever W is) while the expansion's does, so narrow tables trade expander warps for
// epilogue sets (measured, filter ms per 1.2e9 rows at N = 241: <12,2> 65.5, <8,3> 50.7, <4,4> 61.5; per 2e9 rows at
// N = 64: 93.3 / 70.7 / 59.5):
//   split 0  <12, 2>  W > 4 presence words: the expansion of 128 x 64 W presence bits is the larger job
//   split 1  < 8, 3>  W <= 4
//   split 2  < 4, 4>  W <= 2
// as long as NACC x P_pad accumulator columns + two A stages fit the 512 tensor-memory columns.
#define KG_F_SPLITS 3
__host__ __device__ constexpr int kg_filter_split_nexp(int split) { return split == 0 ? 12 : split == 1 ? 8 : 4; }
__host__ __device__ constexpr int kg_filter_split_nacc(int split) { return split == 0 ? 2 : split == 1 ? 3 : 4; }
__host__ __device__ constexpr int kg_filter_split_max_w(int split) { return split == 0 ? 1 << 30 : split == 1 ? 4 : 2; }
#define KG_F_MAX_ACC 4           // barrier slots for the accumulator buffers
#define KG_F_MAX_WPT 3           // presence words per expander thread and stage: a_words <= KG_F_MAX_WPT * (NEXP / 4)
#ifndef KG_F_MMA_WARPS
#define KG_F_MMA_WARPS 2         // issuers that take turns with the A-stage batches of the MMA stream (see the MMA role below)
#endif
#define KG_F_THREADS ((KG_F_EXPAND_WARP0 + 12 + 4 * 2 + KG_F_MMA_WARPS) * 32)
#define KG_F_NO_Q INT32_MIN      // ent_q marker: this (row, group) entry has no recorded accumulators
#define KG_F_ONE 1               // accumulator units per presence bit: A holds -1 (0xFF, s8), B holds the NEGATED phenotype column
// perf-experiment switches (KgFilterParams::dbg) exist only in builds with -DKG_PERF_SWITCHES; a production build
// compiles every such branch away
#ifdef KG_PERF_SWITCHES
#define KG_F_DBG(prm, bits) (((prm).dbg & (bits)) != 0u)
#else
#define KG_F_DBG(prm, bits) (false)
#endif

// Per 16-column group: the loosest bound of its phenotype columns, in accumulator units (x KG_F_ONE).
//   a row is ruled out for the group iff  max |Q| < alpha * sqrt(den) - kappa - slack(m),
//   slack(m) = min_k (line_a[k] + line_b[k] * m)  >=  max over the group's phenotypes of the largest possible
//   |sum of rounding errors| over any m samples (sorted-prefix sums of the positive / negative errors; the lines
//   are upper tangents of that table, see kg_tc.cuh).
struct KgFilterGroupConst {
	float alpha, kappa;
	float line_a[4], line_b[4];
	float pad_[2];
};

struct KgFilterParams {
	const uint64_t *rows;      // raw tile, 16-byte aligned
	uint64_t n_rows;
	uint32_t w_file;           // presence words per row
	uint32_t a_words;          // u64 presence words per A stage (16 TMEM columns, 2 MMAs of K = 32 each); as many as
	                           // fit: every stage hand-off costs a barrier round trip, so few large stages win
	uint32_t raw_stages;       // raw row-block stages in shared memory, 2 .. KG_F_RAW_STAGES
	uint32_t a_stages;         // 2 .. KG_F_MAX_A_STAGES
	uint32_t nc;               // A stages per row block = ceil(w_file / a_words)
	uint32_t p_pad;            // UMMA N (multiple of 16, <= 256): column 0 = all-ones (row popcount), the phenotypes follow
	                           // sorted by their alpha so that the 16 columns of a group have similar bounds
	uint32_t tcols;            // TMEM columns per accumulator buffer (= p_pad)
	const int8_t *yq_image;    // B operand in its shared-memory byte order, b_bytes long
	uint32_t b_bytes;          // (p_pad / 8) * sbo_b
	uint32_t sbo_b;            // ceil(w_file / 2) * 1024: K_pad = 128 * ceil(w_file / 2) columns
	const KgFilterGroupConst *gconst;   // [p_pad / 16]; groups without phenotype columns hold alpha = +inf (kept for diagnostics)
	const int32_t *thr_tab;    // [p_pad / 16][n_used + 1]: the group's bound as a function of the row popcount n1,
	                           //   alpha * sqrt(n1 (N - n1)) - kappa - slack(min(n1, N - n1)), every step rounded down, as the
	                           //   smallest INTEGER max|Q| that is not ruled out (kg_filter_bound_to_int) -- written by
	                           //   kg_filter_retune_kernel, copied to shared memory at kernel start: the epilogue then pays one
	                           //   load and one integer compare per (row, group) instead of ~15 dependent ALU operations (its
	                           //   warps are bound by dependent-issue latency)
	uint32_t n_used, min_count;
	uint32_t *row_list;        // out: rows of the tile (index inside the tile) that could not be ruled out, any order
	unsigned long long *n_listed;   // device counter for row_list (zeroed before the launch); capacity = n_rows
	uint32_t *group_list;      // out: group_list[g * group_cap + k] = position in row_list of the k-th row whose
	unsigned long long *group_count;   // 16-column group g survived; group_count[g] zeroed before the launch
	uint64_t group_cap;        // = capacity of row_list (n_rows)
	// The first qcap entries of every group list also carry what the per-column test (kg_pair_select_kernel) needs:
	int32_t *ent_q;            // [p_pad / 16][qcap][16] the accumulators of the group's 16 columns (q[0] = KG_F_NO_Q: not recorded)
	uint32_t *ent_n1;          // [p_pad / 16][qcap] row popcount
	uint64_t qcap;
	unsigned long long *kept_count;
	int32_t *q_out;            // debug mode: [n_rows][p_pad] accumulators
	uint32_t n_issuers;        // MMA issuer warps in use, 1 .. KG_F_MMA_WARPS
	uint32_t dbg;              // perf experiments only (env KG_FILTER_DEBUG): 1 skip expansion, 2 skip epilogue work, 4 skip MMAs, 8 skip loads, 64 no tcgen05.st, 128 no expansion arithmetic, 256 epilogue reads 32 columns only
};

__host__ __device__ inline uint32_t kg_filter_raw_stage_bytes(uint32_t w_file) { return KG_F_ROWS * 8u * (w_file + 1); }
// per-group threshold table: thr_tab[g][n1], n1 = 0 .. n_used (see KgFilterParams::thr_tab)
__host__ __device__ inline size_t kg_filter_tab_floats(uint32_t p_pad, uint32_t n_used) { return (size_t)(p_pad / 16) * ((size_t)n_used + 1); }
__host__ __device__ inline size_t kg_filter_smem_bytes(uint32_t w_file, uint32_t b_bytes, uint32_t p_pad, uint32_t raw_stages, uint32_t n_used) {
	return 1024 /*alignment slack*/ + (size_t)b_bytes + (size_t)raw_stages * kg_filter_raw_stage_bytes(w_file) +
	       ((kg_filter_tab_floats(p_pad, n_used) * sizeof(float) + 15) & ~(size_t)15) + 640;
}
// K index (byte inside the A / B operands) of file column `col`.  The expander (below) turns 16 presence bits into
// 4 registers with 4 PRMTs: register b holds the samples 4 n + b (n = byte inside the register), i.e. inside every
// 16 columns the two 2-bit fields of the column index are swapped; B is stored with the same permutation.
__host__ __device__ inline uint32_t kg_filter_k_of_column(uint32_t col) { return (col & ~15u) | ((col & 3u) << 2) | ((col >> 2) & 3u); }

// 32 presence bits -> 32 s8 operand bytes (0xFF = -1 for a set bit, 0x00) in 8 registers: 8 PRMTs + 3 shifts, ALU
// pipe only (the former 64-bit multiply per presence byte paid 2 IMADs + 3 ALU operations per 8 bits).
// PRMT picks result byte n from the 8 bytes {b, a} with selector nibble n (its low 16 bits = 16 presence bits); a
// table whose byte i is 0xFF iff bit j of i is set therefore yields bit j of every nibble as a full byte.  Bit 3 of a
// selector nibble means "replicate the sign of the selected byte", which maps 0xFF / 0x00 onto themselves, so the
// other bits of a nibble never disturb the result; bit 3 itself is read as bit 2 of (x >> 1).
__device__ __forceinline__ void kg_expand_u32(uint32_t x, uint32_t *out8) {
	const uint32_t x1 = x >> 1, y = x >> 16, y1 = x >> 17;   // (IMAD.HI shifts on the FMA pipe measured no faster)
	out8[0] = kg_prmt(0xFF00FF00u, 0xFF00FF00u, x);
	out8[1] = kg_prmt(0xFFFF0000u, 0xFFFF0000u, x);
	out8[2] = kg_prmt(0x00000000u, 0xFFFFFFFFu, x);
	out8[3] = kg_prmt(0x00000000u, 0xFFFFFFFFu, x1);
	out8[4] = kg_prmt(0xFF00FF00u, 0xFF00FF00u, y);
	out8[5] = kg_prmt(0xFFFF0000u, 0xFFFF0000u, y);
	out8[6] = kg_prmt(0x00000000u, 0xFFFFFFFFu, y);
	out8[7] = kg_prmt(0x00000000u, 0xFFFFFFFFu, y1);
}

// alpha * g - kappa - slack(m), every step rounded towards -inf (a lower threshold only lists more rows)
__device__ __forceinline__ float kg_filter_group_threshold(const KgFilterGroupConst &gc, float g, float m) {
	float slack = __fmaf_ru(gc.line_b[0], m, gc.line_a[0]);
#pragma unroll
	for (int k = 1; k < 4; k++) slack = fminf(slack, __fmaf_ru(gc.line_b[k], m, gc.line_a[k]));
	return __fsub_rd(__fmaf_rd(gc.alpha, g, -gc.kappa), slack);
}

// max |v[j]| over 16 int32 accumulators: two reduction trees of 3-input max / min (8 + 8 operations, depth 3)
__device__ __forceinline__ int kg_absmax16(const uint32_t (&v)[16]) {
	const int a0 = __vimax3_s32((int)v[0], (int)v[1], (int)v[2]), a1 = __vimax3_s32((int)v[3], (int)v[4], (int)v[5]);
	const int a2 = __vimax3_s32((int)v[6], (int)v[7], (int)v[8]), a3 = __vimax3_s32((int)v[9], (int)v[10], (int)v[11]);
	const int a4 = __vimax3_s32((int)v[12], (int)v[13], (int)v[14]);
	const int b0 = __vimin3_s32((int)v[0], (int)v[1], (int)v[2]), b1 = __vimin3_s32((int)v[3], (int)v[4], (int)v[5]);
	const int b2 = __vimin3_s32((int)v[6], (int)v[7], (int)v[8]), b3 = __vimin3_s32((int)v[9], (int)v[10], (int)v[11]);
	const int b4 = __vimin3_s32((int)v[12], (int)v[13], (int)v[14]);
	const int mx = __vimax3_s32(__vimax3_s32(a0, a1, a2), __vimax3_s32(a3, a4, (int)v[15]), 0);
	const int mn = __vimin3_s32(__vimin3_s32(b0, b1, b2), __vimin3_s32(b3, b4, (int)v[15]), 0);
	return max(mx, -mn);
}

// the group's bound for a row with n1 set presence bits (what thr_tab[g][n1] holds)
__device__ __forceinline__ float kg_filter_group_threshold_n1(const KgFilterGroupConst &gc, uint32_t n1, uint32_t n_used) {
	const float n1f = (float)n1, n0f = (float)n_used - n1f;
	const float hm = fminf(n1f, n0f);                                                      // m = size of the smaller group
	// <= sqrt(den): the product is rounded down, the root is rounded down, and 0.999999 covers the rest (for N > 4096 the
	// product is no longer exact in fp32; its rounding error is 2^-24, far inside the 1e-6 margin)
	const float g = __fmul_rd(__fsqrt_rd(__fmul_rd(n1f, n0f)), 0.999999f);
	return kg_filter_group_threshold(gc, g, hm);
}

// "a row is ruled out iff (float)max|Q| < thr" as an integer test: listed iff max|Q| >= the value returned.
// max|Q| is an integer below 2^24 (exact in fp32), so (float)q >= thr <=> q >= ceil(thr); NaN (0 * inf: rows with an empty
// group, dropped by the MAC filter anyway) lists the row like the float test did, +inf (group without phenotypes) never does.
__device__ __forceinline__ int32_t kg_filter_bound_to_int(float thr) {
	if (thr != thr) return 0;
	if (!(thr > 0.0f)) return 0;
	if (thr >= 2147483520.0f) return INT32_MAX;
	return (int32_t)ceilf(thr);
}

template <int MODE, int NEXP, int NACC>  // MODE 0 = list candidate rows, 1 = debug: dump accumulators; role split (above)
__global__ void __launch_bounds__(KG_F_THREADS, 1) kg_scan_filter_kernel(const KgFilterParams prm) {
	constexpr uint32_t KG_F_EXPAND_WARPS = NEXP, KG_F_NSUB = NEXP / 4;
	constexpr uint32_t KG_F_EPI_WARP0 = KG_F_EXPAND_WARP0 + NEXP, KG_F_MMA_WARP0 = KG_F_EPI_WARP0 + 4 * NACC;
	static_assert(NEXP % 4 == 0 && NACC >= 2 && NACC <= KG_F_MAX_ACC && (KG_F_MMA_WARP0 + KG_F_MMA_WARPS) * 32 == KG_F_THREADS, "role split");
	extern __shared__ uint8_t kg_f_smem_raw[];
	// carve shared memory (1024-byte aligned base)
	uint8_t *base = reinterpret_cast<uint8_t *>(((uintptr_t)kg_f_smem_raw + 1023) & ~(uintptr_t)1023);
	uint8_t *sB = base;
	const uint32_t raw_stage_bytes = kg_filter_raw_stage_bytes(prm.w_file);
	uint8_t *sRaw = sB + prm.b_bytes;
	int32_t *sTab = reinterpret_cast<int32_t *>(sRaw + prm.raw_stages * raw_stage_bytes);
	const uint32_t tab_floats = (uint32_t)kg_filter_tab_floats(prm.p_pad, prm.n_used);
	uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(sTab) + ((tab_floats * sizeof(float) + 15) & ~(size_t)15));
	uint64_t *raw_full = bars, *raw_empty = bars + KG_F_RAW_STAGES;
	uint64_t *a_full = bars + 2 * KG_F_RAW_STAGES, *a_empty = a_full + KG_F_MAX_A_STAGES;
	uint64_t *tm_full = a_empty + KG_F_MAX_A_STAGES, *tm_empty = tm_full + KG_F_MAX_ACC;
	uint64_t *b_full = tm_empty + KG_F_MAX_ACC;
	uint64_t *turn = b_full + 1;   // [KG_F_MMA_WARPS] issue-order token passed round the MMA issuers
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(turn + KG_F_MMA_WARPS);

	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t n_blocks = (uint32_t)((prm.n_rows + KG_F_ROWS - 1) / KG_F_ROWS);
	const uint32_t row_bytes = 8u * (prm.w_file + 1);

	if (threadIdx.x == 0) {
		for (int i = 0; i < KG_F_RAW_STAGES; i++) { kg_mbar_init(&raw_full[i], 1); kg_mbar_init(&raw_empty[i], KG_F_EXPAND_WARPS); }
		for (int i = 0; i < KG_F_MAX_A_STAGES; i++) { kg_mbar_init(&a_full[i], KG_F_EXPAND_WARPS); kg_mbar_init(&a_empty[i], 1); }
		// tm_full: one tcgen05.commit from each issuer that has MMAs in the block (a commit tracks the issuing thread only)
		for (int i = 0; i < NACC; i++) { kg_mbar_init(&tm_full[i], min(prm.nc, prm.n_issuers)); kg_mbar_init(&tm_empty[i], 4); }
		for (int i = 0; i < KG_F_MMA_WARPS; i++) kg_mbar_init(&turn[i], 1);
		kg_mbar_init(b_full, 1);
		kg_fence_mbar_init();
	}
	for (uint32_t i = threadIdx.x; i < tab_floats; i += blockDim.x) sTab[i] = prm.thr_tab[i];
	if (warp == KG_F_MMA_WARP0) kg_tmem_alloc(tmem_slot, KG_F_TMEM_COLS);
	kg_tc_fence_before();
	__syncthreads();
	kg_tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0) {
		// ===================== producer: B once, then raw row blocks =====================
		if (lane == 0) {
			kg_mbar_arrive_expect_tx(b_full, prm.b_bytes);
			for (uint32_t off = 0; off < prm.b_bytes; off += 32768) {
				const uint32_t n = min(32768u, prm.b_bytes - off);
				kg_bulk_g2s(sB + off, prm.yq_image + off, n, b_full);
			}
			uint32_t it = 0;
			for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, it++) {
				const uint32_t st = it % prm.raw_stages, use = it / prm.raw_stages;
				kg_mbar_wait(&raw_empty[st], (use & 1) ^ 1);
				const uint64_t r0 = (uint64_t)blk * KG_F_ROWS;
				const uint32_t valid = (uint32_t)min((uint64_t)KG_F_ROWS, prm.n_rows - r0);
				const uint32_t bytes = valid * row_bytes;
				const uint32_t bulk = bytes & ~15u;
				const uint8_t *src = reinterpret_cast<const uint8_t *>(prm.rows) + r0 * row_bytes;
				uint8_t *dst = sRaw + st * raw_stage_bytes;
				if (bytes != bulk)  // 8 trailing bytes of a ragged last block
					*reinterpret_cast<uint64_t *>(dst + bulk) = *reinterpret_cast<const uint64_t *>(src + bulk);
				if KG_F_DBG(prm, 8) { kg_mbar_arrive(&raw_full[st]); continue; }
				kg_mbar_arrive_expect_tx(&raw_full[st], bulk);
				if (bulk) kg_bulk_g2s(dst, src, bulk, &raw_full[st]);
			}
		}
	} else if (warp >= KG_F_MMA_WARP0) {
		if (warp - KG_F_MMA_WARP0 < prm.n_issuers) {
		// ===================== MMA issuers =====================
		// Measured on B200 (profiles/probes/umma_rate.cu): a tcgen05.commit holds its issuing thread for ~670 cycles, and
		// with N = 112 (58 cycles per MMA) the few MMAs queued behind it cannot cover that -- one issuer loses ~450
		// tensor cycles per commit (80 instead of 58 cycles per MMA at 18 MMAs per batch).  Two issuers that alternate the
		// batches of the stream hide each other's commit (57 cycles per MMA).  Batch j (A stage j mod a_stages) belongs to
		// issuer j mod n_issuers; a token barrier passed round the issuers orders their barrier waits (below).  Whole warps
		// run the loop (descriptors in uniform registers); one elected lane issues.
		const uint32_t k = warp - KG_F_MMA_WARP0;
		const uint32_t idesc = kg_umma_idesc_i8(KG_F_ROWS, prm.p_pad, true, true, false, false);
		const uint32_t a_tmem0 = tmem_base + NACC * prm.tcols;                          // A stage 0, K offset 0
		const uint32_t a_stage_cols = 16 * prm.a_words;
		const uint64_t b_desc0 = kg_umma_smem_desc(kg_smem_u32(sB), 128, prm.sbo_b);   // column chunk 0
		kg_mbar_wait(b_full, 0);
		const uint32_t words_last = prm.w_file - (prm.nc - 1) * prm.a_words;   // words in the last stage of a row block
		const uint32_t n_it = n_blocks > blockIdx.x ? (n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
		const uint32_t NI = prm.n_issuers;
		const uint32_t tail = min(prm.nc, NI);           // the last `tail` batches of a block (one per issuer) commit tm_full
		uint32_t it = k / prm.nc, c = k % prm.nc;                              // row block (of this CTA) and stage inside it of batch j = k
		uint32_t st = k % prm.a_stages, st_par = (k / prm.a_stages) & 1;       // A stage ring position / parity of batch j
		uint32_t turn_par = 0;
		bool first = k == 0;
		while (it < n_it) {
			const uint32_t buf = it % NACC;
			const uint32_t d_tmem = tmem_base + buf * prm.tcols;
			// The token comes FIRST: once every earlier batch has been issued, the previous user of this A stage (and of
			// this accumulator buffer) has passed its own wait, so the barriers below are in the phase this batch waits
			// for -- an issuer running ahead of the token could otherwise match the parity of a phase two uses back.
			if (!first) {
				kg_mbar_wait(&turn[k], turn_par);
				turn_par ^= 1;
			}
			first = false;
			if (c == 0) {
				kg_mbar_wait(&tm_empty[buf], ((it / NACC) & 1) ^ 1);
				kg_tc_fence_after();
			}
			kg_mbar_wait(&a_full[st], st_par);
			kg_tc_fence_after();
			// Integer accumulation commutes, so the only ordering the MMAs of a block need is that its FIRST one (which
			// overwrites the accumulator) is issued before the others: the issuer of stage 0 passes the token after that
			// MMA, every other batch passes it before issuing anything -- the issue of a batch and the ~670-cycle stall
			// of its commits then overlap with the next issuers' batches.
			const uint32_t at = a_tmem0 + st * a_stage_cols;
			const uint64_t bd = b_desc0 + (uint64_t)c * prm.a_words * 32;
			const uint32_t ksteps = KG_F_DBG(prm, 4) ? 0u : 2 * (c + 1 == prm.nc ? words_last : prm.a_words);
			uint32_t kk0 = 0;
			if (c == 0 && ksteps) {
				// A: 8 TMEM columns per K = 32 step, 16 per presence word.  B descriptor address field is in 16-byte
				// units: one K = 32 step = 16 units, one presence word = 32 units.
				if (kg_elect_one()) kg_umma_i8_ts(d_tmem, at, bd, idesc, 0);
				kk0 = 1;
				__syncwarp();
			}
			if (lane == 0) kg_mbar_arrive(&turn[k + 1 == NI ? 0 : k + 1]);
			__syncwarp();
			if (kg_elect_one()) {
				for (uint32_t kk = kk0; kk < ksteps; kk++) kg_umma_i8_ts(d_tmem, at + kk * 8, bd + kk * 16, idesc, 1);
				kg_umma_commit(&a_empty[st]);
				if (c + tail >= prm.nc) kg_umma_commit(&tm_full[buf]);
			}
			__syncwarp();
			// batch j + KG_F_MMA_WARPS
			c += NI;
			while (c >= prm.nc) { c -= prm.nc; it++; }
			st += NI;
			while (st >= prm.a_stages) { st -= prm.a_stages; st_par ^= 1; }
		}
		}
	} else if (warp < KG_F_EPI_WARP0) {
		// ===================== expanders: presence bits -> u8 A operand in tensor memory =====================
		// thread = row (its TMEM lane) x two u64 words of the stage: 128 operand bytes = 32 TMEM columns
		const uint32_t q4 = warp & 3;                                   // TMEM lane quarter of this warp
		const uint32_t sub = (warp - KG_F_EXPAND_WARP0) >> 2;           // the KG_F_NSUB warps of a lane quarter split a stage's words
		const uint32_t r = q4 * 32 + lane;
		const uint32_t a_stage_cols = 16 * prm.a_words;
		const uint32_t words_last = prm.w_file - (prm.nc - 1) * prm.a_words;
		// my words [lo, hi) of a full stage / of the last (possibly shorter) stage of a row block
		const uint32_t per_f = (prm.a_words + KG_F_NSUB - 1) / KG_F_NSUB, lo_f = min(sub * per_f, prm.a_words), hi_f = min(lo_f + per_f, prm.a_words);
		const uint32_t per_l = (words_last + KG_F_NSUB - 1) / KG_F_NSUB, lo_l = min(sub * per_l, words_last), hi_l = min(lo_l + per_l, words_last);
		const uint32_t a_taddr0 = tmem_base + NACC * prm.tcols + ((q4 * 32u) << 16);
		uint32_t it = 0, st = 0, st_par = 1;                            // a_empty parity: the first pass finds every stage free
		uint32_t rst = 0, r_par = 0;
		for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, it++) {
			kg_mbar_wait(&raw_full[rst], r_par);
			const uint64_t *row = reinterpret_cast<const uint64_t *>(sRaw + rst * raw_stage_bytes + r * row_bytes) + 1;
			for (uint32_t c = 0; c < prm.nc; c++) {
				const bool last = c + 1 == prm.nc;
				const uint32_t lo = last ? lo_l : lo_f, hi = last ? hi_l : hi_f;
				const uint64_t *wp = row + c * prm.a_words;
				// Everything that does not need the stage happens BEFORE the wait: this thread's words of the stage
				// (at most KG_F_MAX_WPT) are loaded and expanded into registers, so that the stage is held only for
				// the tcgen05.st instructions themselves.
				uint32_t v[KG_F_MAX_WPT][16];
				if (!KG_F_DBG(prm, 1)) {
#pragma unroll
					for (uint32_t i = 0; i < KG_F_MAX_WPT; i++) {
						if (lo + i < hi) {
							const uint64_t w = wp[lo + i];
							if KG_F_DBG(prm, 128) {   // perf experiment: no expansion arithmetic, stores only
#pragma unroll
								for (int j = 0; j < 16; j++) v[i][j] = (uint32_t)w;
							} else {
								kg_expand_u32((uint32_t)w, v[i]);
								kg_expand_u32((uint32_t)(w >> 32), v[i] + 8);
							}
						}
					}
				}
				kg_mbar_wait(&a_empty[st], st_par);
				kg_tc_fence_after();
				const uint32_t taddr = a_taddr0 + st * a_stage_cols;
				if (!KG_F_DBG(prm, 1)) {
#pragma unroll
					for (uint32_t i = 0; i < KG_F_MAX_WPT; i++)
						if (lo + i < hi) {
							if KG_F_DBG(prm, 64) {   // perf experiment: expansion arithmetic only, no tensor-memory stores
#pragma unroll
								for (int j = 0; j < 16; j++) asm volatile("" ::"r"(v[i][j]));
							} else {
								kg_tmem_st16(taddr + 16 * (lo + i), v[i]);
							}
						}
					kg_tmem_st_wait();
				}
				kg_tc_fence_before();
				__syncwarp();
				if (lane == 0) kg_mbar_arrive(&a_full[st]);
				if (++st == prm.a_stages) { st = 0; st_par ^= 1; }
			}
			__syncwarp();
			if (lane == 0) kg_mbar_arrive(&raw_empty[rst]);
			if (++rst == prm.raw_stages) { rst = 0; r_par ^= 1; }
		}
	} else {
		// ===================== epilogue: MAC filter + bound test =====================
		const uint32_t q4 = warp & 3;                       // TMEM lane quarter this warp may access
		const uint32_t r = q4 * 32 + lane;                  // row of the block = TMEM lane
		const uint32_t eset = (warp - KG_F_EPI_WARP0) >> 2; // this warp's set: it handles the blocks of accumulator buffer eset
		unsigned long long kept_local = 0;
		for (uint32_t it = eset; (uint64_t)blockIdx.x + (uint64_t)it * gridDim.x < n_blocks; it += NACC) {
			const uint32_t blk = blockIdx.x + it * gridDim.x;
			const uint64_t grow = (uint64_t)blk * KG_F_ROWS + r;
			const uint32_t buf = eset;
			kg_mbar_wait(&tm_full[buf], (it / NACC) & 1);
			kg_tc_fence_after();
			const uint32_t taddr = tmem_base + buf * prm.tcols + ((q4 * 32u) << 16);
			if (MODE == 1) {
				for (uint32_t c0 = 0; c0 < prm.p_pad; c0 += 16) {
					uint32_t v[16];
					kg_tmem_ld16(taddr + c0, v);
					kg_tmem_ld_wait();
					if (grow < prm.n_rows) {
#pragma unroll
						for (int j = 0; j < 16; j++) prm.q_out[grow * prm.p_pad + c0 + j] = (int32_t)v[j];
					}
				}
			} else {
				// row popcount (column 0), then per 16-column group: max |accumulator| against the group's loosest bound
				uint32_t n1 = 0;
				const int32_t *tab = sTab;   // -> thr_tab[.][n1] once the row popcount is known
				const uint32_t tab_stride = prm.n_used + 1;
				uint32_t gmask = 0;   // bit k: group k could not be ruled out for this row
				// (dbg 256: perf experiment, only the first 32 accumulator columns are read back)
				for (uint32_t c0 = 0; c0 < (KG_F_DBG(prm, 2) ? 0u : (KG_F_DBG(prm, 256) ? 32u : prm.p_pad)); c0 += 32) {
					uint32_t v[16], u[16];
					kg_tmem_ld16(taddr + c0, v);
					const bool second = c0 + 16 < prm.p_pad;   // warp-uniform
					if (second) kg_tmem_ld16(taddr + c0 + 16, u);
					kg_tmem_ld_wait();
					if (c0 == 0) {
						n1 = v[0] / KG_F_ONE;
						v[0] = 0;
						tab = sTab + min(n1, prm.n_used);
					}
					// max |accumulator| of each group: 3-input min / max TREES (depth 3 instead of a chain of 8: the epilogue
					// warps are bound by dependent-issue latency, profiles/r02_scan_filter_ncu.md), both groups interleaved
					const int amax_v = kg_absmax16(v);
					if (amax_v >= tab[(c0 >> 4) * tab_stride]) gmask |= 1u << (c0 >> 4);
					if (second) {
						const int amax_u = kg_absmax16(u);
						if (amax_u >= tab[((c0 >> 4) + 1) * tab_stride]) gmask |= 1u << ((c0 >> 4) + 1);
					}
				}
				// load_kmers :121  (popcnt >= mac) && (popcnt <= N - mac)
				const bool keep = grow < prm.n_rows && n1 >= prm.min_count && n1 + prm.min_count <= prm.n_used;
				kept_local += __popc(__ballot_sync(0xffffffffu, keep));
				if (!keep) gmask = 0;
				// The 16 accumulators of a row's FIRST surviving group travel with its list entry (per-column re-test,
				// kg_pair_select_kernel): survivors are rare, so they are re-read from tensor memory here, outside the hot
				// loop, by the warps that have one (tcgen05.ld is warp-wide: one load per group some lane of the warp needs).
				const uint32_t kept_k = gmask ? (uint32_t)__ffs(gmask) - 1u : 0xFFFFFFFFu;
				int32_t keepv[16];
#pragma unroll
				for (int j = 0; j < 16; j++) keepv[j] = 0;
				if (prm.qcap) {
					uint32_t need = __reduce_or_sync(0xffffffffu, gmask ? 1u << kept_k : 0u);
					while (need) {
						const uint32_t k = __ffs(need) - 1;
						need &= need - 1;
						uint32_t w[16];
						kg_tmem_ld16(taddr + 16 * k, w);
						kg_tmem_ld_wait();
						if (k == kept_k) {
#pragma unroll
							for (int j = 0; j < 16; j++) keepv[j] = (int32_t)w[j];
							if (k == 0) keepv[0] = 0;   // column 0 is the popcount column
						}
					}
				}
				// accumulators are in registers: hand the TMEM buffer back to the MMA warp before the bookkeeping
				kg_tc_fence_before();
				__syncwarp();
				if (lane == 0) kg_mbar_arrive(&tm_empty[buf]);
				// rows that could not be ruled out go to the exact kernel, which re-scores them (in the reference's fp32
				// order) against the phenotypes of the surviving groups only; one atomic per warp and list
				const uint32_t mb = __ballot_sync(0xffffffffu, gmask != 0);
				if (mb) {
					unsigned long long basep = 0;
					if (lane == 0) basep = atomicAdd(prm.n_listed, (unsigned long long)__popc(mb));
					basep = __shfl_sync(0xffffffffu, basep, 0);
					const uint32_t pos = (uint32_t)basep + __popc(mb & ((1u << lane) - 1u));
					if (gmask) prm.row_list[pos] = (uint32_t)grow;
					uint32_t gor = __reduce_or_sync(0xffffffffu, gmask);
					while (gor) {
						const uint32_t k = __ffs(gor) - 1;
						gor &= gor - 1;
						const uint32_t gb = __ballot_sync(0xffffffffu, (gmask >> k) & 1u);
						unsigned long long gbase = 0;
						if (lane == 0) gbase = atomicAdd(prm.group_count + k, (unsigned long long)__popc(gb));
						gbase = __shfl_sync(0xffffffffu, gbase, 0);
						if ((gmask >> k) & 1u) {
							const uint64_t idx = gbase + __popc(gb & ((1u << lane) - 1u));
							prm.group_list[(size_t)k * prm.group_cap + idx] = pos;
							if (idx < prm.qcap) {
								prm.ent_n1[(size_t)k * prm.qcap + idx] = n1;
								int4 *q = reinterpret_cast<int4 *>(prm.ent_q + ((size_t)k * prm.qcap + idx) * 16);
								if (k == kept_k) {
#pragma unroll
									for (int j = 0; j < 4; j++) q[j] = make_int4(keepv[4 * j], keepv[4 * j + 1], keepv[4 * j + 2], keepv[4 * j + 3]);
								} else {
									q[0] = make_int4(KG_F_NO_Q, 0, 0, 0);   // a second surviving group of the same row: columns untested
								}
							}
						}
					}
				}
			}
			if (MODE == 1) {
				kg_tc_fence_before();
				__syncwarp();
				if (lane == 0) kg_mbar_arrive(&tm_empty[buf]);
			}
		}
		if (lane == 0 && kept_local) atomicAdd(prm.kept_count, kept_local);
	}

	// teardown: every role has drained its loop; the last tm_full wait of the epilogue implies all MMAs completed
	kg_tc_fence_before();
	__syncthreads();
	if (warp == KG_F_MMA_WARP0) kg_tmem_dealloc(tmem_base, KG_F_TMEM_COLS);
}

// ---- per-column test of the listed (row, group) entries ------------------------------------------------------------
// The filter kernel tests 16 columns at a time against the group's loosest bound.  For the groups whose list is short
// (<= dense_limit entries: the normal case once the heaps are warm) this kernel repeats the test per COLUMN with the
// phenotype's own constants -- alpha_p, kappa_p and the exact slack table F_p(m) instead of the group's tangents -- and
// emits (list position, phenotype) pairs for kg_scan_pair_kernel, which re-scores ONE phenotype per pair in the
// reference's fp32 order instead of two tiles of 8.  Groups with longer lists stay with kg_scan_exact_kernel (list mode).
struct KgPairSelectParams {
	const unsigned long long *group_count;   // [n_groups]
	const uint32_t *group_list;              // [n_groups][group_cap] positions in row_list
	uint64_t group_cap;
	const int32_t *ent_q;
	const uint32_t *ent_n1;
	uint64_t qcap, dense_limit;              // dense_limit <= qcap
	uint32_t n_groups, n_used;
	const int32_t *tile_pheno;               // [16 n_groups] phenotype of every filter column, -1 = none
	const float *alpha, *kappa;              // [P] in accumulator units, rounded like the group constants
	const float *slack;                      // [P][n_used / 2 + 1] F_p(m), rounded up
	uint2 *pairs;                            // out: (position in row_list, phenotype)
	unsigned long long *pair_count;          // zeroed before the launch
	uint64_t pair_cap;
	unsigned long long *overflow;            // set if pairs ran out of room (cannot happen with pair_cap = 16 n_groups dense_limit)
};

__global__ void __launch_bounds__(256) kg_pair_select_kernel(const KgPairSelectParams prm) {
	// work items = the entries of the short lists, back to back: first[k] = items before group k (n_groups <= 16)
	__shared__ uint64_t first[17];
	if (threadIdx.x == 0) {
		uint64_t run = 0;
		for (uint32_t k = 0; k < prm.n_groups; k++) {
			first[k] = run;
			const uint64_t cnt = prm.group_count[k];
			if (cnt <= prm.dense_limit) run += cnt;
		}
		for (uint32_t k = prm.n_groups; k <= 16; k++) first[k] = run;
	}
	__syncthreads();
	const uint64_t total = first[16];
	const uint32_t m_stride = prm.n_used / 2 + 1;
	const float Nf = (float)prm.n_used;
	for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t k = 0;
		while (k + 1 < prm.n_groups && e >= first[k + 1]) k++;
		const uint64_t idx = e - first[k];
		const uint32_t pos = prm.group_list[(size_t)k * prm.group_cap + idx];
		const uint32_t n1 = prm.ent_n1[(size_t)k * prm.qcap + idx];
		const int4 *qp = reinterpret_cast<const int4 *>(prm.ent_q + ((size_t)k * prm.qcap + idx) * 16);
		int32_t q[16];
		const int4 t0 = qp[0];
		const bool untested = t0.x == KG_F_NO_Q;   // only the first 16 bytes of such an entry were written
		q[0] = t0.x; q[1] = t0.y; q[2] = t0.z; q[3] = t0.w;
#pragma unroll
		for (int j = 1; j < 4; j++) {
			const int4 t = untested ? make_int4(0, 0, 0, 0) : qp[j];
			q[4 * j] = t.x; q[4 * j + 1] = t.y; q[4 * j + 2] = t.z; q[4 * j + 3] = t.w;
		}
		const float n1f = (float)n1, n0f = Nf - n1f;
		const uint32_t m = (uint32_t)fminf(n1f, n0f);
		const float g = __fmul_rd(__fsqrt_rd(__fmul_rd(n1f, n0f)), 0.999999f);   // as kg_filter_group_threshold_n1
#pragma unroll
		for (int j = 0; j < 16; j++) {
			const int32_t ph = prm.tile_pheno[16 * k + j];
			if (ph < 0) continue;
			bool list = untested;
			if (!list) {
				const float thr = __fsub_rd(__fmaf_rd(prm.alpha[ph], g, -prm.kappa[ph]), prm.slack[(size_t)ph * m_stride + m]);
				list = !((float)abs(q[j]) < thr);
			}
			if (list) {
				const unsigned long long at = atomicAdd(prm.pair_count, 1ull);
				if (at < prm.pair_cap) prm.pairs[at] = make_uint2(pos, (uint32_t)ph);
				else *prm.overflow = 1ull;
			}
		}
	}
}
