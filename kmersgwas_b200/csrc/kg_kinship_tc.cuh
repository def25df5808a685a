// kg_kinship_tc.cuh -- kinship Gram accumulation on the int8 tensor cores (tcgen05.mma kind::i8, TMEM int32
// accumulators).  KG_OPT_KINSHIP_ENGINE = 2.
//
// Reference: update_emma_kinshhip_calculation (/root/reference/src/kmers_multiple_databases.cpp:418-438):
//   K[i][j] += 1 ^ g_i ^ g_j (j < i) over the rows load_kmers keeps.  With G = B^T B (co-presence counts over the
//   kept rows, B = rows x samples 0/1), c_i = G[i][i] and M kept rows: K[i][j] = M - c_i - c_j + 2 G[i][j], exact.
// This is the one dense contraction of the hot path: the contraction index is the ROW.
//
// Work split: the lower triangle of G (in FILE column order) is cut into tiles of 128 x 256 samples; a tile
// group = one 128-sample block I with up to two 256-sample blocks J (2 x 256 int32 TMEM columns = the whole
// tensor memory of an SM).  One CTA = (group, split): it streams the 128-row blocks split, split + n_splits, ... of the
// tile (a group gets row splits in proportion to its number of sample tiles) and keeps its accumulators in TMEM for the whole kernel; at the end it adds them to a global
// u64 delta matrix with atomics (int32 is enough inside a launch: a tile has < 2^31 rows).
//
// Per 128-row block:
//   warp 0      bulk-copies the raw rows (contiguous bytes) into a 2-stage ring
//   warps 2-17  expand the presence bits of the needed sample words to s8 bytes 0x00 / 0xFF (= -1; (-1)(-1) = 1) with
//               4 PRMTs per 16 samples into MN-major core matrices (16 samples x 8 rows), 2 stages; rows that fail the MAC
//               filter (keep bits from kg_prefilter_kernel, once per tile) contribute zeros
//   warp 1      one elected thread issues 4 (K = 32 rows each) x up to 2 tcgen05.mma  D_J += A_I^T-view * B_J
//               (both operands MN-major: element (sample, row))
#pragma once
#include "kg_common.cuh"
#include "kg_tc_ptx.cuh"

#define KG_K_ROWS 128
#define KG_K_EXPAND_WARP0 2
#define KG_K_EXPAND_THREADS 512
#define KG_K_EXPAND_SUBS (KG_K_EXPAND_THREADS / KG_K_ROWS)   // threads per row
#define KG_K_THREADS (KG_K_EXPAND_WARP0 * 32 + KG_K_EXPAND_THREADS)
#define KG_K_STAGES 2
#define KG_K_RAW_STAGES 2
#define KG_K_BT_BYTES (KG_K_ROWS * 256)   // one B tile stage: 128 rows x 256 samples
#define KG_K_AT_BYTES (KG_K_ROWS * 128)   // separate A tile stage: 128 rows x 128 samples
#define KG_K_STAGE_BYTES (2 * KG_K_BT_BYTES + KG_K_AT_BYTES)

struct KgKinGroup {
	int32_t i_blk;     // 128-sample block of the A operand (rows of G)
	int32_t j2[2];     // 256-sample blocks of the B operands (columns of G); -1 = unused
	int32_t a_in;      // B tile (0 / 1) that already contains the samples of block i_blk, or -1: expand A separately
};

// one CTA: row blocks split, split + n_splits, ... of tile group `group` (splits are proportional to the group's tiles)
struct KgKinCta {
	uint32_t group, split, n_splits, pad_;
};

struct KgKinTcParams {
	const uint64_t *rows;       // raw tile, 16-byte aligned
	uint64_t n_rows;
	uint32_t w_file;
	const uint32_t *keep_bits;  // [ceil(n_rows / 32)] MAC filter of load_kmers, one bit per row (kg_prefilter_kernel)
	const KgKinGroup *groups;
	const KgKinCta *ctas;       // [gridDim.x]
	unsigned long long *delta;  // [ld][ld] co-presence counts in FILE column order, entries (a, b <= a)
	uint32_t ld;                // 64 * w_file
};

__host__ __device__ inline size_t kg_kin_tc_smem_bytes(uint32_t w_file) {
	return 1024 + (size_t)KG_K_STAGES * KG_K_STAGE_BYTES + (size_t)KG_K_RAW_STAGES * (KG_K_ROWS * 8u * (w_file + 1)) + 256;
}

// 16 presence bits -> 16 operand bytes 0xFF (= -1) / 0x00 with 4 PRMTs (kg_expand_u32 in kg_scan_filter.cuh has the
// derivation): register b holds the samples 4 n + b of the 16 (n = byte inside the register), i.e. operand index o and
// file column are related by swapping the two 2-bit fields of the low nibble.  Both operands of the Gram use the same
// permutation, so it is undone when the accumulators are written out (kg_kin_col_of_operand).
__device__ __forceinline__ uint4 kg_kin_spread16(uint32_t h) {
	uint4 r;
	r.x = kg_prmt(0xFF00FF00u, 0xFF00FF00u, h);
	r.y = kg_prmt(0xFFFF0000u, 0xFFFF0000u, h);
	r.z = kg_prmt(0x00000000u, 0xFFFFFFFFu, h);
	r.w = kg_prmt(0x00000000u, 0xFFFFFFFFu, h >> 1);
	return r;
}
__host__ __device__ inline uint32_t kg_kin_col_of_operand(uint32_t o) { return (o & ~15u) | ((o & 3u) << 2) | ((o >> 2) & 3u); }

__global__ void __launch_bounds__(KG_K_THREADS, 1) kg_kinship_tc_kernel(const KgKinTcParams prm) {
	extern __shared__ uint8_t kg_k_smem_raw[];
	uint8_t *base = reinterpret_cast<uint8_t *>(((uintptr_t)kg_k_smem_raw + 1023) & ~(uintptr_t)1023);
	uint8_t *sStage = base;
	const uint32_t raw_stage_bytes = KG_K_ROWS * 8u * (prm.w_file + 1);
	uint8_t *sRaw = sStage + KG_K_STAGES * KG_K_STAGE_BYTES;
	uint64_t *bars = reinterpret_cast<uint64_t *>(sRaw + KG_K_RAW_STAGES * raw_stage_bytes);
	uint64_t *raw_full = bars, *raw_empty = bars + KG_K_RAW_STAGES;
	uint64_t *st_full = raw_empty + KG_K_RAW_STAGES, *st_empty = st_full + KG_K_STAGES;
	uint64_t *done = st_empty + KG_K_STAGES;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);

	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const KgKinCta me = prm.ctas[blockIdx.x];
	const KgKinGroup grp = prm.groups[me.group];
	const uint32_t n_blocks = (uint32_t)((prm.n_rows + KG_K_ROWS - 1) / KG_K_ROWS);
	const uint32_t row_bytes = 8u * (prm.w_file + 1);
	const bool has2 = grp.j2[1] >= 0;

	if (threadIdx.x == 0) {
		for (int i = 0; i < KG_K_RAW_STAGES; i++) { kg_mbar_init(&raw_full[i], 1); kg_mbar_init(&raw_empty[i], KG_K_EXPAND_THREADS / 32); }
		for (int i = 0; i < KG_K_STAGES; i++) { kg_mbar_init(&st_full[i], KG_K_EXPAND_THREADS / 32); kg_mbar_init(&st_empty[i], 1); }
		kg_mbar_init(done, 1);
		kg_fence_mbar_init();
	}
	if (warp == 1) kg_tmem_alloc(tmem_slot, 512);
	kg_tc_fence_before();
	__syncthreads();
	kg_tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0) {
		// ===================== producer =====================
		if (lane == 0) {
			uint32_t it = 0;
			for (uint32_t blk = me.split; blk < n_blocks; blk += me.n_splits, it++) {
				const uint32_t st = it % KG_K_RAW_STAGES, use = it / KG_K_RAW_STAGES;
				kg_mbar_wait(&raw_empty[st], (use & 1) ^ 1);
				const uint64_t r0 = (uint64_t)blk * KG_K_ROWS;
				const uint32_t valid = (uint32_t)min((uint64_t)KG_K_ROWS, prm.n_rows - r0);
				const uint32_t bytes = valid * row_bytes, bulk = bytes & ~15u;
				const uint8_t *src = reinterpret_cast<const uint8_t *>(prm.rows) + r0 * row_bytes;
				uint8_t *dst = sRaw + st * raw_stage_bytes;
				if (bytes != bulk) *reinterpret_cast<uint64_t *>(dst + bulk) = *reinterpret_cast<const uint64_t *>(src + bulk);
				kg_mbar_arrive_expect_tx(&raw_full[st], bulk);
				if (bulk) kg_bulk_g2s(dst, src, bulk, &raw_full[st]);
			}
		}
	} else if (warp == 1) {
		// ===================== MMA issuer =====================
		const uint32_t idesc = kg_umma_idesc_i8(128, 256, true, true, true, true);
		const uint32_t s_addr = kg_smem_u32(sStage);
		// MN-major, no swizzle: 16 samples contiguous, 8 rows 16 B apart; SBO = next 16 samples (128 B),
		// LBO = next 8 rows (one row of core matrices = tile samples * 8 bytes)
		const uint64_t b_desc0 = kg_umma_smem_desc(s_addr, 2048, 128);
		const uint32_t a_off = grp.a_in >= 0 ? (uint32_t)grp.a_in * KG_K_BT_BYTES + (uint32_t)(grp.i_blk & 1) * 1024 : 2 * KG_K_BT_BYTES;
		const uint64_t a_desc0 = kg_umma_smem_desc(s_addr + a_off, grp.a_in >= 0 ? 2048 : 1024, 128);
		const uint32_t a_kstep = grp.a_in >= 0 ? (4 * 2048) >> 4 : (4 * 1024) >> 4;   // 32 rows = 4 row blocks, in 16-byte units
		uint32_t it = 0;
		for (uint32_t blk = me.split; blk < n_blocks; blk += me.n_splits, it++) {
			const uint32_t st = it % KG_K_STAGES, use = it / KG_K_STAGES;
			kg_mbar_wait(&st_full[st], use & 1);
			kg_tc_fence_after();
			if (kg_elect_one()) {
				const uint64_t st_units = (uint64_t)(st * (KG_K_STAGE_BYTES >> 4));
#pragma unroll
				for (uint32_t ks = 0; ks < 4; ks++) {
					const uint64_t ad = a_desc0 + st_units + ks * a_kstep;
					kg_umma_i8(tmem_base, ad, b_desc0 + st_units + ks * 512, idesc, (it | ks) != 0);
					if (has2) kg_umma_i8(tmem_base + 256, ad, b_desc0 + st_units + (KG_K_BT_BYTES >> 4) + ks * 512, idesc, (it | ks) != 0);
				}
				kg_umma_commit(&st_empty[st]);
			}
			__syncwarp();
		}
		if (kg_elect_one()) kg_umma_commit(done);
		__syncwarp();
	} else {
		// ===================== expanders (and, at the end, the accumulator flush) =====================
		const uint32_t t = threadIdx.x - KG_K_EXPAND_WARP0 * 32;
		const uint32_t r = t & (KG_K_ROWS - 1), sub = t >> 7;   // sub in [0, KG_K_EXPAND_SUBS)
		// work items of a stage: item i < 4 nb -> word (i % 4) of B tile (i / 4); then the 2 words of a separate A tile
		const uint32_t nb = has2 ? 2u : 1u;
		const uint32_t n_items = 4 * nb + (grp.a_in < 0 ? 2u : 0u);
		const uint32_t s_addr = kg_smem_u32(sStage);
		uint32_t it = 0;
		for (uint32_t blk = me.split; blk < n_blocks; blk += me.n_splits, it++) {
			const uint32_t rst = it % KG_K_RAW_STAGES, ruse = it / KG_K_RAW_STAGES;
			kg_mbar_wait(&raw_full[rst], ruse & 1);
			const uint64_t *row = reinterpret_cast<const uint64_t *>(sRaw + rst * raw_stage_bytes + r * row_bytes) + 1;
			// MAC filter of load_kmers (:117-121): computed once per tile by kg_prefilter_kernel (every tile group reads the
			// same rows); rows outside [mac, N - mac] contribute zeros
			const uint64_t grow = (uint64_t)blk * KG_K_ROWS + r;
			const bool keep = grow < prm.n_rows && ((__ldg(prm.keep_bits + (grow >> 5)) >> (grow & 31)) & 1u);

			const uint32_t st = it % KG_K_STAGES, use = it / KG_K_STAGES;
			kg_mbar_wait(&st_empty[st], (use & 1) ^ 1);
			const uint32_t st_base = s_addr + st * KG_K_STAGE_BYTES + (r & 7) * 16;
			// items are 32-sample halves of the words (2 n_items of them: 8 .. 20, a multiple of the 4 threads of a row)
			for (uint32_t hi = sub; hi < 2 * n_items; hi += KG_K_EXPAND_SUBS) {
				const uint32_t i = hi >> 1, half = hi & 1;
				uint32_t fw, dst_off, lbo;   // file word, byte offset of its first sample chunk in the stage, row-block stride
				if (i < 4 * nb) {
					const uint32_t b = i >> 2, wt = i & 3;
					fw = (uint32_t)grp.j2[b] * 4 + wt;
					dst_off = b * KG_K_BT_BYTES + wt * 512;
					lbo = 2048;
				} else {
					const uint32_t wt = i - 4 * nb;
					fw = (uint32_t)grp.i_blk * 2 + wt;
					dst_off = 2 * KG_K_BT_BYTES + wt * 512;
					lbo = 1024;
				}
				const uint32_t w = (keep && fw < prm.w_file) ? reinterpret_cast<const uint32_t *>(row + fw)[half] : 0u;
				const uint32_t dst = st_base + dst_off + (r >> 3) * lbo + half * 256;
				const uint4 e0 = kg_kin_spread16(w), e1 = kg_kin_spread16(w >> 16);
				asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(e0.x), "r"(e0.y), "r"(e0.z), "r"(e0.w) : "memory");
				asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + 128), "r"(e1.x), "r"(e1.y), "r"(e1.z), "r"(e1.w) : "memory");
			}
			kg_fence_proxy_async();
			__syncwarp();
			if (lane == 0) {
				kg_mbar_arrive(&st_full[st]);
				kg_mbar_arrive(&raw_empty[rst]);
			}
		}

		// ---- flush: TMEM accumulators -> global u64 delta (lower triangle, file column order)
		if (warp < KG_K_EXPAND_WARP0 + 4 && n_blocks > me.split) {
			kg_mbar_wait(done, 0);
			kg_tc_fence_after();
			const uint32_t q4 = warp & 3;
			const uint32_t a = (uint32_t)grp.i_blk * 128 + q4 * 32 + lane;   // operand row index (permuted inside 8)
			const uint32_t a_col = kg_kin_col_of_operand(a);                // file column of that operand row
			for (int b = 0; b < 2; b++) {
				if (grp.j2[b] < 0) continue;
				const uint32_t taddr = tmem_base + b * 256 + ((q4 * 32u) << 16);
				for (uint32_t c0 = 0; c0 < 256; c0 += 16) {
					uint32_t v[16];
					kg_tmem_ld16(taddr + c0, v);
					kg_tmem_ld_wait();
#pragma unroll
					for (int j = 0; j < 16; j++) {
						const uint32_t bb = (uint32_t)grp.j2[b] * 256 + c0 + j;
						const uint32_t b_col = kg_kin_col_of_operand(bb);
						if (v[j] != 0 && b_col <= a_col && a_col < prm.ld)
							atomicAdd(prm.delta + (size_t)a_col * prm.ld + b_col, (unsigned long long)v[j]);
					}
				}
			}
		}
	}
	kg_tc_fence_before();
	__syncthreads();
	if (warp == 1) kg_tmem_dealloc(tmem_base, 512);
}

// accum (memory order, see kg_kinship_begin) += delta (file order); then delta = 0 for the next tile
__global__ void kg_kinship_fold_kernel(unsigned long long *__restrict__ delta, uint32_t ld, const uint32_t *__restrict__ map_mem,
                                       uint32_t n_used, unsigned long long *__restrict__ accum,
                                       unsigned long long *__restrict__ delta_kept) {
	const uint64_t total = (uint64_t)n_used * n_used;
	for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t i = (uint32_t)(idx / n_used), j = (uint32_t)(idx % n_used);
		if (j > i) continue;
		const uint32_t mi = map_mem[i], mj = map_mem[j];
		const uint32_t a = (mi >> 6) * 64 + (mi & 63), b = (mj >> 6) * 64 + (mj & 63);
		const unsigned long long v = delta[(size_t)max(a, b) * ld + min(a, b)];
		if (v) accum[idx] += v;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		accum[total] += *delta_kept;
	}
}
