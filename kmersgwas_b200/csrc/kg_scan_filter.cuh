// kg_scan_filter.cuh -- int8 tensor-core FILTER in front of the exact score kernel (scan engine 2).
//
// With P = 101 phenotype columns the scan is a (rows x N) . (N x P) contraction of presence bits with
// phenotype values (hard part 2).  The reference's result is defined by a float32
// summation order, so the tensor core cannot produce it -- but it can prove, for almost every
// (row, phenotype) pair, that the reference score cannot exceed the heap threshold.  Only the pairs it
// cannot rule out go to the exact kernel (kg_scan_pairs_kernel), so the reported hits stay bit-identical.
//
// Bound (DESIGN.md section 4 has the derivation).  Per phenotype p, host side (kg_tc.cuh):
//   ybar = sum_ref / N,  c_i = y_i - ybar,  s = max|c_i| / 127,  q_i = rint(c_i / s) in [-127, 127],
//   e_i = c_i - s q_i (|e_i| <= s/2),  e_tot = sum e_i,  A = sum |y_i|,
//   gamma = float32 error of the reference's lane sums.
// Per row (bits S, N1 = |S|, m = min(N1, N - N1), den = N1 (N - N1)) the tensor core gives the EXACT integer
//   Q = sum_{i in S} q_i, and
//   |r_ref| <= N s ( |Q| + m/2 + kappa ),      kappa = (|e_tot| + gamma A + |N ybar - sum_ref|) / s + 1
//   score_ref = r_ref^2 / den > thr   ==>   |Q| >= alpha sqrt(den) - m/2 - kappa,   alpha = sqrt(thr) / (N s)
// with alpha rounded down and sqrt(den) rounded down, so no qualifying pair is ever dropped.
//
// Kernel: persistent, one CTA per SM, 128-row blocks, warp-specialised:
//   warp 0      bulk-async-copies raw 128-row blocks (contiguous 128 * 8(1+W) bytes) into a 2-stage ring
//   warps 2-5   expand presence bits -> u8 {0,1} into K-major core-matrix A stages (16 KB = 128 rows x 128 columns)
//   warp 1      one thread issues tcgen05.mma kind::i8 (M = 128, N = P_pad, K = 32 per instruction) into a
//               double-buffered TMEM accumulator; B (quantised phenotypes, P_pad x K_pad s8) stays in smem
//   warps 6-9   epilogue: masked popcount + MAC filter from the raw rows, tcgen05.ld of the 128 x P_pad
//               accumulators, the bound test above, rare candidate pairs appended to a global list
#pragma once
#include "kg_common.cuh"
#include "kg_tc_ptx.cuh"

#define KG_F_ROWS 128            // rows per block = UMMA M
#define KG_F_CHUNK_COLS 128      // presence columns per A stage (two u64 words per row)
#define KG_F_A_STAGE_BYTES (KG_F_ROWS * KG_F_CHUNK_COLS)
#define KG_F_A_STAGES 3
#define KG_F_RAW_STAGES 2
#define KG_F_THREADS 320
#define KG_F_EXPAND_WARP0 2
#define KG_F_EPI_WARP0 6

struct KgFilterParams {
	const uint64_t *rows;      // raw tile, 16-byte aligned
	uint64_t n_rows;
	uint32_t w_file;           // presence words per row
	uint32_t nc;               // A stages per block = ceil(w_file / 2)
	uint32_t p_pad;            // UMMA N (multiple of 16, <= 256)
	uint32_t tcols;            // TMEM columns per accumulator buffer (power of two >= p_pad, >= 32)
	const int8_t *yq_image;    // B operand in its shared-memory byte order, b_bytes long
	uint32_t b_bytes;          // (p_pad / 8) * sbo_b
	uint32_t sbo_b;            // nc * 1024
	const float2 *pconst;      // [p_pad] (alpha, kappa); padding columns hold (+inf, 0)
	const uint64_t *file_mask; // [w_file] used-column mask (m_map_mask)
	uint32_t n_used, min_count;
	uint2 *pairs;              // out: (row in tile, phenotype)
	unsigned long long *n_pairs;
	uint64_t pair_capacity;
	unsigned long long *kept_count;
	int32_t *q_out;            // debug mode: [n_rows][p_pad] accumulators
};

__host__ __device__ inline uint32_t kg_filter_raw_stage_bytes(uint32_t w_file) { return KG_F_ROWS * 8u * (w_file + 1); }
__host__ __device__ inline size_t kg_filter_smem_bytes(uint32_t w_file, uint32_t b_bytes, uint32_t p_pad) {
	return 1024 /*alignment slack*/ + (size_t)b_bytes + (size_t)KG_F_A_STAGES * KG_F_A_STAGE_BYTES +
	       (size_t)KG_F_RAW_STAGES * kg_filter_raw_stage_bytes(w_file) + (size_t)p_pad * 8 + (size_t)w_file * 8 + 256;
}

// 4 presence bits -> 4 bytes of 0/1 (bit k -> byte k): the partial products land on distinct bit positions
__device__ __forceinline__ uint32_t kg_spread4(uint32_t nibble) { return (nibble * 0x00204081u) & 0x01010101u; }

__device__ __forceinline__ void kg_expand16(uint32_t h, uint32_t smem_dst) {
	const uint32_t a = kg_spread4(h & 15u), b = kg_spread4((h >> 4) & 15u), c = kg_spread4((h >> 8) & 15u),
	               d = kg_spread4((h >> 12) & 15u);
	asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int MODE>  // 0 = candidate pairs, 1 = debug: dump accumulators
__global__ void __launch_bounds__(KG_F_THREADS, 1) kg_scan_filter_kernel(const KgFilterParams prm) {
	extern __shared__ uint8_t kg_f_smem_raw[];
	// carve shared memory (1024-byte aligned base)
	uint8_t *base = reinterpret_cast<uint8_t *>(((uintptr_t)kg_f_smem_raw + 1023) & ~(uintptr_t)1023);
	uint8_t *sB = base;
	uint8_t *sA = sB + prm.b_bytes;
	const uint32_t raw_stage_bytes = kg_filter_raw_stage_bytes(prm.w_file);
	uint8_t *sRaw = sA + KG_F_A_STAGES * KG_F_A_STAGE_BYTES;
	float2 *sConst = reinterpret_cast<float2 *>(sRaw + KG_F_RAW_STAGES * raw_stage_bytes);
	uint64_t *sMask = reinterpret_cast<uint64_t *>(sConst + prm.p_pad);
	uint64_t *bars = sMask + prm.w_file;
	uint64_t *raw_full = bars, *raw_empty = bars + KG_F_RAW_STAGES;
	uint64_t *a_full = bars + 2 * KG_F_RAW_STAGES, *a_empty = a_full + KG_F_A_STAGES;
	uint64_t *tm_full = a_empty + KG_F_A_STAGES, *tm_empty = tm_full + 2;
	uint64_t *b_full = tm_empty + 2;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(b_full + 1);

	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t n_blocks = (uint32_t)((prm.n_rows + KG_F_ROWS - 1) / KG_F_ROWS);
	const uint32_t row_bytes = 8u * (prm.w_file + 1);

	if (threadIdx.x == 0) {
		for (int i = 0; i < KG_F_RAW_STAGES; i++) { kg_mbar_init(&raw_full[i], 1); kg_mbar_init(&raw_empty[i], 8); }
		for (int i = 0; i < KG_F_A_STAGES; i++) { kg_mbar_init(&a_full[i], 4); kg_mbar_init(&a_empty[i], 1); }
		for (int i = 0; i < 2; i++) { kg_mbar_init(&tm_full[i], 1); kg_mbar_init(&tm_empty[i], 4); }
		kg_mbar_init(b_full, 1);
		kg_fence_mbar_init();
	}
	for (uint32_t i = threadIdx.x; i < prm.p_pad; i += blockDim.x) sConst[i] = prm.pconst[i];
	for (uint32_t i = threadIdx.x; i < prm.w_file; i += blockDim.x) sMask[i] = prm.file_mask[i];
	if (warp == 1) kg_tmem_alloc(tmem_slot, 2 * prm.tcols);
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
				const uint32_t st = it % KG_F_RAW_STAGES, use = it / KG_F_RAW_STAGES;
				kg_mbar_wait(&raw_empty[st], (use & 1) ^ 1);
				const uint64_t r0 = (uint64_t)blk * KG_F_ROWS;
				const uint32_t valid = (uint32_t)min((uint64_t)KG_F_ROWS, prm.n_rows - r0);
				const uint32_t bytes = valid * row_bytes;
				const uint32_t bulk = bytes & ~15u;
				const uint8_t *src = reinterpret_cast<const uint8_t *>(prm.rows) + r0 * row_bytes;
				uint8_t *dst = sRaw + st * raw_stage_bytes;
				if (bytes != bulk)  // 8 trailing bytes of a ragged last block
					*reinterpret_cast<uint64_t *>(dst + bulk) = *reinterpret_cast<const uint64_t *>(src + bulk);
				kg_mbar_arrive_expect_tx(&raw_full[st], bulk);
				if (bulk) kg_bulk_g2s(dst, src, bulk, &raw_full[st]);
			}
		}
	} else if (warp == 1) {
		// ===================== MMA issuer (one thread) =====================
		if (lane == 0) {
			const uint32_t idesc = kg_umma_idesc_i8(KG_F_ROWS, prm.p_pad, false, true, false, false);
			const uint32_t sA_addr = kg_smem_u32(sA), sB_addr = kg_smem_u32(sB);
			kg_mbar_wait(b_full, 0);
			uint32_t it = 0, ait = 0;
			for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, it++) {
				const uint32_t buf = it & 1;
				kg_mbar_wait(&tm_empty[buf], ((it >> 1) & 1) ^ 1);
				kg_tc_fence_after();
				const uint32_t d_tmem = tmem_base + buf * prm.tcols;
				for (uint32_t c = 0; c < prm.nc; c++, ait++) {
					const uint32_t st = ait % KG_F_A_STAGES, use = ait / KG_F_A_STAGES;
					kg_mbar_wait(&a_full[st], use & 1);
					kg_tc_fence_after();
#pragma unroll
					for (uint32_t kk = 0; kk < 4; kk++) {
						const uint64_t ad = kg_umma_smem_desc(sA_addr + st * KG_F_A_STAGE_BYTES + kk * 256, 128, 1024);
						const uint64_t bd = kg_umma_smem_desc(sB_addr + (c * 8 + kk * 2) * 128, 128, prm.sbo_b);
						kg_umma_i8(d_tmem, ad, bd, idesc, (c | kk) != 0);
					}
					kg_umma_commit(&a_empty[st]);
				}
				kg_umma_commit(&tm_full[buf]);
			}
		}
	} else if (warp < KG_F_EPI_WARP0) {
		// ===================== expanders: bits -> u8 core matrices =====================
		const uint32_t r = threadIdx.x - KG_F_EXPAND_WARP0 * 32;  // row of the block
		const uint32_t dst_row = (r & 7) * 16 + (r >> 3) * 1024;
		const uint32_t sA_addr = kg_smem_u32(sA);
		uint32_t it = 0, ait = 0;
		for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, it++) {
			const uint32_t rst = it % KG_F_RAW_STAGES, ruse = it / KG_F_RAW_STAGES;
			kg_mbar_wait(&raw_full[rst], ruse & 1);
			const uint64_t *row = reinterpret_cast<const uint64_t *>(sRaw + rst * raw_stage_bytes + r * row_bytes) + 1;
			for (uint32_t c = 0; c < prm.nc; c++, ait++) {
				const uint32_t st = ait % KG_F_A_STAGES, use = ait / KG_F_A_STAGES;
				const uint64_t w0 = row[2 * c];
				const uint64_t w1 = (2 * c + 1 < prm.w_file) ? row[2 * c + 1] : 0ull;
				kg_mbar_wait(&a_empty[st], (use & 1) ^ 1);
				const uint32_t dst = sA_addr + st * KG_F_A_STAGE_BYTES + dst_row;
				kg_expand16((uint32_t)w0 & 0xFFFFu, dst);
				kg_expand16((uint32_t)(w0 >> 16) & 0xFFFFu, dst + 128);
				kg_expand16((uint32_t)(w0 >> 32) & 0xFFFFu, dst + 256);
				kg_expand16((uint32_t)(w0 >> 48), dst + 384);
				kg_expand16((uint32_t)w1 & 0xFFFFu, dst + 512);
				kg_expand16((uint32_t)(w1 >> 16) & 0xFFFFu, dst + 640);
				kg_expand16((uint32_t)(w1 >> 32) & 0xFFFFu, dst + 768);
				kg_expand16((uint32_t)(w1 >> 48), dst + 896);
				kg_fence_proxy_async();
				__syncwarp();
				if (lane == 0) kg_mbar_arrive(&a_full[st]);
			}
			__syncwarp();
			if (lane == 0) kg_mbar_arrive(&raw_empty[rst]);
		}
	} else {
		// ===================== epilogue: MAC filter + bound test =====================
		const uint32_t q4 = warp & 3;                       // TMEM lane quarter this warp may access
		const uint32_t r = q4 * 32 + lane;                  // row of the block = TMEM lane
		const float Nf = (float)prm.n_used;
		unsigned long long kept_local = 0;
		uint32_t it = 0;
		for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, it++) {
			const uint32_t rst = it % KG_F_RAW_STAGES, ruse = it / KG_F_RAW_STAGES;
			const uint64_t grow = (uint64_t)blk * KG_F_ROWS + r;
			kg_mbar_wait(&raw_full[rst], ruse & 1);
			uint32_t n1 = 0;
			if (grow < prm.n_rows) {
				const uint64_t *row = reinterpret_cast<const uint64_t *>(sRaw + rst * raw_stage_bytes + r * row_bytes) + 1;
				for (uint32_t k = 0; k < prm.w_file; k++) n1 += __popcll(row[k] & sMask[k]);
			}
			__syncwarp();
			if (lane == 0) kg_mbar_arrive(&raw_empty[rst]);
			// load_kmers :121  (popcnt >= mac) && (popcnt <= N - mac)
			const bool keep = grow < prm.n_rows && n1 >= prm.min_count && n1 + prm.min_count <= prm.n_used;
			kept_local += __popc(__ballot_sync(0xffffffffu, keep));
			const float n1f = (float)n1, n0f = Nf - n1f;
			const float hm = 0.5f * fminf(n1f, n0f);
			const float g = __fmul_rd(__fsqrt_rd(n1f * n0f), 0.999999f);   // <= sqrt(den), den exact in fp32 (< 2^24)

			const uint32_t buf = it & 1;
			kg_mbar_wait(&tm_full[buf], (it >> 1) & 1);
			kg_tc_fence_after();
			const uint32_t taddr = tmem_base + buf * prm.tcols + ((q4 * 32u) << 16);
			for (uint32_t c0 = 0; c0 < prm.p_pad; c0 += 16) {
				uint32_t v[16];
				kg_tmem_ld16(taddr + c0, v);
				kg_tmem_ld_wait();
				if (MODE == 1) {
					if (grow < prm.n_rows) {
#pragma unroll
						for (int j = 0; j < 16; j++) prm.q_out[grow * prm.p_pad + c0 + j] = (int32_t)v[j];
					}
				} else {
#pragma unroll
					for (int j = 0; j < 16; j++) {
						const float2 ak = sConst[c0 + j];
						const float thr = __fsub_rd(__fmaf_rd(ak.x, g, -ak.y), hm);
						const float qa = (float)abs((int32_t)v[j]);
						if (keep && qa >= thr) {
							const unsigned long long pos = atomicAdd(prm.n_pairs, 1ull);
							if (pos < prm.pair_capacity) prm.pairs[pos] = make_uint2((uint32_t)grow, c0 + j);
						}
					}
				}
			}
			kg_tc_fence_before();
			__syncwarp();
			if (lane == 0) kg_mbar_arrive(&tm_empty[buf]);
		}
		if (lane == 0 && kept_local && q4 == 0) { /* each row is counted by exactly one epilogue warp */ }
		if (lane == 0 && kept_local) atomicAdd(prm.kept_count, kept_local);
	}

	// teardown: every role has drained its loop; the last tm_full wait of the epilogue implies all MMAs completed
	kg_tc_fence_before();
	__syncthreads();
	if (warp == 1) kg_tmem_dealloc(tmem_base, 2 * prm.tcols);
}
