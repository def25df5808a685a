// kg_scan_exact.cuh -- bit-exact score kernels (fp32 summation ORDER of the reference is part of
// the result; see SURVEY.md Appendix A and /root/reference/src/kmers_multiple_databases.cpp:327-363).
//
// Reference semantics restated for the GPU:
//   memory row = N_pad bits (N_pad = 128 * NB); float lane L in 0..3 of 128-sample block b walks
//   samples 128b + 32L + 31 - t for t = 0..31 (sign bit of a 32-bit half, shifted left every step),
//   adding y[sample] when the bit is set and +0.0f otherwise.  A predicated fp32 add is identical to
//   adding +0.0f (an accumulator that starts at +0 can never become -0), for every y incl. inf/nan.
//   score = ((l0+l1)+l2)+l3 in fp32 -> double epilogue with individually rounded mul/sub/div.
//
// Thread mapping (dense kernel): lane = (row_sub << 2) | L.  A thread owns R rows x PT phenotypes
// for one L; the four L lanes of a row are adjacent lanes, combined with shuffles in the epilogue.
// The y tile of PT phenotypes lives in shared memory as ys[g][t][q] (g = 4b + L) with a 4-float pad
// per g so the four distinct L addresses of a warp-wide LDS.128 fall in different banks.
#pragma once
#include "kg_common.cuh"
#include "kg_select.cuh"

struct KgScanParams {
	KgRowView view;          // memory-order rows (identity map: the raw tile itself)
	uint32_t nb;             // 128-sample blocks (W_mem / 2)
	uint32_t n_used;         // N
	uint32_t n_pheno;        // P
	uint32_t min_count;
	const float *y_lane;     // [P_pad][nb*128]  y_lane[p][(4b+L)*32 + t] = y[p][128b + 32L + 31 - t]
	const float *y_pair;     // [P_pad][nb*128]  y_pair[p][((8b+t4)*4+L)*4+k] = y_lane[p][(4b+L)*32 + 4 t4 + k]  (pair mode)
	const float *sums;       // [P_pad] sequential fp32 sum of the permuted padded vector (:288-295)
	const uint32_t *mask32;  // [nb*4] valid-sample mask of u32 word 4b+L
	const double *thr;       // [P_pad]
	// outputs
	kg_hit *hits;
	unsigned long long *hit_count;   // device counter
	uint64_t hit_capacity;
	unsigned long long *kept_count;  // device counter (rows passing the MAC filter)
	uint64_t first_row_id;
	uint8_t *keep_out;       // dense mode
	double *scores_out;      // dense mode [P][n_rows]
	// list mode (MODE 2), fed by the tensor-core filter.  row_list[pos] = tile row of every row the filter could not
	// rule out; the filter tests 16 phenotype columns at a time, and group_list[g * group_cap + k] = pos of the k-th
	// row that survived the test of group g.  blockIdx.y = tile of 8 filter columns: tile tau re-scores the rows of
	// group tau / 2 against the phenotypes tile_pheno[8 tau .. 8 tau + 7] (-1 = no phenotype in that column).
	const uint32_t *row_list;
	const uint32_t *group_list;
	const unsigned long long *group_count;   // [n_groups]
	uint64_t group_cap;
	const int32_t *tile_pheno;
	unsigned int *tile_chunk_counter;   // [tiles] zeroed before the launch: CTAs of a tile pull chunks of its list dynamically
	uint32_t list_compact;   // 1: the view holds the listed rows back to back (squeezed copies), indexed by pos
	uint64_t dense_limit;    // list mode: groups with at most this many entries are left to kg_scan_pair_kernel (0: none)
	// pair mode (kg_scan_pair_kernel): (position in row_list, phenotype) pairs from kg_pair_select_kernel
	const uint2 *pairs;
	const unsigned long long *pair_count;
	// device-selection mode (kg_select.cuh): hits go to per-phenotype candidate segments instead of `hits`
	KgCand *cand;            // [P][cand_cap], NULL = host mode
	uint32_t *cand_count;    // [P]
	uint32_t cand_cap;
};

// one admitted (row, phenotype): host mode appends a kg_hit to the interval buffer, device-selection mode appends a
// candidate to the phenotype's segment (counts may run past the capacity: the round is then discarded, kg_select.cuh)
__device__ __forceinline__ void kg_emit_hit(const KgScanParams &prm, uint32_t p, uint64_t row_id, uint64_t kmer, double score) {
	if (prm.cand) {
		const uint32_t pos = atomicAdd(prm.cand_count + p, 1u);
		if (pos < prm.cand_cap) {
			KgCand c;
			c.row = row_id; c.kmer = kmer; c.score = score;
			prm.cand[(size_t)p * prm.cand_cap + pos] = c;
		}
		return;
	}
	const unsigned long long pos = atomicAdd(prm.hit_count, 1ull);
	if (pos < prm.hit_capacity) {
		kg_hit h;
		h.row = row_id; h.kmer = kmer; h.score = score; h.pheno = p; h.pad_ = 0;
		prm.hits[pos] = h;
	}
}

__device__ __forceinline__ double kg_score_epilogue(float l0, float l1, float l2, float l3, double Nd,
                                                    double N1d, float sum) {
	// :358  yigi = sumsf[0] + sumsf[1] + sumsf[2] + sumsf[3]  (float, left to right)
	const float s = __fadd_rn(__fadd_rn(__fadd_rn(l0, l1), l2), l3);
	const double yigi = (double)s;
	// :359-361, every operation individually rounded (the reference build has no FMA)
	double r = __dsub_rn(__dmul_rn(Nd, yigi), __dmul_rn(N1d, (double)sum));
	r = __dmul_rn(r, r);
	const double den = __dsub_rn(__dmul_rn(Nd, N1d), __dmul_rn(N1d, N1d));
	return __ddiv_rn(r, den);
}

// One predicate per (row, step), PT predicated adds under it.
template <int PT>
__device__ __forceinline__ void kg_pred_add(float (&acc)[PT], const float (&y)[PT], uint32_t bit) {
	if constexpr (PT == 8) {
		asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %16, 0;\n\t"
		    "@p add.rn.f32 %0, %0, %8;\n\t@p add.rn.f32 %1, %1, %9;\n\t"
		    "@p add.rn.f32 %2, %2, %10;\n\t@p add.rn.f32 %3, %3, %11;\n\t"
		    "@p add.rn.f32 %4, %4, %12;\n\t@p add.rn.f32 %5, %5, %13;\n\t"
		    "@p add.rn.f32 %6, %6, %14;\n\t@p add.rn.f32 %7, %7, %15;\n\t}"
		    : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]),
		      "+f"(acc[6]), "+f"(acc[7])
		    : "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]), "f"(y[4]), "f"(y[5]), "f"(y[6]), "f"(y[7]),
		      "r"(bit));
	} else if constexpr (PT == 4) {
		asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %8, 0;\n\t"
		    "@p add.rn.f32 %0, %0, %4;\n\t@p add.rn.f32 %1, %1, %5;\n\t"
		    "@p add.rn.f32 %2, %2, %6;\n\t@p add.rn.f32 %3, %3, %7;\n\t}"
		    : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
		    : "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]), "r"(bit));
	} else if constexpr (PT == 2) {
		asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t"
		    "@p add.rn.f32 %0, %0, %2;\n\t@p add.rn.f32 %1, %1, %3;\n\t}"
		    : "+f"(acc[0]), "+f"(acc[1]) : "f"(y[0]), "f"(y[1]), "r"(bit));
	} else {
		static_assert(PT == 1, "PT must be 1, 2, 4 or 8");
		asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p add.rn.f32 %0, %0, %1;\n\t}"
		    : "+f"(acc[0]) : "f"(y[0]), "r"(bit));
	}
}

#ifndef KG_EXACT_UNROLL
#define KG_EXACT_UNROLL 4
#endif

template <int PT>
__host__ __device__ constexpr int kg_ys_group_stride() { return 32 * PT + 4; }

// grid = (row workers, p tiles); block = 256 threads.
// MODE 0: append hits (score > thr[p], or thr[p] < 0);  MODE 1: dense keep/scores output;
// MODE 2: like 0 but over the rows listed by the tensor-core filter (kept rows are counted by the filter).
template <int R, int PT, int MODE>
__global__ void __launch_bounds__(256) kg_scan_exact_kernel(const KgScanParams prm) {
	extern __shared__ __align__(16) float ys[];
	constexpr int GS = kg_ys_group_stride<PT>();
	const uint32_t ng = prm.nb * 4;
	const uint32_t p0 = blockIdx.y * PT;
	const uint64_t n_work = (MODE == 2) ? (uint64_t)prm.group_count[blockIdx.y >> 1] : prm.view.n_rows;
	if (MODE == 2 && (n_work == 0 || n_work <= prm.dense_limit)) return;   // nothing survived in this tile's group, or few: pair mode

	// stage this CTA's y tile:  ys[g*GS + t*PT + q] = y_lane[phenotype of slot q][g*32 + t]
	for (uint32_t i = threadIdx.x; i < ng * 32 * PT; i += blockDim.x) {
		const uint32_t q = i / (ng * 32);
		const uint32_t gt = i - q * (ng * 32);
		float v;
		if (MODE == 2) {
			const int32_t ph = prm.tile_pheno[p0 + q];
			v = ph >= 0 ? prm.y_lane[(size_t)ph * (ng * 32) + gt] : 0.0f;
		} else {
			v = prm.y_lane[(size_t)(p0 + q) * (ng * 32) + gt];
		}
		ys[(gt >> 5) * GS + (gt & 31) * PT + q] = v;
	}
	__syncthreads();

	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warp = threadIdx.x >> 5;
	const uint32_t L = lane & 3;
	const uint32_t row_sub = lane >> 2;
	const uint32_t grp_base = lane & ~3u;
	constexpr uint32_t ROWS_PER_CTA = 8 * 8 * R;
	const double Nd = (double)prm.n_used;
	const uint32_t *row32 = reinterpret_cast<const uint32_t *>(prm.view.base);
	const uint32_t stride32 = prm.view.stride * 2;
	const uint32_t w32_in = prm.view.w_in * 2;

	const uint32_t *glist = (MODE == 2) ? prm.group_list + (size_t)(blockIdx.y >> 1) * prm.group_cap : nullptr;
	__shared__ unsigned int s_next_chunk;
	for (uint64_t chunk = blockIdx.x;; chunk += gridDim.x) {
		if (MODE == 2) {
			// list lengths differ per tile and are short: dynamic chunk scheduling keeps the CTAs of a tile balanced
			__syncthreads();
			if (threadIdx.x == 0) s_next_chunk = atomicAdd(prm.tile_chunk_counter + blockIdx.y, 1u);
			__syncthreads();
			chunk = s_next_chunk;
		}
		if (chunk * ROWS_PER_CTA >= n_work) break;
		const uint64_t row0 = chunk * ROWS_PER_CTA + (uint64_t)warp * (8 * R) + row_sub;
		uint64_t rows_of[R];   // view row of work item row0 + 8 r
		uint64_t ids_of[R];    // row id inside the submitted tile (differs from rows_of for compacted lists)
		bool valid_of[R];
#pragma unroll
		for (int r = 0; r < R; r++) {
			const uint64_t item = row0 + (uint64_t)r * 8;
			valid_of[r] = item < n_work;
			if (MODE == 2) {
				const uint32_t pos = valid_of[r] ? glist[item] : 0u;
				ids_of[r] = valid_of[r] ? (uint64_t)prm.row_list[pos] : 0ull;
				rows_of[r] = prm.list_compact ? (uint64_t)pos : ids_of[r];
			} else {
				ids_of[r] = item;
				rows_of[r] = item;
			}
		}
		float acc[R][PT];
		uint32_t n1[R];
#pragma unroll
		for (int r = 0; r < R; r++) {
			n1[r] = 0;
#pragma unroll
			for (int q = 0; q < PT; q++) acc[r][q] = 0.0f;
		}
		for (uint32_t b = 0; b < prm.nb; b++) {
			const uint32_t g = b * 4 + L;
			const uint32_t m32 = prm.mask32[g];
			uint32_t w[R];
#pragma unroll
			for (int r = 0; r < R; r++) {
				const uint64_t row = rows_of[r];
				uint32_t v = 0;
				if (valid_of[r] && g < w32_in) v = __ldg(row32 + row * stride32 + 2 + g);
				w[r] = v & m32;
				n1[r] += __popc(w[r]);
			}
			const float *yg = ys + g * GS;
			// 32 steps in groups of KG_EXACT_UNROLL: the fully unrolled body (2 k+ FADDs) overflowed the
			// instruction caches (ncu: "no_instruction" was the top stall); a short body stays resident
#pragma unroll 1
			for (int t0 = 0; t0 < 32; t0 += KG_EXACT_UNROLL) {
				const uint32_t m0 = 0x80000000u >> t0;
#pragma unroll
				for (int k = 0; k < KG_EXACT_UNROLL; k++) {
					const int t = t0 + k;
					float yv[PT];
					if constexpr (PT >= 4) {
#pragma unroll
						for (int q = 0; q < PT; q += 4) {
							const float4 f = *reinterpret_cast<const float4 *>(yg + t * PT + q);
							yv[q] = f.x; yv[q + 1] = f.y; yv[q + 2] = f.z; yv[q + 3] = f.w;
						}
					} else {
#pragma unroll
						for (int q = 0; q < PT; q++) yv[q] = yg[t * PT + q];
					}
#pragma unroll
					for (int r = 0; r < R; r++) kg_pred_add<PT>(acc[r], yv, w[r] & (m0 >> k));
				}
			}
		}
		// ---- epilogue: N1 over the 4 lanes of the row, MAC filter, lane combine, double score
#pragma unroll
		for (int r = 0; r < R; r++) {
			uint32_t c = n1[r];
			c += __shfl_xor_sync(0xffffffffu, c, 1);
			c += __shfl_xor_sync(0xffffffffu, c, 2);
			const uint64_t row = rows_of[r];
			const bool in_range = valid_of[r];
			// :121  (popcnt >= mac) && (popcnt <= N - mac)
			const bool keep = in_range && c >= prm.min_count && c + prm.min_count <= prm.n_used;
			if (blockIdx.y == 0 && L == 0) {
				if (MODE == 1) { if (in_range) prm.keep_out[row] = keep ? 1 : 0; }
				else if (MODE == 0 && keep) atomicAdd(prm.kept_count, 1ull);
			}
			const double N1d = (double)c;
#pragma unroll
			for (int q = 0; q < PT; q++) {
				const float l0 = __shfl_sync(0xffffffffu, acc[r][q], grp_base + 0);
				const float l1 = __shfl_sync(0xffffffffu, acc[r][q], grp_base + 1);
				const float l2 = __shfl_sync(0xffffffffu, acc[r][q], grp_base + 2);
				const float l3 = __shfl_sync(0xffffffffu, acc[r][q], grp_base + 3);
				const int32_t ph = (MODE == 2) ? prm.tile_pheno[p0 + q] : (int32_t)(p0 + q);
				const uint32_t p = (uint32_t)ph;
				if ((q & 3) == (int)L && keep && ph >= 0 && p < prm.n_pheno) {
					const double score = kg_score_epilogue(l0, l1, l2, l3, Nd, N1d, prm.sums[p]);
					if (MODE == 1) {
						prm.scores_out[(size_t)p * prm.view.n_rows + row] = score;
					} else {
						const double th = prm.thr[p];
						if (th < 0.0 || score > th) kg_emit_hit(prm, p, prm.first_row_id + ids_of[r], prm.view.base[row * prm.view.stride], score);
					}
				}
			}
		}
	}
}

// ---- pair mode: ONE phenotype per listed row ------------------------------------------------------------------
// Four adjacent lanes (L = 0..3, the reference's SSE lanes) re-score one (row, phenotype) pair: lane L walks the
// 32-sample quarters 4 b + L of every 128-sample block in the reference's order (kg_scan_exact_kernel has the
// derivation) with y read straight from global memory (y_pair: L2 / L1 resident, P x N_pad floats, laid out so that
// the four lanes of a pair read 64 contiguous bytes per load).  Work per pair is N_pad predicated
// adds instead of the 16 N_pad of the two 8-phenotype tiles a listed (row, group) costs in list mode.
__global__ void __launch_bounds__(256) kg_scan_pair_kernel(const KgScanParams prm) {
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t L = lane & 3;
	const uint32_t grp_base = lane & ~3u;
	const uint64_t n_pairs = *prm.pair_count;
	const uint32_t *row32 = reinterpret_cast<const uint32_t *>(prm.view.base);
	const uint32_t stride32 = prm.view.stride * 2;
	const uint32_t w32_in = prm.view.w_in * 2;
	const uint32_t ng = prm.nb * 4;
	const double Nd = (double)prm.n_used;
	const uint64_t per_pass = ((uint64_t)gridDim.x * blockDim.x) >> 2;
	// whole warps stay in the loop (shuffles below): the bound is rounded up to the 8 pairs of a warp
	for (uint64_t item0 = 0; item0 < n_pairs; item0 += per_pass) {
		const uint64_t item = item0 + ((((uint64_t)blockIdx.x * blockDim.x + threadIdx.x)) >> 2);
		const bool valid = item < n_pairs;
		uint32_t pos = 0, p = 0;
		if (valid) { const uint2 pr = prm.pairs[item]; pos = pr.x; p = pr.y; }
		const uint64_t id = valid ? (uint64_t)prm.row_list[pos] : 0ull;
		const uint64_t row = prm.list_compact ? (uint64_t)pos : id;
		const float4 *yp = reinterpret_cast<const float4 *>(prm.y_pair + (size_t)p * (ng * 32)) + L;
		float acc = 0.0f;
		uint32_t n1 = 0;
		for (uint32_t b = 0; b < prm.nb; b++) {
			const uint32_t g = b * 4 + L;
			uint32_t w = 0;
			if (valid && g < w32_in) w = __ldg(row32 + row * stride32 + 2 + g);
			w &= prm.mask32[g];
			n1 += __popc(w);
#pragma unroll
			for (int t4 = 0; t4 < 8; t4++) {
				const float4 f = __ldg(yp + (b * 8 + t4) * 4);
				const float yv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
				for (int k = 0; k < 4; k++) {
					float a1[1] = {acc};
					const float y1[1] = {yv[k]};
					kg_pred_add<1>(a1, y1, w & (0x80000000u >> (4 * t4 + k)));
					acc = a1[0];
				}
			}
		}
		uint32_t c = n1;
		c += __shfl_xor_sync(0xffffffffu, c, 1);
		c += __shfl_xor_sync(0xffffffffu, c, 2);
		const float l0 = __shfl_sync(0xffffffffu, acc, grp_base + 0);
		const float l1 = __shfl_sync(0xffffffffu, acc, grp_base + 1);
		const float l2 = __shfl_sync(0xffffffffu, acc, grp_base + 2);
		const float l3 = __shfl_sync(0xffffffffu, acc, grp_base + 3);
		// :121  (popcnt >= mac) && (popcnt <= N - mac)  (the filter already applied it; kept for symmetry with list mode)
		const bool keep = valid && c >= prm.min_count && c + prm.min_count <= prm.n_used;
		if (L == 0 && keep && p < prm.n_pheno) {
			const double score = kg_score_epilogue(l0, l1, l2, l3, Nd, (double)c, prm.sums[p]);
			const double th = prm.thr[p];
			if (th < 0.0 || score > th) kg_emit_hit(prm, p, prm.first_row_id + id, prm.view.base[row * prm.view.stride], score);
		}
	}
}

// ---- squeeze: raw file rows -> memory-order rows (load_kmers :125-132) -------------------------
// map_lane[(4b+L)*32 + t] = (file_word << 6) | bit of memory sample 128b + 32L + 31 - t, 0xFFFFFFFF = pad
// Here we need memory sample order directly: map_mem[i] for memory column i.
// With row_list != NULL only the listed rows are squeezed, back to back (out row i = raw row row_list[i]).
__global__ void kg_squeeze_kernel(KgRowView raw, const uint32_t *__restrict__ map_mem, uint32_t n_used,
                                  uint32_t w_mem, uint64_t *__restrict__ out, const uint32_t *__restrict__ row_list,
                                  const unsigned long long *__restrict__ row_list_count) {
	const uint64_t n_out = row_list ? (uint64_t)*row_list_count : raw.n_rows;
	const uint64_t total = n_out * (uint64_t)(w_mem + 1);
	for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
	     idx += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t r = idx / (w_mem + 1);
		const uint32_t k = (uint32_t)(idx - r * (w_mem + 1));
		const uint64_t *row = raw.base + (row_list ? (uint64_t)row_list[r] : r) * raw.stride;
		if (k == 0) { out[idx] = row[0]; continue; }
		const uint32_t mw = k - 1;
		uint64_t v = 0;
		const uint32_t c0 = mw * 64;
#pragma unroll 4
		for (uint32_t j = 0; j < 64; j++) {
			const uint32_t col = c0 + j;
			if (col < n_used) {
				const uint32_t m = __ldg(map_mem + col);
				v |= ((row[1 + (m >> 6)] >> (m & 63)) & 1ull) << j;
			}
		}
		out[idx] = v;
	}
}

