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

struct KgScanParams {
	KgRowView view;          // memory-order rows (identity map: the raw tile itself)
	uint32_t nb;             // 128-sample blocks (W_mem / 2)
	uint32_t n_used;         // N
	uint32_t n_pheno;        // P
	uint32_t min_count;
	const float *y_lane;     // [P_pad][nb*128]  y_lane[p][(4b+L)*32 + t] = y[p][128b + 32L + 31 - t]
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
};

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

template <int PT>
__host__ __device__ constexpr int kg_ys_group_stride() { return 32 * PT + 4; }

// grid = (row workers, p tiles); block = 256 threads.
// MODE 0: append hits (score > thr[p], or thr[p] < 0);  MODE 1: dense keep/scores output.
template <int R, int PT, int MODE>
__global__ void __launch_bounds__(256) kg_scan_exact_kernel(const KgScanParams prm) {
	extern __shared__ __align__(16) float ys[];
	constexpr int GS = kg_ys_group_stride<PT>();
	const uint32_t ng = prm.nb * 4;
	const uint32_t p0 = blockIdx.y * PT;

	// stage this CTA's y tile:  ys[g*GS + t*PT + q] = y_lane[p0+q][g*32 + t]
	for (uint32_t i = threadIdx.x; i < ng * 32 * PT; i += blockDim.x) {
		const uint32_t q = i / (ng * 32);
		const uint32_t gt = i - q * (ng * 32);
		ys[(gt >> 5) * GS + (gt & 31) * PT + q] = prm.y_lane[(size_t)(p0 + q) * (ng * 32) + gt];
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

	for (uint64_t chunk = blockIdx.x; chunk * ROWS_PER_CTA < prm.view.n_rows; chunk += gridDim.x) {
		const uint64_t row0 = chunk * ROWS_PER_CTA + (uint64_t)warp * (8 * R) + row_sub;
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
				const uint64_t row = row0 + (uint64_t)r * 8;
				uint32_t v = 0;
				if (row < prm.view.n_rows && g < w32_in) v = __ldg(row32 + row * stride32 + 2 + g);
				w[r] = v & m32;
				n1[r] += __popc(w[r]);
			}
			const float *yg = ys + g * GS;
#pragma unroll
			for (int t = 0; t < 32; t++) {
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
				for (int r = 0; r < R; r++) kg_pred_add<PT>(acc[r], yv, w[r] & (0x80000000u >> t));
			}
		}
		// ---- epilogue: N1 over the 4 lanes of the row, MAC filter, lane combine, double score
#pragma unroll
		for (int r = 0; r < R; r++) {
			uint32_t c = n1[r];
			c += __shfl_xor_sync(0xffffffffu, c, 1);
			c += __shfl_xor_sync(0xffffffffu, c, 2);
			const uint64_t row = row0 + (uint64_t)r * 8;
			const bool in_range = row < prm.view.n_rows;
			// :121  (popcnt >= mac) && (popcnt <= N - mac)
			const bool keep = in_range && c >= prm.min_count && c + prm.min_count <= prm.n_used;
			if (blockIdx.y == 0 && L == 0) {
				if (MODE == 1) { if (in_range) prm.keep_out[row] = keep ? 1 : 0; }
				else if (keep) atomicAdd(prm.kept_count, 1ull);
			}
			const double N1d = (double)c;
#pragma unroll
			for (int q = 0; q < PT; q++) {
				const float l0 = __shfl_sync(0xffffffffu, acc[r][q], grp_base + 0);
				const float l1 = __shfl_sync(0xffffffffu, acc[r][q], grp_base + 1);
				const float l2 = __shfl_sync(0xffffffffu, acc[r][q], grp_base + 2);
				const float l3 = __shfl_sync(0xffffffffu, acc[r][q], grp_base + 3);
				const uint32_t p = p0 + q;
				if ((q & 3) == (int)L && keep && p < prm.n_pheno) {
					const double score = kg_score_epilogue(l0, l1, l2, l3, Nd, N1d, prm.sums[p]);
					if (MODE == 1) {
						prm.scores_out[(size_t)p * prm.view.n_rows + row] = score;
					} else {
						const double th = prm.thr[p];
						if (th < 0.0 || score > th) {
							const unsigned long long pos = atomicAdd(prm.hit_count, 1ull);
							if (pos < prm.hit_capacity) {
								kg_hit h;
								h.row = prm.first_row_id + row;
								h.kmer = prm.view.base[row * prm.view.stride];
								h.score = score;
								h.pheno = p;
								h.pad_ = 0;
								prm.hits[pos] = h;
							}
						}
					}
				}
			}
		}
	}
}

// ---- squeeze: raw file rows -> memory-order rows (load_kmers :125-132) -------------------------
// map_lane[(4b+L)*32 + t] = (file_word << 6) | bit of memory sample 128b + 32L + 31 - t, 0xFFFFFFFF = pad
// Here we need memory sample order directly: map_mem[i] for memory column i.
__global__ void kg_squeeze_kernel(KgRowView raw, const uint32_t *__restrict__ map_mem, uint32_t n_used,
                                  uint32_t w_mem, uint64_t *__restrict__ out) {
	const uint64_t total = raw.n_rows * (uint64_t)(w_mem + 1);
	for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
	     idx += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t r = idx / (w_mem + 1);
		const uint32_t k = (uint32_t)(idx - r * (w_mem + 1));
		const uint64_t *row = raw.base + r * raw.stride;
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

// ---- exact refine of (row, phenotype) candidate pairs straight from RAW rows --------------------
// 4 lanes (L) per pair, 288-step sequential chains; bits gathered through map_lane.
struct KgPairParams {
	KgRowView raw;               // raw file rows
	uint32_t nb, n_used, n_pheno, min_count;
	const float *y_lane;         // [P_pad][nb*128]
	const float *sums;
	const uint32_t *map_lane;    // [nb*128]
	const uint64_t *file_mask;   // [W_file] m_map_mask
	const double *thr;
	const uint2 *pairs;          // (row in tile, phenotype)
	const unsigned long long *n_pairs;  // device counter written by the filter kernel (end of this tile's pairs)
	const unsigned long long *pair_begin;  // first pair of this tile (pairs of earlier tiles were refined already)
	uint64_t pair_capacity;
	kg_hit *hits;
	unsigned long long *hit_count;
	uint64_t hit_capacity;
	uint64_t first_row_id;
};

__global__ void __launch_bounds__(256) kg_scan_pairs_kernel(const KgPairParams prm) {
	const uint32_t L = threadIdx.x & 3;
	const uint32_t grp_base = (threadIdx.x & 31) & ~3u;
	unsigned long long n = *prm.n_pairs;
	if (n > prm.pair_capacity) n = prm.pair_capacity;
	const unsigned long long begin = *prm.pair_begin;
	const uint64_t n_iter = begin + (n - begin + 63) / 64 * 64;  // keep all lanes of a warp in the loop for the shuffles
	for (uint64_t i = begin + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2); i < n_iter;
	     i += ((uint64_t)gridDim.x * blockDim.x) >> 2) {
		const bool valid = i < n;
		uint2 pr = valid ? prm.pairs[i] : make_uint2(0, 0);
		const uint64_t *row = prm.raw.base + (uint64_t)pr.x * prm.raw.stride;
		const float *y = prm.y_lane + (size_t)pr.y * (prm.nb * 128);
		float acc = 0.0f;
		if (valid) {
			for (uint32_t b = 0; b < prm.nb; b++) {
				const uint32_t g = b * 4 + L;
#pragma unroll 4
				for (uint32_t t = 0; t < 32; t++) {
					const uint32_t m = __ldg(prm.map_lane + g * 32 + t);
					if (m != 0xFFFFFFFFu) {
						const uint64_t wv = __ldg(row + 1 + (m >> 6));
						if ((wv >> (m & 63)) & 1ull) acc = __fadd_rn(acc, __ldg(y + g * 32 + t));
					}
				}
			}
		}
		const float l0 = __shfl_sync(0xffffffffu, acc, grp_base + 0);
		const float l1 = __shfl_sync(0xffffffffu, acc, grp_base + 1);
		const float l2 = __shfl_sync(0xffffffffu, acc, grp_base + 2);
		const float l3 = __shfl_sync(0xffffffffu, acc, grp_base + 3);
		if (valid && L == 0) {
			uint32_t c = 0;
			for (uint32_t k = 0; k < prm.raw.w_in; k++) c += __popcll(row[1 + k] & prm.file_mask[k]);
			const double score =
			    kg_score_epilogue(l0, l1, l2, l3, (double)prm.n_used, (double)c, prm.sums[pr.y]);
			const double th = prm.thr[pr.y];
			if (th < 0.0 || score > th) {
				const unsigned long long pos = atomicAdd(prm.hit_count, 1ull);
				if (pos < prm.hit_capacity) {
					kg_hit h;
					h.row = prm.first_row_id + pr.x;
					h.kmer = row[0];
					h.score = score;
					h.pheno = pr.y;
					h.pad_ = 0;
					prm.hits[pos] = h;
				}
			}
		}
	}
}

// after a tile's pairs are refined: the next tile's pairs start where this tile's ended
__global__ void kg_scan_pairs_advance_kernel(const unsigned long long *n_pairs, unsigned long long *pair_begin) {
	*pair_begin = *n_pairs;
}
