// kg_filter_retune.cuh -- device-side maintenance of the tensor filter's bound constants (device-selection mode).
//
// With the heaps on the device (kg_select.cuh) the thresholds never visit the host, so what kg_tc_update_thresholds
// (kg_tc.cuh) does on the host per round -- per-phenotype (alpha, kappa), the column order sorted by alpha, the B
// operand image in that order, the per-group loosest constants and the upper tangents of the groups' slack tables --
// runs here, stream-ordered behind the heap replay of the round.  One CTA per filter pass (<= 127 phenotype columns).
// Every rounding is directed so that the filter only ever lists MORE rows (kg_scan_filter.cuh has the bound).
#pragma once
#include "kg_scan_filter.cuh"
#include "kg_select.cuh"

struct KgRetuneParams {
	uint32_t n_pheno;            // phenotypes of this context
	uint32_t n_used;             // N
	uint32_t p_pad, sbo_b, b_bytes;
	uint32_t m_half;             // N / 2: slack tables hold m = 0 .. m_half
	uint32_t cols_per_pass;      // pass (= blockIdx.x) owns the phenotypes [pass * cols_per_pass, ...) (<= 127 of them)
	const double *thr;           // [P] current thresholds (-1: heap not full)
	const double *scale;         // [P] quantisation step s_p
	const float *kappa0;         // [P]
	const uint8_t *degenerate;   // [P]
	const int8_t *q;             // [P][N] quantised centred phenotypes, memory (phenotype-file) column order
	const uint32_t *kidx;        // [N] K index (byte inside the operand rows) of memory column i
	const float *slack;          // [P][m_half + 1] F_p(m)
	uint32_t *col_of;            // [P] in/out: B / accumulator column of phenotype p inside its pass (0 = not assigned yet)
	float *group_lines;          // [n_pass][16][8] in/out: 4 intercepts + 4 slopes per group
	int8_t *yq_image;            // [n_pass][b_bytes] out (on reorder)
	int32_t *tile_pheno;         // [n_pass][p_pad] out (on reorder)
	KgFilterGroupConst *gconst;  // [n_pass][16] group slots
	int32_t *thr_tab;            // [n_pass][p_pad / 16][n_used + 1] group bound per row popcount (KgFilterParams::thr_tab)
	float *alpha_out, *kappa_out;   // [P] per-phenotype constants of the per-column test
	unsigned long long *status;  // KG_SEL_ST_* (reorder counter) or NULL
	uint32_t force;              // 1: rebuild regardless of the tightness test
};

__device__ __forceinline__ size_t kg_retune_b_offset(uint32_t sbo_b, uint32_t n, uint32_t k) {
	return (size_t)(n % 8) * 16 + (size_t)(n / 8) * sbo_b + (size_t)(k / 16) * 128 + (k % 16);
}

__global__ void __launch_bounds__(256) kg_filter_retune_kernel(const KgRetuneParams prm_in) {
	// this CTA's pass: shift the per-pass tables
	KgRetuneParams prm = prm_in;
	const uint32_t pass = blockIdx.x;
	prm.group_lines += (size_t)pass * 16 * 8;
	prm.yq_image += (size_t)pass * prm.b_bytes;
	prm.tile_pheno += (size_t)pass * prm.p_pad;
	prm.gconst += (size_t)pass * 16;
	prm.thr_tab += (size_t)pass * kg_filter_tab_floats(prm.p_pad, prm.n_used);
	__shared__ float s_alpha[128], s_kappa[128];
	__shared__ uint32_t s_col_new[128], s_col_cur[128];
	__shared__ float s_amin[2][16];
	__shared__ double s_tight[2];
	__shared__ int s_reorder;
	__shared__ float s_red[256];
	__shared__ float s_slope;
	__shared__ KgFilterGroupConst s_gc[16];
	const uint32_t P0 = pass * prm.cols_per_pass, PC = min(prm.cols_per_pass, prm.n_pheno - P0), N = prm.n_used;
	const uint32_t n_groups = prm.p_pad / 16;
	const uint32_t tid = threadIdx.x;

	// 1. per-phenotype constants (host twin: kg_tc_update_thresholds)
	for (uint32_t j = tid; j < PC; j += blockDim.x) {
		const uint32_t p = P0 + j;
		const double thr = prm.thr[p];
		float a = 0.0f, k = 3.0e38f;
		if (!prm.degenerate[p] && thr >= 0.0 && isfinite(thr)) {
			const double ad = sqrt(thr) / ((double)N * prm.scale[p]) * (1.0 - 1e-6);
			a = __double2float_rd(ad);
			k = __double2float_ru((double)prm.kappa0[p]);
		}
		s_alpha[j] = a;
		s_kappa[j] = k;
		prm.alpha_out[p] = a;
		prm.kappa_out[p] = k;
		s_col_cur[j] = prm.col_of[p];
	}
	__syncthreads();
	// 2. columns sorted by alpha (stable): column = 1 + rank
	for (uint32_t j = tid; j < PC; j += blockDim.x) {
		const float a = s_alpha[j];
		uint32_t rank = 0;
		for (uint32_t i = 0; i < PC; i++) {
			const float b = s_alpha[i];
			rank += (b < a || (b == a && i < j)) ? 1u : 0u;
		}
		s_col_new[j] = 1 + rank;
	}
	__syncthreads();
	// 3. tightness of both assignments = sum over phenotypes of the alpha their group tests with
	if (tid < 2) {
		const uint32_t *cols = tid == 0 ? s_col_new : s_col_cur;
		bool valid = true;
		for (uint32_t g = 0; g < 16; g++) s_amin[tid][g] = INFINITY;
		for (uint32_t j = 0; j < PC; j++) {
			if (cols[j] == 0 || cols[j] >= prm.p_pad) { valid = false; break; }
			s_amin[tid][cols[j] / 16] = fminf(s_amin[tid][cols[j] / 16], s_alpha[j]);
		}
		double t = -1.0;
		if (valid) {
			t = 0.0;
			for (uint32_t j = 0; j < PC; j++) t += (double)s_amin[tid][cols[j] / 16];
		}
		s_tight[tid] = t;
	}
	__syncthreads();
	if (tid == 0) {
		s_reorder = prm.force || s_tight[1] < 0.0 || s_tight[0] > 1.01 * s_tight[1];
		if (s_reorder && prm.status) atomicAdd(prm.status + KG_SEL_ST_REORDERS, 1ull);
	}
	__syncthreads();
	const bool reorder = s_reorder != 0;
	const uint32_t *cols = reorder ? s_col_new : s_col_cur;

	if (reorder) {
		// 4a. B image: zero, column 0 = -1 over the used columns (row popcount), phenotype columns = -q (A holds -1 per set bit)
		for (uint32_t i = tid; i < prm.b_bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(prm.yq_image)[i] = make_uint4(0, 0, 0, 0);
		for (uint32_t i = tid; i < prm.p_pad; i += blockDim.x) prm.tile_pheno[i] = -1;
		__syncthreads();
		for (uint32_t i = tid; i < N; i += blockDim.x) prm.yq_image[kg_retune_b_offset(prm.sbo_b, 0, prm.kidx[i])] = (int8_t)-1;
		for (uint32_t e = tid; e < PC * N; e += blockDim.x) {
			const uint32_t j = e / N, i = e - j * N;
			prm.yq_image[kg_retune_b_offset(prm.sbo_b, cols[j], prm.kidx[i])] = (int8_t)-prm.q[(size_t)(P0 + j) * N + i];
		}
		for (uint32_t j = tid; j < PC; j += blockDim.x) {
			prm.col_of[P0 + j] = cols[j];
			prm.tile_pheno[cols[j]] = (int32_t)(P0 + j);
		}
		// 4b. slack lines per group: upper tangents of U_g(m) = max over the group's phenotypes of F_p(m), m in [0, N/2]
		const uint32_t M = prm.m_half;
		for (uint32_t g = 0; g < n_groups; g++) {
			uint32_t anchor[4] = {max(1u, M / 16), max(1u, M / 5), max(1u, M / 2), max(1u, M > 0 ? M - 1 : 1u)};
			for (int k = 0; k < 4; k++) {
				const uint32_t m0 = min(anchor[k], M > 0 ? M - 1 : 0u);
				// U(m0), U(m0 + 1) -> slope; then intercept = max_m (U(m) - slope m)
				if (tid == 0) {
					float u0 = 0.f, u1 = 0.f;
					bool any = false;
					for (uint32_t j = 0; j < PC; j++) {
						if (cols[j] / 16 != g || prm.degenerate[P0 + j]) continue;
						any = true;
						const float *F = prm.slack + (size_t)(P0 + j) * (M + 1);
						u0 = fmaxf(u0, F[m0]);
						u1 = fmaxf(u1, F[min(m0 + 1, M)]);
					}
					s_slope = (any && M > 0) ? fmaxf(0.0f, __fsub_ru(u1, u0)) : (any ? 0.0f : -1.0f);
				}
				__syncthreads();
				const float slope = s_slope;
				float icpt = 0.0f;
				if (slope >= 0.0f) {
					for (uint32_t m = tid; m <= M; m += blockDim.x) {
						float u = 0.f;
						for (uint32_t j = 0; j < PC; j++) {
							if (cols[j] / 16 != g || prm.degenerate[P0 + j]) continue;
							u = fmaxf(u, prm.slack[(size_t)(P0 + j) * (M + 1) + m]);
						}
						icpt = fmaxf(icpt, __fsub_ru(u, __fmul_rd(slope, (float)m)));
					}
				}
				s_red[tid] = icpt;
				__syncthreads();
				for (uint32_t o = 128; o > 0; o >>= 1) {
					if (tid < o) s_red[tid] = fmaxf(s_red[tid], s_red[tid + o]);
					__syncthreads();
				}
				if (tid == 0) {
					const bool any = slope >= 0.0f;
					prm.group_lines[g * 8 + k] = any ? __fmul_ru(s_red[0], 1.000002f) : 0.0f;
					prm.group_lines[g * 8 + 4 + k] = any ? __fmul_ru(slope, 1.000002f) : 0.0f;
				}
				__syncthreads();
			}
		}
	}
	__syncthreads();
	// 5. group slots: loosest constants of the group's phenotype columns
	if (tid < n_groups) {
		KgFilterGroupConst gc;
		gc.alpha = INFINITY;
		gc.kappa = 0.0f;
		for (uint32_t j = 0; j < PC; j++)
			if (cols[j] / 16 == tid) {
				gc.alpha = fminf(gc.alpha, s_alpha[j]);
				gc.kappa = fmaxf(gc.kappa, s_kappa[j]);
			}
		for (int k = 0; k < 4; k++) {
			gc.line_a[k] = prm.group_lines[tid * 8 + k];
			gc.line_b[k] = prm.group_lines[tid * 8 + 4 + k];
		}
		gc.pad_[0] = gc.pad_[1] = 0.0f;
		prm.gconst[tid] = gc;
		s_gc[tid] = gc;
	}
	__syncthreads();
	// 6. the groups' bounds as a table over the row popcount: the filter's epilogue looks them up
	for (uint32_t e = tid; e < n_groups * (N + 1); e += blockDim.x) {
		const uint32_t g = e / (N + 1), n1 = e - g * (N + 1);
		prm.thr_tab[e] = kg_filter_bound_to_int(kg_filter_group_threshold_n1(s_gc[g], n1, N));
	}
}
