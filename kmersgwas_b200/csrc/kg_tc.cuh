// kg_tc.cuh -- host side of the tcgen05 / TMEM int8 engines (scan filter, kinship Gram).
// Included by kg_abi.cu after kg_ctx, KG_FAIL, KG_CUDA, KG_LAUNCH_CHECK and the timing helpers are defined.
#pragma once
#include <cfloat>

#include "kg_scan_filter.cuh"

static void kg_tc_free(KgTcState *tc) {
	cudaFree(tc->d_pairs); cudaFree(tc->d_yq); cudaFree(tc->d_pconst); cudaFree(tc->d_scratch); cudaFree(tc->d_aligned);
	tc->d_pairs = nullptr; tc->d_yq = nullptr; tc->d_pconst = nullptr; tc->d_scratch = nullptr; tc->d_aligned = nullptr;
	tc->aligned_cap = 0;
}
static bool kg_tc_scan_available(const kg_ctx *c) { return c->tc.scan_ready; }
static bool kg_tc_kinship_available(const kg_ctx *c) { return c->tc.kin_ready; }

// Auto engine choice: the filter pays off once candidate pairs are rare (every pair costs an exact
// 4-lane re-score); while heaps are cold or thresholds low, the dense exact kernel is cheaper.
static bool kg_tc_scan_profitable(const kg_ctx *c) {
	if (!c->tc.scan_ready) return false;
	for (uint32_t p = 0; p < c->n_pheno; p++)
		if (c->h_thr[p] < 0.0) return false;
	return c->tc.use_filter;
}

static float kg_float_down(double x) {  // largest float <= x
	float f = (float)x;
	if ((double)f > x) f = nextafterf(f, -INFINITY);
	return f;
}
static float kg_float_up(double x) {  // smallest float >= x
	float f = (float)x;
	if ((double)f < x) f = nextafterf(f, INFINITY);
	return f;
}

// (alpha, kappa) of every phenotype column from the current thresholds (see kg_scan_filter.cuh header)
static kg_status kg_tc_update_thresholds(kg_ctx *c) {
	KgTcState &tc = c->tc;
	if (!tc.scan_ready) return KG_OK;
	std::vector<float2> pc(tc.p_pad);
	for (uint32_t p = 0; p < tc.p_pad; p++) {
		if (p >= c->n_pheno) { pc[p] = make_float2(INFINITY, 0.0f); continue; }
		const double thr = c->h_thr[p];
		if (tc.degenerate[p] || !(thr >= 0.0) || !std::isfinite(thr)) { pc[p] = make_float2(0.0f, 3.0e38f); continue; }
		const double alpha = std::sqrt(thr) / ((double)c->n_used * tc.scale[p]) * (1.0 - 1e-6);
		pc[p] = make_float2(kg_float_down(alpha), tc.kappa[p]);
	}
	KG_CUDA(c, cudaMemcpyAsync(tc.d_pconst, pc.data(), pc.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	return KG_OK;
}

static kg_status kg_tc_prepare_scan(kg_ctx *c) {
	KgTcState &tc = c->tc;
	tc.scan_ready = false;
	tc.use_filter = false;
	cudaFree(tc.d_yq); cudaFree(tc.d_pconst);
	tc.d_yq = nullptr; tc.d_pconst = nullptr;
	const uint32_t P = c->n_pheno, N = (uint32_t)c->n_used;
	tc.p_pad = (P + 15) / 16 * 16;
	tc.nc = (c->w_file + 1) / 2;
	tc.sbo_b = tc.nc * 1024;
	tc.b_bytes = (tc.p_pad / 8) * tc.sbo_b;
	tc.tcols = 32;
	while (tc.tcols < tc.p_pad) tc.tcols *= 2;
	if (c->min_count < 1) { tc.why_unavailable = "min_count = 0 (rows with an empty group have no finite bound)"; return KG_OK; }
	if (tc.p_pad > 256) { tc.why_unavailable = "more than 256 phenotype columns per pass"; return KG_OK; }
	if (tc.sbo_b > 0x3FFFu * 16) { tc.why_unavailable = "table too wide for the B descriptor stride"; return KG_OK; }
	const size_t smem = kg_filter_smem_bytes(c->w_file, tc.b_bytes, tc.p_pad);
	if (smem > 227u * 1024) {
		tc.why_unavailable = "phenotype tile (P_pad x K_pad int8) does not fit shared memory";
		return KG_OK;
	}
	tc.smem_bytes = smem;

	// quantise: centred, symmetric int8, in FILE column order (unused columns stay 0)
	std::vector<int8_t> img(tc.b_bytes, 0);
	tc.scale.assign(P, 0.0);
	tc.kappa.assign(P, 0.0f);
	tc.degenerate.assign(P, 0);
	const double gamma = (double)(c->nb * 32 + 3) * ldexp(1.0, -24) * 1.001;  // fp32 lane-sum error of the reference
	std::vector<double> cen(N);
	for (uint32_t p = 0; p < P; p++) {
		const float *y = c->h_y.data() + (size_t)p * N;
		const double sum_ref = (double)c->h_sums[p];
		const double ybar = sum_ref / (double)N;
		double amax = 0.0, A = 0.0;
		bool bad = !std::isfinite(sum_ref);
		for (uint32_t i = 0; i < N; i++) {
			const double v = (double)y[i];
			if (!std::isfinite(v) || std::fabs(v) > 1e30) bad = true;
			cen[i] = v - ybar;
			amax = std::max(amax, std::fabs(cen[i]));
			A += std::fabs(v);
		}
		const double s = amax / 127.0;
		if (bad || !(s > 1e-30)) { tc.degenerate[p] = 1; continue; }
		double e_tot = 0.0;
		for (uint32_t i = 0; i < N; i++) {
			double q = std::nearbyint(cen[i] / s);
			q = std::max(-127.0, std::min(127.0, q));
			e_tot += cen[i] - s * q;
			const uint32_t k = c->map_word[i] * 64 + c->map_bit[i];
			img[(size_t)(p % 8) * 16 + (size_t)(p / 8) * tc.sbo_b + (size_t)(k / 16) * 128 + (k % 16)] = (int8_t)q;
		}
		const double t = std::fabs((double)N * ybar - sum_ref);
		// + 1e-9 A: double rounding in the centring; + 1: float32 evaluation slack of the device-side test
		const double kappa = (std::fabs(e_tot) + gamma * A + t + 1e-9 * A) / s + 1.0;
		if (!(kappa < 1e30)) { tc.degenerate[p] = 1; continue; }
		tc.scale[p] = s;
		tc.kappa[p] = kg_float_up(kappa);
	}
	cudaError_t e = cudaMalloc((void **)&tc.d_yq, img.size());
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc quantised phenotypes: %s", cudaGetErrorString(e));
	KG_CUDA(c, cudaMemcpy(tc.d_yq, img.data(), img.size(), cudaMemcpyHostToDevice));
	tc.h_yq_image.swap(img);
	e = cudaMalloc((void **)&tc.d_pconst, tc.p_pad * sizeof(float2));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc filter constants: %s", cudaGetErrorString(e));
	if (!tc.d_pairs) {
		e = cudaMalloc((void **)&tc.d_pairs, tc.pair_capacity * sizeof(uint2));
		if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc candidate pairs: %s", cudaGetErrorString(e));
	}
	KG_CUDA(c, cudaFuncSetAttribute(kg_scan_filter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	KG_CUDA(c, cudaFuncSetAttribute(kg_scan_filter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	tc.scan_ready = true;
	tc.why_unavailable.clear();
	return kg_tc_update_thresholds(c);
}

// the bulk-copy producer needs a 16-byte aligned tile: realign odd device pointers through a scratch copy
static kg_status kg_tc_aligned_tile(kg_ctx *c, const uint64_t *dev, uint64_t n_rows, const uint64_t **out) {
	if (((uintptr_t)dev & 15) == 0) { *out = dev; return KG_OK; }
	const size_t bytes = (size_t)n_rows * (c->w_file + 1) * 8;
	if (c->tc.aligned_cap < bytes) {
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(c->tc.d_aligned);
		c->tc.d_aligned = nullptr;
		c->tc.aligned_cap = 0;
		cudaError_t e = cudaMalloc((void **)&c->tc.d_aligned, bytes);
		if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc aligned tile copy: %s", cudaGetErrorString(e));
		c->tc.aligned_cap = bytes;
	}
	KG_CUDA(c, cudaMemcpyAsync(c->tc.d_aligned, dev, bytes, cudaMemcpyDeviceToDevice, c->stream));
	*out = c->tc.d_aligned;
	return KG_OK;
}

static KgFilterParams kg_tc_filter_params(kg_ctx *c, const uint64_t *dev, uint64_t n_rows) {
	KgTcState &tc = c->tc;
	KgFilterParams f;
	memset(&f, 0, sizeof f);
	f.rows = dev;
	f.n_rows = n_rows;
	f.w_file = c->w_file;
	f.nc = tc.nc;
	f.p_pad = tc.p_pad;
	f.tcols = tc.tcols;
	f.yq_image = tc.d_yq;
	f.b_bytes = tc.b_bytes;
	f.sbo_b = tc.sbo_b;
	f.pconst = tc.d_pconst;
	f.file_mask = c->d_file_mask;
	f.n_used = (uint32_t)c->n_used;
	f.min_count = (uint32_t)std::min<uint64_t>(c->min_count, 0xFFFFFFFFull);
	f.pairs = tc.d_pairs;
	f.n_pairs = c->d_counters + 2;
	f.pair_capacity = tc.pair_capacity;
	f.kept_count = c->d_counters + 1;
	return f;
}

static kg_status kg_tc_scan_tile(kg_ctx *c, const uint64_t *dev_in, uint64_t n_rows, uint64_t first_row_id) {
	KgTcState &tc = c->tc;
	const uint64_t *dev = nullptr;
	kg_status st = kg_tc_aligned_tile(c, dev_in, n_rows, &dev);
	if (st != KG_OK) return st;
	KgFilterParams f = kg_tc_filter_params(c, dev, n_rows);
	const uint32_t n_blocks = (uint32_t)((n_rows + KG_F_ROWS - 1) / KG_F_ROWS);
	const unsigned grid = std::max(1u, std::min<uint32_t>(n_blocks, (uint32_t)c->sm_count));
	timing_begin(c, KG_KERNEL_SCAN_FILTER, n_rows);
	kg_scan_filter_kernel<0><<<grid, KG_F_THREADS, tc.smem_bytes, c->stream>>>(f);
	timing_end(c);
	KG_LAUNCH_CHECK(c);

	KgPairParams pp;
	memset(&pp, 0, sizeof pp);
	pp.raw = KgRowView{dev, n_rows, c->w_file + 1, c->w_file};
	pp.nb = c->nb;
	pp.n_used = (uint32_t)c->n_used;
	pp.n_pheno = c->n_pheno;
	pp.min_count = f.min_count;
	pp.y_lane = c->d_y_lane;
	pp.sums = c->d_sums;
	pp.map_lane = c->d_map_lane;
	pp.file_mask = c->d_file_mask;
	pp.thr = c->d_thr;
	pp.pairs = tc.d_pairs;
	pp.n_pairs = c->d_counters + 2;
	pp.pair_capacity = tc.pair_capacity;
	pp.pair_begin = c->d_counters + 3;
	pp.hits = c->d_hits;
	pp.hit_count = c->d_counters + 0;
	pp.hit_capacity = c->hit_capacity;
	pp.first_row_id = first_row_id;
	timing_begin(c, KG_KERNEL_SCAN_REFINE, 0);
	kg_scan_pairs_kernel<<<(unsigned)c->sm_count * 4, 256, 0, c->stream>>>(pp);
	timing_end(c);
	KG_LAUNCH_CHECK(c);
	kg_scan_pairs_advance_kernel<<<1, 1, 0, c->stream>>>(c->d_counters + 2, c->d_counters + 3);
	KG_LAUNCH_CHECK(c);
	return KG_OK;
}

// testing aid: exact int32 sums Q[row][p] of the filter and the int8 phenotype quantisation behind them
static kg_status kg_tc_filter_debug(kg_ctx *c, const uint64_t *dev_in, uint64_t n_rows, int32_t *q_host, int8_t *yq_host) {
	KgTcState &tc = c->tc;
	if (!tc.scan_ready) KG_FAIL(c, KG_ERR_INVALID, "tensor filter engine unavailable: %s", tc.why_unavailable.c_str());
	const uint64_t *dev = nullptr;
	kg_status st = kg_tc_aligned_tile(c, dev_in, n_rows, &dev);
	if (st != KG_OK) return st;
	KgFilterParams f = kg_tc_filter_params(c, dev, n_rows);
	int32_t *d_q = nullptr;
	cudaError_t e = cudaMalloc((void **)&d_q, (size_t)n_rows * tc.p_pad * sizeof(int32_t));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc debug sums: %s", cudaGetErrorString(e));
	f.q_out = d_q;
	unsigned long long dummy_kept_host = 0;
	(void)dummy_kept_host;
	f.kept_count = c->d_counters + 4;  // scratch counter: the debug pass must not change rows_kept
	const uint32_t n_blocks = (uint32_t)((n_rows + KG_F_ROWS - 1) / KG_F_ROWS);
	const unsigned grid = std::max(1u, std::min<uint32_t>(n_blocks, (uint32_t)c->sm_count));
	kg_scan_filter_kernel<1><<<grid, KG_F_THREADS, tc.smem_bytes, c->stream>>>(f);
	c->launches++;
	cudaError_t e1 = cudaGetLastError();
	cudaError_t e2 = cudaStreamSynchronize(c->stream);
	std::vector<int32_t> q((size_t)n_rows * tc.p_pad);
	cudaError_t e3 = cudaMemcpy(q.data(), d_q, q.size() * sizeof(int32_t), cudaMemcpyDeviceToHost);
	cudaFree(d_q);
	KG_CUDA(c, e1); KG_CUDA(c, e2); KG_CUDA(c, e3);
	for (uint64_t r = 0; r < n_rows; r++)
		for (uint32_t p = 0; p < c->n_pheno; p++) q_host[r * c->n_pheno + p] = q[r * tc.p_pad + p];
	if (yq_host) {
		const uint32_t kpad = 64 * c->w_file;
		for (uint32_t p = 0; p < c->n_pheno; p++)
			for (uint32_t k = 0; k < kpad; k++)
				yq_host[(size_t)p * kpad + k] =
				    tc.h_yq_image[(size_t)(p % 8) * 16 + (size_t)(p / 8) * tc.sbo_b + (size_t)(k / 16) * 128 + (k % 16)];
	}
	return KG_OK;
}

static kg_status kg_tc_prepare_kinship(kg_ctx *c) { (void)c; return KG_OK; }
static kg_status kg_tc_kinship_tile(kg_ctx *c, const KgRowView &view) {
	(void)view;
	KG_FAIL(c, KG_ERR_INVALID, "tensor kinship engine not built");
}
