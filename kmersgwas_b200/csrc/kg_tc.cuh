// kg_tc.cuh -- host side of the tcgen05 / TMEM int8 engines (scan filter, kinship Gram).
// Included by kg_abi.cu after kg_ctx, KG_FAIL, KG_CUDA, KG_LAUNCH_CHECK and the timing helpers are defined.
#pragma once
#include <cfloat>

#include "kg_filter_retune.cuh"
#include "kg_kinship_tc.cuh"
#include "kg_scan_filter.cuh"

static void kg_tc_free(KgTcState *tc) {
	cudaFree(tc->d_row_list); cudaFree(tc->d_group_list); cudaFree(tc->d_group_count); cudaFree(tc->d_tile_pheno);
	cudaFree(tc->d_ent_q); cudaFree(tc->d_ent_n1); cudaFree(tc->d_pairs); cudaFree(tc->d_slack);
	tc->d_ent_q = nullptr; tc->d_ent_n1 = nullptr; tc->d_pairs = nullptr; tc->d_slack = nullptr; tc->qcap = 0;
	tc->d_group_list = nullptr; tc->d_group_count = nullptr; tc->d_tile_pheno = nullptr;
	cudaFree(tc->d_yq); cudaFree(tc->d_gconst); cudaFree(tc->d_kin_groups); cudaFree(tc->d_kin_delta); cudaFree(tc->d_kin_ctas);
	tc->d_kin_groups = nullptr; tc->d_kin_delta = nullptr; tc->d_kin_ctas = nullptr; cudaFree(tc->d_scratch); cudaFree(tc->d_aligned);
	tc->d_row_list = nullptr; tc->d_yq = nullptr; tc->d_gconst = nullptr; tc->d_scratch = nullptr; tc->d_aligned = nullptr;
	tc->aligned_cap = 0;
	tc->row_list_cap = 0;
	cudaFree(tc->d_scale); cudaFree(tc->d_kappa0); cudaFree(tc->d_degenerate); cudaFree(tc->d_q); cudaFree(tc->d_kidx);
	cudaFree(tc->d_col_of); cudaFree(tc->d_group_lines); cudaFree(tc->d_thr_tab);
	tc->d_thr_tab = nullptr;
	tc->d_scale = nullptr; tc->d_kappa0 = nullptr; tc->d_degenerate = nullptr; tc->d_q = nullptr; tc->d_kidx = nullptr;
	tc->d_col_of = nullptr; tc->d_group_lines = nullptr;
}
static bool kg_tc_scan_available(const kg_ctx *c) { return c->tc.scan_ready; }
static bool kg_tc_kinship_available(const kg_ctx *c) { return c->tc.kin_ready; }

// Auto engine choice (host-driven mode): the filter needs a threshold for every phenotype (heaps full); it then pays off
// unless it fails to rule out most rows (every listed row is re-scored by the exact kernel anyway).
static bool kg_tc_scan_profitable(const kg_ctx *c) {
	if (!c->tc.scan_ready) return false;
	for (uint32_t p = 0; p < c->n_pheno; p++)
		if (c->h_thr[p] < 0.0) return false;
	return c->tc.use_filter;
}

// byte offset of B element (column n, K index k) in the shared-memory image (K-major core matrices, no swizzle)
static size_t kg_tc_b_offset(const KgTcState &tc, uint32_t n, uint32_t k) {
	return (size_t)(n % 8) * 16 + (size_t)(n / 8) * tc.sbo_b + (size_t)(k / 16) * 128 + (k % 16);
}

static float kg_float_up(double x) {  // smallest float >= x
	float f = (float)x;
	if ((double)f < x) f = nextafterf(f, INFINITY);
	return f;
}

// Bound constants of the tensor filter from the thresholds in d_thr: per-phenotype (alpha, kappa), the column order
// sorted by alpha, the B operand image in that order, the per-group loosest constants and slack tangents.  All of it
// is computed ON THE DEVICE (kg_filter_retune_kernel, one CTA per pass), stream-ordered behind whatever wrote d_thr
// (kg_scan_set_thresholds in host-driven mode, the heap replay in device-selection mode).
static kg_status kg_tc_retune(kg_ctx *c, bool force) {
	KgTcState &tc = c->tc;
	if (!tc.scan_ready) return KG_OK;
	KgRetuneParams r;
	memset(&r, 0, sizeof r);
	r.n_pheno = c->n_pheno;
	r.n_used = (uint32_t)c->n_used;
	r.p_pad = tc.p_pad;
	r.sbo_b = tc.sbo_b;
	r.b_bytes = tc.b_bytes;
	r.m_half = (uint32_t)c->n_used / 2;
	r.cols_per_pass = tc.cols_per_pass;
	r.thr = c->d_thr;
	r.scale = tc.d_scale;
	r.kappa0 = tc.d_kappa0;
	r.degenerate = tc.d_degenerate;
	r.q = tc.d_q;
	r.kidx = tc.d_kidx;
	r.slack = tc.d_slack;
	r.col_of = tc.d_col_of;
	r.group_lines = tc.d_group_lines;
	r.yq_image = tc.d_yq;
	r.tile_pheno = tc.d_tile_pheno;
	r.gconst = tc.d_gconst;
	r.thr_tab = tc.d_thr_tab;
	r.alpha_out = reinterpret_cast<float *>(tc.d_gconst + 16 * (size_t)tc.n_pass);
	r.kappa_out = r.alpha_out + c->n_pheno;
	r.status = c->sel.active ? c->sel.d_status : nullptr;
	r.force = force ? 1u : 0u;
	kg_filter_retune_kernel<<<tc.n_pass, 256, 0, c->stream>>>(r);
	KG_LAUNCH_CHECK(c);
	return KG_OK;
}

// the filter kernel instance for a role split / mode (kg_scan_filter.cuh)
typedef void (*KgFilterKernel)(const KgFilterParams);
template <int SPLIT>
static KgFilterKernel kg_filter_kernel_pick(int mode) {
	constexpr int E = kg_filter_split_nexp(SPLIT), A = kg_filter_split_nacc(SPLIT);
	return mode ? kg_scan_filter_kernel<1, E, A> : kg_scan_filter_kernel<0, E, A>;
}
static KgFilterKernel kg_filter_kernel_of(int split, int mode) {
	return split == 2 ? kg_filter_kernel_pick<2>(mode) : split == 1 ? kg_filter_kernel_pick<1>(mode) : kg_filter_kernel_pick<0>(mode);
}

static kg_status kg_tc_prepare_scan(kg_ctx *c) {
	KgTcState &tc = c->tc;
	tc.scan_ready = false;
	tc.use_filter = true;
	cudaFree(tc.d_yq); cudaFree(tc.d_gconst);
	tc.d_yq = nullptr; tc.d_gconst = nullptr;
	// the filter's list buffers are sized for the phenotype count they were allocated with ([p_pad / 16][capacity]):
	// The list buffers ([groups][capacity]) survive a new phenotype set: kg_tc_ensure_row_list reallocates them when the
	// new set needs more 16-column groups than they were sized for (tc.row_list_groups), or a longer tile arrives.
	cudaFree(tc.d_scale); cudaFree(tc.d_kappa0); cudaFree(tc.d_degenerate); cudaFree(tc.d_q); cudaFree(tc.d_kidx);
	cudaFree(tc.d_col_of); cudaFree(tc.d_group_lines); cudaFree(tc.d_thr_tab);
	tc.d_thr_tab = nullptr;
	tc.d_scale = nullptr; tc.d_kappa0 = nullptr; tc.d_degenerate = nullptr; tc.d_q = nullptr; tc.d_kidx = nullptr;
	tc.d_col_of = nullptr; tc.d_group_lines = nullptr;
	const uint32_t P = c->n_pheno, N = (uint32_t)c->n_used;
	if (c->min_count < 1) { tc.why_unavailable = "min_count = 0 (rows with an empty group have no finite bound)"; return KG_OK; }
	tc.sbo_b = ((c->w_file + 1) / 2) * 1024;       // K_pad = 128 * ceil(w_file / 2) bytes per B column, 1024 B per 128
	if (tc.sbo_b > 0x3FFFu * 16) { tc.why_unavailable = "table too wide for the B descriptor stride"; return KG_OK; }
	// Shape of a pass: P_pad = 16 .. 128 accumulator columns (column 0 = all-ones, so up to 127 phenotypes) whose B image
	// (P_pad x K_pad int8) must fit shared memory next to 2 .. 4 raw row-block stages.  More phenotypes than a pass holds
	// -> several passes over the same tile (BASELINE config 5: 1001 phenotypes; wide tables: few columns per pass).
	const uint32_t pp_want = std::min<uint32_t>(128, (P + 1 + 15) / 16 * 16);
	uint32_t best_pp = 0, best_rs = 0;
	for (uint32_t rs = KG_F_RAW_STAGES; rs >= 2; rs--)
		for (uint32_t pp = pp_want; pp >= 16; pp -= 16)
			if (kg_filter_smem_bytes(c->w_file, (pp / 8) * tc.sbo_b, pp, rs, N) <= 227u * 1024) {
				if (pp > best_pp) { best_pp = pp; best_rs = rs; }
				break;
			}
	if (best_pp == 0) { tc.why_unavailable = "a 16-column phenotype tile (P_pad x K_pad int8) does not fit shared memory"; return KG_OK; }
	tc.n_pass = (P + (best_pp - 1) - 1) / (best_pp - 1);
	tc.cols_per_pass = (P + tc.n_pass - 1) / tc.n_pass;
	tc.p_pad = (tc.cols_per_pass + 1 + 15) / 16 * 16;
	tc.raw_stages = best_rs;
	tc.b_bytes = (tc.p_pad / 8) * tc.sbo_b;
	tc.tcols = tc.p_pad;
	// role split of the kernel (kg_scan_filter.cuh): the narrower the table, the more accumulator buffers / epilogue sets
	tc.split = 0;
	for (int sp = KG_F_SPLITS - 1; sp > 0; sp--)
		if ((int)c->w_file <= kg_filter_split_max_w(sp) &&
		    kg_filter_split_nacc(sp) * tc.p_pad + 2 * 16 * std::min<uint32_t>(c->w_file, 2) <= KG_F_TMEM_COLS) { tc.split = sp; break; }
#ifdef KG_PERF_SWITCHES
	if (const char *e = getenv("KG_FILTER_SPLIT")) tc.split = std::max(0, std::min(KG_F_SPLITS - 1, atoi(e)));
#endif
	// tensor memory: the accumulator buffers + the A stages (16 columns per presence word); as few, as large stages as fit
	{
		const uint32_t KG_F_NSUB = (uint32_t)kg_filter_split_nexp(tc.split) / 4;
		const uint32_t a_cols = KG_F_TMEM_COLS - (uint32_t)kg_filter_split_nacc(tc.split) * tc.p_pad;
		tc.a_words = std::max(1u, std::min<uint32_t>(c->w_file, a_cols / 32));           // at least two stages (few large stages measured best)
		tc.a_words = std::min<uint32_t>(tc.a_words, KG_F_MAX_WPT * KG_F_NSUB);            // register budget of the expanders
#ifdef KG_PERF_SWITCHES   // perf experiments only (profiles/filter_dbg_sweep.sh builds with -DKG_PERF_SWITCHES)
		if (const char *e = getenv("KG_FILTER_A_WORDS")) {
			const uint32_t v = (uint32_t)atoi(e);
			if (v >= 1 && v <= c->w_file && v <= a_cols / 32 && v <= KG_F_MAX_WPT * KG_F_NSUB) tc.a_words = v;
		}
#endif
		tc.a_stages = std::max(2u, std::min<uint32_t>(KG_F_MAX_A_STAGES, a_cols / (16 * tc.a_words)));
		tc.nc = (c->w_file + tc.a_words - 1) / tc.a_words;
	}
	const size_t smem = kg_filter_smem_bytes(c->w_file, tc.b_bytes, tc.p_pad, tc.raw_stages, N);
	tc.smem_bytes = smem;

	// quantise: centred, symmetric int8 per phenotype
	tc.h_q.assign((size_t)P * N, 0);
	tc.scale.assign(P, 0.0);
	tc.kappa.assign(P, 0.0f);
	tc.degenerate.assign(P, 0);
	const double gamma = (double)(c->nb * 32 + 3) * ldexp(1.0, -24) * 1.001;  // fp32 lane-sum error of the reference
	std::vector<double> cen(N);
	// F_p(m), m = 0 .. N/2: the largest |sum of e_i / s| any m samples can have
	const uint32_t Mhalf = N / 2;
	tc.slack_table.assign((size_t)P * (Mhalf + 1), 0.0f);
	std::vector<double> epos, eneg;
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
		int8_t *qrow = tc.h_q.data() + (size_t)p * N;
		epos.clear();
		eneg.clear();
		for (uint32_t i = 0; i < N; i++) {
			double q = std::nearbyint(cen[i] / s);
			q = std::max(-127.0, std::min(127.0, q));
			const double e = cen[i] - s * q;
			e_tot += e;
			qrow[i] = (int8_t)q;
			if (e > 0) epos.push_back(e / s); else if (e < 0) eneg.push_back(-e / s);
		}
		{
			std::sort(epos.begin(), epos.end(), std::greater<double>());
			std::sort(eneg.begin(), eneg.end(), std::greater<double>());
			float *F = tc.slack_table.data() + (size_t)p * (Mhalf + 1);
			double sp = 0.0, sn = 0.0;
			for (uint32_t m = 1; m <= Mhalf; m++) {
				if (m <= epos.size()) sp += epos[m - 1];
				if (m <= eneg.size()) sn += eneg[m - 1];
				F[m] = kg_float_up(std::max(sp, sn) * (1.0 + 1e-9) + 1e-9 * (double)m);  // + double rounding of e_i
			}
		}
		const double t = std::fabs((double)N * ybar - sum_ref);
		// + 1e-9 A: double rounding in the centring; + 1: float32 evaluation slack of the device-side test
		const double kappa = (std::fabs(e_tot) + gamma * A + t + 1e-9 * A) / s + 1.0;
		if (!(kappa < 1e30)) {
			tc.degenerate[p] = 1;
			std::fill(qrow, qrow + N, (int8_t)0);
			continue;
		}
		tc.scale[p] = s;
		tc.kappa[p] = kg_float_up(kappa);
	}
	cudaError_t e = cudaMalloc((void **)&tc.d_yq, (size_t)tc.n_pass * tc.b_bytes);
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc quantised phenotypes: %s", cudaGetErrorString(e));
	cudaFree(tc.d_tile_pheno); cudaFree(tc.d_group_count);
	tc.d_tile_pheno = nullptr; tc.d_group_count = nullptr;
	e = cudaMalloc((void **)&tc.d_tile_pheno, (size_t)tc.n_pass * tc.p_pad * sizeof(int32_t));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc tile table: %s", cudaGetErrorString(e));
	e = cudaMalloc((void **)&tc.d_group_count, (16 + 16) * sizeof(unsigned long long));   // + 32 u32 tile chunk counters
	if (e == cudaSuccess) e = cudaMemset(tc.d_group_count, 0, (16 + 16) * sizeof(unsigned long long));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc group counters: %s", cudaGetErrorString(e));
	e = cudaMalloc((void **)&tc.d_gconst, (size_t)tc.n_pass * 16 * sizeof(KgFilterGroupConst) + 2 * (size_t)P * sizeof(float));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc filter constants: %s", cudaGetErrorString(e));
	cudaFree(tc.d_slack);
	tc.d_slack = nullptr;
	static_assert(KG_F_ONE == 1, "the slack table is uploaded in accumulator units");
	e = cudaMalloc((void **)&tc.d_slack, tc.slack_table.size() * sizeof(float));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc slack table: %s", cudaGetErrorString(e));
	KG_CUDA(c, cudaMemcpyAsync(tc.d_slack, tc.slack_table.data(), tc.slack_table.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
	{
		// device twins of the per-phenotype tables (kg_filter_retune_kernel)
		std::vector<uint32_t> kidx(N);
		for (uint32_t i = 0; i < N; i++) kidx[i] = kg_filter_k_of_column(c->map_word[i] * 64 + c->map_bit[i]);
		std::vector<uint32_t> zero_cols(P, 0u);
		std::vector<float> zero_lines((size_t)tc.n_pass * 16 * 8, 0.0f);
		KG_CUDA(c, dev_alloc_copy(&tc.d_scale, tc.scale));
		KG_CUDA(c, dev_alloc_copy(&tc.d_kappa0, tc.kappa));
		KG_CUDA(c, dev_alloc_copy(&tc.d_degenerate, tc.degenerate));
		KG_CUDA(c, dev_alloc_copy(&tc.d_q, tc.h_q));
		KG_CUDA(c, dev_alloc_copy(&tc.d_kidx, kidx));
		KG_CUDA(c, dev_alloc_copy(&tc.d_col_of, zero_cols));
		KG_CUDA(c, dev_alloc_copy(&tc.d_group_lines, zero_lines));
		KG_CUDA(c, cudaMalloc((void **)&tc.d_thr_tab, (size_t)tc.n_pass * kg_filter_tab_floats(tc.p_pad, N) * sizeof(float)));
	}
	tc.use_pairs = true;
	tc.dbg_flags = 0u;
	tc.n_issuers = 0u;
	tc.print_stats = false;
#ifdef KG_PERF_SWITCHES   // perf experiments only: a production build never reads these (results are wrong with KG_FILTER_DEBUG)
	tc.use_pairs = !(getenv("KG_FILTER_NO_PAIRS") && atoi(getenv("KG_FILTER_NO_PAIRS")));
	tc.dbg_flags = getenv("KG_FILTER_DEBUG") ? (uint32_t)atoi(getenv("KG_FILTER_DEBUG")) : 0u;
	tc.n_issuers = getenv("KG_FILTER_ISSUERS") ? (uint32_t)std::max(1, std::min(KG_F_MMA_WARPS, atoi(getenv("KG_FILTER_ISSUERS")))) : 0u;
	tc.print_stats = getenv("KG_FILTER_STATS") != nullptr;
#endif
	KG_CUDA(c, cudaFuncSetAttribute(kg_filter_kernel_of(tc.split, 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	KG_CUDA(c, cudaFuncSetAttribute(kg_filter_kernel_of(tc.split, 1), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	tc.scan_ready = true;
	tc.why_unavailable.clear();
	// first images: every threshold is -1 (d_thr as kg_scan_set_phenotypes left it) -> alpha = 0, every kept row listed
	kg_status st = kg_tc_retune(c, true);
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	return st;
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

static KgFilterParams kg_tc_filter_params(kg_ctx *c, const uint64_t *dev, uint64_t n_rows, uint32_t pass) {
	KgTcState &tc = c->tc;
	KgFilterParams f;
	memset(&f, 0, sizeof f);
	f.rows = dev;
	f.n_rows = n_rows;
	f.w_file = c->w_file;
	f.nc = tc.nc;
	f.a_words = tc.a_words;
	f.raw_stages = tc.raw_stages;
	f.a_stages = tc.a_stages;
	f.p_pad = tc.p_pad;
	f.tcols = tc.tcols;
	f.yq_image = tc.d_yq + (size_t)pass * tc.b_bytes;
	f.b_bytes = tc.b_bytes;
	f.sbo_b = tc.sbo_b;
	f.gconst = tc.d_gconst + (size_t)pass * 16;
	f.thr_tab = tc.d_thr_tab + (size_t)pass * kg_filter_tab_floats(tc.p_pad, (uint32_t)c->n_used);
	f.n_used = (uint32_t)c->n_used;
	f.min_count = (uint32_t)std::min<uint64_t>(c->min_count, 0xFFFFFFFFull);
	f.row_list = tc.d_row_list;
	f.n_listed = c->d_counters + 2;
	f.group_list = tc.d_group_list;
	f.group_count = tc.d_group_count;
	f.group_cap = tc.row_list_cap;
	f.ent_q = tc.d_ent_q;
	f.ent_n1 = tc.d_ent_n1;
	f.qcap = tc.use_pairs ? tc.qcap : 0;
	// kept rows are counted once per tile (pass 0); the other passes count into a scratch word
	f.kept_count = pass != 0 ? c->d_counters + 4 : (c->sel.active ? c->sel.d_status + KG_SEL_ST_ROUND_KEPT : c->d_counters + 1);
	f.n_issuers = tc.n_issuers ? tc.n_issuers : KG_F_MMA_WARPS;
	f.dbg = tc.dbg_flags;
	return f;
}

static kg_status kg_tc_ensure_row_list(kg_ctx *c, uint64_t n_rows) {
	KgTcState &tc = c->tc;
	const uint32_t n_groups_now = tc.p_pad / 16;
	if (tc.row_list_cap >= n_rows && tc.row_list_groups >= n_groups_now) return KG_OK;
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaFree(tc.d_row_list);
	cudaFree(tc.d_group_list);
	tc.d_row_list = nullptr;
	tc.d_group_list = nullptr;
	// Capacity grows by doubling: the device selection submits rounds that grow 1.5x each in the cold phase, and a
	// reallocation (stream sync + cudaFree + cudaMalloc) per round was part of every job's first 10^8 rows
	uint64_t want = std::max<uint64_t>({n_rows, tc.row_list_cap, (uint64_t)1 << 20});
	if (n_rows > tc.row_list_cap && tc.row_list_cap) want = std::max<uint64_t>(want, 2 * tc.row_list_cap);
	tc.row_list_cap = 0;
	tc.row_list_groups = std::max(tc.row_list_groups, n_groups_now);
	const uint64_t cap = want;
	cudaError_t e = cudaMalloc((void **)&tc.d_row_list, cap * sizeof(uint32_t));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc filter row list: %s", cudaGetErrorString(e));
	e = cudaMalloc((void **)&tc.d_group_list, (size_t)tc.row_list_groups * cap * sizeof(uint32_t));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc filter group lists: %s", cudaGetErrorString(e));
	tc.row_list_cap = cap;
	// entries that carry their accumulators: the first qcap of every group list (longer lists stay in list mode)
	cudaFree(tc.d_ent_q); cudaFree(tc.d_ent_n1); cudaFree(tc.d_pairs);
	tc.d_ent_q = nullptr; tc.d_ent_n1 = nullptr; tc.d_pairs = nullptr;
	// measured (B200, P = 101): a pair costs ~1.3 ns, a list-mode entry ~3.9 ns, and the per-column test leaves about
	// one pair per entry, so pair mode wins for every list that is not most of the tile
	tc.qcap = std::max<uint64_t>(4096, cap / 16);
	const size_t n_groups = tc.row_list_groups;
	e = cudaMalloc((void **)&tc.d_ent_q, n_groups * tc.qcap * 16 * sizeof(int32_t));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc filter entry sums: %s", cudaGetErrorString(e));
	e = cudaMalloc((void **)&tc.d_ent_n1, n_groups * tc.qcap * sizeof(uint32_t));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc filter entry counts: %s", cudaGetErrorString(e));
	e = cudaMalloc((void **)&tc.d_pairs, n_groups * tc.qcap * 16 * sizeof(uint2));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc pair list: %s", cudaGetErrorString(e));
	return KG_OK;
}

// End of a filtered tile pass: interval totals += the tile's counts ([5] listed rows, [6] (row, group) entries, [7] (row,
// phenotype) pairs), then the per-tile counters are zeroed for the next pass / tile (no memsets in front of the filter launch).
__global__ void kg_tile_end_kernel(unsigned long long *interval_cnt, unsigned long long *tile_cnt, uint32_t n_groups) {
	if (threadIdx.x == 0) {
		unsigned long long t = 0;
		for (uint32_t i = 0; i < n_groups; i++) t += tile_cnt[i];
		interval_cnt[5] += interval_cnt[2];
		interval_cnt[6] += t;
		interval_cnt[7] += tile_cnt[24];
		interval_cnt[2] = 0;
	}
	__syncthreads();
	if (threadIdx.x < 32) tile_cnt[threadIdx.x] = 0;
}

// filter the tile on the tensor cores (one pass per <= 127 phenotype columns), then re-score the (row, phenotype) pairs
// it could not rule out with the exact kernels
static kg_status kg_tc_scan_tile(kg_ctx *c, const uint64_t *dev_in, uint64_t n_rows, uint64_t first_row_id) {
	KgTcState &tc = c->tc;
	const uint64_t *dev = nullptr;
	kg_status st = kg_tc_aligned_tile(c, dev_in, n_rows, &dev);
	if (st != KG_OK) return st;
	st = kg_tc_ensure_row_list(c, n_rows);
	if (st != KG_OK) return st;
	if (!c->identity) {
		st = ensure_squeeze_scratch(c, n_rows);
		if (st != KG_OK) return st;
	}
	const uint32_t n_blocks = (uint32_t)((n_rows + KG_F_ROWS - 1) / KG_F_ROWS);
	const unsigned grid = std::max(1u, std::min<uint32_t>(n_blocks, (uint32_t)c->sm_count));
	const float *alpha_all = reinterpret_cast<const float *>(tc.d_gconst + 16 * (size_t)tc.n_pass);
	for (uint32_t pass = 0; pass < tc.n_pass; pass++) {
		// per-tile counters (d_counters[2], tc.d_group_count[0..31]) are zero here: zeroed at allocation / interval open and
		// by kg_tile_end_kernel / kg_select_round_end_kernel after every pass
		KgFilterParams f = kg_tc_filter_params(c, dev, n_rows, pass);
		timing_begin(c, KG_KERNEL_SCAN_FILTER, n_rows);
		kg_filter_kernel_of(tc.split, 0)<<<grid, KG_F_THREADS, tc.smem_bytes, c->stream>>>(f);
		timing_end(c);
		KG_LAUNCH_CHECK(c);

		KgRowView view{dev, n_rows, c->w_file + 1, c->w_file};
		uint32_t compact = 0;
		if (!c->identity) {
			// memory-order copies of the listed rows only
			const unsigned sg = (unsigned)c->sm_count * 8;
			timing_begin(c, KG_KERNEL_AUX, 0);
			kg_squeeze_kernel<<<sg, 256, 0, c->stream>>>(view, c->d_map_mem, (uint32_t)c->n_used, c->w_mem, c->d_squeezed,
			                                            tc.d_row_list, c->d_counters + 2);
			timing_end(c);
			KG_LAUNCH_CHECK(c);
			view = KgRowView{c->d_squeezed, n_rows, c->w_mem + 1, c->w_mem};
			compact = 1;
		}
		KgScanParams prm = scan_params(c, view, first_row_id);
		prm.row_list = tc.d_row_list;
		prm.group_list = tc.d_group_list;
		prm.group_count = tc.d_group_count;
		prm.group_cap = tc.row_list_cap;
		prm.tile_pheno = tc.d_tile_pheno + (size_t)pass * tc.p_pad;
		prm.tile_chunk_counter = reinterpret_cast<unsigned int *>(tc.d_group_count + 16);
		prm.list_compact = compact;
		timing_begin(c, KG_KERNEL_SCAN_REFINE, 0);
		if (tc.use_pairs && tc.pair_limit != 0) {
			// short group lists: per-column re-test, then one phenotype per (row, phenotype) pair
			unsigned long long *pair_count = tc.d_group_count + 24;   // [24] pairs [25] overflow flag (both zeroed above)
			KgPairSelectParams ps;
			memset(&ps, 0, sizeof ps);
			ps.group_count = tc.d_group_count;
			ps.group_list = tc.d_group_list;
			ps.group_cap = tc.row_list_cap;
			ps.ent_q = tc.d_ent_q;
			ps.ent_n1 = tc.d_ent_n1;
			ps.qcap = tc.qcap;
			const uint64_t dense_limit = tc.pair_limit < 0 ? tc.qcap : std::min<uint64_t>(tc.qcap, (uint64_t)tc.pair_limit);
			ps.dense_limit = dense_limit;
			ps.n_groups = tc.p_pad / 16;
			ps.n_used = (uint32_t)c->n_used;
			ps.tile_pheno = prm.tile_pheno;
			ps.alpha = alpha_all;
			ps.kappa = ps.alpha + c->n_pheno;
			ps.slack = tc.d_slack;
			ps.pairs = tc.d_pairs;
			ps.pair_count = pair_count;
			ps.pair_cap = (uint64_t)ps.n_groups * tc.qcap * 16;
			ps.overflow = pair_count + 1;
			const uint64_t entries = std::min<uint64_t>((uint64_t)ps.n_groups * tc.qcap, (uint64_t)ps.n_groups * n_rows);
			const unsigned sel_grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((entries + 255) / 256, (uint64_t)c->sm_count * 8));
			kg_pair_select_kernel<<<sel_grid, 256, 0, c->stream>>>(ps);
			KG_LAUNCH_CHECK(c);
			prm.dense_limit = dense_limit;
			prm.pairs = tc.d_pairs;
			prm.pair_count = pair_count;
			kg_scan_pair_kernel<<<(unsigned)c->sm_count * 8, 256, 0, c->stream>>>(prm);
			KG_LAUNCH_CHECK(c);
		}
		st = launch_exact_list(c, prm, tc.p_pad / 8);
		timing_end(c);
		if (st != KG_OK) return st;
		if (tc.print_stats) {   // diagnosis only: synchronises the stream
			unsigned long long h[32];
			cudaStreamSynchronize(c->stream);
			cudaMemcpy(h, tc.d_group_count, sizeof h, cudaMemcpyDeviceToHost);
			fprintf(stderr, "[kg filter] pass %u rows %llu qcap %llu groups:", pass, (unsigned long long)n_rows, (unsigned long long)tc.qcap);
			for (uint32_t g = 0; g < tc.p_pad / 16; g++) fprintf(stderr, " %llu", h[g]);
			fprintf(stderr, "  pairs %llu\n", h[24]);
		}
		if (c->sel.active && pass + 1 == tc.n_pass) return kg_sel_finish_round(c, n_rows, first_row_id, true);
		kg_tile_end_kernel<<<1, 32, 0, c->stream>>>(c->d_counters, tc.d_group_count, tc.p_pad / 16);
		KG_LAUNCH_CHECK(c);
	}
	return KG_OK;
}

// testing aid: exact int32 sums Q[row][p] of the filter and the int8 phenotype quantisation behind them
static kg_status kg_tc_filter_debug(kg_ctx *c, const uint64_t *dev_in, uint64_t n_rows, int32_t *q_host, int8_t *yq_host) {
	KgTcState &tc = c->tc;
	if (!tc.scan_ready) KG_FAIL(c, KG_ERR_INVALID, "tensor filter engine unavailable: %s", tc.why_unavailable.c_str());
	const uint64_t *dev = nullptr;
	kg_status st = kg_tc_aligned_tile(c, dev_in, n_rows, &dev);
	if (st != KG_OK) return st;
	st = kg_tc_ensure_row_list(c, n_rows);
	if (st != KG_OK) return st;
	int32_t *d_q = nullptr;
	cudaError_t e = cudaMalloc((void **)&d_q, (size_t)n_rows * tc.p_pad * sizeof(int32_t));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc debug sums: %s", cudaGetErrorString(e));
	// the column assignment and the operand images as the device holds them (the re-tune kernel wrote them)
	std::vector<uint32_t> col_of(c->n_pheno);
	std::vector<int8_t> image((size_t)tc.n_pass * tc.b_bytes);
	std::vector<int32_t> q((size_t)n_rows * tc.p_pad);
	const uint32_t n_blocks = (uint32_t)((n_rows + KG_F_ROWS - 1) / KG_F_ROWS);
	const unsigned grid = std::max(1u, std::min<uint32_t>(n_blocks, (uint32_t)c->sm_count));
	cudaError_t e1 = cudaStreamSynchronize(c->stream);
	cudaError_t e2 = cudaMemcpy(col_of.data(), tc.d_col_of, col_of.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost);
	cudaError_t e3 = cudaMemcpy(image.data(), tc.d_yq, image.size(), cudaMemcpyDeviceToHost);
	for (uint32_t pass = 0; pass < tc.n_pass && e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess; pass++) {
		KgFilterParams f = kg_tc_filter_params(c, dev, n_rows, pass);
		f.q_out = d_q;
		f.kept_count = c->d_counters + 4;  // scratch counter: the debug pass must not change rows_kept
		kg_filter_kernel_of(tc.split, 1)<<<grid, KG_F_THREADS, tc.smem_bytes, c->stream>>>(f);
		c->launches++;
		e1 = cudaGetLastError();
		if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(c->stream);
		if (e1 == cudaSuccess) e1 = cudaMemcpy(q.data(), d_q, q.size() * sizeof(int32_t), cudaMemcpyDeviceToHost);
		if (e1 != cudaSuccess) break;
		const uint32_t p0 = pass * tc.cols_per_pass, p1 = std::min(c->n_pheno, p0 + tc.cols_per_pass);
		for (uint64_t r = 0; r < n_rows; r++)
			for (uint32_t p = p0; p < p1; p++) q_host[r * c->n_pheno + p] = q[r * tc.p_pad + col_of[p]];
	}
	cudaFree(d_q);
	KG_CUDA(c, e1); KG_CUDA(c, e2); KG_CUDA(c, e3);
	if (yq_host) {
		const uint32_t kpad = 64 * c->w_file;
		for (uint32_t p = 0; p < c->n_pheno; p++) {
			const int8_t *img = image.data() + (size_t)(p / tc.cols_per_pass) * tc.b_bytes;
			for (uint32_t col = 0; col < kpad; col++)
				yq_host[(size_t)p * kpad + col] = (int8_t)-img[kg_tc_b_offset(tc, col_of[p], kg_filter_k_of_column(col))];
		}
	}
	return KG_OK;
}

// ------------------------------------------------------------------------------------- kinship (tensor engine)
static kg_status kg_tc_prepare_kinship(kg_ctx *c) {
	KgTcState &tc = c->tc;
	const uint32_t ld = 64 * c->w_file;
	tc.kin_dirty = false;
	if (tc.kin_ready && tc.kin_tables_w_file == c->w_file && tc.d_kin_delta) {
		// same shape as the last kg_kinship_begin: the tile-group / CTA tables are still valid, only the delta restarts
		KG_CUDA(c, cudaMemsetAsync(tc.d_kin_delta, 0, ((size_t)ld * ld + 1) * sizeof(unsigned long long), c->stream));
		return KG_OK;
	}
	tc.kin_ready = false;
	tc.kin_tables_w_file = 0;
	const size_t smem = kg_kin_tc_smem_bytes(c->w_file);
	if (smem > 227u * 1024) { tc.why_unavailable = "row block does not fit shared memory next to the operand stages"; return KG_OK; }
	// tile groups over the lower triangle in file column order
	std::vector<KgKinGroup> groups;
	const uint32_t n_i = (ld + 127) / 128;
	for (uint32_t I = 0; I < n_i; I++) {
		const uint32_t n_j = I / 2 + 1;   // 256-sample blocks 0 .. I/2 intersect the columns <= the rows of block I
		for (uint32_t j = 0; j < n_j; j += 2) {
			KgKinGroup g;
			g.i_blk = (int32_t)I;
			g.j2[0] = (int32_t)j;
			g.j2[1] = j + 1 < n_j ? (int32_t)(j + 1) : -1;
			g.a_in = g.j2[0] == (int32_t)(I / 2) ? 0 : (g.j2[1] == (int32_t)(I / 2) ? 1 : -1);
			groups.push_back(g);
		}
	}
	// CTA table.  Up to one wave: every group gets row splits in proportion to its tiles (a 2-tile group does twice the
	// MMAs).  More groups than SMs (wide tables): one CTA per group, the launch runs in several waves.
	std::vector<KgKinCta> ctas;
	if (groups.size() > (size_t)c->sm_count) {
		for (uint32_t gi = 0; gi < groups.size(); gi++) ctas.push_back(KgKinCta{gi, 0, 1, 0});
	} else {
		uint32_t total_tiles = 0;
		for (const KgKinGroup &g : groups) total_tiles += g.j2[1] >= 0 ? 2 : 1;
		uint32_t left = (uint32_t)c->sm_count, tiles_left = total_tiles;
		for (uint32_t gi = 0; gi < groups.size(); gi++) {
			const uint32_t t = groups[gi].j2[1] >= 0 ? 2 : 1;
			const uint32_t groups_after = (uint32_t)groups.size() - gi - 1;
			uint32_t n = std::max(1u, (left * t + tiles_left / 2) / tiles_left);
			n = std::min(n, left - groups_after);   // every later group still gets a CTA
			for (uint32_t s_ = 0; s_ < n; s_++) ctas.push_back(KgKinCta{gi, s_, n, 0});
			left -= n;
			tiles_left -= t;
		}
	}
	cudaFree(tc.d_kin_groups); cudaFree(tc.d_kin_delta); cudaFree(tc.d_kin_ctas);
	tc.d_kin_groups = nullptr; tc.d_kin_delta = nullptr; tc.d_kin_ctas = nullptr;
	cudaError_t e0 = cudaMalloc((void **)&tc.d_kin_ctas, ctas.size() * sizeof(KgKinCta));
	if (e0 != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc kinship CTA table: %s", cudaGetErrorString(e0));
	KG_CUDA(c, cudaMemcpy(tc.d_kin_ctas, ctas.data(), ctas.size() * sizeof(KgKinCta), cudaMemcpyHostToDevice));
	tc.kin_ctas = (uint32_t)ctas.size();
	cudaError_t e = cudaMalloc((void **)&tc.d_kin_groups, groups.size() * sizeof(KgKinGroup));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc kinship groups: %s", cudaGetErrorString(e));
	KG_CUDA(c, cudaMemcpy(tc.d_kin_groups, groups.data(), groups.size() * sizeof(KgKinGroup), cudaMemcpyHostToDevice));
	e = cudaMalloc((void **)&tc.d_kin_delta, ((size_t)ld * ld + 1) * sizeof(unsigned long long));
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc kinship delta: %s", cudaGetErrorString(e));
	KG_CUDA(c, cudaMemsetAsync(tc.d_kin_delta, 0, ((size_t)ld * ld + 1) * sizeof(unsigned long long), c->stream));
	KG_CUDA(c, cudaFuncSetAttribute(kg_kinship_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	tc.kin_groups = (uint32_t)groups.size();
	tc.kin_smem = smem;
	tc.kin_ready = true;
	tc.kin_tables_w_file = c->w_file;
	return KG_OK;
}

// Fold the tensor engine's co-presence delta (file column order, accumulated over every tile since the last flush) into
// the caller-visible accumulator (memory order) and restart the delta.  Called before the accumulator is read or
// all-reduced; the int32 TMEM sums are widened per launch, so the u64 delta itself never overflows.
static kg_status kg_tc_kinship_flush(kg_ctx *c) {
	KgTcState &tc = c->tc;
	if (!tc.kin_dirty) return KG_OK;
	const uint32_t ld = 64 * c->w_file;
	const uint64_t n2 = (uint64_t)c->n_used * c->n_used;
	const unsigned fg = (unsigned)std::min<uint64_t>((n2 + 255) / 256, (uint64_t)c->sm_count * 8);
	timing_begin(c, KG_KERNEL_AUX, 0);
	kg_kinship_fold_kernel<<<std::max(fg, 1u), 256, 0, c->stream>>>(tc.d_kin_delta, ld, c->d_map_mem, (uint32_t)c->n_used,
	                                                                 c->d_accum, tc.d_kin_delta + (size_t)ld * ld);
	timing_end(c);
	KG_LAUNCH_CHECK(c);
	KG_CUDA(c, cudaMemsetAsync(tc.d_kin_delta, 0, ((size_t)ld * ld + 1) * sizeof(unsigned long long), c->stream));
	tc.kin_dirty = false;
	return KG_OK;
}

// G (file order) of one raw tile on the tensor cores, folded into the caller-visible accumulator (memory order)
static kg_status kg_tc_kinship_tile(kg_ctx *c, const uint64_t *dev_in, uint64_t n_rows) {
	KgTcState &tc = c->tc;
	const uint64_t *dev = nullptr;
	kg_status st = kg_tc_aligned_tile(c, dev_in, n_rows, &dev);
	if (st != KG_OK) return st;
	const uint32_t ld = 64 * c->w_file;
	KgKinTcParams k;
	memset(&k, 0, sizeof k);
	k.rows = dev;
	k.n_rows = n_rows;
	k.w_file = c->w_file;
	// MAC filter bits + kept-row count, once per tile (every tile group of the Gram kernel reads the same rows)
	const size_t kb_need = (size_t)((n_rows + 31) / 32);
	if (c->keep_bits_cap < kb_need) {
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(c->d_keep_bits);
		c->d_keep_bits = nullptr;
		c->keep_bits_cap = 0;
		cudaError_t me = cudaMalloc((void **)&c->d_keep_bits, kb_need * 4);
		if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc keep bits: %s", cudaGetErrorString(me));
		c->keep_bits_cap = kb_need;
	}
	{
		const KgRowView raw{dev, n_rows, c->w_file + 1, c->w_file};
		const unsigned grid = (unsigned)std::min<uint64_t>((n_rows + 255) / 256, (uint64_t)c->sm_count * 8);
		timing_begin(c, KG_KERNEL_AUX, n_rows);
		kg_prefilter_kernel<<<std::max(grid, 1u), 256, 0, c->stream>>>(raw, c->d_file_mask, (uint32_t)c->n_used,
		                                                               (uint32_t)std::min<uint64_t>(c->kin_min_count, 0xFFFFFFFFull),
		                                                               c->d_keep_bits, tc.d_kin_delta + (size_t)ld * ld);
		timing_end(c);
		KG_LAUNCH_CHECK(c);
	}
	k.keep_bits = c->d_keep_bits;
	k.groups = tc.d_kin_groups;
	k.ctas = tc.d_kin_ctas;
	k.delta = tc.d_kin_delta;
	k.ld = ld;
	timing_begin(c, KG_KERNEL_KINSHIP, n_rows);
	kg_kinship_tc_kernel<<<tc.kin_ctas, KG_K_THREADS, tc.kin_smem, c->stream>>>(k);
	timing_end(c);
	KG_LAUNCH_CHECK(c);
	tc.kin_dirty = true;
	return KG_OK;
}
