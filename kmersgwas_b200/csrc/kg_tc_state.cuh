// kg_tc_state.cuh -- per-context state of the tcgen05 (int8 tensor-core) engines.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

struct KgTcState {
	bool scan_ready = false;
	bool kin_ready = false;
	bool use_filter = false;               // auto engine: candidate density is low enough for the filter
	std::string why_unavailable = "tensor-core engine not initialised";
	uint32_t *d_row_list = nullptr;        // rows of the current tile the filter could not rule out
	uint64_t row_list_cap = 0;
	uint32_t row_list_groups = 0;          // 16-column groups the list buffers below were sized for
	uint32_t *d_group_list = nullptr;      // [p_pad / 16][row_list_cap] positions in d_row_list, per 16-column group
	unsigned long long *d_group_count = nullptr;   // [16]
	int32_t *d_tile_pheno = nullptr;       // [p_pad] phenotype of every filter column (-1: none)
	// per-column re-test of the short group lists (kg_pair_select_kernel) and pair-mode re-scoring
	int32_t *d_ent_q = nullptr;            // [p_pad / 16][qcap][16]
	uint32_t *d_ent_n1 = nullptr;          // [p_pad / 16][qcap]
	uint2 *d_pairs = nullptr;              // [16 (p_pad / 16) qcap]
	uint64_t qcap = 0;
	float *d_slack = nullptr;              // [P][n_used / 2 + 1] = slack_table
	bool use_pairs = true;
	int64_t pair_limit = -1;               // KG_OPT_FILTER_PAIR_LIMIT
	// perf-experiment switches, read from the environment once per kg_scan_set_phenotypes (kg_tc_prepare_scan)
	uint32_t dbg_flags = 0;                // KG_FILTER_DEBUG (results are wrong when set)
	uint32_t n_issuers = 0;                // KG_FILTER_ISSUERS (0 = default)
	bool print_stats = false;              // KG_FILTER_STATS: per-tile list sizes on stderr (synchronises the stream)
	// scan filter: quantised centred phenotypes in UMMA (K-major core matrix) layout + per-phenotype constants
	int8_t *d_yq = nullptr;
	struct KgFilterGroupConst *d_gconst = nullptr;   // [p_pad / 16] per column group, see kg_scan_filter.cuh
	std::vector<float> slack_table;        // [P][N/2 + 1] F_p(m): largest |sum of rounding errors / s| over m samples
	std::vector<int8_t> h_q;               // [P][n_used] quantised phenotypes, memory (phenotype-file) order
	std::vector<double> scale;             // [P] quantisation step s_p
	std::vector<float> kappa;              // [P]
	std::vector<uint8_t> degenerate;       // [P] phenotype column the bound cannot handle: always a candidate
	uint32_t p_pad = 0, nc = 0, sbo_b = 0, b_bytes = 0, tcols = 0, a_words = 0, a_stages = 0;
	int split = 0;                         // role split of the filter kernel (kg_filter_split_*): 0 wide, 1 / 2 narrow tables
	uint32_t n_pass = 1, cols_per_pass = 0, raw_stages = 4;   // phenotype columns are scanned in passes of <= 127 (P_pad <= 128)
	size_t smem_bytes = 0;
	uint64_t *d_aligned = nullptr;         // realigned copy of a tile whose device pointer is not 16-byte aligned
	size_t aligned_cap = 0;
	// kinship tensor engine: tile groups, per-tile co-presence delta in file column order (+ kept-row counter)
	struct KgKinGroup *d_kin_groups = nullptr;
	unsigned long long *d_kin_delta = nullptr;
	struct KgKinCta *d_kin_ctas = nullptr;
	uint32_t kin_groups = 0, kin_ctas = 0;
	bool kin_dirty = false;                // the delta holds counts that are not folded into the caller's accumulator yet
	uint32_t kin_tables_w_file = 0;        // shape the kinship tables were built for (0 = none)
	size_t kin_smem = 0;
	void *d_scratch = nullptr;
	// device twins of the host-side filter tables, for kg_filter_retune_kernel (device-selection mode)
	double *d_scale = nullptr;             // [P]
	float *d_kappa0 = nullptr;             // [P]
	uint8_t *d_degenerate = nullptr;       // [P]
	int8_t *d_q = nullptr;                 // [P][n_used]
	uint32_t *d_kidx = nullptr;            // [n_used] operand K index of memory column i
	uint32_t *d_col_of = nullptr;          // [P] column assignment the device image was built with
	float *d_group_lines = nullptr;        // [16][8]
	int32_t *d_thr_tab = nullptr;            // [n_pass][p_pad / 16][n_used + 1] group bounds per row popcount
};

// Device-resident BestAssociationsHeap set (kg_select.cuh); owned by kg_ctx.
struct KgSelState {
	bool active = false;
	bool log_enabled = false;
	uint32_t n_pheno = 0, kmax = 0, cand_cap = 0, sort_stride = 0, sort_smem = 0, log_cap = 0;
	size_t smem = 0;
	std::vector<uint32_t> kbest;
	uint32_t *d_kbest = nullptr, *d_hsize = nullptr, *d_hslot = nullptr, *d_cand_count = nullptr, *d_order = nullptr, *d_log_count = nullptr;
	double *d_hscore = nullptr, *d_floor = nullptr;
	uint64_t *d_pay_kmer = nullptr, *d_pay_row = nullptr;
	unsigned long long *d_hstat = nullptr, *d_sort_buf = nullptr, *d_status = nullptr, *d_digest = nullptr;
	struct KgCand *d_cand = nullptr, *d_log = nullptr;
	unsigned long long *h_status = nullptr;   // pinned copy of d_status
	bool floor_set = false;
	// round schedule (host side, deterministic: the host never sees the thresholds)
	uint64_t rows_submitted = 0;           // rows handed to the device since kg_select_begin / the last overflow recovery
	uint64_t fill_rows = 0;                // rows scanned by the dense exact kernel before the filter takes over
	double growth = 0.5;                   // round length = growth x rows submitted so far (candidates per phenotype ~ growth x K)
	uint64_t max_round = 1ull << 24;   // measured on the config-2 job: 2^23 0.368 s, 2^24 0.358 s (fewer rounds: less per-round overhead, slightly staler thresholds)
	uint64_t min_round = 4096;
	uint64_t cand_cap_opt = 0, log_cap_opt = 0;   // KG_OPT_SELECT_CAND_CAP / KG_OPT_SELECT_LOG_CAP (0 = default)
};
