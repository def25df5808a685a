// kg_tc_state.cuh -- per-context state of the tcgen05 (int8 tensor-core) engines.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

struct KgTcState {
	bool scan_ready = false;
	bool kin_ready = false;
	bool use_filter = false;               // auto engine: candidate density is low enough for the filter
	std::string why_unavailable = "tensor-core engine not initialised";
	uint64_t pair_capacity = 1ull << 22;   // candidate (row, phenotype) pairs per fetch interval
	uint2 *d_pairs = nullptr;
	// scan filter: quantised centred phenotypes in UMMA (K-major core matrix) layout + per-phenotype constants
	int8_t *d_yq = nullptr;
	float2 *d_pconst = nullptr;            // [p_pad] (alpha, kappa)
	std::vector<int8_t> h_yq_image;
	std::vector<double> scale;             // [P] quantisation step s_p
	std::vector<float> kappa;              // [P]
	std::vector<uint8_t> degenerate;       // [P] phenotype column the bound cannot handle: always a candidate
	uint32_t p_pad = 0, nc = 0, sbo_b = 0, b_bytes = 0, tcols = 0;
	size_t smem_bytes = 0;
	uint64_t *d_aligned = nullptr;         // realigned copy of a tile whose device pointer is not 16-byte aligned
	size_t aligned_cap = 0;
	// kinship: int32 partial Gram flush scratch
	void *d_scratch = nullptr;
};
