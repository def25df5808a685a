// kg_tc_state.cuh -- per-context state of the tcgen05 (int8 tensor-core) engines.
#pragma once
#include <stdint.h>
#include <string>

struct KgTcState {
	bool scan_ready = false;
	bool kin_ready = false;
	std::string why_unavailable = "tensor-core engine not initialised";
	uint64_t pair_capacity = 1ull << 22;   // candidate (row, phenotype) pairs per fetch interval
	uint2 *d_pairs = nullptr;
	// scan filter: quantised centred phenotypes in UMMA layout + per-phenotype constants
	int8_t *d_yq = nullptr;
	float *d_pconst = nullptr;
	uint32_t n_pad_p = 0;
	// kinship: int32 partial Gram flush scratch
	void *d_scratch = nullptr;
};
