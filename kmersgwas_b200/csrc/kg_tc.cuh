// kg_tc.cuh -- tcgen05 / TMEM int8 engines (scan filter and kinship Gram).  Included by kg_abi.cu
// after kg_ctx is defined.
#pragma once

static void kg_tc_free(KgTcState *tc) {
	cudaFree(tc->d_pairs); cudaFree(tc->d_yq); cudaFree(tc->d_pconst); cudaFree(tc->d_scratch);
	tc->d_pairs = nullptr; tc->d_yq = nullptr; tc->d_pconst = nullptr; tc->d_scratch = nullptr;
}
static bool kg_tc_scan_available(const kg_ctx *c) { return c->tc.scan_ready; }
static bool kg_tc_scan_profitable(const kg_ctx *c) { return false; }
static bool kg_tc_kinship_available(const kg_ctx *c) { return c->tc.kin_ready; }
static kg_status kg_tc_prepare_scan(kg_ctx *c) { (void)c; return KG_OK; }
static kg_status kg_tc_update_thresholds(kg_ctx *c) { (void)c; return KG_OK; }
static kg_status kg_tc_prepare_kinship(kg_ctx *c) { (void)c; return KG_OK; }
static kg_status kg_tc_scan_tile(kg_ctx *c, const uint64_t *dev, uint64_t n_rows, uint64_t first_row_id) {
	(void)dev; (void)n_rows; (void)first_row_id;
	KG_FAIL(c, KG_ERR_INVALID, "tensor filter engine not built");
}
static kg_status kg_tc_kinship_tile(kg_ctx *c, const KgRowView &view) {
	(void)view;
	KG_FAIL(c, KG_ERR_INVALID, "tensor kinship engine not built");
}
