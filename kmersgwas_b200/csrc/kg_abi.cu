// kg_abi.cu -- implementation of include/kmersgwas_b200.h (the C ABI) on top of the sm_100a kernels.
// No CPU fallback: every entry point needs a CUDA device.
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <cctype>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "kg_common.cuh"
#include "kg_kinship_popc.cuh"
#include "kg_patterns.cuh"
#include "kg_probe.cuh"
#include "kg_scan_exact.cuh"
#include "kg_snps.cuh"
#include "kg_synth.cuh"
#include "kg_table_build.cuh"
#include "kg_tc_state.cuh"

// ------------------------------------------------------------------------------------- context
struct kg_ctx {
	int device = 0;
	int sm_count = KG_SM_COUNT_DEFAULT;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	cudaStream_t copy_stream = nullptr;
	std::string err;
	uint64_t launches = 0;

	// geometry
	uint64_t n_file = 0, n_used = 0;
	uint32_t w_file = 0, w_mem = 0, nb = 0;
	bool identity = false;
	std::vector<uint32_t> map_word, map_bit;
	uint32_t *d_map_mem = nullptr;   // [n_used] (word<<6)|bit of memory column i
	uint32_t *d_map_lane = nullptr;  // [nb*128] same, in lane order (4b+L)*32+t ; 0xFFFFFFFF = pad
	uint64_t *d_file_mask = nullptr; // [w_file]
	uint64_t *d_mem_mask = nullptr;  // [w_mem]
	uint32_t *d_mask32 = nullptr;    // [nb*4]

	// phenotypes
	uint32_t n_pheno = 0, p_alloc = 0, pt = 8;
	uint64_t min_count = 0;
	float *d_y_lane = nullptr;
	float *d_y_pair = nullptr;   // the same values in the order kg_scan_pair_kernel reads them (see kg_scan_exact.cuh)
	float *d_sums = nullptr;
	double *d_thr = nullptr;
	std::vector<float> h_y;      // [P][n_used] as given (tensor filter quantises from it)
	std::vector<float> h_sums;
	std::vector<double> h_thr;

	// hits / counters.  Hits are produced in intervals (kg_scan_mark): two sets of device buffers so that the
	// device fills interval i+1 while the host fetches interval i.  d_hits / d_counters point at the OPEN interval.
	struct ScanInterval {
		kg_hit *d_hits = nullptr;
		unsigned long long *d_cnt = nullptr;   // [0] hits [1] kept rows [2] rows listed by the filter (current tile)
		                                       // [4] debug scratch [5] rows listed by the filter (whole interval)
		                                       // [6] (row, column group) pairs listed (whole interval)
		unsigned long long *h_cnt = nullptr;   // pinned copy of d_cnt, valid once `done` has completed
		cudaEvent_t done = nullptr;
		cudaEvent_t closed_ev = nullptr;       // recorded on the compute stream when the interval is marked
		uint64_t rows = 0;                     // rows submitted into this interval
		bool closed = false;                   // marked, waiting for kg_scan_fetch
		bool used_filter = false;
	} iv[2];
	int cur = 0;
	kg_hit *d_hits = nullptr;
	uint64_t hit_capacity = 1ull << 22;
	unsigned long long *d_counters = nullptr;
	uint64_t rows_seen_total = 0, kept_total = 0;   // over consumed intervals
	cudaStream_t d2h_stream = nullptr;
	// pinned staging ring for stream-ordered threshold updates (no host sync)
	struct ThrStage { double *h_thr = nullptr; cudaEvent_t ev = nullptr; } thr_stage[4];
	int thr_next = 0;
	size_t thr_stage_p = 0, thr_stage_g = 0;

	// tile upload
	uint64_t *d_tile[2] = {nullptr, nullptr};
	size_t tile_cap[2] = {0, 0};
	cudaEvent_t slot_free[2] = {nullptr, nullptr}, slot_ready[2] = {nullptr, nullptr};
	int next_slot = 0, cur_slot = -1;
	void *h_stage[2] = {nullptr, nullptr};
	cudaEvent_t stage_free[2] = {nullptr, nullptr};
	static constexpr size_t kStageBytes = 16u << 20;

	// squeeze scratch
	uint64_t *d_squeezed = nullptr;
	size_t squeezed_cap = 0;
	// kg_scan_scores_dense output scratch
	void *d_dense = nullptr;
	size_t dense_cap = 0;

	// kinship
	unsigned long long *d_accum = nullptr;
	bool own_accum = false, kin_active = false;
	uint64_t kin_min_count = 0;
	uint32_t *d_keep_bits = nullptr;
	size_t keep_bits_cap = 0;
	uint8_t *d_keep_bytes = nullptr;   // kg_mac_filter scratch: [cap] keep flags, then one u64 counter (8-byte aligned)
	size_t keep_bytes_cap = 0;
	unsigned long long *d_ibs = nullptr;

	// tensor-core engine state (kg_tc.cuh)
	KgTcState tc;
	// device-resident heaps (kg_select_host.cuh)
	KgSelState sel;

	int scan_engine = 0, kin_engine = 0;

	// distinct presence/absence patterns (kg_patterns_*)
	unsigned long long *d_pat_table = nullptr, *d_pat_counters = nullptr;
	uint64_t pat_slots = 0, pat_count = 0;
	bool pat_attached = false;             // kg_patterns_attach: every tile of kg_scan_submit is counted as well
	uint64_t pat_min_count = 0, pat_budget = 0;   // rows the table still has guaranteed room for

	// stream tickets (kg_stream_mark / kg_stream_wait): events on the copy and compute streams
	struct Mark { cudaEvent_t copy_ev = nullptr, compute_ev = nullptr; };
	std::vector<Mark> marks;
	// NCCL communicator of this context (kg_comm_init_rank / kg_comm_init_all), opaque here
	void *nccl_comm = nullptr;

	// per-launch device timing (KG_OPT_KERNEL_TIMING)
	bool timing = false;
	struct TimedLaunch { cudaEvent_t beg, end; int cls; uint64_t rows; };
	std::vector<TimedLaunch> timed_pending;
	std::vector<cudaEvent_t> event_pool;
	double timed_ms[KG_KERNEL_CLASSES] = {0};
	uint64_t timed_launches[KG_KERNEL_CLASSES] = {0}, timed_rows[KG_KERNEL_CLASSES] = {0};
};

static std::string g_create_error;

#define KG_FAIL(ctx, code, ...)                                   \
	do {                                                          \
		char buf_[512];                                           \
		snprintf(buf_, sizeof buf_, __VA_ARGS__);                 \
		(ctx)->err = buf_;                                        \
		return (code);                                            \
	} while (0)

#define KG_CUDA(ctx, expr)                                                                   \
	do {                                                                                     \
		cudaError_t e_ = (expr);                                                             \
		if (e_ != cudaSuccess) {                                                             \
			char buf_[512];                                                                  \
			snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
			         __FILE__, __LINE__);                                                    \
			(ctx)->err = buf_;                                                               \
			return KG_ERR_CUDA;                                                              \
		}                                                                                    \
	} while (0)

#define KG_LAUNCH_CHECK(ctx)            \
	do {                                \
		(ctx)->launches++;              \
		KG_CUDA(ctx, cudaGetLastError()); \
	} while (0)

// ---- per-launch timing: events on the launching stream, resolved after the stream is synchronised
static cudaEvent_t timing_event(kg_ctx *c) {
	if (!c->event_pool.empty()) {
		cudaEvent_t e = c->event_pool.back();
		c->event_pool.pop_back();
		return e;
	}
	cudaEvent_t e = nullptr;
	cudaEventCreate(&e);
	return e;
}
static void timing_begin(kg_ctx *c, int cls, uint64_t rows) {
	if (!c->timing) return;
	kg_ctx::TimedLaunch t{timing_event(c), timing_event(c), cls, rows};
	cudaEventRecord(t.beg, c->stream);
	c->timed_pending.push_back(t);
}
static void timing_end(kg_ctx *c) {
	if (!c->timing || c->timed_pending.empty()) return;
	cudaEventRecord(c->timed_pending.back().end, c->stream);
}
// call only after cudaStreamSynchronize(c->stream)
static void timing_resolve(kg_ctx *c) {
	for (const kg_ctx::TimedLaunch &t : c->timed_pending) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, t.beg, t.end) == cudaSuccess) {
			c->timed_ms[t.cls] += ms;
			c->timed_launches[t.cls]++;
			c->timed_rows[t.cls] += t.rows;
		} else {
			cudaGetLastError();
		}
		c->event_pool.push_back(t.beg);
		c->event_pool.push_back(t.end);
	}
	c->timed_pending.clear();
}

struct KgScanParams;
static KgScanParams scan_params(kg_ctx *c, const KgRowView &view, uint64_t first_row_id);
template <int MODE> static kg_status launch_exact_pt(kg_ctx *c, const KgScanParams &prm);
static kg_status launch_exact_list(kg_ctx *c, const KgScanParams &prm, uint32_t n_tiles);
static kg_status ensure_squeeze_scratch(kg_ctx *c, uint64_t n_rows);

template <typename T>
static cudaError_t dev_alloc_copy(T **dst, const std::vector<T> &src) {
	cudaError_t e = cudaMalloc((void **)dst, std::max<size_t>(src.size(), 1) * sizeof(T));
	if (e != cudaSuccess) return e;
	if (!src.empty()) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
	return e;
}

static bool is_device_pointer(const void *p);
static kg_status acquire_tile(kg_ctx *c, const uint64_t *rows, uint64_t n_rows, const uint64_t **dev);
static kg_status release_tile(kg_ctx *c);
static kg_status memory_view(kg_ctx *c, const uint64_t *dev, uint64_t n_rows, KgRowView *view);
static kg_status kg_sel_finish_round(kg_ctx *c, uint64_t n_rows, uint64_t first_row_id, bool filter_counters);
static kg_status patterns_attached_tile(kg_ctx *c, const uint64_t *dev, uint64_t n_rows);
static const uint64_t kHostSubTileRows = 1ull << 20;

static void kg_comm_destroy(kg_ctx *c);

// tensor-core engine (needs kg_ctx and the macros above)
#include "kg_tc.cuh"
// device-resident heaps
#include "kg_select_host.cuh"

extern "C" int kg_abi_version(void) { return KG_ABI_VERSION; }

extern "C" const char *kg_last_error(const kg_ctx *ctx) {
	return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" uint64_t kg_launch_count(const kg_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------ host placement
static bool kg_read_small_file(const char *path, char *buf, size_t cap) {
	FILE *f = fopen(path, "r");
	if (!f) return false;
	const size_t n = fread(buf, 1, cap - 1, f);
	fclose(f);
	buf[n] = 0;
	return n > 0;
}

extern "C" int kg_bind_host_to_device(int device, int *n_cpus) {
	if (n_cpus) *n_cpus = 0;
	if (const char *e = getenv("KMERSGWAS_NUMA_BIND"))
		if (atoi(e) == 0) return -1;
	char bus[32] = {0}, path[128], text[4096];
	if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
	for (char *q = bus; *q; q++) *q = (char)tolower((unsigned char)*q);
	snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
	if (!kg_read_small_file(path, text, sizeof text)) return -1;
	const int node = atoi(text);
	if (node < 0 || node >= 1024) return -1;
	snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
	if (!kg_read_small_file(path, text, sizeof text)) return -1;
	// "0-31,64-95" -> the part of it inside the thread's current affinity mask (a container's cpuset)
	cpu_set_t now, want;
	CPU_ZERO(&want);
	if (sched_getaffinity(0, sizeof now, &now) != 0) return -1;
	for (const char *q = text; *q && *q != '\n';) {
		char *end;
		long a = strtol(q, &end, 10), b = a;
		if (end == q) break;
		if (*end == '-') { q = end + 1; b = strtol(q, &end, 10); }
		for (long cpu = a; cpu <= b && cpu < CPU_SETSIZE; cpu++)
			if (CPU_ISSET(cpu, &now)) CPU_SET(cpu, &want);
		q = (*end == ',') ? end + 1 : end;
	}
	const int cpus = CPU_COUNT(&want);
	if (cpus == 0 || sched_setaffinity(0, sizeof want, &want) != 0) return -1;
	if (n_cpus) *n_cpus = cpus;
	// set_mempolicy(MPOL_PREFERRED, {node}): memory the thread faults in (and the pages the driver pins for it) comes
	// from the device's node while it has room; no libnuma in the image, so the raw system call
	unsigned long mask[1024 / (8 * sizeof(unsigned long))] = {0};
	mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
	(void)syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, (unsigned long)(sizeof mask * 8));
	return node;
}

static kg_status ctx_init(kg_ctx *c, int device, const kg_shape *shape, void *stream) {
	if (!shape || !shape->map_word || !shape->map_bit || shape->n_used == 0 || shape->n_file == 0)
		KG_FAIL(c, KG_ERR_INVALID, "kg_ctx_create: empty shape");
	if (shape->n_used > 65000) KG_FAIL(c, KG_ERR_INVALID, "kg_ctx_create: n_used > 65000 not supported");
	int n_dev = 0;
	cudaError_t e = cudaGetDeviceCount(&n_dev);
	if (e != cudaSuccess || n_dev == 0)
		KG_FAIL(c, KG_ERR_CUDA, "kg_ctx_create: no CUDA device (%s); this library has no CPU fallback",
		        cudaGetErrorString(e));
	if (device < 0 || device >= n_dev) KG_FAIL(c, KG_ERR_INVALID, "kg_ctx_create: bad device %d", device);
	KG_CUDA(c, cudaSetDevice(device));
	c->device = device;
	cudaDeviceProp prop;
	KG_CUDA(c, cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10)
		KG_FAIL(c, KG_ERR_CUDA, "kg_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
		        device, prop.major, prop.minor);
	c->sm_count = prop.multiProcessorCount;
	if (stream) {
		c->stream = (cudaStream_t)stream;
	} else {
		KG_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
		c->own_stream = true;
	}
	KG_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	for (int i = 0; i < 2; i++) {
		KG_CUDA(c, cudaEventCreateWithFlags(&c->slot_free[i], cudaEventDisableTiming));
		KG_CUDA(c, cudaEventCreateWithFlags(&c->slot_ready[i], cudaEventDisableTiming));
		KG_CUDA(c, cudaEventCreateWithFlags(&c->stage_free[i], cudaEventDisableTiming));
	}

	c->n_file = shape->n_file;
	c->n_used = shape->n_used;
	c->w_file = (uint32_t)((c->n_file + 63) / 64);
	c->w_mem = (uint32_t)(2 * ((c->n_used + 127) / 128));  // kmers_multiple_databases.cpp:51
	c->nb = c->w_mem / 2;
	c->map_word.assign(shape->map_word, shape->map_word + c->n_used);
	c->map_bit.assign(shape->map_bit, shape->map_bit + c->n_used);
	c->identity = (c->n_used == c->n_file);
	std::vector<uint64_t> file_mask(c->w_file, 0), mem_mask(c->w_mem, 0);
	std::vector<uint32_t> map_mem(c->n_used), map_lane((size_t)c->nb * 128, 0xFFFFFFFFu), mask32((size_t)c->nb * 4, 0);
	for (uint64_t i = 0; i < c->n_used; i++) {
		const uint32_t w = c->map_word[i], b = c->map_bit[i];
		if (w >= c->w_file || b >= 64 || (uint64_t)w * 64 + b >= c->n_file)
			KG_FAIL(c, KG_ERR_INVALID, "kg_ctx_create: column map entry %llu out of range", (unsigned long long)i);
		if (file_mask[w] & (1ull << b))
			KG_FAIL(c, KG_ERR_INVALID, "kg_ctx_create: file column used twice (entry %llu)", (unsigned long long)i);
		file_mask[w] |= 1ull << b;  // :308 m_map_mask
		mem_mask[i >> 6] |= 1ull << (i & 63);
		map_mem[i] = (w << 6) | b;
		if ((uint64_t)w * 64 + b != i) c->identity = false;
		// lane order: sample i = 128 blk + 32 L + (31 - t)
		const uint32_t blk = (uint32_t)(i / 128), L = (uint32_t)((i % 128) / 32), t = 31 - (uint32_t)(i % 32);
		map_lane[(size_t)(blk * 4 + L) * 32 + t] = map_mem[i];
		mask32[blk * 4 + L] |= 1u << (i % 32);
	}
	KG_CUDA(c, dev_alloc_copy(&c->d_map_mem, map_mem));
	KG_CUDA(c, dev_alloc_copy(&c->d_map_lane, map_lane));
	KG_CUDA(c, dev_alloc_copy(&c->d_file_mask, file_mask));
	KG_CUDA(c, dev_alloc_copy(&c->d_mem_mask, mem_mask));
	KG_CUDA(c, dev_alloc_copy(&c->d_mask32, mask32));
	KG_CUDA(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
	for (int i = 0; i < 2; i++) {
		KG_CUDA(c, cudaMalloc((void **)&c->iv[i].d_cnt, 8 * sizeof(unsigned long long)));
		KG_CUDA(c, cudaMemset(c->iv[i].d_cnt, 0, 8 * sizeof(unsigned long long)));
		KG_CUDA(c, cudaMallocHost((void **)&c->iv[i].h_cnt, 8 * sizeof(unsigned long long)));
		KG_CUDA(c, cudaEventCreateWithFlags(&c->iv[i].done, cudaEventDisableTiming));
		KG_CUDA(c, cudaEventCreateWithFlags(&c->iv[i].closed_ev, cudaEventDisableTiming));
	}
	for (int i = 0; i < 4; i++) KG_CUDA(c, cudaEventCreateWithFlags(&c->thr_stage[i].ev, cudaEventDisableTiming));
	c->cur = 0;
	c->d_counters = c->iv[0].d_cnt;
	return KG_OK;
}

extern "C" kg_status kg_ctx_create(int device, const kg_shape *shape, void *stream, kg_ctx **out) {
	if (!out) return KG_ERR_INVALID;
	*out = nullptr;
	kg_ctx *c = new (std::nothrow) kg_ctx();
	if (!c) { g_create_error = "out of host memory"; return KG_ERR_NOMEM; }
	kg_status st = ctx_init(c, device, shape, stream);
	if (st != KG_OK) {
		g_create_error = c->err;
		kg_ctx_destroy(c);
		return st;
	}
	*out = c;
	return KG_OK;
}

extern "C" void kg_ctx_destroy(kg_ctx *c) {
	if (!c) return;
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
	cudaFree(c->d_map_mem); cudaFree(c->d_map_lane); cudaFree(c->d_file_mask); cudaFree(c->d_mem_mask);
	cudaFree(c->d_keep_bytes);
	cudaFree(c->d_mask32); cudaFree(c->d_y_lane); cudaFree(c->d_y_pair); cudaFree(c->d_sums); cudaFree(c->d_thr);
	for (int i = 0; i < 2; i++) {
		cudaFree(c->iv[i].d_hits); cudaFree(c->iv[i].d_cnt);
		if (c->iv[i].h_cnt) cudaFreeHost(c->iv[i].h_cnt);
		if (c->iv[i].done) cudaEventDestroy(c->iv[i].done);
		if (c->iv[i].closed_ev) cudaEventDestroy(c->iv[i].closed_ev);
	}
	for (int i = 0; i < 4; i++) {
		if (c->thr_stage[i].h_thr) cudaFreeHost(c->thr_stage[i].h_thr);
		if (c->thr_stage[i].ev) cudaEventDestroy(c->thr_stage[i].ev);
	}
	if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
	cudaFree(c->d_squeezed); cudaFree(c->d_keep_bits); cudaFree(c->d_dense);
	cudaFree(c->d_ibs);
	if (c->own_accum) cudaFree(c->d_accum);
	kg_tc_free(&c->tc);
	kg_sel_free(&c->sel);
	cudaFree(c->d_pat_table); cudaFree(c->d_pat_counters);
	for (int i = 0; i < 2; i++) {
		cudaFree(c->d_tile[i]);
		if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
		if (c->slot_free[i]) cudaEventDestroy(c->slot_free[i]);
		if (c->slot_ready[i]) cudaEventDestroy(c->slot_ready[i]);
		if (c->stage_free[i]) cudaEventDestroy(c->stage_free[i]);
	}
	for (kg_ctx::Mark &m : c->marks) { if (m.copy_ev) cudaEventDestroy(m.copy_ev); if (m.compute_ev) cudaEventDestroy(m.compute_ev); }
	kg_comm_destroy(c);
	for (const kg_ctx::TimedLaunch &t : c->timed_pending) { cudaEventDestroy(t.beg); cudaEventDestroy(t.end); }
	for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

extern "C" kg_status kg_set_option(kg_ctx *c, int option, int64_t value) {
	if (!c) return KG_ERR_INVALID;
	switch (option) {
	case KG_OPT_SCAN_ENGINE:
		if (value < 0 || value > 2) KG_FAIL(c, KG_ERR_INVALID, "scan engine must be 0, 1 or 2");
		c->scan_engine = (int)value;
		return KG_OK;
	case KG_OPT_HIT_CAPACITY:
		if (value < 1) KG_FAIL(c, KG_ERR_INVALID, "hit capacity must be >= 1");
		if (c->iv[0].d_hits) {
			// already allocated (kg_scan_set_phenotypes): only between intervals, after the device has drained
			if (c->iv[0].closed || c->iv[1].closed || c->iv[c->cur].rows)
				KG_FAIL(c, KG_ERR_STATE, "hit capacity can only change while no hit interval is open or waiting (fetch first)");
			KG_CUDA(c, cudaSetDevice(c->device));
			KG_CUDA(c, cudaStreamSynchronize(c->stream));
			for (int i = 0; i < 2; i++) {
				cudaFree(c->iv[i].d_hits);
				c->iv[i].d_hits = nullptr;
				cudaError_t me = cudaMalloc((void **)&c->iv[i].d_hits, (uint64_t)value * sizeof(kg_hit));
				if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc hit buffer: %s", cudaGetErrorString(me));
			}
			c->d_hits = c->iv[c->cur].d_hits;
		}
		c->hit_capacity = (uint64_t)value;
		return KG_OK;
	case KG_OPT_KINSHIP_ENGINE:
		if (value < 0 || value > 2) KG_FAIL(c, KG_ERR_INVALID, "kinship engine must be 0, 1 or 2");
		c->kin_engine = (int)value;
		return KG_OK;
	case KG_OPT_KERNEL_TIMING:
		c->timing = value != 0;
		return KG_OK;
	case KG_OPT_FILTER_PAIR_LIMIT:
		if (value < -1) KG_FAIL(c, KG_ERR_INVALID, "filter pair limit must be >= -1");
		c->tc.pair_limit = value;
		return KG_OK;
	case KG_OPT_SELECT_GROWTH_PERMILLE:
		if (value < 1 || value > 100000) KG_FAIL(c, KG_ERR_INVALID, "selection round growth must be in [1, 100000] per mille");
		c->sel.growth = (double)value / 1000.0;
		return KG_OK;
	case KG_OPT_SELECT_MAX_ROUND:
		if (value < 1 || value >= (1ll << 31)) KG_FAIL(c, KG_ERR_INVALID, "selection round length must be in [1, 2^31)");
		c->sel.max_round = (uint64_t)value;
		return KG_OK;
	case KG_OPT_SELECT_CAND_CAP:
		if (value < 1) KG_FAIL(c, KG_ERR_INVALID, "candidate capacity must be >= 1");
		if (c->sel.active) KG_FAIL(c, KG_ERR_STATE, "candidate capacity must be set before kg_select_begin");
		c->sel.cand_cap_opt = (uint64_t)value;
		return KG_OK;
	case KG_OPT_SELECT_LOG_CAP:
		if (value < 1) KG_FAIL(c, KG_ERR_INVALID, "log capacity must be >= 1");
		if (c->sel.active) KG_FAIL(c, KG_ERR_STATE, "log capacity must be set before kg_select_begin");
		c->sel.log_cap_opt = (uint64_t)value;
		return KG_OK;
	default:
		KG_FAIL(c, KG_ERR_INVALID, "unknown option %d", option);
	}
}

extern "C" kg_status kg_sync(kg_ctx *c) {
	if (!c) return KG_ERR_INVALID;
	KG_CUDA(c, cudaSetDevice(c->device));
	if (c->kin_active) {   // a caller-owned kinship accumulator is current after kg_sync
		kg_status st = kg_tc_kinship_flush(c);
		if (st != KG_OK) return st;
	}
	KG_CUDA(c, cudaStreamSynchronize(c->copy_stream));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	timing_resolve(c);
	return KG_OK;
}

// ---- stream tickets: "everything submitted so far" as a waitable point, for callers that recycle input buffers
extern "C" kg_status kg_stream_mark(kg_ctx *c, uint64_t *ticket) {
	if (!c || !ticket) return KG_ERR_INVALID;
	KG_CUDA(c, cudaSetDevice(c->device));
	// reuse a completed slot if there is one
	size_t slot = c->marks.size();
	for (size_t i = 0; i < c->marks.size(); i++)
		if (cudaEventQuery(c->marks[i].copy_ev) == cudaSuccess && cudaEventQuery(c->marks[i].compute_ev) == cudaSuccess) { slot = i; break; }
	cudaGetLastError();
	if (slot == c->marks.size()) {
		kg_ctx::Mark m;
		KG_CUDA(c, cudaEventCreateWithFlags(&m.copy_ev, cudaEventDisableTiming));
		KG_CUDA(c, cudaEventCreateWithFlags(&m.compute_ev, cudaEventDisableTiming));
		c->marks.push_back(m);
	}
	KG_CUDA(c, cudaEventRecord(c->marks[slot].copy_ev, c->copy_stream));
	KG_CUDA(c, cudaEventRecord(c->marks[slot].compute_ev, c->stream));
	*ticket = slot;
	return KG_OK;
}

extern "C" kg_status kg_stream_wait(kg_ctx *c, uint64_t ticket) {
	if (!c) return KG_ERR_INVALID;
	if (ticket >= c->marks.size()) KG_FAIL(c, KG_ERR_INVALID, "kg_stream_wait: unknown ticket");
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaEventSynchronize(c->marks[ticket].copy_ev));
	KG_CUDA(c, cudaEventSynchronize(c->marks[ticket].compute_ev));
	return KG_OK;
}

extern "C" kg_status kg_kernel_time(kg_ctx *c, int cls, double *ms_total, uint64_t *launches, uint64_t *rows) {
	if (!c) return KG_ERR_INVALID;
	if (cls < 0 || cls >= KG_KERNEL_CLASSES) KG_FAIL(c, KG_ERR_INVALID, "kg_kernel_time: bad kernel class %d", cls);
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	timing_resolve(c);
	if (ms_total) *ms_total = c->timed_ms[cls];
	if (launches) *launches = c->timed_launches[cls];
	if (rows) *rows = c->timed_rows[cls];
	return KG_OK;
}

extern "C" kg_status kg_kernel_time_reset(kg_ctx *c) {
	if (!c) return KG_ERR_INVALID;
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	timing_resolve(c);
	for (int i = 0; i < KG_KERNEL_CLASSES; i++) { c->timed_ms[i] = 0; c->timed_launches[i] = 0; c->timed_rows[i] = 0; }
	return KG_OK;
}

extern "C" kg_status kg_host_alloc(kg_ctx *c, size_t bytes, void **out) {
	if (!c || !out) return KG_ERR_INVALID;
	KG_CUDA(c, cudaSetDevice(c->device));
	cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
	return KG_OK;
}

extern "C" void kg_host_free(kg_ctx *c, void *p) {
	(void)c;
	if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------- tile upload
// Returns a device pointer holding the tile.  Device pointers are used in place; host pointers are
// copied on copy_stream into one of two device slots (so the copy of tile i+1 overlaps the kernels
// of tile i) and the compute stream is made to wait for the copy.
static kg_status acquire_tile(kg_ctx *c, const uint64_t *rows, uint64_t n_rows, const uint64_t **dev) {
	c->cur_slot = -1;
	const size_t bytes = (size_t)n_rows * (c->w_file + 1) * 8;
	cudaPointerAttributes attr;
	cudaError_t e = cudaPointerGetAttributes(&attr, rows);
	if (e != cudaSuccess) { cudaGetLastError(); attr.type = cudaMemoryTypeUnregistered; }
	if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) {
		if (attr.type == cudaMemoryTypeDevice && attr.device != c->device)
			KG_FAIL(c, KG_ERR_INVALID, "tile is on device %d, context on %d", attr.device, c->device);
		*dev = rows;
		return KG_OK;
	}
	const int s = c->next_slot;
	c->next_slot ^= 1;
	if (c->tile_cap[s] < bytes) {
		KG_CUDA(c, cudaEventSynchronize(c->slot_free[s]));
		if (c->d_tile[s]) KG_CUDA(c, cudaFree(c->d_tile[s]));
		c->d_tile[s] = nullptr;
		c->tile_cap[s] = 0;
		const size_t cap = std::max<size_t>(bytes, 1u << 20);
		cudaError_t me = cudaMalloc((void **)&c->d_tile[s], cap);
		if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc(%zu) for tile: %s", cap, cudaGetErrorString(me));
		c->tile_cap[s] = cap;
	}
	KG_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->slot_free[s], 0));
	if (attr.type == cudaMemoryTypeHost) {
		KG_CUDA(c, cudaMemcpyAsync(c->d_tile[s], rows, bytes, cudaMemcpyHostToDevice, c->copy_stream));
	} else {
		// pageable memory: bounce through two pinned staging buffers
		const char *src = reinterpret_cast<const char *>(rows);
		size_t off = 0;
		int k = 0;
		while (off < bytes) {
			if (!c->h_stage[k]) {
				cudaError_t he = cudaMallocHost(&c->h_stage[k], kg_ctx::kStageBytes);
				if (he != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMallocHost staging: %s", cudaGetErrorString(he));
			}
			const size_t n = std::min(kg_ctx::kStageBytes, bytes - off);
			KG_CUDA(c, cudaEventSynchronize(c->stage_free[k]));
			memcpy(c->h_stage[k], src + off, n);
			KG_CUDA(c, cudaMemcpyAsync(reinterpret_cast<char *>(c->d_tile[s]) + off, c->h_stage[k], n,
			                           cudaMemcpyHostToDevice, c->copy_stream));
			KG_CUDA(c, cudaEventRecord(c->stage_free[k], c->copy_stream));
			off += n;
			k ^= 1;
		}
	}
	KG_CUDA(c, cudaEventRecord(c->slot_ready[s], c->copy_stream));
	KG_CUDA(c, cudaStreamWaitEvent(c->stream, c->slot_ready[s], 0));
	c->cur_slot = s;
	*dev = c->d_tile[s];
	return KG_OK;
}

static kg_status release_tile(kg_ctx *c) {
	if (c->cur_slot >= 0) KG_CUDA(c, cudaEventRecord(c->slot_free[c->cur_slot], c->stream));
	c->cur_slot = -1;
	return KG_OK;
}

// Memory-order view of a device tile: the raw tile itself when the column map is the identity,
// else a squeezed copy (load_kmers :125-132) in the context's scratch buffer.
static kg_status memory_view(kg_ctx *c, const uint64_t *dev, uint64_t n_rows, KgRowView *view) {
	if (c->identity) {
		*view = KgRowView{dev, n_rows, c->w_file + 1, c->w_file};
		return KG_OK;
	}
	kg_status st0 = ensure_squeeze_scratch(c, n_rows);
	if (st0 != KG_OK) return st0;
	KgRowView raw{dev, n_rows, c->w_file + 1, c->w_file};
	const uint64_t total = n_rows * (uint64_t)(c->w_mem + 1);
	const unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)c->sm_count * 16);
	timing_begin(c, KG_KERNEL_AUX, n_rows);
	kg_squeeze_kernel<<<std::max(grid, 1u), 256, 0, c->stream>>>(raw, c->d_map_mem, (uint32_t)c->n_used, c->w_mem,
	                                                             c->d_squeezed, nullptr, nullptr);
	timing_end(c);
	KG_LAUNCH_CHECK(c);
	*view = KgRowView{c->d_squeezed, n_rows, c->w_mem + 1, c->w_mem};
	return KG_OK;
}

static kg_status ensure_squeeze_scratch(kg_ctx *c, uint64_t n_rows) {
	const size_t need = (size_t)n_rows * (c->w_mem + 1) * 8;
	if (c->squeezed_cap < need) {
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		if (c->d_squeezed) KG_CUDA(c, cudaFree(c->d_squeezed));
		c->d_squeezed = nullptr;
		c->squeezed_cap = 0;
		cudaError_t me = cudaMalloc((void **)&c->d_squeezed, need);
		if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc(%zu) squeeze scratch: %s", need, cudaGetErrorString(me));
		c->squeezed_cap = need;
	}
	return KG_OK;
}

// ------------------------------------------------------------------------------------- scan
extern "C" kg_status kg_scan_set_phenotypes(kg_ctx *c, const float *y, uint32_t n_pheno, uint64_t min_count) {
	if (!c || !y || n_pheno == 0) { if (c) c->err = "kg_scan_set_phenotypes: bad arguments"; return KG_ERR_INVALID; }
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	kg_sel_free(&c->sel);   // heaps belong to a phenotype set
	c->n_pheno = n_pheno;
	c->min_count = min_count;
	c->pt = n_pheno > 4 ? 8 : (n_pheno > 2 ? 4 : n_pheno);
	// shared memory of the y tile must fit: shrink PT for very wide tables
	while (c->pt > 1 && (size_t)c->nb * 4 * (32 * c->pt + 4) * 4 > 200u * 1024) c->pt /= 2;
	if ((size_t)c->nb * 4 * (32 * c->pt + 4) * 4 > 200u * 1024)
		KG_FAIL(c, KG_ERR_INVALID, "n_used = %llu too wide for the exact kernel", (unsigned long long)c->n_used);
	c->p_alloc = (n_pheno + 7) / 8 * 8;
	const size_t lane_len = (size_t)c->nb * 128;
	std::vector<float> y_lane((size_t)c->p_alloc * lane_len, 0.0f);
	c->h_sums.assign(c->p_alloc, 0.0f);
	c->h_y.assign(y, y + (size_t)n_pheno * c->n_used);
	for (uint32_t p = 0; p < n_pheno; p++) {
		const float *yp = y + (size_t)p * c->n_used;
		float *yl = y_lane.data() + (size_t)p * lane_len;
		for (uint64_t i = 0; i < c->n_used; i++) {
			const uint32_t blk = (uint32_t)(i / 128), L = (uint32_t)((i % 128) / 32), t = 31 - (uint32_t)(i % 32);
			yl[(size_t)(blk * 4 + L) * 32 + t] = yp[i];
		}
		// update_scores_and_sum (:288-295): sequential fp32 sum over the PERMUTED padded vector,
		// permuted index = 128 blk + 4 t + L  (kmer_general.cpp:155-167)
		volatile float sum = 0.0f;
		for (uint32_t blk = 0; blk < c->nb; blk++)
			for (uint32_t t = 0; t < 32; t++)
				for (uint32_t L = 0; L < 4; L++) sum = sum + yl[(size_t)(blk * 4 + L) * 32 + t];
		c->h_sums[p] = sum;
	}
	cudaFree(c->d_y_lane); cudaFree(c->d_y_pair); cudaFree(c->d_sums); cudaFree(c->d_thr);
	c->d_y_lane = nullptr; c->d_y_pair = nullptr; c->d_sums = nullptr; c->d_thr = nullptr;
	KG_CUDA(c, dev_alloc_copy(&c->d_y_lane, y_lane));
	{
		// pair mode: y_pair[p][((b * 8 + t4) * 4 + L) * 4 + k] = y_lane[p][(4 b + L) * 32 + 4 t4 + k]: the four lanes of a
		// pair read 64 contiguous bytes per load instead of four different 128-byte lines
		std::vector<float> y_pair(y_lane.size());
		for (uint32_t p = 0; p < c->p_alloc; p++)
			for (uint32_t b = 0; b < c->nb; b++)
				for (uint32_t t4 = 0; t4 < 8; t4++)
					for (uint32_t L = 0; L < 4; L++)
						for (uint32_t k = 0; k < 4; k++)
							y_pair[(size_t)p * lane_len + ((size_t)(b * 8 + t4) * 4 + L) * 4 + k] =
							    y_lane[(size_t)p * lane_len + (size_t)(4 * b + L) * 32 + 4 * t4 + k];
		KG_CUDA(c, dev_alloc_copy(&c->d_y_pair, y_pair));
	}
	KG_CUDA(c, dev_alloc_copy(&c->d_sums, c->h_sums));
	c->h_thr.assign(c->p_alloc, -1.0);
	KG_CUDA(c, dev_alloc_copy(&c->d_thr, c->h_thr));
	for (int i = 0; i < 2; i++) {
		if (!c->iv[i].d_hits) {
			cudaError_t me = cudaMalloc((void **)&c->iv[i].d_hits, c->hit_capacity * sizeof(kg_hit));
			if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc hit buffer: %s", cudaGetErrorString(me));
		}
		KG_CUDA(c, cudaMemset(c->iv[i].d_cnt, 0, 8 * sizeof(unsigned long long)));
		c->iv[i].rows = 0;
		c->iv[i].closed = false;
		c->iv[i].used_filter = false;
	}
	c->cur = 0;
	c->d_hits = c->iv[0].d_hits;
	c->d_counters = c->iv[0].d_cnt;
	c->rows_seen_total = c->kept_total = 0;
	// (re)size the pinned threshold staging ring
	for (int i = 0; i < 4; i++) {
		KG_CUDA(c, cudaEventSynchronize(c->thr_stage[i].ev));
		if (c->thr_stage[i].h_thr) cudaFreeHost(c->thr_stage[i].h_thr);
		c->thr_stage[i].h_thr = nullptr;
		KG_CUDA(c, cudaMallocHost((void **)&c->thr_stage[i].h_thr, (size_t)c->p_alloc * sizeof(double)));
	}
	kg_status st = kg_tc_prepare_scan(c);
	return st;
}

// Host-driven mode: the thresholds travel as KERNEL ARGUMENTS (copied at launch, no staging buffer, no copy-engine round
// trip in front of the filter launch) when they fit the 4 KB argument space (P <= 128); more phenotypes go through
// pinned staging + cudaMemcpyAsync.  The filter's bound constants are then recomputed on the device (kg_tc_retune).
#define KG_RC_MAX_P 128
struct KgRoundConsts {
	double thr[KG_RC_MAX_P];
	uint32_t p_alloc;
};
__global__ void kg_round_constants_kernel(const KgRoundConsts rc, double *__restrict__ d_thr) {
	for (uint32_t i = threadIdx.x; i < rc.p_alloc; i += blockDim.x) d_thr[i] = rc.thr[i];
}

extern "C" kg_status kg_scan_set_thresholds(kg_ctx *c, const double *thr, uint32_t n_pheno) {
	if (!c || !thr) return KG_ERR_INVALID;
	if (n_pheno != c->n_pheno || !c->d_thr) KG_FAIL(c, KG_ERR_STATE, "kg_scan_set_thresholds: phenotypes not set / count mismatch");
	if (c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_scan_set_thresholds: thresholds live on the device while a selection is active (kg_select_begin)");
	KG_CUDA(c, cudaSetDevice(c->device));
	for (uint32_t p = 0; p < n_pheno; p++) c->h_thr[p] = thr[p];
	if (c->p_alloc <= KG_RC_MAX_P) {
		KgRoundConsts rc;
		for (uint32_t p = 0; p < c->p_alloc; p++) rc.thr[p] = c->h_thr[p];
		rc.p_alloc = c->p_alloc;
		kg_round_constants_kernel<<<1, 128, 0, c->stream>>>(rc, c->d_thr);
		KG_LAUNCH_CHECK(c);
	} else {
		// stream-ordered through a pinned staging slot: tiles already queued keep the thresholds they were submitted
		// with, and the host never waits for the device here
		kg_ctx::ThrStage &stg = c->thr_stage[c->thr_next];
		c->thr_next = (c->thr_next + 1) & 3;
		KG_CUDA(c, cudaEventSynchronize(stg.ev));
		memcpy(stg.h_thr, c->h_thr.data(), (size_t)c->p_alloc * sizeof(double));
		KG_CUDA(c, cudaMemcpyAsync(c->d_thr, stg.h_thr, (size_t)c->p_alloc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		KG_CUDA(c, cudaEventRecord(stg.ev, c->stream));
	}
	return kg_tc_retune(c, false);
}

template <int R, int PT, int MODE>
static kg_status launch_exact(kg_ctx *c, const KgScanParams &prm) {
	constexpr int GS = kg_ys_group_stride<PT>();
	const size_t smem = (size_t)c->nb * 4 * GS * sizeof(float);
	auto kern = kg_scan_exact_kernel<R, PT, MODE>;
	KG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int occ = 1;
	KG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
	if (occ < 1) occ = 1;
	const uint32_t p_tiles = (c->n_pheno + PT - 1) / PT;
	const uint64_t n_chunks = (prm.view.n_rows + (64 * R) - 1) / (64 * R);
	// one resident wave: never more CTAs than (SMs x CTAs per SM), or the few extra CTAs double the time
	uint64_t gx = ((uint64_t)c->sm_count * occ) / p_tiles;
	gx = std::max<uint64_t>(1, std::min(gx, n_chunks));
	dim3 grid((unsigned)gx, p_tiles);
	if (MODE != 2) timing_begin(c, KG_KERNEL_SCAN_EXACT, prm.view.n_rows);   // MODE 2 is timed by the caller (refine class)
	kern<<<grid, 256, smem, c->stream>>>(prm);
	if (MODE != 2) timing_end(c);
	KG_LAUNCH_CHECK(c);
	return KG_OK;
}

#ifndef KG_LIST_R
#define KG_LIST_R 4
#endif
// list mode (MODE 2): one grid column per tile of 8 filter columns; the lists' lengths are only known on the device,
// so the grid is one full wave and every CTA loops over its tile's list
static kg_status launch_exact_list(kg_ctx *c, const KgScanParams &prm, uint32_t n_tiles) {
	constexpr int GS = kg_ys_group_stride<8>();
	const size_t smem = (size_t)c->nb * 4 * GS * sizeof(float);
	auto kern = kg_scan_exact_kernel<KG_LIST_R, 8, 2>;
	KG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int occ = 1;
	KG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
	if (occ < 1) occ = 1;
	const uint64_t gx = std::max<uint64_t>(1, ((uint64_t)c->sm_count * occ) / n_tiles);
	dim3 grid((unsigned)gx, n_tiles);
	kern<<<grid, 256, smem, c->stream>>>(prm);
	KG_LAUNCH_CHECK(c);
	return KG_OK;
}

template <int MODE>
static kg_status launch_exact_pt(kg_ctx *c, const KgScanParams &prm) {
	switch (c->pt) {
	case 8: return launch_exact<8, 8, MODE>(c, prm);
	case 4: return launch_exact<8, 4, MODE>(c, prm);
	case 2: return launch_exact<8, 2, MODE>(c, prm);
	default: return launch_exact<8, 1, MODE>(c, prm);
	}
}

static KgScanParams scan_params(kg_ctx *c, const KgRowView &view, uint64_t first_row_id) {
	KgScanParams prm;
	memset(&prm, 0, sizeof prm);
	prm.view = view;
	prm.nb = c->nb;
	prm.n_used = (uint32_t)c->n_used;
	prm.n_pheno = c->n_pheno;
	prm.min_count = (uint32_t)std::min<uint64_t>(c->min_count, 0xFFFFFFFFull);
	prm.y_lane = c->d_y_lane;
	prm.y_pair = c->d_y_pair;
	prm.sums = c->d_sums;
	prm.mask32 = c->d_mask32;
	prm.thr = c->d_thr;
	prm.hits = c->d_hits;
	prm.hit_count = c->d_counters + 0;
	prm.hit_capacity = c->hit_capacity;
	prm.kept_count = c->d_counters + 1;
	prm.first_row_id = first_row_id;
	if (c->sel.active) {
		prm.cand = c->sel.d_cand;
		prm.cand_count = c->sel.d_cand_count;
		prm.cand_cap = c->sel.cand_cap;
		prm.kept_count = c->sel.d_status + KG_SEL_ST_ROUND_KEPT;
	}
	return prm;
}

static bool is_device_pointer(const void *p) {
	cudaPointerAttributes attr;
	if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

static kg_status scan_submit_one(kg_ctx *c, const uint64_t *rows, uint64_t n_rows, uint64_t first_row_id);

// Host tiles are cut into sub-tiles (kHostSubTileRows) so that the H2D copy of sub-tile i+1 runs under the kernels of
// sub-tile i (two device slots); device tiles are scanned in place, in one launch.

extern "C" kg_status kg_scan_submit(kg_ctx *c, const uint64_t *rows, uint64_t n_rows, uint64_t first_row_id) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_y_lane) KG_FAIL(c, KG_ERR_STATE, "kg_scan_submit: call kg_scan_set_phenotypes first");
	if (n_rows == 0) return KG_OK;
	if (!rows) KG_FAIL(c, KG_ERR_INVALID, "kg_scan_submit: null rows");
	if (n_rows >= (1ull << 31)) KG_FAIL(c, KG_ERR_INVALID, "kg_scan_submit: tile of %llu rows too large (max 2^31-1)", (unsigned long long)n_rows);
	KG_CUDA(c, cudaSetDevice(c->device));
	if (c->sel.active) return kg_sel_submit(c, rows, n_rows, first_row_id);
	if (n_rows <= kHostSubTileRows || is_device_pointer(rows)) return scan_submit_one(c, rows, n_rows, first_row_id);
	const size_t stride = (size_t)c->w_file + 1;
	for (uint64_t off = 0; off < n_rows; off += kHostSubTileRows) {
		const uint64_t n = std::min<uint64_t>(kHostSubTileRows, n_rows - off);
		kg_status st = scan_submit_one(c, rows + off * stride, n, first_row_id + off);
		if (st != KG_OK) return st;
	}
	return KG_OK;
}

static kg_status scan_submit_one(kg_ctx *c, const uint64_t *rows, uint64_t n_rows, uint64_t first_row_id) {
	const uint64_t *dev = nullptr;
	kg_status st = acquire_tile(c, rows, n_rows, &dev);
	if (st != KG_OK) return st;
	if (c->pat_attached) {
		st = patterns_attached_tile(c, dev, n_rows);
		if (st != KG_OK) return st;
	}
	bool use_tc = false;
	if (c->scan_engine == 2) use_tc = true;
	else if (c->scan_engine == 0) use_tc = kg_tc_scan_profitable(c);
	if (use_tc && !kg_tc_scan_available(c))
		KG_FAIL(c, KG_ERR_INVALID, "tensor filter engine unavailable for this shape: %s", c->tc.why_unavailable.c_str());
	if (use_tc) {
		st = kg_tc_scan_tile(c, dev, n_rows, first_row_id);
		if (st != KG_OK) return st;
		c->iv[c->cur].used_filter = true;
	} else {
		KgRowView view;
		st = memory_view(c, dev, n_rows, &view);
		if (st != KG_OK) return st;
		KgScanParams prm = scan_params(c, view, first_row_id);
		st = launch_exact_pt<0>(c, prm);
		if (st != KG_OK) return st;
	}
	c->iv[c->cur].rows += n_rows;
	return release_tile(c);
}

// resolve the timed launches that have completed (no stream sync)
static void timing_resolve_completed(kg_ctx *c) {
	size_t done = 0;
	while (done < c->timed_pending.size() && cudaEventQuery(c->timed_pending[done].end) == cudaSuccess) done++;
	cudaGetLastError();
	if (!done) return;
	std::vector<kg_ctx::TimedLaunch> rest(c->timed_pending.begin() + done, c->timed_pending.end());
	c->timed_pending.resize(done);
	timing_resolve(c);
	c->timed_pending.swap(rest);
}

extern "C" kg_status kg_scan_mark(kg_ctx *c) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_y_lane) KG_FAIL(c, KG_ERR_STATE, "kg_scan_mark: call kg_scan_set_phenotypes first");
	if (c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_scan_mark: no hit intervals while a selection is active (kg_select_sync / kg_select_export)");
	if (c->iv[c->cur ^ 1].closed)
		KG_FAIL(c, KG_ERR_STATE, "kg_scan_mark: the previous interval has not been fetched yet");
	KG_CUDA(c, cudaSetDevice(c->device));
	kg_ctx::ScanInterval &v = c->iv[c->cur];
	// the counters are read back on the D2H stream behind an event, so that the copy engine's round trip does not sit
	// between this interval's last kernel and the next interval's first one on the compute stream
	KG_CUDA(c, cudaEventRecord(v.closed_ev, c->stream));
	KG_CUDA(c, cudaStreamWaitEvent(c->d2h_stream, v.closed_ev, 0));
	KG_CUDA(c, cudaMemcpyAsync(v.h_cnt, v.d_cnt, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->d2h_stream));
	KG_CUDA(c, cudaEventRecord(v.done, c->d2h_stream));
	v.closed = true;
	c->cur ^= 1;
	kg_ctx::ScanInterval &n = c->iv[c->cur];
	n.rows = 0;
	n.used_filter = false;
	KG_CUDA(c, cudaMemsetAsync(n.d_cnt, 0, 8 * sizeof(unsigned long long), c->stream));
	c->d_hits = n.d_hits;
	c->d_counters = n.d_cnt;
	return KG_OK;
}

// an interval leaves the pipeline: totals, auto-engine policy
static void consume_interval(kg_ctx *c, kg_ctx::ScanInterval &v, bool count_rows) {
	if (count_rows) {
		c->rows_seen_total += v.rows;
		c->kept_total += v.h_cnt[1];
		if (c->timing) c->timed_rows[KG_KERNEL_SCAN_REFINE] += v.h_cnt[6];  // "rows" of the refine class = (row, 16-phenotype group) pairs re-scored
		// auto engine: the exact kernel re-scores 2 tiles of 8 phenotypes per listed (row, group); if that is more
		// work than scoring every row against every phenotype, the next interval runs dense
		if (v.used_filter && v.rows > 0)
			c->tc.use_filter = 16.0 * (double)v.h_cnt[6] < 0.8 * (double)v.rows * (double)c->n_pheno;
		else c->tc.use_filter = true;
	}
	v.closed = false;
}

extern "C" kg_status kg_scan_fetch(kg_ctx *c, kg_hit *out, size_t cap, size_t *n_hits, uint64_t *rows_seen,
                                   uint64_t *rows_kept) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_y_lane) KG_FAIL(c, KG_ERR_STATE, "kg_scan_fetch: call kg_scan_set_phenotypes first");
	if (c->sel.active) KG_FAIL(c, KG_ERR_STATE, "kg_scan_fetch: no hit intervals while a selection is active (kg_select_sync / kg_select_export)");
	KG_CUDA(c, cudaSetDevice(c->device));
	if (!c->iv[c->cur ^ 1].closed) {
		kg_status st = kg_scan_mark(c);
		if (st != KG_OK) return st;
	}
	kg_ctx::ScanInterval &v = c->iv[c->cur ^ 1];
	KG_CUDA(c, cudaEventSynchronize(v.done));
	timing_resolve_completed(c);
	const unsigned long long n = v.h_cnt[0];
	if (n > c->hit_capacity) {
		// the interval is dropped (its rows are not counted): the caller resubmits them in smaller pieces
		consume_interval(c, v, false);
		KG_FAIL(c, KG_ERR_HITS_OVERFLOW, "%llu hits exceed the hit buffer (%llu): resubmit smaller tiles", n,
		        (unsigned long long)c->hit_capacity);
	}
	if (n_hits) *n_hits = (size_t)n;
	if (rows_seen) *rows_seen = c->rows_seen_total + v.rows;
	if (rows_kept) *rows_kept = c->kept_total + v.h_cnt[1];
	if (out && cap >= n) {
		if (n) {
			KG_CUDA(c, cudaMemcpyAsync(out, v.d_hits, n * sizeof(kg_hit), cudaMemcpyDeviceToHost, c->d2h_stream));
			KG_CUDA(c, cudaStreamSynchronize(c->d2h_stream));
		}
		consume_interval(c, v, true);
	} else if (n == 0) {
		consume_interval(c, v, true);
	}
	return KG_OK;
}

extern "C" kg_status kg_scan_discard(kg_ctx *c) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_y_lane) return KG_OK;
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	timing_resolve(c);
	for (int i = 0; i < 2; i++) {
		c->iv[i].closed = false;
		c->iv[i].rows = 0;
		c->iv[i].used_filter = false;
		KG_CUDA(c, cudaMemsetAsync(c->iv[i].d_cnt, 0, 8 * sizeof(unsigned long long), c->stream));
	}
	return KG_OK;
}

extern "C" kg_status kg_scan_clear_hits(kg_ctx *c) {
	if (!c) return KG_ERR_INVALID;
	kg_ctx::ScanInterval &v = c->iv[c->cur ^ 1];
	if (v.closed) {
		KG_CUDA(c, cudaSetDevice(c->device));
		KG_CUDA(c, cudaEventSynchronize(v.done));
		consume_interval(c, v, v.h_cnt[0] <= c->hit_capacity);
	}
	return KG_OK;
}

// keep bits -> one byte per row
__global__ void kg_keep_bytes_kernel(const uint32_t *__restrict__ keep_bits, uint64_t n_rows, uint8_t *__restrict__ keep) {
	for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (uint64_t)gridDim.x * blockDim.x)
		keep[r] = (uint8_t)((keep_bits[r >> 5] >> (r & 31)) & 1u);
}

extern "C" kg_status kg_mac_filter(kg_ctx *c, const uint64_t *rows, uint64_t n_rows, uint64_t min_count, uint8_t *keep, uint64_t *kept) {
	if (!c) return KG_ERR_INVALID;
	if (kept) *kept = 0;
	if (n_rows == 0) return KG_OK;
	if (!rows || !keep) KG_FAIL(c, KG_ERR_INVALID, "kg_mac_filter: null argument");
	if (n_rows >= (1ull << 31)) KG_FAIL(c, KG_ERR_INVALID, "kg_mac_filter: tile too large (max 2^31-1 rows)");
	KG_CUDA(c, cudaSetDevice(c->device));
	const uint64_t *dev = nullptr;
	kg_status st = acquire_tile(c, rows, n_rows, &dev);
	if (st != KG_OK) return st;
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
	if (c->keep_bytes_cap < n_rows) {
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(c->d_keep_bytes);
		c->d_keep_bytes = nullptr;
		c->keep_bytes_cap = 0;
		const size_t cap = (size_t)((n_rows + 7) & ~7ull);
		cudaError_t me = cudaMalloc((void **)&c->d_keep_bytes, cap + sizeof(unsigned long long));
		if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc keep bytes: %s", cudaGetErrorString(me));
		c->keep_bytes_cap = cap;
	}
	uint8_t *d_keep = c->d_keep_bytes;
	unsigned long long *d_kept = reinterpret_cast<unsigned long long *>(c->d_keep_bytes + c->keep_bytes_cap);
	KG_CUDA(c, cudaMemsetAsync(d_kept, 0, sizeof(unsigned long long), c->stream));
	const KgRowView raw{dev, n_rows, c->w_file + 1, c->w_file};
	const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_rows + 255) / 256, (uint64_t)c->sm_count * 8));
	timing_begin(c, KG_KERNEL_AUX, n_rows);
	kg_prefilter_kernel<<<grid, 256, 0, c->stream>>>(raw, c->d_file_mask, (uint32_t)c->n_used,
	                                                 (uint32_t)std::min<uint64_t>(min_count, 0xFFFFFFFFull), c->d_keep_bits, d_kept);
	kg_keep_bytes_kernel<<<grid, 256, 0, c->stream>>>(c->d_keep_bits, n_rows, d_keep);
	timing_end(c);
	c->launches += 2;
	cudaError_t e0 = cudaGetLastError();
	st = release_tile(c);
	cudaError_t e1 = cudaStreamSynchronize(c->stream);
	cudaError_t e2 = cudaMemcpy(keep, d_keep, n_rows, cudaMemcpyDeviceToHost);
	unsigned long long h_kept = 0;
	cudaError_t e3 = cudaMemcpy(&h_kept, d_kept, sizeof h_kept, cudaMemcpyDeviceToHost);
	if (st != KG_OK) return st;
	KG_CUDA(c, e0); KG_CUDA(c, e1); KG_CUDA(c, e2); KG_CUDA(c, e3);
	if (kept) *kept = h_kept;
	return KG_OK;
}

extern "C" kg_status kg_scan_scores_dense(kg_ctx *c, const uint64_t *rows, uint64_t n_rows, uint8_t *keep,
                                          double *scores) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_y_lane) KG_FAIL(c, KG_ERR_STATE, "kg_scan_scores_dense: call kg_scan_set_phenotypes first");
	if (n_rows == 0) return KG_OK;
	if (!rows || !keep || !scores) KG_FAIL(c, KG_ERR_INVALID, "kg_scan_scores_dense: null argument");
	KG_CUDA(c, cudaSetDevice(c->device));
	const uint64_t *dev = nullptr;
	kg_status st = acquire_tile(c, rows, n_rows, &dev);
	if (st != KG_OK) return st;
	KgRowView view;
	st = memory_view(c, dev, n_rows, &view);
	if (st != KG_OK) return st;
	// context-owned output scratch, grown on demand: [n_pheno][n_rows] doubles, then n_rows keep flags
	const size_t need = (size_t)c->n_pheno * n_rows * sizeof(double) + n_rows;
	if (c->dense_cap < need) {
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(c->d_dense);
		c->d_dense = nullptr;
		c->dense_cap = 0;
		cudaError_t me = cudaMalloc((void **)&c->d_dense, need);
		if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc dense scores (%zu bytes): %s", need, cudaGetErrorString(me));
		c->dense_cap = need;
	}
	double *d_scores = reinterpret_cast<double *>(c->d_dense);
	uint8_t *d_keep = reinterpret_cast<uint8_t *>(c->d_dense) + (size_t)c->n_pheno * n_rows * sizeof(double);
	cudaMemsetAsync(d_scores, 0, (size_t)c->n_pheno * n_rows * sizeof(double), c->stream);
	KgScanParams prm = scan_params(c, view, 0);
	prm.cand = nullptr;   // dense mode writes scores, never candidates
	prm.keep_out = d_keep;
	prm.scores_out = d_scores;
	st = launch_exact_pt<1>(c, prm);
	if (st == KG_OK) st = release_tile(c);
	cudaError_t e1 = cudaStreamSynchronize(c->stream);
	cudaError_t e2 = cudaMemcpy(keep, d_keep, n_rows, cudaMemcpyDeviceToHost);
	cudaError_t e3 = cudaMemcpy(scores, d_scores, (size_t)c->n_pheno * n_rows * sizeof(double), cudaMemcpyDeviceToHost);
	if (st != KG_OK) return st;
	KG_CUDA(c, e1); KG_CUDA(c, e2); KG_CUDA(c, e3);
	return KG_OK;
}

extern "C" kg_status kg_scan_filter_shape(kg_ctx *c, uint32_t *n_pass, uint32_t *p_pad, uint32_t *k_pad, uint32_t *raw_stages) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_y_lane) KG_FAIL(c, KG_ERR_STATE, "kg_scan_filter_shape: call kg_scan_set_phenotypes first");
	const bool ok = c->tc.scan_ready;
	if (n_pass) *n_pass = ok ? c->tc.n_pass : 0;
	if (p_pad) *p_pad = ok ? c->tc.p_pad : 0;
	if (k_pad) *k_pad = ok ? 128 * ((c->w_file + 1) / 2) : 0;
	if (raw_stages) *raw_stages = ok ? c->tc.raw_stages : 0;
	return KG_OK;
}

extern "C" kg_status kg_scan_filter_sums(kg_ctx *c, const uint64_t *rows, uint64_t n_rows, int32_t *q, int8_t *yq) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_y_lane) KG_FAIL(c, KG_ERR_STATE, "kg_scan_filter_sums: call kg_scan_set_phenotypes first");
	if (n_rows == 0) return KG_OK;
	if (!rows || !q) KG_FAIL(c, KG_ERR_INVALID, "kg_scan_filter_sums: null argument");
	KG_CUDA(c, cudaSetDevice(c->device));
	const uint64_t *dev = nullptr;
	kg_status st = acquire_tile(c, rows, n_rows, &dev);
	if (st != KG_OK) return st;
	st = kg_tc_filter_debug(c, dev, n_rows, q, yq);
	if (st != KG_OK) return st;
	return release_tile(c);
}

// ------------------------------------------------------------------------------------- kinship
extern "C" size_t kg_kinship_accum_len(const kg_ctx *c) { return c ? (size_t)c->n_used * c->n_used + 1 : 0; }

extern "C" kg_status kg_kinship_begin(kg_ctx *c, uint64_t min_count, uint64_t *accum_dev) {
	if (!c) return KG_ERR_INVALID;
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	const size_t len = kg_kinship_accum_len(c);
	if (c->own_accum) { cudaFree(c->d_accum); c->own_accum = false; }
	c->d_accum = nullptr;
	if (accum_dev) {
		c->d_accum = reinterpret_cast<unsigned long long *>(accum_dev);
	} else {
		cudaError_t me = cudaMalloc((void **)&c->d_accum, len * 8);
		if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc kinship accumulator: %s", cudaGetErrorString(me));
		c->own_accum = true;
	}
	KG_CUDA(c, cudaMemsetAsync(c->d_accum, 0, len * 8, c->stream));
	c->kin_min_count = min_count;
	c->kin_active = true;
	return kg_tc_prepare_kinship(c);
}

extern "C" kg_status kg_kinship_submit(kg_ctx *c, const uint64_t *rows, uint64_t n_rows) {
	if (!c) return KG_ERR_INVALID;
	if (!c->kin_active) KG_FAIL(c, KG_ERR_STATE, "kg_kinship_submit: call kg_kinship_begin first");
	if (n_rows == 0) return KG_OK;
	if (!rows) KG_FAIL(c, KG_ERR_INVALID, "kg_kinship_submit: null rows");
	if (n_rows >= (1ull << 31)) KG_FAIL(c, KG_ERR_INVALID, "kg_kinship_submit: tile too large (max 2^31-1 rows)");
	KG_CUDA(c, cudaSetDevice(c->device));
	const uint64_t *dev = nullptr;
	kg_status st = acquire_tile(c, rows, n_rows, &dev);
	if (st != KG_OK) return st;
	bool use_tc = false;
	if (c->kin_engine == 2) use_tc = true;
	else if (c->kin_engine == 0) use_tc = kg_tc_kinship_available(c);
	if (use_tc && !kg_tc_kinship_available(c))
		KG_FAIL(c, KG_ERR_INVALID, "tensor kinship engine unavailable: %s", c->tc.why_unavailable.c_str());
	if (use_tc) {
		// raw rows straight to the tensor cores (file column order; MAC filter inside the kernel)
		st = kg_tc_kinship_tile(c, dev, n_rows);
		if (st != KG_OK) return st;
		return release_tile(c);
	}
	KgRowView view;
	st = memory_view(c, dev, n_rows, &view);
	if (st != KG_OK) return st;
	// MAC filter bits + kept count
	const size_t kb_need = (size_t)((n_rows + 31) / 32);
	if (c->keep_bits_cap < kb_need) {
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(c->d_keep_bits);
		c->d_keep_bits = nullptr;
		cudaError_t me = cudaMalloc((void **)&c->d_keep_bits, kb_need * 4);
		if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc keep bits: %s", cudaGetErrorString(me));
		c->keep_bits_cap = kb_need;
	}
	const uint64_t *mask = c->identity ? c->d_file_mask : c->d_mem_mask;
	const unsigned long long n2 = (unsigned long long)c->n_used * c->n_used;
	{
		const unsigned grid = (unsigned)std::min<uint64_t>((n_rows + 255) / 256, (uint64_t)c->sm_count * 8);
		timing_begin(c, KG_KERNEL_AUX, n_rows);
		kg_prefilter_kernel<<<std::max(grid, 1u), 256, 0, c->stream>>>(view, mask, (uint32_t)c->n_used,
		                                                               (uint32_t)std::min<uint64_t>(c->kin_min_count, 0xFFFFFFFFull),
		                                                               c->d_keep_bits, c->d_accum + n2);
		timing_end(c);
		KG_LAUNCH_CHECK(c);
	}
	{
		const uint32_t t64 = (uint32_t)((c->n_used + 63) / 64);
		const uint32_t pair_tiles = t64 * (t64 + 1) / 2;
		const uint64_t n_chunks = (n_rows + KG_KIN_CHUNK_ROWS - 1) / KG_KIN_CHUNK_ROWS;
		uint64_t splits = ((uint64_t)c->sm_count * 4 + pair_tiles - 1) / pair_tiles;
		splits = std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(splits, n_chunks), 65535));
		dim3 grid(pair_tiles, (unsigned)splits);
		timing_begin(c, KG_KERNEL_KINSHIP, n_rows);
		kg_kinship_popc_kernel<<<grid, 256, 0, c->stream>>>(view, c->d_keep_bits, (uint32_t)c->n_used, t64, c->d_accum);
		timing_end(c);
		KG_LAUNCH_CHECK(c);
	}
	return release_tile(c);
}

extern "C" kg_status kg_kinship_fetch(kg_ctx *c, uint64_t *ibs, uint64_t *kept_rows) {
	if (!c) return KG_ERR_INVALID;
	if (!c->kin_active) KG_FAIL(c, KG_ERR_STATE, "kg_kinship_fetch: call kg_kinship_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	{
		kg_status st = kg_tc_kinship_flush(c);
		if (st != KG_OK) return st;
	}
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	const unsigned long long n2 = (unsigned long long)c->n_used * c->n_used;
	unsigned long long M = 0;
	KG_CUDA(c, cudaMemcpy(&M, c->d_accum + n2, 8, cudaMemcpyDeviceToHost));
	if (kept_rows) *kept_rows = M;
	if (ibs) {
		if (!c->d_ibs) {
			cudaError_t me = cudaMalloc((void **)&c->d_ibs, n2 * 8);
			if (me != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc ibs: %s", cudaGetErrorString(me));
		}
		const unsigned grid = (unsigned)std::min<uint64_t>((n2 + 255) / 256, (uint64_t)c->sm_count * 8);
		kg_kinship_finalize_kernel<<<std::max(grid, 1u), 256, 0, c->stream>>>(c->d_accum, (uint32_t)c->n_used, M, c->d_ibs);
		KG_LAUNCH_CHECK(c);
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		KG_CUDA(c, cudaMemcpy(ibs, c->d_ibs, n2 * 8, cudaMemcpyDeviceToHost));
	}
	return KG_OK;
}

// ------------------------------------------------------------------------------------- synthetic
extern "C" kg_status kg_synth_rows_device(kg_ctx *c, uint64_t seed, uint64_t first_row, uint64_t n_rows,
                                          uint64_t *rows_dev) {
	if (!c) return KG_ERR_INVALID;
	if (n_rows == 0) return KG_OK;
	if (!rows_dev) KG_FAIL(c, KG_ERR_INVALID, "kg_synth_rows_device: null output");
	KG_CUDA(c, cudaSetDevice(c->device));
	const uint64_t last_mask = (c->n_file % 64) ? ((1ull << (c->n_file % 64)) - 1ull) : ~0ull;
	const uint64_t total = n_rows * (uint64_t)(c->w_file + 1);
	const unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)c->sm_count * 32);
	kg_synth_rows_kernel<<<std::max(grid, 1u), 256, 0, c->stream>>>(seed, first_row, n_rows, c->w_file, last_mask, rows_dev);
	KG_LAUNCH_CHECK(c);
	return KG_OK;
}

// ------------------------------------------------------------------------------------- distinct patterns
static kg_status patterns_read_counters(kg_ctx *c, unsigned long long *h3) {
	KG_CUDA(c, cudaMemcpyAsync(h3, c->d_pat_counters, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	c->pat_count = h3[0] + h3[1];
	return KG_OK;
}

// make room for `extra` more keys at a load factor <= 1/2 (rehash into a larger table if needed)
static kg_status patterns_reserve(kg_ctx *c, uint64_t extra) {
	uint64_t need = 2 * (c->pat_count + extra) + 1024;
	if (need <= c->pat_slots) return KG_OK;
	uint64_t slots = std::max<uint64_t>(c->pat_slots, 1ull << 16);
	while (slots < need) slots <<= 1;
	unsigned long long *nt = nullptr;
	cudaError_t e = cudaMalloc((void **)&nt, slots * 8);
	if (e != cudaSuccess) KG_FAIL(c, KG_ERR_NOMEM, "cudaMalloc pattern set (%llu slots): %s", (unsigned long long)slots, cudaGetErrorString(e));
	KG_CUDA(c, cudaMemsetAsync(nt, 0xFF, slots * 8, c->stream));
	if (c->d_pat_table) {
		KG_CUDA(c, cudaMemsetAsync(c->d_pat_counters, 0, sizeof(unsigned long long), c->stream));   // distinct is rebuilt; the other two stay
		const unsigned grid = (unsigned)std::min<uint64_t>((c->pat_slots + 255) / 256, (uint64_t)c->sm_count * 16);
		kg_patterns_rehash_kernel<<<grid, 256, 0, c->stream>>>(c->d_pat_table, c->pat_slots, nt, slots, c->d_pat_counters);
		KG_LAUNCH_CHECK(c);
		KG_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(c->d_pat_table);
	}
	c->d_pat_table = nt;
	c->pat_slots = slots;
	return KG_OK;
}

extern "C" kg_status kg_patterns_begin(kg_ctx *c, uint64_t expected) {
	if (!c) return KG_ERR_INVALID;
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaFree(c->d_pat_table);
	c->d_pat_table = nullptr;
	c->pat_slots = 0;
	c->pat_count = 0;
	if (!c->d_pat_counters) KG_CUDA(c, cudaMalloc((void **)&c->d_pat_counters, 4 * sizeof(unsigned long long)));
	KG_CUDA(c, cudaMemsetAsync(c->d_pat_counters, 0, 4 * sizeof(unsigned long long), c->stream));
	return patterns_reserve(c, expected);
}

extern "C" kg_status kg_patterns_submit(kg_ctx *c, const uint64_t *rows, uint64_t n_rows, uint64_t min_count) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_pat_counters) KG_FAIL(c, KG_ERR_STATE, "kg_patterns_submit: call kg_patterns_begin first");
	if (n_rows == 0) return KG_OK;
	if (!rows) KG_FAIL(c, KG_ERR_INVALID, "kg_patterns_submit: null rows");
	if (n_rows >= (1ull << 31)) KG_FAIL(c, KG_ERR_INVALID, "kg_patterns_submit: tile too large (max 2^31-1 rows)");
	KG_CUDA(c, cudaSetDevice(c->device));
	unsigned long long h3[3];
	kg_status st = patterns_read_counters(c, h3);   // exact fill before the tile (the previous tile has completed)
	if (st != KG_OK) return st;
	st = patterns_reserve(c, n_rows);
	if (st != KG_OK) return st;
	const uint64_t *dev = nullptr;
	st = acquire_tile(c, rows, n_rows, &dev);
	if (st != KG_OK) return st;
	KgRowView view;
	st = memory_view(c, dev, n_rows, &view);
	if (st != KG_OK) return st;
	const uint64_t *mask = c->identity ? c->d_file_mask : c->d_mem_mask;
	const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_rows + 255) / 256, (uint64_t)c->sm_count * 16));
	timing_begin(c, KG_KERNEL_AUX, n_rows);
	kg_patterns_rows_kernel<<<grid, 256, 0, c->stream>>>(view, mask, c->w_mem, (uint32_t)c->n_used,
	                                                    (uint32_t)std::min<uint64_t>(min_count, 0xFFFFFFFFull), c->d_pat_table, c->pat_slots, c->d_pat_counters);
	timing_end(c);
	KG_LAUNCH_CHECK(c);
	return release_tile(c);
}

// the tile of a kg_scan_submit, already on the device: count its patterns too (no growth here: kg_patterns_attach
// made room for every row it was promised)
static kg_status patterns_attached_tile(kg_ctx *c, const uint64_t *dev, uint64_t n_rows) {
	if (n_rows > c->pat_budget)
		KG_FAIL(c, KG_ERR_STATE, "kg_patterns_attach: more rows submitted than the pattern set was sized for");
	c->pat_budget -= n_rows;
	KgRowView view;
	kg_status st = memory_view(c, dev, n_rows, &view);
	if (st != KG_OK) return st;
	const uint64_t *mask = c->identity ? c->d_file_mask : c->d_mem_mask;
	const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_rows + 255) / 256, (uint64_t)c->sm_count * 16));
	timing_begin(c, KG_KERNEL_AUX, n_rows);
	kg_patterns_rows_kernel<<<grid, 256, 0, c->stream>>>(view, mask, c->w_mem, (uint32_t)c->n_used,
	                                                    (uint32_t)std::min<uint64_t>(c->pat_min_count, 0xFFFFFFFFull), c->d_pat_table, c->pat_slots, c->d_pat_counters);
	timing_end(c);
	KG_LAUNCH_CHECK(c);
	return KG_OK;
}

extern "C" kg_status kg_patterns_attach(kg_ctx *c, uint64_t min_count, uint64_t max_rows) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_pat_counters) KG_FAIL(c, KG_ERR_STATE, "kg_patterns_attach: call kg_patterns_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	if (max_rows == 0) { c->pat_attached = false; return KG_OK; }
	unsigned long long h3[3];
	kg_status st = patterns_read_counters(c, h3);
	if (st != KG_OK) return st;
	st = patterns_reserve(c, max_rows);
	if (st != KG_OK) return st;
	c->pat_attached = true;
	c->pat_min_count = min_count;
	c->pat_budget = max_rows;
	return KG_OK;
}

extern "C" kg_status kg_patterns_count(kg_ctx *c, uint64_t *distinct, uint64_t *rows_kept) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_pat_counters) KG_FAIL(c, KG_ERR_STATE, "kg_patterns_count: call kg_patterns_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	unsigned long long h3[3];
	kg_status st = patterns_read_counters(c, h3);
	if (st != KG_OK) return st;
	if (distinct) *distinct = h3[0] + h3[1];
	if (rows_kept) *rows_kept = h3[2];
	return KG_OK;
}

extern "C" kg_status kg_patterns_export(kg_ctx *c, uint64_t *keys, uint64_t cap, uint64_t *n) {
	if (!c || !n) return KG_ERR_INVALID;
	if (!c->d_pat_counters) KG_FAIL(c, KG_ERR_STATE, "kg_patterns_export: call kg_patterns_begin first");
	KG_CUDA(c, cudaSetDevice(c->device));
	unsigned long long h3[3];
	kg_status st = patterns_read_counters(c, h3);
	if (st != KG_OK) return st;
	*n = h3[0] + h3[1];
	if (!keys || cap < *n) return KG_OK;   // size query
	unsigned long long *dst = reinterpret_cast<unsigned long long *>(keys);
	const bool dev = is_device_pointer(keys);
	unsigned long long *tmp = nullptr;
	if (!dev) {
		KG_CUDA(c, cudaMalloc((void **)&tmp, std::max<uint64_t>(*n, 1) * 8));
		dst = tmp;
	}
	KG_CUDA(c, cudaMemsetAsync(c->d_pat_counters + 3, 0, sizeof(unsigned long long), c->stream));
	const unsigned grid = (unsigned)std::min<uint64_t>((c->pat_slots + 255) / 256, (uint64_t)c->sm_count * 16);
	kg_patterns_export_kernel<<<grid, 256, 0, c->stream>>>(c->d_pat_table, c->pat_slots, dst, h3[0], c->d_pat_counters + 3);
	c->launches++;
	cudaError_t e0 = cudaGetLastError();
	cudaError_t e1 = cudaStreamSynchronize(c->stream);
	cudaError_t e2 = cudaSuccess;
	if (!dev && e0 == cudaSuccess && e1 == cudaSuccess) {
		e2 = cudaMemcpy(keys, tmp, h3[0] * 8, cudaMemcpyDeviceToHost);
		if (h3[1]) keys[h3[0]] = KG_PAT_EMPTY;
	} else if (dev && h3[1]) {
		const unsigned long long k = KG_PAT_EMPTY;
		e2 = cudaMemcpy(dst + h3[0], &k, 8, cudaMemcpyHostToDevice);
	}
	cudaFree(tmp);
	KG_CUDA(c, e0); KG_CUDA(c, e1); KG_CUDA(c, e2);
	return KG_OK;
}

extern "C" kg_status kg_patterns_insert(kg_ctx *c, const uint64_t *keys, uint64_t n) {
	if (!c) return KG_ERR_INVALID;
	if (!c->d_pat_counters) KG_FAIL(c, KG_ERR_STATE, "kg_patterns_insert: call kg_patterns_begin first");
	if (n == 0) return KG_OK;
	if (!keys) KG_FAIL(c, KG_ERR_INVALID, "kg_patterns_insert: null keys");
	KG_CUDA(c, cudaSetDevice(c->device));
	unsigned long long h3[3];
	kg_status st = patterns_read_counters(c, h3);
	if (st != KG_OK) return st;
	st = patterns_reserve(c, n);
	if (st != KG_OK) return st;
	const unsigned long long *src = reinterpret_cast<const unsigned long long *>(keys);
	unsigned long long *tmp = nullptr;
	if (!is_device_pointer(keys)) {
		KG_CUDA(c, cudaMalloc((void **)&tmp, n * 8));
		cudaError_t e = cudaMemcpyAsync(tmp, keys, n * 8, cudaMemcpyHostToDevice, c->stream);
		if (e != cudaSuccess) { cudaFree(tmp); KG_CUDA(c, e); }
		src = tmp;
	}
	const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 16));
	kg_patterns_keys_kernel<<<grid, 256, 0, c->stream>>>(src, n, c->d_pat_table, c->pat_slots, c->d_pat_counters);
	c->launches++;
	cudaError_t e0 = cudaGetLastError();
	cudaError_t e1 = cudaStreamSynchronize(c->stream);
	cudaFree(tmp);
	KG_CUDA(c, e0); KG_CUDA(c, e1);
	return KG_OK;
}

// ------------------------------------------------------------------------------------- NCCL (kinship all-reduce)
// The library does not link NCCL: the symbols are resolved at run time from libnccl.so.2 (the copy the process
// already has loaded -- e.g. PyTorch's -- or the system one).
#include <dlfcn.h>
#include <nccl.h>
namespace {
struct KgNccl {
	void *lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	std::string err;
	bool load() {
		if (lib) return true;
		// NCCL writes its version banner (NCCL_DEBUG=VERSION or higher) to stdout unless told otherwise; stdout belongs to
		// the caller (emma_kinship_kmers prints the matrix there)
		setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
		for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
			lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (lib) break;
		}
		if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define KG_NCCL_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); if (!field) { err = std::string("libnccl lacks ") + name; lib = nullptr; return false; }
		KG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
		KG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
		KG_NCCL_SYM(CommInitAll, "ncclCommInitAll")
		KG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
		KG_NCCL_SYM(AllReduce, "ncclAllReduce")
		KG_NCCL_SYM(GroupStart, "ncclGroupStart")
		KG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
		KG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef KG_NCCL_SYM
		return true;
	}
} g_nccl;
}  // namespace

// NCCL prints its version banner with the NCCL_DEBUG=VERSION level straight to stdout (NCCL_DEBUG_FILE only applies to
// higher levels); stdout belongs to the caller (emma_kinship_kmers prints the matrix there), so fd 1 points at stderr
// while a communicator is being created.
#include <unistd.h>
struct KgStdoutToStderr {
	int saved;
	KgStdoutToStderr() { fflush(stdout); saved = dup(1); if (saved >= 0) dup2(2, 1); }
	~KgStdoutToStderr() { fflush(stdout); if (saved >= 0) { dup2(saved, 1); close(saved); } }
};

#define KG_NCCL(ctx, expr)                                                                         \
	do {                                                                                           \
		ncclResult_t r_ = (expr);                                                                  \
		if (r_ != ncclSuccess) KG_FAIL(ctx, KG_ERR_CUDA, "%s failed: %s", #expr, g_nccl.GetErrorString(r_)); \
	} while (0)

static void kg_comm_destroy(kg_ctx *c) {
	if (c->nccl_comm && g_nccl.lib) g_nccl.CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
	c->nccl_comm = nullptr;
}

extern "C" kg_status kg_comm_unique_id(void *id128) {
	if (!id128) return KG_ERR_INVALID;
	if (!g_nccl.load()) { g_create_error = g_nccl.err; return KG_ERR_CUDA; }
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	KgStdoutToStderr quiet;
	if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return KG_ERR_CUDA; }
	memcpy(id128, &id, sizeof id);
	return KG_OK;
}

extern "C" kg_status kg_comm_init_rank(kg_ctx *c, const void *id128, int n_ranks, int rank) {
	if (!c || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return KG_ERR_INVALID;
	if (!g_nccl.load()) KG_FAIL(c, KG_ERR_CUDA, "%s", g_nccl.err.c_str());
	KG_CUDA(c, cudaSetDevice(c->device));
	kg_comm_destroy(c);
	ncclUniqueId id;
	memcpy(&id, id128, sizeof id);
	ncclComm_t comm = nullptr;
	KgStdoutToStderr quiet;
	KG_NCCL(c, g_nccl.CommInitRank(&comm, n_ranks, id, rank));
	c->nccl_comm = comm;
	return KG_OK;
}

extern "C" kg_status kg_comm_init_all(kg_ctx *const *ctxs, int n) {
	if (!ctxs || n < 1) return KG_ERR_INVALID;
	kg_ctx *c0 = ctxs[0];
	if (!g_nccl.load()) KG_FAIL(c0, KG_ERR_CUDA, "%s", g_nccl.err.c_str());
	std::vector<int> devs(n);
	for (int i = 0; i < n; i++) { devs[i] = ctxs[i]->device; kg_comm_destroy(ctxs[i]); }
	std::vector<ncclComm_t> comms(n, nullptr);
	KgStdoutToStderr quiet;
	KG_NCCL(c0, g_nccl.CommInitAll(comms.data(), n, devs.data()));
	for (int i = 0; i < n; i++) ctxs[i]->nccl_comm = comms[i];
	return KG_OK;
}

// sum the kinship accumulators ([n_used^2] Gram counts + kept rows, u64: exact) over the communicator's ranks, in place
static kg_status kinship_allreduce_enqueue(kg_ctx *c) {
	if (!c->kin_active) KG_FAIL(c, KG_ERR_STATE, "kg_kinship_allreduce: call kg_kinship_begin first");
	if (!c->nccl_comm) KG_FAIL(c, KG_ERR_STATE, "kg_kinship_allreduce: no communicator (kg_comm_init_rank / kg_comm_init_all)");
	KG_CUDA(c, cudaSetDevice(c->device));
	kg_status st = kg_tc_kinship_flush(c);
	if (st != KG_OK) return st;
	KG_NCCL(c, g_nccl.AllReduce(c->d_accum, c->d_accum, kg_kinship_accum_len(c), ncclUint64, ncclSum, static_cast<ncclComm_t>(c->nccl_comm), c->stream));
	return KG_OK;
}

extern "C" kg_status kg_kinship_allreduce(kg_ctx *c) {
	if (!c) return KG_ERR_INVALID;
	return kinship_allreduce_enqueue(c);
}

extern "C" kg_status kg_kinship_allreduce_all(kg_ctx *const *ctxs, int n) {
	if (!ctxs || n < 1) return KG_ERR_INVALID;
	if (!g_nccl.load()) KG_FAIL(ctxs[0], KG_ERR_CUDA, "%s", g_nccl.err.c_str());
	KG_NCCL(ctxs[0], g_nccl.GroupStart());
	kg_status st = KG_OK;
	for (int i = 0; i < n && st == KG_OK; i++) st = kinship_allreduce_enqueue(ctxs[i]);
	ncclResult_t r = g_nccl.GroupEnd();
	if (st != KG_OK) return st;
	if (r != ncclSuccess) KG_FAIL(ctxs[0], KG_ERR_CUDA, "ncclGroupEnd failed: %s", g_nccl.GetErrorString(r));
	return KG_OK;
}

// ------------------------------------------------------------------------------------- SNP twin of the scan
extern "C" kg_status kg_snps_scores(int device, const uint8_t *bed, uint64_t n_snps, uint32_t bytes_per_snp, const uint32_t *map_byte,
                                    const uint32_t *map_shift, uint32_t n_samples, const float *y, uint32_t n_pheno, double mac, double *scores) {
	auto fail = [](const char *what, cudaError_t e) { g_create_error = std::string(what) + ": " + cudaGetErrorString(e); return KG_ERR_CUDA; };
	if (!bed || !map_byte || !map_shift || !y || !scores || n_samples == 0 || n_pheno == 0) { g_create_error = "kg_snps_scores: bad arguments"; return KG_ERR_INVALID; }
	if (n_snps == 0) return KG_OK;
	int n_dev = 0;
	cudaError_t e = cudaGetDeviceCount(&n_dev);
	if (e != cudaSuccess || n_dev == 0) { g_create_error = "kg_snps_scores: no CUDA device; this library has no CPU fallback"; return KG_ERR_CUDA; }
	if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
	const uint32_t nb = (n_samples + 127) / 128, w_mem = 2 * nb;
	for (uint32_t i = 0; i < n_samples; i++)
		if (map_byte[i] >= bytes_per_snp || map_shift[i] > 6) { g_create_error = "kg_snps_scores: sample map out of range"; return KG_ERR_INVALID; }
	// phenotypes in lane order (permute_scores, kmer_general.cpp:155-167): y_lane[(4 b + L) * 32 + t] = y[128 b + 32 L + 31 - t]
	std::vector<float> y_lane((size_t)n_pheno * nb * 128, 0.0f);
	for (uint32_t p = 0; p < n_pheno; p++)
		for (uint32_t i = 0; i < n_samples; i++) {
			const uint32_t blk = i / 128, L = (i % 128) / 32, t = 31 - (i % 32);
			y_lane[(size_t)p * nb * 128 + (size_t)(blk * 4 + L) * 32 + t] = y[(size_t)p * n_samples + i];
		}
	uint8_t *d_bed = nullptr;
	uint32_t *d_mb = nullptr, *d_ms = nullptr;
	uint64_t *d_planes = nullptr;
	double *d_sums = nullptr, *d_scores = nullptr;
	float *d_y = nullptr;
	kg_status st = KG_OK;
	auto cleanup = [&] { cudaFree(d_bed); cudaFree(d_mb); cudaFree(d_ms); cudaFree(d_planes); cudaFree(d_sums); cudaFree(d_scores); cudaFree(d_y); };
#define KG_SNP_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { st = fail(#expr, e_); cleanup(); return st; } } while (0)
	KG_SNP_TRY(cudaMalloc((void **)&d_bed, n_snps * bytes_per_snp));
	KG_SNP_TRY(cudaMemcpy(d_bed, bed, n_snps * bytes_per_snp, cudaMemcpyHostToDevice));
	KG_SNP_TRY(cudaMalloc((void **)&d_mb, n_samples * 4));
	KG_SNP_TRY(cudaMalloc((void **)&d_ms, n_samples * 4));
	KG_SNP_TRY(cudaMemcpy(d_mb, map_byte, n_samples * 4, cudaMemcpyHostToDevice));
	KG_SNP_TRY(cudaMemcpy(d_ms, map_shift, n_samples * 4, cudaMemcpyHostToDevice));
	KG_SNP_TRY(cudaMalloc((void **)&d_planes, 3 * n_snps * w_mem * 8));
	KG_SNP_TRY(cudaMalloc((void **)&d_sums, 3 * n_snps * 8));
	KG_SNP_TRY(cudaMalloc((void **)&d_scores, (size_t)n_pheno * n_snps * 8));
	KG_SNP_TRY(cudaMalloc((void **)&d_y, y_lane.size() * 4));
	KG_SNP_TRY(cudaMemcpy(d_y, y_lane.data(), y_lane.size() * 4, cudaMemcpyHostToDevice));
	uint64_t *d_pa = d_planes, *d_nm = d_planes + n_snps * w_mem, *d_het = d_planes + 2 * n_snps * w_mem;
	double *d_sg = d_sums, *d_sn = d_sums + n_snps, *d_sg2 = d_sums + 2 * n_snps;
	const unsigned g1 = (unsigned)std::min<uint64_t>((n_snps + 127) / 128, 148 * 16);
	kg_snp_planes_kernel<<<std::max(g1, 1u), 128>>>(d_bed, n_snps, bytes_per_snp, d_mb, d_ms, n_samples, w_mem, d_pa, d_nm, d_het, d_sg, d_sn, d_sg2);
	KG_SNP_TRY(cudaGetLastError());
	const unsigned g2 = (unsigned)std::min<uint64_t>((n_snps * 4 + 255) / 256, 148 * 8);
	const size_t smem = (size_t)nb * 128 * sizeof(float);
	KG_SNP_TRY(cudaFuncSetAttribute(kg_snp_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	kg_snp_scores_kernel<<<std::max(g2, 1u), 256, smem>>>(d_pa, d_nm, d_het, d_sg, d_sn, d_sg2, n_snps, nb, d_y, n_pheno, mac, d_scores);
	KG_SNP_TRY(cudaGetLastError());
	KG_SNP_TRY(cudaDeviceSynchronize());
	KG_SNP_TRY(cudaMemcpy(scores, d_scores, (size_t)n_pheno * n_snps * 8, cudaMemcpyDeviceToHost));
#undef KG_SNP_TRY
	cleanup();
	return KG_OK;
}

// ------------------------------------------------------------------------------------- table construction
extern "C" kg_status kg_table_build(int device, const uint64_t *all_kmers, uint64_t n_all, uint32_t n_acc, const uint64_t *packed,
                                    const uint64_t *offsets, uint64_t *table) {
	auto fail = [](const char *what, cudaError_t e) { g_create_error = std::string(what) + ": " + cudaGetErrorString(e); return KG_ERR_CUDA; };
	if (n_all == 0) return KG_OK;
	if (!all_kmers || !offsets || !table || n_acc == 0) { g_create_error = "kg_table_build: bad arguments"; return KG_ERR_INVALID; }
	const uint64_t total = offsets[n_acc];
	if (total && !packed) { g_create_error = "kg_table_build: null k-mer lists"; return KG_ERR_INVALID; }
	int n_dev = 0;
	cudaError_t e = cudaGetDeviceCount(&n_dev);
	if (e != cudaSuccess || n_dev == 0) { g_create_error = "kg_table_build: no CUDA device; this library has no CPU fallback"; return KG_ERR_CUDA; }
	if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
	const uint32_t w = (n_acc + 63) / 64;
	uint64_t *d_all = nullptr, *d_packed = nullptr, *d_off = nullptr, *d_table = nullptr;
	kg_status st = KG_OK;
	auto cleanup = [&] { cudaFree(d_all); cudaFree(d_packed); cudaFree(d_off); cudaFree(d_table); };
#define KG_TB_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { st = fail(#expr, e_); cleanup(); return st; } } while (0)
	KG_TB_TRY(cudaMalloc((void **)&d_all, n_all * 8));
	KG_TB_TRY(cudaMemcpy(d_all, all_kmers, n_all * 8, cudaMemcpyHostToDevice));
	KG_TB_TRY(cudaMalloc((void **)&d_off, (size_t)(n_acc + 1) * 8));
	KG_TB_TRY(cudaMemcpy(d_off, offsets, (size_t)(n_acc + 1) * 8, cudaMemcpyHostToDevice));
	KG_TB_TRY(cudaMalloc((void **)&d_packed, std::max<uint64_t>(total, 1) * 8));
	if (total) KG_TB_TRY(cudaMemcpy(d_packed, packed, total * 8, cudaMemcpyHostToDevice));
	KG_TB_TRY(cudaMalloc((void **)&d_table, n_all * (uint64_t)(w + 1) * 8));
	const unsigned g1 = (unsigned)std::min<uint64_t>((n_all * (w + 1) + 255) / 256, 148 * 16);
	kg_table_init_kernel<<<std::max(g1, 1u), 256>>>(d_all, n_all, w, d_table);
	KG_TB_TRY(cudaGetLastError());
	if (total) {
		const unsigned g2 = (unsigned)std::min<uint64_t>((total + 255) / 256, 148 * 16);
		kg_table_mark_kernel<<<std::max(g2, 1u), 256>>>(d_all, n_all, d_packed, d_off, n_acc, w, reinterpret_cast<unsigned long long *>(d_table));
		KG_TB_TRY(cudaGetLastError());
	}
	KG_TB_TRY(cudaDeviceSynchronize());
	KG_TB_TRY(cudaMemcpy(table, d_table, n_all * (uint64_t)(w + 1) * 8, cudaMemcpyDeviceToHost));
#undef KG_TB_TRY
	cleanup();
	return KG_OK;
}

// ------------------------------------------------------------------------------------- probes
extern "C" kg_status kg_probe_int8_peak(kg_ctx *c, double *tops) {
	if (!c || !tops) return KG_ERR_INVALID;
	KG_CUDA(c, cudaSetDevice(c->device));
	KG_CUDA(c, cudaFuncSetAttribute(kg_probe_umma_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KG_PROBE_SMEM));
	cudaEvent_t e0, e1;
	KG_CUDA(c, cudaEventCreate(&e0));
	KG_CUDA(c, cudaEventCreate(&e1));
	const int n_mma = 16384;
	double best = 0.0;
	for (int rep = 0; rep < 4; rep++) {   // rep 0 = warm-up
		cudaEventRecord(e0, c->stream);
		kg_probe_umma_i8_kernel<<<c->sm_count, 128, KG_PROBE_SMEM, c->stream>>>(n_mma);
		cudaEventRecord(e1, c->stream);
		c->launches++;
		cudaError_t e = cudaStreamSynchronize(c->stream);
		if (e != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); KG_CUDA(c, e); }
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		const double ops = 2.0 * 128 * 256 * 32 * (double)n_mma * c->sm_count;
		if (rep > 0 && ms > 0.f) best = std::max(best, ops / (ms * 1e-3) / 1e12);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	*tops = best;
	return KG_OK;
}
