/* oracle/oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the kmersGWAS association hot path, used as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under kmersgwas_b200/ (the product) may include, link or call this file.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors of its own
 * (SURVEY.md section 4), so this restatement is pinned against the UNMODIFIED reference
 * compiled here by oracle/Makefile into oracle/_ref/ (tests/test_oracle_vs_ref.py) and against
 * fixtures generated from that build (tests/golden/, made by tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/).  Arithmetic notes:
 *   - float32 adds are performed in exactly the reference's order; this file must be built
 *     with -ffp-contract=off (the reference is built for SSE4.2: no FMA, FLT_EVAL_METHOD 0).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------
 * src/kmer_general.cpp:155-167  permute_scores
 *   R[offset + 4*i + j/32] = V[offset + 31 - i + j],  i in 0..31, j in {0,32,64,96}
 * ------------------------------------------------------------------------------------- */
void kgo_permute_scores(const float *v, size_t n_pad, float *out) {
	size_t index = 0;
	for (size_t offset = 0; offset < n_pad; offset += 128)
		for (size_t i = 0; i < 32; i++)
			for (size_t j = 0; j < 128; j += 32)
				out[index++] = v[31 - i + j + offset];
}

/* src/kmers_multiple_databases.cpp:288-295  update_scores_and_sum
 *   resize to n_pad with zeros, permute, sum sequentially in float32 over the PERMUTED vector */
float kgo_update_scores_and_sum(const float *y, size_t n, size_t n_pad, float *y_perm_out) {
	float *tmp = (float *)calloc(n_pad, sizeof(float));
	memcpy(tmp, y, n * sizeof(float));
	kgo_permute_scores(tmp, n_pad, y_perm_out);
	free(tmp);
	volatile float sum = 0.0f;
	for (size_t i = 0; i < n_pad; i++) sum = sum + y_perm_out[i];
	return sum;
}

/* src/kmers_multiple_databases.cpp:149-154  calculate_unsqueezed_popcnt */
uint64_t kgo_masked_popcount(const uint64_t *file_row, const uint64_t *mask, size_t w_file) {
	uint64_t res = 0;
	for (size_t i = 0; i < w_file; i++) res += (uint64_t)__builtin_popcountll(file_row[1 + i] & mask[i]);
	return res;
}

/* src/kmers_multiple_databases.cpp:125-132  the "squeeze": memory column i = file column
 * (map_word[i], map_bit[i]); memory row is w_mem zero-padded words. */
void kgo_squeeze_row(const uint64_t *file_row, const uint32_t *map_word, const uint32_t *map_bit,
                     size_t n, size_t w_mem, uint64_t *mem_row) {
	memset(mem_row, 0, w_mem * sizeof(uint64_t));
	for (size_t col = 0; col < n; col++) {
		uint64_t bit = (file_row[map_word[col] + 1] >> map_bit[col]) & 1ull;
		mem_row[col >> 6] |= bit << (col & 63);
	}
}

/* src/kmers_multiple_databases.cpp:327-363  calculate_kmer_score
 * The SSE4.1 loop, scalarised: the 128-bit mask register holds words h, h+1; float lane L is
 * 32-bit half L of that register; blendv selects on the sign bit (bit 31) of each half; the
 * mask shifts left by one every step; four independent float32 accumulators; final
 * ((l0+l1)+l2)+l3 in float32 (:358), then the double epilogue (:359-361) with no FMA. */
double kgo_score_row(const uint64_t *mem_row, size_t w_mem, const float *y_perm, float sum,
                     size_t n, double n1, uint64_t min_in_group) {
	double N = (double)n;
	double N1 = n1;
	double N0 = N - N1;
	if (!(((double)min_in_group <= N0) && ((double)min_in_group <= N1))) return 0.0;
	volatile float lane[4] = {0.0f, 0.0f, 0.0f, 0.0f};
	size_t j = 0;
	for (size_t h = 0; h < w_mem; h += 2, j += 128) {
		uint32_t m[4];
		m[0] = (uint32_t)(mem_row[h] & 0xffffffffu);
		m[1] = (uint32_t)(mem_row[h] >> 32);
		m[2] = (uint32_t)(mem_row[h + 1] & 0xffffffffu);
		m[3] = (uint32_t)(mem_row[h + 1] >> 32);
		for (size_t i = 0; i < 128; i += 4) {
			for (int L = 0; L < 4; L++) {
				float f = y_perm[j + i + L];
				float z = (m[L] & 0x80000000u) ? f : 0.0f;
				lane[L] = lane[L] + z;
				m[L] <<= 1;
			}
		}
	}
	volatile float s01 = lane[0] + lane[1];
	volatile float s012 = s01 + lane[2];
	volatile float s = s012 + lane[3];
	double yigi = (double)s;
	volatile double a = N * yigi;
	volatile double b = N1 * (double)sum;
	volatile double r = a - b;
	volatile double r2 = r * r;
	volatile double d1 = N * N1;
	volatile double d2 = N1 * N1;
	volatile double den = d1 - d2;
	return r2 / den;
}

static void build_mask(const uint32_t *map_word, const uint32_t *map_bit, size_t n, size_t w_file,
                       uint64_t *mask) {
	/* src/kmers_multiple_databases.cpp:297-311  create_map_from_all_DBs */
	memset(mask, 0, w_file * sizeof(uint64_t));
	for (size_t i = 0; i < n; i++) mask[map_word[i]] |= 1ull << map_bit[i];
}

/* src/kmers_multiple_databases.cpp:103-146 (load_kmers) + :275-284 (add_kmers_to_heap), heap left
 * to the caller. */
uint64_t kgo_scan_scores(const uint64_t *table, uint64_t n_rows, size_t w_file,
                         const uint32_t *map_word, const uint32_t *map_bit, size_t n,
                         const float *y, size_t n_pheno, uint64_t min_count,
                         uint8_t *keep, double *scores) {
	size_t w_mem = 2 * ((n + 127) / 128); /* :51 */
	size_t n_pad = w_mem * 64;
	uint64_t *mask = (uint64_t *)malloc(w_file * sizeof(uint64_t));
	uint64_t *mem = (uint64_t *)malloc(w_mem * sizeof(uint64_t));
	float *yperm = (float *)malloc(n_pheno * n_pad * sizeof(float));
	float *sums = (float *)malloc(n_pheno * sizeof(float));
	build_mask(map_word, map_bit, n, w_file, mask);
	for (size_t p = 0; p < n_pheno; p++)
		sums[p] = kgo_update_scores_and_sum(y + p * n, n, n_pad, yperm + p * n_pad);
	uint64_t kept = 0;
	size_t stride = w_file + 1;
	for (uint64_t r = 0; r < n_rows; r++) {
		const uint64_t *row = table + r * stride;
		uint64_t cnt = kgo_masked_popcount(row, mask, w_file);
		int k = (cnt >= min_count) && (cnt <= (uint64_t)n - min_count); /* :121 (unsigned compare) */
		if ((uint64_t)n < min_count) k = 0; /* n - mac would wrap in the reference; never used */
		keep[r] = (uint8_t)k;
		if (!k) continue;
		kept++;
		kgo_squeeze_row(row, map_word, map_bit, n, w_mem, mem);
		for (size_t p = 0; p < n_pheno; p++)
			scores[p * n_rows + r] =
			    kgo_score_row(mem, w_mem, yperm + p * n_pad, sums[p], n, (double)cnt, min_count);
	}
	free(mask); free(mem); free(yperm); free(sums);
	return kept;
}

/* src/kmers_multiple_databases.cpp:418-438  update_emma_kinshhip_calculation, applied to the rows
 * load_kmers keeps (:121). */
uint64_t kgo_kinship(const uint64_t *table, uint64_t n_rows, size_t w_file,
                     const uint32_t *map_word, const uint32_t *map_bit, size_t n,
                     uint64_t min_count, uint64_t *K, uint64_t *count) {
	size_t w_mem = 2 * ((n + 127) / 128);
	uint64_t *mask = (uint64_t *)malloc(w_file * sizeof(uint64_t));
	uint64_t *mem = (uint64_t *)malloc(w_mem * sizeof(uint64_t));
	uint8_t *g = (uint8_t *)malloc(n);
	build_mask(map_word, map_bit, n, w_file, mask);
	uint64_t kept = 0;
	size_t stride = w_file + 1;
	for (uint64_t r = 0; r < n_rows; r++) {
		const uint64_t *row = table + r * stride;
		uint64_t cnt = kgo_masked_popcount(row, mask, w_file);
		if (!((cnt >= min_count) && (cnt <= (uint64_t)n - min_count)) || (uint64_t)n < min_count) continue;
		kept++;
		kgo_squeeze_row(row, map_word, map_bit, n, w_mem, mem);
		for (size_t i = 0; i < n; i++) g[i] = (uint8_t)((mem[i >> 6] >> (i & 63)) & 1ull);
		for (size_t i = 0; i < n; i++) {
			uint64_t *Ki = K + i * n;
			uint8_t gi = g[i];
			for (size_t j = 0; j < i; j++) Ki[j] += (uint64_t)(1u ^ gi ^ g[j]);
		}
	}
	*count += kept;
	free(mask); free(mem); free(g);
	return kept;
}

/* ---------------------------------------------------------------------------------------
 * BestAssociationsHeap over std::priority_queue<tuple<u64,double,size_t>, vector, cmp_second>
 * (src/kmer_general.h:113-128).  cmp_second(l, r) = l.score > r.score, so the queue is a
 * min-heap on score.  libstdc++ (bits/stl_heap.h, GCC 13) algorithms restated:
 *   push:  c.push_back(v); __push_heap(first, hole=len-1, top=0, v)
 *   pop :  v = c.back(); c.back() = c.front(); __adjust_heap(first, 0, len-1, v); c.pop_back()
 * ------------------------------------------------------------------------------------- */
typedef struct { uint64_t kmer; double score; uint64_t row; } kgo_entry;
struct kgo_heap {
	uint64_t max_results, size, cap;
	uint64_t cnt_kmers, cnt_pops, cnt_push;
	double lowest_score;
	kgo_entry *c;
};

static int cmp_second(const kgo_entry *l, const kgo_entry *r) { return l->score > r->score; }

static void push_heap_(kgo_entry *first, int64_t hole, int64_t top, kgo_entry v) {
	int64_t parent = (hole - 1) / 2;
	while (hole > top && cmp_second(&first[parent], &v)) {
		first[hole] = first[parent];
		hole = parent;
		parent = (hole - 1) / 2;
	}
	first[hole] = v;
}

static void adjust_heap_(kgo_entry *first, int64_t hole, int64_t len, kgo_entry v) {
	const int64_t top = hole;
	int64_t child = hole;
	while (child < (len - 1) / 2) {
		child = 2 * (child + 1);
		if (cmp_second(&first[child], &first[child - 1])) child--;
		first[hole] = first[child];
		hole = child;
	}
	if ((len & 1) == 0 && child == (len - 2) / 2) {
		child = 2 * (child + 1);
		first[hole] = first[child - 1];
		hole = child - 1;
	}
	push_heap_(first, hole, top, v);
}

static void q_push(kgo_heap *h, kgo_entry v) {
	if (h->size == h->cap) {
		h->cap = h->cap ? h->cap * 2 : 1024;
		h->c = (kgo_entry *)realloc(h->c, h->cap * sizeof(kgo_entry));
	}
	h->c[h->size++] = v;
	push_heap_(h->c, (int64_t)h->size - 1, 0, v);
}

static void q_pop(kgo_entry *c, uint64_t *size) {
	if (*size > 1) {
		kgo_entry v = c[*size - 1];
		c[*size - 1] = c[0];
		adjust_heap_(c, 0, (int64_t)*size - 1, v);
	}
	(*size)--;
}

kgo_heap *kgo_heap_new(uint64_t max_results) {
	kgo_heap *h = (kgo_heap *)calloc(1, sizeof(kgo_heap));
	h->max_results = max_results;
	return h;
}
void kgo_heap_free(kgo_heap *h) { if (h) { free(h->c); free(h); } }

/* src/best_associations_heap.cpp:43-59  add_association */
void kgo_heap_add(kgo_heap *h, uint64_t kmer, double score, uint64_t row) {
	kgo_entry v = {kmer, score, row};
	h->cnt_kmers++;
	if (h->size < h->max_results) {
		q_push(h, v);
		h->cnt_push++;
		h->lowest_score = h->c[0].score;
	} else if (score > h->lowest_score) {
		h->cnt_pops++;
		h->cnt_push++;
		q_pop(h->c, &h->size);
		q_push(h, v);
		h->lowest_score = h->c[0].score;
	}
}
uint64_t kgo_heap_size(const kgo_heap *h) { return h->size; }
uint64_t kgo_heap_insertions(const kgo_heap *h) { return h->cnt_kmers; }

/* src/best_associations_heap.cpp:82-92 / :110-127: copy the queue and pop until empty. */
void kgo_heap_dump(const kgo_heap *h, uint64_t *kmers, double *scores, uint64_t *rows) {
	uint64_t n = h->size;
	kgo_entry *c = (kgo_entry *)malloc((n ? n : 1) * sizeof(kgo_entry));
	memcpy(c, h->c, n * sizeof(kgo_entry));
	uint64_t i = 0;
	while (n > 0) {
		kmers[i] = c[0].kmer; scores[i] = c[0].score; rows[i] = c[0].row; i++;
		q_pop(c, &n);
	}
	free(c);
}

/* ---------------------------------------------------------------------------------------
 * Synthetic table rows (OURS).  Counter-based: every word is a pure function of
 * (seed, row, word index), so host and device produce identical tables of any size.
 *   base(r)   = mix64(seed ^ mix64(r))
 *   dup(r)    = r > 0 and (mix64(base(r) ^ 0xD00D) & 63) == 0    -> pattern copied from row r-1's
 *               own generator (no chaining): about 1.6 % exact duplicate patterns (score ties)
 *   level(s)  = (mix64(base(s) ^ 0x1E7E1) >> 7) & 15 -> bit density via AND/OR of 5 random words:
 *               1/32, 1/8, 1/4, 3/8, 1/2, 1/2, 1/2, 5/8, 3/4, 7/8, 31/32, 1/4, 3/8, 1/2, 5/8, 3/4
 *               (levels 0 and 10 fail a 5 % MAF filter and exercise the compaction)
 *   kmer(r)   = r * 1024 + (mix64(base(r) ^ 0x4B3A) & 1023)       strictly increasing, < 2^62
 * ------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t x) {
	x += 0x9e3779b97f4a7c15ull;
	x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
	x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
	return x ^ (x >> 31);
}

void kgo_synth_rows(uint64_t seed, uint64_t first_row, uint64_t n_rows, uint64_t n_file, uint64_t *out) {
	size_t w_file = (size_t)((n_file + 63) / 64);
	uint64_t last_mask = (n_file % 64) ? ((1ull << (n_file % 64)) - 1ull) : ~0ull;
	for (uint64_t i = 0; i < n_rows; i++) {
		uint64_t r = first_row + i;
		uint64_t base_r = mix64(seed ^ mix64(r));
		uint64_t *row = out + i * (w_file + 1);
		row[0] = r * 1024ull + (mix64(base_r ^ 0x4B3Aull) & 1023ull);
		uint64_t s = r;
		if (r > 0 && (mix64(base_r ^ 0xD00Dull) & 63ull) == 0) s = r - 1;
		uint64_t b = mix64(seed ^ mix64(s));
		unsigned level = (unsigned)((mix64(b ^ 0x1E7E1ull) >> 7) & 15ull);
		for (size_t w = 0; w < w_file; w++) {
			uint64_t q[5];
			for (int k = 0; k < 5; k++) q[k] = mix64(b + 8ull * (uint64_t)w + (uint64_t)k + 1ull);
			uint64_t v;
			switch (level) {
			case 0: v = q[0] & q[1] & q[2] & q[3] & q[4]; break;
			case 1: v = q[0] & q[1] & q[2]; break;
			case 2: case 11: v = q[0] & q[1]; break;
			case 3: case 12: v = q[0] & (q[1] | q[2]); break;
			case 7: case 14: v = q[0] | (q[1] & q[2]); break;
			case 8: case 15: v = q[0] | q[1]; break;
			case 9: v = q[0] | q[1] | q[2]; break;
			case 10: v = q[0] | q[1] | q[2] | q[3] | q[4]; break;
			default: v = q[0]; break; /* 4, 5, 6, 13 */
			}
			if (w == w_file - 1) v &= last_mask;
			row[1 + w] = v;
		}
	}
}

/* ---------------------------------------------------------------------------------------
 * SNP twin (src/snps_multiple_databases.cpp): bit planes of one .bed row over the used samples
 * (ctor :95-135), the three SSE4-order dot products (dot_product_SSE4 :41-61) and
 * calculate_grammmar_approx_association (:143-158).  y_perm: permuted, zero-padded phenotypes.
 * ------------------------------------------------------------------------------------- */
static float dot_lane_order(const uint64_t *row, size_t w_mem, const float *y_perm) {
	float lane[4] = {0.f, 0.f, 0.f, 0.f};
	for (size_t b = 0; b < w_mem / 2; b++) {
		uint32_t m[4];
		memcpy(m, row + 2 * b, 16);
		for (int t = 0; t < 32; t++)
			for (int L = 0; L < 4; L++) {
				float add = (m[L] & (0x80000000u >> t)) ? y_perm[128 * b + 4 * t + L] : 0.0f;
				lane[L] = lane[L] + add;
			}
	}
	return ((lane[0] + lane[1]) + lane[2]) + lane[3];
}

void kgo_snp_scores(const uint8_t *bed, uint64_t n_snps, size_t bytes_per_snp, const uint32_t *map_byte,
                    const uint32_t *map_shift, size_t n, const float *y, double mac, double *scores) {
	size_t w_mem = 2 * ((n + 127) / 128), n_pad = 64 * w_mem;
	float *y_perm = (float *)malloc(n_pad * sizeof(float));
	kgo_update_scores_and_sum(y, n, n_pad, y_perm);
	uint64_t *pa = (uint64_t *)malloc(w_mem * 8), *nm = (uint64_t *)malloc(w_mem * 8), *het = (uint64_t *)malloc(w_mem * 8);
	for (uint64_t i = 0; i < n_snps; i++) {
		const uint8_t *row = bed + i * bytes_per_snp;
		memset(pa, 0, w_mem * 8); memset(nm, 0, w_mem * 8); memset(het, 0, w_mem * 8);
		double S = 0, S2 = 0, N = 0;
		static const double to_cnt[4] = {0, 0, 0.5, 1};
		for (size_t si = 0; si < n; si++) {
			unsigned d = (row[map_byte[si]] >> map_shift[si]) & 3u;
			S += to_cnt[d];
			S2 += to_cnt[d] * to_cnt[d];
			N += (d != 1u) ? 1.0 : 0.0;
			pa[si >> 6] |= (uint64_t)(d == 3u) << (si & 63);
			nm[si >> 6] |= (uint64_t)(d != 1u) << (si & 63);
			het[si >> 6] |= (uint64_t)(d == 2u) << (si & 63);
		}
		if (mac > S || mac > N - S) { scores[i] = 0; continue; }
		double yigi = (double)dot_lane_order(pa, w_mem, y_perm) + (double)dot_lane_order(het, w_mem, y_perm) * 0.5;
		double ss = (double)dot_lane_order(nm, w_mem, y_perm);
		double r = N * yigi - S * ss;
		r = r * r;
		scores[i] = r / (N * (N * S2 - S * S));
	}
	free(y_perm); free(pa); free(nm); free(het);
}
