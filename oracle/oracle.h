/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY (see oracle.c header). */
#ifndef KMERSGWAS_ORACLE_H
#define KMERSGWAS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kmer_general.cpp:155-167 */
void kgo_permute_scores(const float *v, size_t n_pad, float *out);
/* kmers_multiple_databases.cpp:288-295: pad to n_pad, permute, sequential fp32 sum. */
float kgo_update_scores_and_sum(const float *y, size_t n, size_t n_pad, float *y_perm_out);
/* kmers_multiple_databases.cpp:149-154 */
uint64_t kgo_masked_popcount(const uint64_t *file_row, const uint64_t *mask, size_t w_file);
/* kmers_multiple_databases.cpp:125-132 */
void kgo_squeeze_row(const uint64_t *file_row, const uint32_t *map_word, const uint32_t *map_bit,
                     size_t n, size_t w_mem, uint64_t *mem_row);
/* kmers_multiple_databases.cpp:327-363 */
double kgo_score_row(const uint64_t *mem_row, size_t w_mem, const float *y_perm, float sum,
                     size_t n, double n1, uint64_t min_in_group);

/* load_kmers + add_kmers_to_heap over a whole in-memory table (associate_kmers.cpp:123-148),
 * without the heap: writes, for every FILE row r, keep[r] (MAC filter, :121) and, for kept
 * rows, scores[p * n_rows + r].  Returns the number of kept rows. */
uint64_t kgo_scan_scores(const uint64_t *table, uint64_t n_rows, size_t w_file,
                         const uint32_t *map_word, const uint32_t *map_bit, size_t n,
                         const float *y, size_t n_pheno, uint64_t min_count,
                         uint8_t *keep, double *scores);

/* update_emma_kinshhip_calculation (:418-438) over a whole in-memory table with the MAC filter
 * of load_kmers.  K is n*n u64 (only j<i written). Returns kept rows added to *count. */
uint64_t kgo_kinship(const uint64_t *table, uint64_t n_rows, size_t w_file,
                     const uint32_t *map_word, const uint32_t *map_bit, size_t n,
                     uint64_t min_count, uint64_t *K, uint64_t *count);

/* BestAssociationsHeap (best_associations_heap.cpp:43-59, 82-92, 110-127) on top of a restatement
 * of libstdc++'s std::priority_queue<.., cmp_second> (push_heap/pop_heap, bits/stl_heap.h). */
typedef struct kgo_heap kgo_heap;
kgo_heap *kgo_heap_new(uint64_t max_results);
void kgo_heap_free(kgo_heap *h);
void kgo_heap_add(kgo_heap *h, uint64_t kmer, double score, uint64_t row);
uint64_t kgo_heap_size(const kgo_heap *h);
uint64_t kgo_heap_insertions(const kgo_heap *h);
/* pops a COPY in ascending-score order: kmers[i], scores[i], rows[i], i = 0..size-1 */
void kgo_heap_dump(const kgo_heap *h, uint64_t *kmers, double *scores, uint64_t *rows);

/* Deterministic synthetic table generator (OURS, not the reference's): the same function is
 * implemented on the device in kmersgwas_b200/csrc/kg_synth.cuh; tests check they agree. */
void kgo_synth_rows(uint64_t seed, uint64_t first_row, uint64_t n_rows, uint64_t n_file, uint64_t *out);

/* SNP twin: scores of every SNP of a .bed payload for ONE phenotype (src/snps_multiple_databases.cpp:95-158). */
void kgo_snp_scores(const uint8_t *bed, uint64_t n_snps, size_t bytes_per_snp, const uint32_t *map_byte,
                    const uint32_t *map_shift, size_t n, const float *y, double mac, double *scores);

#ifdef __cplusplus
}
#endif
#endif
