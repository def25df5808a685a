/* kmersgwas_b200.h -- C ABI of the B200 (sm_100a) association hot path of kmersGWAS.
 *
 * This is the drop-in boundary.  The reference has no plugin/FFI layer: its
 * boundary is the public C++ surface of MultipleKmersDataBases / BestAssociationsHeap called from
 * the two CLIs.  The host-side C++ mirror of that surface lives in kmersgwas_b200/host/ (same class
 * and method names) and calls ONLY the functions below; a reference maintainer would bind exactly
 * these from the reference classes (INTEGRATION.md shows the patch).
 *
 * Conventions
 *   - plain C: pointers + sizes, no STL / torch types, no exceptions cross this boundary;
 *   - every function returns kg_status (0 = ok); kg_last_error(ctx) gives the message
 *     (kg_last_error(NULL) for failures of kg_ctx_create);
 *   - one kg_ctx per GPU, driven by one host thread at a time; different contexts are fully
 *     concurrent; device work is issued on the context's stream and is asynchronous until a
 *     *_fetch / kg_sync call;
 *   - "rows" arguments are RAW .table rows exactly as on disk
 *     (/root/reference/src/kmers_merge_multiple_databaes.cpp:60-73): per row one u64 k-mer followed
 *     by W_file = ceil(N_file/64) u64 presence words, no padding, little endian.  The pointer may
 *     be host memory (pageable or pinned; copied through the context's pinned staging ring) or
 *     device memory of the context's GPU (used in place);
 *   - there is NO CPU fallback: without a CUDA device kg_ctx_create fails with KG_ERR_CUDA.
 */
#ifndef KMERSGWAS_B200_H
#define KMERSGWAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KG_ABI_VERSION 1

typedef int kg_status;
enum {
	KG_OK = 0,
	KG_ERR_INVALID = 1,       /* bad argument / shape */
	KG_ERR_CUDA = 2,          /* CUDA runtime error (message has the cudaError string) */
	KG_ERR_NOMEM = 3,         /* host or device allocation failed */
	KG_ERR_HITS_OVERFLOW = 4, /* more hits than the hit buffer holds: resubmit the tile in smaller pieces */
	KG_ERR_STATE = 5          /* call sequence error (e.g. fetch without submit) */
};

typedef struct kg_ctx kg_ctx;

/* Table / column geometry.  Replaces the members MultipleKmersDataBases derives in its ctor and in
 * create_map_from_all_DBs (/root/reference/src/kmers_multiple_databases.cpp:39-94, 297-311). */
typedef struct kg_shape {
	uint64_t n_file;          /* accessions (columns) in the .table file          (m_accessions_db_file) */
	uint64_t n_used;          /* accessions used = memory columns, phenotype order (m_accessions)        */
	const uint32_t *map_word; /* [n_used] file word index of memory column i       (m_map_word_index)    */
	const uint32_t *map_bit;  /* [n_used] bit index inside that word               (m_map_bit_index)     */
} kg_shape;

/* One admitted association: what add_kmers_to_heap hands to BestAssociationsHeap::add_association
 * (/root/reference/src/kmers_multiple_databases.cpp:281-283). */
typedef struct kg_hit {
	uint64_t row;    /* caller's row id: first_row_id + index of the row inside the submitted tile */
	uint64_t kmer;   /* k-mer word of the row */
	double score;    /* bit-identical to calculate_kmer_score (:327-363) */
	uint32_t pheno;  /* phenotype column index */
	uint32_t pad_;
} kg_hit;

/* Options for kg_set_option */
enum {
	KG_OPT_SCAN_ENGINE = 1, /* 0 = auto, 1 = exact fp32-order kernel on every row, 2 = int8 tensor filter + exact refine */
	KG_OPT_HIT_CAPACITY = 2, /* number of kg_hit a device hit interval holds (default 1<<22); may be changed whenever no
	                            interval is open or waiting for kg_scan_fetch */
	KG_OPT_KINSHIP_ENGINE = 3, /* 0 = auto, 1 = popcount kernel, 2 = int8 tensor-core Gram */
	KG_OPT_KERNEL_TIMING = 4,  /* 1 = bracket every hot-path kernel launch with CUDA events on the context's stream
	                              (read back with kg_kernel_time); 0 = off (default) */
	KG_OPT_FILTER_PAIR_LIMIT = 5, /* tensor filter: column groups whose list of surviving rows has at most this many
	                              entries are re-tested per phenotype column and re-scored as single (row, phenotype)
	                              pairs; longer lists are re-scored 16 phenotypes at a time.  0 = never use pair
	                              mode, -1 = default (1/16 of the tile capacity).  Results are identical either way. */
	/* device-side selection (kg_select_*): set before kg_select_begin */
	KG_OPT_SELECT_GROWTH_PERMILLE = 6, /* round length = growth x rows scanned so far, in 1/1000 (default 500); the expected
	                              number of candidates per phenotype and round is growth x K */
	KG_OPT_SELECT_MAX_ROUND = 7,  /* longest round in rows (default 1 << 24) */
	KG_OPT_SELECT_CAND_CAP = 8,   /* candidates a phenotype's segment holds per round (default max(4 K, 65536)) */
	KG_OPT_SELECT_LOG_CAP = 9     /* admission-log entries per phenotype (default 32 K + 65536); KG_SELECT_LOG only */
};

/* Kernel classes reported by kg_kernel_time */
enum {
	KG_KERNEL_SCAN_EXACT = 0,  /* exact fp32-order score kernel over every (row, phenotype) */
	KG_KERNEL_SCAN_FILTER = 1, /* int8 tensor-core bound kernel (lists the rows it cannot rule out) */
	KG_KERNEL_SCAN_REFINE = 2, /* exact re-score of the listed rows (reported "rows" = rows re-scored) */
	KG_KERNEL_KINSHIP = 3,     /* Gram accumulation (popcount or tensor-core engine) */
	KG_KERNEL_AUX = 4,         /* squeeze / MAC prefilter / finalize / synthetic generator */
	KG_KERNEL_SCAN_SELECT = 5, /* device heaps: round bookkeeping + candidate sort + heap replay + filter re-tuning */
	KG_KERNEL_CLASSES = 6
};

/* ---- lifecycle ------------------------------------------------------------------------------ */
int kg_abi_version(void);
/* device: CUDA ordinal.  stream: a cudaStream_t (as void*) to issue all work on, or NULL to let the
 * context create its own non-blocking stream. */
kg_status kg_ctx_create(int device, const kg_shape *shape, void *stream, kg_ctx **out);
void kg_ctx_destroy(kg_ctx *ctx);
const char *kg_last_error(const kg_ctx *ctx);
kg_status kg_set_option(kg_ctx *ctx, int option, int64_t value);
kg_status kg_sync(kg_ctx *ctx);

/* ---- association scan ------------------------------------------------------------------------
 * Replaces, per batch, the P calls of MultipleKmersDataBases::add_kmers_to_heap
 * (/root/reference/src/associate_kmers.cpp:134-141; kmers_multiple_databases.cpp:275-295, 327-363)
 * and the MAC filter of load_kmers (:117-121). */

/* y: [n_pheno][n_used] float32, UN-permuted, memory (phenotype-file) order -- the vectors the
 * reference passes to add_kmers_to_heap.  min_count: the effective minor allele count
 * (associate_kmers.cpp:99-102).  May be called again to change phenotypes. */
kg_status kg_scan_set_phenotypes(kg_ctx *ctx, const float *y, uint32_t n_pheno, uint64_t min_count);

/* Stream-ordered and asynchronous: tiles already submitted keep the thresholds they were submitted with.
 * thresholds[p]: only (row, p) with score > thresholds[p] are reported; a negative threshold reports
 * every row passing the MAC filter (heap not yet full).  The caller passes the current
 * BestAssociationsHeap::lowest_score of each phenotype's heap
 * (/root/reference/src/best_associations_heap.cpp:43-59: strict '>'); because that value never
 * decreases, stale (lower) thresholds are always safe. */
kg_status kg_scan_set_thresholds(kg_ctx *ctx, const double *thresholds, uint32_t n_pheno);

/* Score one tile of raw rows against all phenotypes; asynchronous.  Hits are appended to the
 * context's hit buffer.  first_row_id: id given to the tile's first row (kg_hit.row). */
kg_status kg_scan_submit(kg_ctx *ctx, const uint64_t *rows, uint64_t n_rows, uint64_t first_row_id);

/* Hits are produced in INTERVALS.  Every kg_scan_submit appends to the open interval; kg_scan_mark closes it
 * without waiting for the device and opens the next one, so that the device scans interval i+1 while the host
 * fetches and replays interval i.  At most one closed interval may be waiting for kg_scan_fetch. */
kg_status kg_scan_mark(kg_ctx *ctx);

/* Wait for the oldest closed interval (the open one is closed first when none is waiting).
 * *n_hits = its number of hits.  If out != NULL and cap >= *n_hits, the hits are copied to out -- in NO particular
 * order; a buffer from kg_host_alloc makes this a single DMA -- and the interval is consumed.  Otherwise it stays
 * pending: call again with a large enough buffer, or drop it with kg_scan_clear_hits.
 * rows_seen / rows_kept (either may be NULL): totals since kg_scan_set_phenotypes over every fetched interval
 * including this one -- rows_kept is the reference's number_of_insertion() (.tested_kmers,
 * associate_kmers.cpp:203-205).
 * KG_ERR_HITS_OVERFLOW: the interval produced more hits than the hit buffer holds; it is dropped, its rows are not
 * counted, and the caller must submit them again in smaller intervals. */
kg_status kg_scan_fetch(kg_ctx *ctx, kg_hit *out, size_t cap, size_t *n_hits,
                        uint64_t *rows_seen, uint64_t *rows_kept);
/* Drop the hits of a fetched-but-not-copied interval (its rows still count as seen). */
kg_status kg_scan_clear_hits(kg_ctx *ctx);
/* Wait for the device and drop every interval that has not been consumed (closed or open) WITHOUT counting
 * its rows: the state is as it was after the last consumed interval (recovery after KG_ERR_HITS_OVERFLOW). */
kg_status kg_scan_discard(kg_ctx *ctx);

/* The MAC filter of load_kmers alone (/root/reference/src/kmers_multiple_databases.cpp:117-121 with
 * calculate_unsqueezed_popcnt :149-154): keep[r] = 1 iff min_count <= popcount(row r & used columns) <= N_used - min_count.
 * No phenotypes needed.  rows: host or device; keep: HOST buffer of n_rows bytes; *kept (may be NULL) = number of ones.
 * Used by the kmers_table_to_bed CLI (table -> PLINK conversion of every kept row). */
kg_status kg_mac_filter(kg_ctx *ctx, const uint64_t *rows, uint64_t n_rows, uint64_t min_count, uint8_t *keep, uint64_t *kept);

/* Testing / --k_mers_scores aid: exact scores of EVERY row of one tile.
 * keep[r] = row passes the MAC filter; scores[p * n_rows + r] valid where keep[r].  Host outputs. */
kg_status kg_scan_scores_dense(kg_ctx *ctx, const uint64_t *rows, uint64_t n_rows,
                               uint8_t *keep, double *scores);

/* Geometry of scan engine 2 for the phenotypes set (all 0 when the engine is unavailable for the shape): passes over a
 * tile, accumulator columns per pass (UMMA N), contraction length (UMMA K total), raw row-block stages.  The tensor work
 * of one filter launch over R rows is 2 * k_pad * p_pad * R int8 operations (bench.py's roofline). */
kg_status kg_scan_filter_shape(kg_ctx *ctx, uint32_t *n_pass, uint32_t *p_pad, uint32_t *k_pad, uint32_t *raw_stages);

/* Testing aid for scan engine 2 (int8 tensor-core filter): the filter's exact integer sums
 * q[r * n_pheno + p] = sum over the set presence bits of row r of the int8-quantised, centred phenotype p,
 * and (optional, may be NULL) the quantised values yq[p * 64 * W_file + file_column].  Host outputs.
 * Fails with KG_ERR_INVALID when the engine is unavailable for the context's shape. */
kg_status kg_scan_filter_sums(kg_ctx *ctx, const uint64_t *rows, uint64_t n_rows, int32_t *q, int8_t *yq);

/* ---- device-side selection: BestAssociationsHeap on the GPU -------------------------------------
 * Replaces BestAssociationsHeap::add_association (/root/reference/src/best_associations_heap.cpp:43-59) for all
 * phenotypes: after kg_select_begin, kg_scan_submit feeds P device-resident heaps that perform exactly the
 * std::priority_queue<.., cmp_second> element moves of the reference (/root/reference/src/kmer_general.h:113-128;
 * libstdc++ push_heap / pop_heap), in row order, so set, ties and pop order (ranks) are the reference's.  Thresholds and
 * the tensor filter's constants stay on the device: the host is not in the scan loop (no kg_scan_set_thresholds /
 * kg_scan_mark / kg_scan_fetch in this mode; they return KG_ERR_STATE).  Tiles must be submitted in ascending row order. */
enum {
	KG_SELECT_LOG = 1   /* keep, per phenotype, every candidate the heap admitted (row order): the hit log of a row shard
	                       other than the first, replayed on the merging GPU with kg_select_replay */
};
/* k_best[p] = capacity of heap p (associate_kmers.cpp:92-96).  Requires kg_scan_set_phenotypes with the same n_pheno.
 * Fails with KG_ERR_INVALID when a heap does not fit shared memory (k_best > ~14000: use the host replay path). */
kg_status kg_select_begin(kg_ctx *ctx, const uint64_t *k_best, uint32_t n_pheno, uint32_t flags);
/* Leave selection mode and free its buffers. */
kg_status kg_select_end(kg_ctx *ctx);
/* Wait for the device.  *rows_applied = rows whose candidates are in the heaps, *rows_kept = those of them that passed
 * the MAC filter (number_of_insertion(), .tested_kmers).  KG_ERR_HITS_OVERFLOW: a round produced more candidates for one
 * phenotype than its segment holds (scores rising along the table); that round and every later one were NOT applied:
 * resubmit the rows from *rows_applied on (the library has already shortened its rounds). */
kg_status kg_select_sync(kg_ctx *ctx, uint64_t *rows_applied, uint64_t *rows_kept);
/* Heap state image, u64 words: [P][4] {size, cnt_push, cnt_pops, 0}, then [P][k_max][3] {kmer, score bits, row} in
 * libstdc++ layout order (position 0 = top; pushing a phenotype's entries in this order into an empty
 * std::priority_queue reproduces the layout).  state: host or device memory.  import also sets the row / kept totals. */
size_t kg_select_state_len(const kg_ctx *ctx);
kg_status kg_select_export(kg_ctx *ctx, uint64_t *state);
kg_status kg_select_import(kg_ctx *ctx, const uint64_t *state, uint64_t rows_applied, uint64_t rows_kept);
/* Order-sensitive 64-bit digest of all heaps (layout order, k-mer + score bits + row): equal digests <=> equal heaps. */
kg_status kg_select_digest(kg_ctx *ctx, uint64_t *digest);
/* Counters since kg_select_begin (any pointer may be NULL): rounds applied, candidates replayed, candidates the heaps
 * admitted (sum of cnt_push over the phenotypes), rebuilds of the tensor filter's column order.  Waits for the device. */
kg_status kg_select_stats(kg_ctx *ctx, uint64_t *rounds, uint64_t *candidates, uint64_t *admitted, uint64_t *reorders);
/* Current thresholds (lowest kept score, -1 while a heap is not full), host output [n_pheno]. */
kg_status kg_select_thresholds(kg_ctx *ctx, double *thr);
/* Admission log (KG_SELECT_LOG).  kg_select_log_reset drops what was logged so far and zeroes the kept-row total (call
 * after scanning the shared prefix); counts: host [n_pheno]; export packs the entries {row, kmer, score bits} of
 * phenotype p at entries[3 * offsets[p] ...] (offsets: host [n_pheno + 1], entries: host or device). */
kg_status kg_select_log_reset(kg_ctx *ctx);
kg_status kg_select_log_counts(kg_ctx *ctx, uint64_t *counts);
kg_status kg_select_log_export(kg_ctx *ctx, uint64_t *entries, const uint64_t *offsets);
/* Replay candidates that are already in ascending row order per phenotype (a shard's exported log) through the
 * heaps; rows / kept are added to the totals.  entries: host or device, packed as kg_select_log_export writes them. */
kg_status kg_select_replay(kg_ctx *ctx, const uint64_t *entries, const uint64_t *offsets, uint64_t rows, uint64_t kept);
/* Multi-GPU threshold exchange (optional; shrinks the shards' logs).  kg_select_export_scores writes, position by
 * position, the scores of this context's heap entries whose row id is >= min_row into scores[P][k_max] (DEVICE memory;
 * -1 elsewhere): a shard passes the first row id of its own block, so that the shards contribute disjoint row sets.
 * kg_select_set_floor takes n_heaps such arrays back to back (scores[g][p][k_max], e.g. an NCCL all-gather restricted
 * to this shard and the shards BEFORE it in row order): the k_best-th largest score of their union is a lower bound of
 * the sequential heap's threshold for every row this context still scans, and becomes a floor under its thresholds
 * (candidates at or below it are dropped before they reach the heap or the log).  Floors never decrease. */
kg_status kg_select_export_scores(kg_ctx *ctx, uint64_t min_row, double *scores);
kg_status kg_select_set_floor(kg_ctx *ctx, const double *scores, uint32_t n_heaps);
uint32_t kg_select_kmax(const kg_ctx *ctx);

/* ---- kinship ----------------------------------------------------------------------------------
 * Replaces MultipleKmersDataBases::update_emma_kinshhip_calculation
 * (/root/reference/src/kmers_multiple_databases.cpp:418-438) over the rows load_kmers keeps. */

/* Start (or restart) accumulation.  accum_dev: optional caller-owned DEVICE buffer of
 * kg_kinship_accum_len(ctx) u64 (e.g. a torch int64 tensor, so that the caller can all-reduce it
 * with NCCL between kg_kinship_submit and kg_kinship_fetch); NULL = context-owned.
 * The buffer is zeroed. Layout: [n_used*n_used] Gram counts G[i][j] = #kept rows with both bits
 * set (lower triangle j<=i; the diagonal G[i][i] is the column count c[i]), then [1] kept rows M.
 * Every entry is a plain sum over rows, so shards add: all-reduce(sum) the whole buffer (kg_kinship_allreduce).
 * A caller-owned buffer is current after kg_sync, kg_kinship_allreduce or kg_kinship_fetch (the tensor engine folds
 * its per-tile counts into it lazily). */
kg_status kg_kinship_begin(kg_ctx *ctx, uint64_t min_count, uint64_t *accum_dev);
size_t kg_kinship_accum_len(const kg_ctx *ctx);
kg_status kg_kinship_submit(kg_ctx *ctx, const uint64_t *rows, uint64_t n_rows);
/* Wait; convert the (possibly all-reduced) accumulator into the reference's matrix:
 * ibs[i*n_used + j] = M - c[i] - c[j] + 2 G[i][j] for j < i (other entries 0), *kept_rows = M. */
kg_status kg_kinship_fetch(kg_ctx *ctx, uint64_t *ibs, uint64_t *kept_rows);

/* ---- table construction -------------------------------------------------------------------------------
 * Replaces MultipleKmersDataBasesMerger::load_kmers (/root/reference/src/kmers_merge_multiple_databaes.cpp:85-121) for one
 * range of the sorted list of all k-mers: table row i = {all_kmers[i], presence words}, bit a set iff all_kmers[i] is
 * in accession a's sorted list.  packed: the accessions' k-mers of the range back to back, accession a =
 * packed[offsets[a] .. offsets[a + 1]) (k-mers that are not in all_kmers are ignored); everything host memory;
 * table: n_all x (1 + ceil(n_acc / 64)) u64 -- the bytes MultipleKmersDataBasesMerger::output_to_table (:60-73) writes. */
kg_status kg_table_build(int device, const uint64_t *all_kmers, uint64_t n_all, uint32_t n_acc, const uint64_t *packed,
                         const uint64_t *offsets, uint64_t *table);

/* ---- SNP twin of the scan -------------------------------------------------------------------------------
 * Replaces the scoring loop of MultipleSNPsDataBases::get_most_associated_snps
 * (/root/reference/src/snps_multiple_databases.cpp:225-236 with calculate_grammmar_approx_association :143-158 and the
 * bit-plane construction of the ctor :95-135) for all phenotypes at once.  bed: the PLINK .bed payload after its 3-byte
 * magic, n_snps rows of bytes_per_snp bytes (host); sample i of the phenotype order sits in byte map_byte[i] at bit
 * map_shift[i] (:205-221); y: [n_pheno][n_samples] un-permuted; scores: host [n_pheno][n_snps], bit-identical to the
 * reference (0 for SNPs that fail the minor-allele-count test).  No kg_ctx needed; error text via kg_last_error(NULL). */
kg_status kg_snps_scores(int device, const uint8_t *bed, uint64_t n_snps, uint32_t bytes_per_snp, const uint32_t *map_byte,
                         const uint32_t *map_shift, uint32_t n_samples, const float *y, uint32_t n_pheno, double mac, double *scores);

/* ---- distinct presence/absence patterns (--pattern_counter) -----------------------------------------
 * Replaces MultipleKmersDataBases::update_presence_absence_pattern_counter
 * (/root/reference/src/kmers_multiple_databases.cpp:367-380): every row that passes the MAC filter is hashed over its
 * memory-order words exactly as the reference does (Hash64 + boost-style combine) and the 64-bit hashes are kept in a
 * device hash set that grows as needed; the count equals the reference's KmersSet::size().
 * expected: number of distinct patterns to make room for up front (0 = grow on demand).  Shards: export the keys of one
 * context (kg_patterns_export with keys = NULL returns the count) and kg_patterns_insert them into another. */
kg_status kg_patterns_begin(kg_ctx *ctx, uint64_t expected);
kg_status kg_patterns_submit(kg_ctx *ctx, const uint64_t *rows, uint64_t n_rows, uint64_t min_count);
/* Count the patterns of every tile kg_scan_submit sees from now on as well (same device copy of the rows: the table
 * is not transferred twice).  max_rows: rows that will be submitted while attached (room is made for all of them up
 * front, so the scan stays asynchronous); 0 detaches. */
kg_status kg_patterns_attach(kg_ctx *ctx, uint64_t min_count, uint64_t max_rows);
kg_status kg_patterns_count(kg_ctx *ctx, uint64_t *distinct, uint64_t *rows_kept);
kg_status kg_patterns_export(kg_ctx *ctx, uint64_t *keys, uint64_t cap, uint64_t *n);
kg_status kg_patterns_insert(kg_ctx *ctx, const uint64_t *keys, uint64_t n);

/* ---- NCCL all-reduce of the kinship accumulator (the path's only exchange step) ------------------
 * The accumulator is a plain sum over rows (u64 Gram counts + kept rows), so row shards add exactly.  The library
 * resolves NCCL at run time (dlopen of libnccl.so.2); it does not link it.
 *   multi-process (one rank per GPU): rank 0 calls kg_comm_unique_id, ships the 128 bytes to the other ranks by any
 *     means, every rank calls kg_comm_init_rank; then kg_kinship_allreduce between kg_kinship_submit and
 *     kg_kinship_fetch (stream-ordered, asynchronous);
 *   single process, several contexts: kg_comm_init_all once, then kg_kinship_allreduce_all (one NCCL group call).
 * After the all-reduce every context's kg_kinship_fetch returns the matrix of the whole table. */
kg_status kg_comm_unique_id(void *id128);
kg_status kg_comm_init_rank(kg_ctx *ctx, const void *id128, int n_ranks, int rank);
kg_status kg_comm_init_all(kg_ctx *const *ctxs, int n);
kg_status kg_kinship_allreduce(kg_ctx *ctx);
kg_status kg_kinship_allreduce_all(kg_ctx *const *ctxs, int n);

/* ---- stream tickets ---------------------------------------------------------------------------
 * kg_stream_mark returns a ticket for "everything submitted to this context so far" (copies and kernels);
 * kg_stream_wait blocks until that point has passed.  For callers that recycle their host input buffers (the tile
 * reader of MultipleKmersDataBases keeps three and waits for the ticket of the batch that used a buffer last). */
kg_status kg_stream_mark(kg_ctx *ctx, uint64_t *ticket);
kg_status kg_stream_wait(kg_ctx *ctx, uint64_t ticket);

/* ---- pinned host memory ----------------------------------------------------------------------
 * Page-locked buffers for the host-side tile reader (so that it needs no CUDA headers); tiles handed
 * to *_submit from such buffers are copied with one asynchronous DMA instead of the staging ring.
 * Host rows passed to *_submit must stay valid and unchanged until the next *_fetch / kg_sync. */
kg_status kg_host_alloc(kg_ctx *ctx, size_t bytes, void **out);
void kg_host_free(kg_ctx *ctx, void *p);

/* ---- synthetic tiles (bench / tests) -----------------------------------------------------------
 * Fill DEVICE memory with n_rows raw rows of the counter-based generator documented in
 * oracle/oracle.c (kgo_synth_rows): bit-identical on host and device. */
kg_status kg_synth_rows_device(kg_ctx *ctx, uint64_t seed, uint64_t first_row, uint64_t n_rows,
                               uint64_t *rows_dev);

/* Number of kernels this library launched since the context was created (bench "gpu_launches"). */
uint64_t kg_launch_count(const kg_ctx *ctx);

/* With KG_OPT_KERNEL_TIMING = 1: device time (CUDA events recorded on the context's stream around each
 * launch), launch count and table rows processed, summed per kernel class since the last
 * kg_kernel_time_reset.  Waits for the stream.  Any output pointer may be NULL. */
kg_status kg_kernel_time(kg_ctx *ctx, int kernel_class, double *ms_total, uint64_t *launches, uint64_t *rows);
kg_status kg_kernel_time_reset(kg_ctx *ctx);

/* Measured int8 tensor-pipe peak of the context's GPU in TOP/s: back-to-back tcgen05.mma.kind::i8 (M 128, N 256, K 32)
 * on every SM, timed with CUDA events (the roofline denominator bench.py reports for the tcgen05 kernels). */
kg_status kg_probe_int8_peak(kg_ctx *ctx, double *tops);

/* Host placement for the pinned tile loader (the reference's reader, kmers_multiple_databases.cpp:103-146, has no
 * such notion: it reads into pageable memory on whatever core runs it).  Restricts the CALLING thread (and the threads
 * it creates afterwards) to the CPUs of the NUMA node the device's PCIe slot hangs off, and makes that node the
 * preferred one for the memory the thread allocates from now on -- call it before allocating the pinned row buffers
 * (cudaHostAlloc / kg-owned staging) and before starting reader threads.  With 4+ GPUs streaming 55 GB/s each, buffers
 * on the wrong socket are limited by the inter-socket link.  Returns the node, or -1 when nothing was changed (one
 * node, no sysfs entry, KMERSGWAS_NUMA_BIND=0, or the node's CPUs are outside the process's cpuset).  n_cpus (may be
 * NULL) receives the number of CPUs the thread may now run on. */
int kg_bind_host_to_device(int device, int *n_cpus);

#ifdef __cplusplus
}
#endif
#endif /* KMERSGWAS_B200_H */
