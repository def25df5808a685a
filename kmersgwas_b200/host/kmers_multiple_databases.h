// kmers_multiple_databases.h -- presence/absence table reader + scorer, B200 build.
//
// Same class name and public methods as /root/reference/src/kmers_multiple_databases.h:32-72 for the
// hot path (ctor, load_kmers, add_kmers_to_heap, update_emma_kinshhip_calculation,
// output_plink_bed_file(handle, list, index), update_presence_absence_pattern_counter, clear,
// get_dbs_names).  What differs underneath:
//   * load_kmers reads RAW file rows into a pinned host buffer; the MAC filter, the column squeeze
//     and the popcount of the reference's load loop (:110-143) run on the GPU inside the scan /
//     kinship kernels.  Row ids handed to the heaps are FILE row indices (monotone, like the
//     reference's m_row_offset + index; SURVEY.md section 7, hard part 5).
//   * scoring calls the C ABI (include/kmersgwas_b200.h); add_kmers_to_heaps scores ALL phenotypes
//     of a batch in one device pass instead of one CTPL task per phenotype
//     (associate_kmers.cpp:134-141).
//   * dead code of the reference (gamma precalculation, textual dump) is not mirrored.
#ifndef KGH_KMER_MULTIPLEDB_H
#define KGH_KMER_MULTIPLEDB_H

#include <future>
#include <stdexcept>

#include "association_driver.h"
#include "best_associations_heap.h"
#include "kmer_general.h"
#include "kmersgwas_b200.h"

// Thrown by finish_device_selection when a round overflowed a candidate segment on the device (scores rising along
// the table): the caller re-runs the scan through the host replay path (add_kmers_to_heaps), which has no such limit.
struct DeviceSelectionOverflow : public std::runtime_error {
	explicit DeviceSelectionOverflow(const std::string &m) : std::runtime_error(m) {}
};

class MultipleKmersDataBases {
	public:
		MultipleKmersDataBases(const std::string &kmers_table_base, const std::vector<std::string> &db_to_use,
		                       const uint32_t &kmer_len);
		MultipleKmersDataBases() = delete;
		MultipleKmersDataBases(const MultipleKmersDataBases &) = delete;
		MultipleKmersDataBases &operator=(const MultipleKmersDataBases &) = delete;
		~MultipleKmersDataBases();

		// Load the next batch of up to batch_size file rows.  Returns false iff the file was already
		// exhausted when called (same contract as the reference :103-108,145).
		bool load_kmers(const uint64_t &batch_size, const std::size_t &minor_allele_count = 0);
		inline bool load_kmers() { return load_kmers(NULL_KEY); }

		// One phenotype (reference signature, :275-284).
		void add_kmers_to_heap(BestAssociationsHeap &kmers_and_scores, std::vector<float> scores,
		                       const std::size_t &min_cnt) const;
		// All phenotypes of the batch in one device pass; heaps[j] <-> scores[j].
		void add_kmers_to_heaps(std::vector<BestAssociationsHeap> &heaps, const std::vector<std::vector<float> > &scores,
		                        const std::size_t &min_cnt) const;

		// ---- device-resident heaps (kg_select_*): the streaming form of add_kmers_to_heaps used by the CLI.  The heaps
		// live on the GPU for the whole scan, batches are only submitted (asynchronously: the pinned tile reader keeps
		// reading ahead), and the BestAssociationsHeap objects are filled once at the end, layout and all.
		// Returns false (and changes nothing) when a capacity does not fit the device heaps: use add_kmers_to_heaps then.
		bool begin_device_selection(const std::vector<std::size_t> &capacities, const std::vector<std::vector<float> > &scores,
		                            const std::size_t &min_cnt, bool log_admissions);
		void add_loaded_kmers_to_device_heaps();
		void device_selection_log_reset();      // row shards > 0: forget the shared prefix's log entries and kept rows
		// the shard's admission log, packed: offsets[P + 1], entries[3 * offsets[P]] {row, k-mer, score bits}; + kept rows
		void device_selection_log(std::vector<uint64_t> &offsets, std::vector<uint64_t> &entries, uint64_t &rows, uint64_t &kept);
		void device_selection_replay(const std::vector<uint64_t> &offsets, const std::vector<uint64_t> &entries, uint64_t rows, uint64_t kept);
		void finish_device_selection(std::vector<BestAssociationsHeap> &heaps);

		// K[i][j] += IBS count over the loaded batch for j < i; count += kept rows (:418-438).
		void update_emma_kinshhip_calculation(std::vector<std::vector<uint64_t> > &K, uint64_t &count) const;
		// Streaming form used by the CLI: accumulate on the device over all batches, read back once.
		void kinship_begin(const std::size_t &min_count);
		void kinship_accumulate_loaded();
		void kinship_finish(std::vector<std::vector<uint64_t> > &K, uint64_t &count);

		// PLINK output of the selected rows that fall inside the loaded batch (:241-252).
		std::size_t output_plink_bed_file(BedBimFilesHandle &f, const std::vector<AssociationOutputInfo> &kmer_list,
		                                  std::size_t index) const;
		// Same output without re-streaming the table: reads only the selected rows from the file.
		void output_plink_bed_file_selected(BedBimFilesHandle &f, const std::vector<AssociationOutputInfo> &kmer_list) const;

		void update_presence_absence_pattern_counter(KmersSet &pa_pattern_counter) const;
		// Device form of the same counter (kg_patterns_*): the hashes of the kept rows go into a device hash set while the
		// scan's own copy of the rows is on the GPU (no second transfer); legal with row shards (sets are merged by key).
		void pattern_counter_begin(uint64_t max_rows, const std::size_t &min_count);   // counts every row submitted to the scan from now on
		uint64_t pattern_counter_size();
		void pattern_counter_export(std::vector<uint64_t> &keys);
		void pattern_counter_insert(const std::vector<uint64_t> &keys);

		// kmers_table_to_bed (reference :204-216, :262-272): PLINK output of EVERY row load_kmers keeps.  The reference's
		// load_kmers stops after batch_size KEPT rows; here a batch is a range of raw file rows, so the caller walks the
		// loaded rows with the keep flags of the device MAC filter and cuts its output files itself.
		void mac_filter_loaded(const std::size_t &min_count, std::vector<uint8_t> &keep) const;
		void output_plink_loaded_row(BedBimFilesHandle &f, std::size_t row_in_batch) const;      // name = the k-mer (:208)
		uint64_t presence_absence_pattern_hash_loaded_row(std::size_t row_in_batch) const;       // (:367-374)
		inline const std::vector<std::string> get_dbs_names() { return m_db_names_table; }
		void clear() { m_rows_loaded = 0; }

		// B200 additions
		static void set_device(int device) { s_device = device; }
		uint64_t rows_in_file() const { return m_file_rows ? m_file_rows : m_kmer_number; }
		uint64_t rows_loaded() const { return m_rows_loaded; }
		std::size_t file_words() const { return m_hash_words_db_file; }
		uint64_t row_offset() const { return m_row_offset; }
		const uint64_t *loaded_rows() const { return m_batch; }
		kg_ctx *context() const { return m_ctx; }
		int device() const { return m_device; }
		void set_scan_engine(int engine) const;
		void set_kinship_engine(int engine) const;
		// Restrict this object to file rows [first, first + count): one shard of a multi-GPU run.
		void restrict_to_rows(uint64_t first, uint64_t count);

	private:
		std::vector<std::string> m_db_names_db_file;  // accessions in the table file
		std::vector<std::string> m_db_names_table;    // accessions used (memory / phenotype order)
		std::size_t m_accessions_db_file, m_accessions;
		std::size_t m_hash_words_db_file, m_hash_words;
		uint32_t m_kmer_len;
		std::string m_table_path;
		int m_fd;
		uint64_t m_file_rows = 0; // rows in the whole file (set by the first restrict_to_rows)
		uint64_t m_kmer_number;   // rows in (this shard of) the file
		uint64_t m_first_row;     // first file row of this shard
		uint64_t m_kmer_loaded;   // rows read so far (incl. current batch)
		uint64_t m_row_offset;    // file row index of the first row of the current batch
		uint64_t *m_batch;        // pinned host buffer with the raw rows of the current batch (= m_buf[m_cur])
		// tile reader: three pinned buffers; while the device works on the current batch (and may still copy the previous
		// one), a reader thread preads the next batch with several threads (load_kmers then only swaps buffers)
		static const int kBuffers = 3;
		uint64_t *m_buf[kBuffers];
		std::size_t m_buf_cap[kBuffers];      // capacity in rows
		uint64_t m_buf_ticket[kBuffers];      // kg_stream_mark of the last submit that read the buffer
		bool m_buf_busy[kBuffers];
		int m_cur;
		std::future<uint64_t> m_prefetch;     // rows read into m_buf[(m_cur + 1) % kBuffers]
		bool m_prefetch_pending;
		uint64_t m_prefetch_batch, m_prefetch_first;
		uint64_t read_batch_into(int b, uint64_t first_row, uint64_t n_rows);
		void start_prefetch(uint64_t batch_size);
		void cancel_prefetch();
		bool m_device_selection;
		uint64_t m_rows_submitted_sel;
		uint64_t m_rows_loaded;   // rows in the current batch
		std::size_t m_load_mac;   // MAC given to the last load_kmers
		std::vector<uint32_t> m_map_word_index, m_map_bit_index;
		std::vector<uint64_t> m_map_mask;
		mutable kg_ctx *m_ctx;
		int m_device = 0;
		mutable std::vector<float> m_pheno_flat;   // phenotypes currently resident on the device
		mutable std::size_t m_pheno_min_cnt;
		mutable AssociationDriverState m_driver;   // rows scored since the phenotypes were set, hit buffer
		mutable bool m_kinship_streaming;
		static int s_device;

		void create_map_from_all_DBs();
		void squeeze_row(const uint64_t *file_row, std::vector<uint64_t> &mem_row) const;
		void write_PA(const std::string &name, const std::vector<uint64_t> &mem_row, BedBimFilesHandle &f) const;
		void ensure_phenotypes(const std::vector<std::vector<float> > &scores, std::size_t min_cnt) const;
		void check(kg_status st, const char *what) const;
};

#endif
