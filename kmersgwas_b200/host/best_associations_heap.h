// best_associations_heap.h -- bounded best-K store of (k-mer, score, row).
//
// Same public surface as /root/reference/src/best_associations_heap.h:30-52.  The container is
// std::priority_queue<AssociationScoreHeap, vector, cmp_second> exactly as in the reference
// (kmer_general.h:128), because which of several equal-score minima gets evicted -- and therefore
// the reported set and ranks under ties -- is decided by libstdc++'s heap layout.
// On the B200 path the GPU proposes candidates (kg_hit, score > threshold) and this class replays
// them in row order, which reproduces the reference heap state exactly; see add_hits().
#ifndef KGH_BEST_ASSOCIATIONS_H
#define KGH_BEST_ASSOCIATIONS_H

#include "kmer_general.h"
#include "kmersgwas_b200.h"

class BestAssociationsHeap {
	public:
		explicit BestAssociationsHeap(std::size_t max_results);
		void add_association(const uint64_t &k, const double &score, const uint64_t &kmer_row);

		void output_to_file(const std::string &filename) const;
		void output_to_file_with_scores(const std::string &filename) const;
		void plot_stat() const;
		inline void empty_heap() { AssociationsPriorityQueue().swap(m_best_kmers); }
		KmersSet get_KmersSet() const;
		kmers_output_list get_kmers_for_output(const std::size_t &kmer_len) const;
		std::vector<std::size_t> get_rows_sorted_indices() const;
		std::size_t number_of_insertion() const { return cnt_kmers; }

		// ---- additions for the GPU path -------------------------------------------------------
		// Replay device candidates of ONE phenotype (ascending row order) through add_association.
		// Rows the device filtered out (score <= threshold) could never have changed the heap, but the
		// reference counts every tested row: account for them with note_tested_rows().
		void add_hits(const kg_hit *hits, std::size_t n);
		void note_tested_rows(std::size_t n_rows_not_replayed) { cnt_kmers += n_rows_not_replayed; }
		// Threshold to give the device: lowest kept score once the heap is full, else -1 (report all).
		double device_threshold() const { return m_best_kmers.size() < m_n_res ? -1.0 : lowest_score; }
		std::size_t size() const { return m_best_kmers.size(); }
		// All entries in pop order (ascending score), as output_to_file_with_scores would write them.
		std::vector<AssociationScoreHeap> entries_in_pop_order() const;
		std::size_t capacity() const { return m_n_res; }
		// Adopt a heap kept on the device (kg_select_export): n entries {k-mer, score bits, row} in libstdc++ LAYOUT order
		// (position 0 = top).  Pushing a valid heap array front to back moves no element (every parent <= its child), so the
		// queue ends up with exactly the device's layout -- and therefore the reference's pop order under ties.
		void load_layout(const uint64_t *entries, std::size_t n, std::size_t tested, std::size_t pushes, std::size_t pops);
	private:
		std::size_t m_n_res;
		AssociationsPriorityQueue m_best_kmers;
		std::size_t cnt_kmers, cnt_pops, cnt_push;
		double lowest_score;
};

#endif
