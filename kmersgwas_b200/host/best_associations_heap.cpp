// best_associations_heap.cpp -- see best_associations_heap.h.
// Behaviour follows /root/reference/src/best_associations_heap.cpp (cited per method).
#include "best_associations_heap.h"

#include <algorithm>
#include <cstring>
#include <iostream>

using std::get;

BestAssociationsHeap::BestAssociationsHeap(std::size_t max_results)
    : m_n_res(max_results), m_best_kmers(), cnt_kmers(0), cnt_pops(0), cnt_push(0), lowest_score(0) {}

// reference :43-59 -- push while not full; afterwards replace the minimum only on a STRICTLY larger score
void BestAssociationsHeap::add_association(const uint64_t &k, const double &score, const uint64_t &kmer_row) {
	++cnt_kmers;
	const bool full = !(m_best_kmers.size() < m_n_res);
	if (full) {
		if (!(score > lowest_score)) return;
		m_best_kmers.pop();
		++cnt_pops;
	}
	m_best_kmers.push(AssociationScoreHeap(k, score, kmer_row));
	++cnt_push;
	lowest_score = get<1>(m_best_kmers.top());
}

void BestAssociationsHeap::add_hits(const kg_hit *hits, std::size_t n) {
	for (std::size_t i = 0; i < n; i++) add_association(hits[i].kmer, hits[i].score, hits[i].row);
}

void BestAssociationsHeap::load_layout(const uint64_t *entries, std::size_t n, std::size_t tested, std::size_t pushes, std::size_t pops) {
	empty_heap();
	for (std::size_t i = 0; i < n; i++) {
		double score;
		static_assert(sizeof(double) == sizeof(uint64_t), "score bits");
		memcpy(&score, &entries[3 * i + 1], sizeof score);
		m_best_kmers.push(AssociationScoreHeap(entries[3 * i], score, (std::size_t)entries[3 * i + 2]));
	}
	cnt_kmers = tested;
	cnt_push = pushes;
	cnt_pops = pops;
	lowest_score = m_best_kmers.empty() ? 0.0 : get<1>(m_best_kmers.top());
}

// Pops a copy of the queue in ascending score order, handing each entry and the queue size
// before its pop (= the entry's rank, 1 = best) to fn.
template <class Fn>
static void drain_copy(const AssociationsPriorityQueue &q, Fn fn) {
	AssociationsPriorityQueue tmp(q);
	while (!tmp.empty()) {
		fn(tmp.top(), tmp.size());
		tmp.pop();
	}
}

// reference :67-76 -- k-mers only
void BestAssociationsHeap::output_to_file(const std::string &filename) const {
	std::ofstream of(filename, std::ios::binary);
	drain_copy(m_best_kmers, [&](const AssociationScoreHeap &e, std::size_t) {
		of.write(reinterpret_cast<const char *>(&get<0>(e)), sizeof(uint64_t));
	});
}

// reference :82-92 -- (u64 k-mer, f64 score) records, ascending score
void BestAssociationsHeap::output_to_file_with_scores(const std::string &filename) const {
	std::ofstream of(filename, std::ios::binary);
	drain_copy(m_best_kmers, [&](const AssociationScoreHeap &e, std::size_t) {
		of.write(reinterpret_cast<const char *>(&get<0>(e)), sizeof(uint64_t));
		of.write(reinterpret_cast<const char *>(&get<1>(e)), sizeof(double));
	});
}

std::vector<AssociationScoreHeap> BestAssociationsHeap::entries_in_pop_order() const {
	std::vector<AssociationScoreHeap> out;
	out.reserve(m_best_kmers.size());
	drain_copy(m_best_kmers, [&](const AssociationScoreHeap &e, std::size_t) { out.push_back(e); });
	return out;
}

// reference :97-108
KmersSet BestAssociationsHeap::get_KmersSet() const {
	KmersSet res;
	drain_copy(m_best_kmers, [&](const AssociationScoreHeap &e, std::size_t) { res.insert(get<0>(e)); });
	return res;
}

// reference :110-127 -- (k-mer, rank, row) sorted by row; rank = queue size at pop time.
// std::sort on the row key alone (rows are unique, so the order is fully determined).
kmers_output_list BestAssociationsHeap::get_kmers_for_output(const std::size_t &kmer_len) const {
	(void)kmer_len;
	kmers_output_list res;
	res.next_index = 0;
	drain_copy(m_best_kmers, [&](const AssociationScoreHeap &e, std::size_t rank) {
		res.list.push_back(std::make_tuple(get<0>(e), (uint64_t)rank, get<2>(e)));
	});
	std::sort(res.list.begin(), res.list.end(),
	          [](const AssociationOutputInfo &a, const AssociationOutputInfo &b) { return get<2>(a) < get<2>(b); });
	return res;
}

// reference :135-147
std::vector<std::size_t> BestAssociationsHeap::get_rows_sorted_indices() const {
	std::vector<std::size_t> rows;
	drain_copy(m_best_kmers, [&](const AssociationScoreHeap &e, std::size_t) { rows.push_back(get<2>(e)); });
	std::sort(rows.begin(), rows.end());
	return rows;
}

// reference :150-156
void BestAssociationsHeap::plot_stat() const {
	std::cerr << "[heap-stat] max\t" << m_n_res << "\tsize\t" << m_best_kmers.size() << "\tkmers\t" << cnt_kmers
	          << "\tpops\t" << cnt_pops << "\tpush\t" << cnt_push << std::endl;
}
