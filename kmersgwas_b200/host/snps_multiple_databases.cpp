// snps_multiple_databases.cpp -- see snps_multiple_databases.h.
#include "snps_multiple_databases.h"

#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "best_associations_heap.h"
#include "kmer_general.h"
#include "kmersgwas_b200.h"

using std::string;
using std::vector;

int MultipleSNPsDataBases::s_device = 0;

namespace {
// first space-separated field of every .fam line (reference :166-180)
vector<string> fam_sample_names(const string &fam_fn) {
	vector<string> names;
	std::ifstream fin(fam_fn);
	string line;
	while (std::getline(fin, line)) {
		std::stringstream ls(line);
		string cell;
		std::getline(ls, cell, ' ');
		names.push_back(cell);
	}
	return names;
}
}  // namespace

MultipleSNPsDataBases::MultipleSNPsDataBases(const string &base_name_bedbim, const vector<string> &samples_to_use)
    : m_base_name(base_name_bedbim), m_samples_names(samples_to_use), m_n_snps(0), m_n_bytes_per_snp(0) {
	const vector<string> all_samples = fam_sample_names(m_base_name + ".fam");
	// sample i of the phenotype order -> (byte, bit) of a .bed row (reference :205-221)
	for (const string &s : m_samples_names) {
		const size_t i_full = (size_t)(std::find(all_samples.begin(), all_samples.end(), s) - all_samples.begin());
		if (i_full == all_samples.size()) throw std::logic_error("All accessions should be in fam file: " + s);
		m_map_byte.push_back((uint32_t)(i_full / 4));
		m_map_shift.push_back((uint32_t)((i_full % 4) * 2));
	}
	std::ifstream bed(m_base_name + ".bed", std::ios::binary | std::ios::ate);
	const size_t bed_size = bed ? (size_t)bed.tellg() : 0;
	if (bed_size < 3) throw std::logic_error("Bed file is too small");
	m_n_bytes_per_snp = (4 + all_samples.size() - 1) / 4;
	m_n_snps = (bed_size - 3) / m_n_bytes_per_snp;
	if (bed_size != m_n_snps * m_n_bytes_per_snp + 3) throw std::logic_error("Ilegal size of bed file");
	std::cerr << base_name_bedbim << "\t(snps,samples) = " << m_n_snps << ", " << all_samples.size() << std::endl;
	bed.seekg(3, std::ios::beg);
	m_bed.resize(m_n_snps * m_n_bytes_per_snp);
	bed.read(reinterpret_cast<char *>(m_bed.data()), (std::streamsize)m_bed.size());
}

vector<vector<size_t> > MultipleSNPsDataBases::get_most_associated_snps(const vector<vector<float> > &phenotypes, const size_t &n_best,
                                                                        const double &mac) const {
	const size_t P = phenotypes.size(), N = m_samples_names.size();
	vector<float> flat;
	for (const auto &y : phenotypes) {
		if (y.size() != N) throw std::logic_error("phenotype vector length differs from the number of samples used");
		flat.insert(flat.end(), y.begin(), y.end());
	}
	vector<double> scores(P * m_n_snps);
	if (P && m_n_snps &&
	    kg_snps_scores(s_device, m_bed.data(), m_n_snps, (uint32_t)m_n_bytes_per_snp, m_map_byte.data(), m_map_shift.data(), (uint32_t)N,
	                   flat.data(), (uint32_t)P, mac, scores.data()) != KG_OK)
		throw std::runtime_error(string("kg_snps_scores: ") + kg_last_error(nullptr));
	// every SNP goes through the heap, zero scores included (reference :230-234): ties are resolved by its layout
	vector<vector<size_t> > res(P);
	for (size_t p = 0; p < P; p++) {
		BestAssociationsHeap best(n_best);
		const double *s = scores.data() + p * m_n_snps;
		for (size_t i = 0; i < m_n_snps; i++) best.add_association(0, s[i], i);
		res[p] = best.get_rows_sorted_indices();
	}
	return res;
}

vector<size_t> MultipleSNPsDataBases::get_most_associated_snps(vector<float> phenotypes, const size_t &n_best, const double &mac) const {
	return get_most_associated_snps(vector<vector<float> >(1, phenotypes), n_best, mac)[0];
}

void MultipleSNPsDataBases::output_plink_bed_file(const vector<string> &files_base_names, vector<vector<size_t> > SNPs_indices) const {
	vector<BedBimFilesHandle> out;
	for (const string &b : files_base_names) out.emplace_back(b);
	std::ifstream bim(m_base_name + ".bim");
	string bim_line;
	vector<size_t> next(SNPs_indices.size(), 0);
	for (size_t i = 0; i < m_n_snps; i++) {
		std::getline(bim, bim_line);
		for (size_t l = 0; l < SNPs_indices.size(); l++) {
			if (next[l] < SNPs_indices[l].size() && i == SNPs_indices[l][next[l]]) {
				out[l].f_bim << bim_line << std::endl;
				out[l].f_bed.write(reinterpret_cast<const char *>(m_bed.data() + i * m_n_bytes_per_snp), (std::streamsize)m_n_bytes_per_snp);
				next[l]++;
			}
		}
	}
	for (auto &h : out) h.close();
}
