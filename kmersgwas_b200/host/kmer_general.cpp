// kmer_general.cpp -- see kmer_general.h.  Own implementation; behaviour (formats, errors) follows the
// reference functions cited at each definition.
#include "kmer_general.h"

#include <sys/time.h>

#include <stdexcept>

using std::string;
using std::vector;

// ---- small helpers -------------------------------------------------------------------------------
// Split like repeated std::getline(stream, cell, '\t'): interior empty fields are kept, a trailing
// delimiter does not produce a final empty field, an empty line produces no fields.
static vector<string> split_tabs(const string &line) {
	vector<string> out;
	size_t pos = 0;
	while (pos < line.size()) {
		size_t tab = line.find('\t', pos);
		if (tab == string::npos) tab = line.size();
		out.emplace_back(line, pos, tab - pos);
		pos = tab + 1;
	}
	return out;
}

// ---- BedBimFilesHandle (kmer_general.h:134-145, kmer_general.cpp:146-151) -------------------------
BedBimFilesHandle::BedBimFilesHandle(const string &base_name)
    : f_bed(base_name + ".bed", std::ios::binary), f_bim(base_name + ".bim", std::ios::out) {
	static const char magic[3] = {0x6C, 0x1B, 0x01};
	f_bed.write(magic, 3);
}

void BedBimFilesHandle::close() {
	if (f_bed.is_open()) f_bed.close();
	if (f_bim.is_open()) f_bim.close();
}

// ---- <base>.names: whitespace separated accession names (kmer_general.cpp:45-53) ------------------
vector<string> load_kmers_talbe_column_names(const string &kmers_table_base) {
	std::ifstream in(kmers_table_base + ".names");
	vector<string> names;
	for (string tok; in >> tok;) names.push_back(tok);
	return names;
}

// ---- phenotype TSV (kmer_general.cpp:175-205) ------------------------------------------------------
// line 1: <anything>\t<name_1>...<name_P>; then <accession>\t<v_1>...; values parsed with stof.
std::pair<vector<string>, vector<PhenotypeList> > load_phenotypes_file(const string &filename) {
	std::ifstream in(filename);
	vector<string> pheno_names;
	vector<PhenotypeList> lists;
	string line;
	bool header = true;
	while (std::getline(in, line)) {
		const vector<string> cells = split_tabs(line);
		if (header) {
			pheno_names.assign(cells.size() > 1 ? cells.begin() + 1 : cells.end(), cells.end());
			lists.resize(pheno_names.size());
			header = false;
			continue;
		}
		if (cells.size() != pheno_names.size() + 1)
			throw std::logic_error("File should have the same number of fields in each row | " + filename);
		for (size_t p = 0; p < pheno_names.size(); p++) {
			lists[p].first.push_back(cells[0]);
			lists[p].second.push_back(std::stof(cells[p + 1]));
		}
	}
	return std::make_pair(pheno_names, lists);
}

// ---- name lookup; duplicates are an error (kmer_general.cpp:227-237) -------------------------------
size_t get_index_DB(const string &name, const vector<string> &names) {
	const size_t none = (~0u);
	size_t found = none;
	for (size_t j = 0; j < names.size(); j++) {
		if (names[j] != name) continue;
		if (found != none) throw std::logic_error("Two DBs with the same name! " + name);
		found = j;
	}
	return found;
}

// ---- keep the accessions present in the table, in phenotype-file order (kmer_general.cpp:239-253) --
PhenotypeList intersect_phenotypes_to_present_DBs(const PhenotypeList &pl, const string &kmers_table_base,
                                                  const bool &must_be_present) {
	const vector<string> table_names = load_kmers_talbe_column_names(kmers_table_base);
	PhenotypeList out;
	for (size_t i = 0; i < pl.first.size(); i++) {
		if (get_index_DB(pl.first[i], table_names) == (size_t)(~0u)) {
			if (must_be_present) throw std::logic_error("Couldn't find path for DB: " + pl.first[i]);
			continue;
		}
		out.first.push_back(pl.first[i]);
		out.second.push_back(pl.second[i]);
	}
	return out;
}

// ---- PLINK .fam (kmer_general.cpp:207-225): "<acc> <acc> 0 0 0 <v...>", default float formatting ----
void write_fam_file(const vector<PhenotypeList> &phenotypes, const string &fn) {
	std::ofstream f(fn, std::ios::out);
	const vector<string> &acc = phenotypes[0].first;
	for (size_t i = 0; i < acc.size(); i++) {
		f << acc[i] << " " << acc[i] << " 0 0 0";
		for (size_t j = 0; j < phenotypes.size(); j++) {
			if (phenotypes[j].first[i] != acc[i])
				throw std::logic_error("phenotypes should have the same order " + phenotypes[j].first[i] + "!=" + acc[i]);
			f << " " << phenotypes[j].second[i];
		}
		f << std::endl;
	}
}

void write_fam_file(const PhenotypeList &phenotype, const string &fn) {
	write_fam_file(vector<PhenotypeList>(1, phenotype), fn);
}

// ---- 2-bit decode, most significant base first (kmer_general.cpp:77-87) ----------------------------
string bits2kmer31(uint64_t w, const size_t &k) {
	static const char bases[4] = {'A', 'C', 'G', 'T'};
	string s(k, 'X');
	for (size_t i = k; i-- > 0;) {
		s[i] = bases[w & 3u];
		w >>= 2;
	}
	return s;
}

double get_time(void) {
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return (double)tv.tv_sec + (double)tv.tv_usec / 1e6;
}

bool is_file_exist(const string &file_name) {
	std::ifstream f(file_name);
	return f.good();
}
