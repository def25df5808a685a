// build_kmers_table -- build the k-mers presence/absence table from the accessions' sorted k-mer lists (B200 build).
//
// Same flags and output files (<o>.table, <o>.names) as the reference CLI (/root/reference/src/build_kmers_table.cpp:
// flags :32-38, loop :98-104; table format kmers_merge_multiple_databaes.cpp:54-73).  The reference walks the k-mer space
// in 5000 threshold steps, looking every accession's k-mers up in a hash map; here the list of all k-mers is cut into
// ranges of --range_kmers rows, each range's accession sub-lists are streamed from their files and matched on the GPU
// (kg_table_build).  The bytes written are the same.
#include <algorithm>
#include <cstring>
#include <exception>
#include <fstream>
#include <iostream>

#include "cli_options.h"
#include "kmer_general.h"
#include "kmersgwas_b200.h"

using namespace std;

namespace {
const uint64_t kKmerMask = 0x3FFFFFFFFFFFFFFFull;   // the two top bits of a stored k-mer are flags (kmers_single_database.cpp:146-147)

// sequential reader of a sorted k-mer file with a one-k-mer look-ahead (KmersSingleDataBaseSortedFile, :144-177)
struct SortedKmerFile {
	ifstream fin;
	vector<uint64_t> buf;
	size_t at = 0, have = 0;
	bool open(const string &fn) {
		fin.open(fn, ios::binary);
		buf.resize(1 << 16);
		return fin.is_open();
	}
	bool peek(uint64_t &k) {
		if (at == have) {
			fin.read(reinterpret_cast<char *>(buf.data()), (streamsize)(buf.size() * 8));
			have = (size_t)fin.gcount() / 8;
			at = 0;
			if (have == 0) return false;
		}
		k = buf[at] & kKmerMask;
		return true;
	}
	void pop() { at++; }
	// append every k-mer <= threshold
	void take_upto(uint64_t threshold, vector<uint64_t> &out) {
		uint64_t k;
		while (peek(k) && k <= threshold) { out.push_back(k); pop(); }
	}
};

struct AccessionPath { string path, name; };
vector<AccessionPath> read_accessions_path_list(const string &fn) {   // kmer_general.cpp:32-43
	ifstream fin(fn);
	vector<AccessionPath> res;
	AccessionPath a;
	while (fin >> a.path) { fin >> a.name; res.push_back(a); }
	return res;
}
}  // namespace

int main(int argc, char *argv[]) {
	CliOptions options("build_kmers_table", "Build the k-mers table");
	options.add('l', "list_kmers_files", "list of separate k-mers files");
	options.add('k', "kmers_len", "length of k-mers");
	options.add('a', "all_kmers", "path to file with all k-mers");
	options.add('o', "output", "prefix for kmers-table files");
	options.add(0, "range_kmers", "k-mers (table rows) per GPU pass", false, "4194304");
	options.add(0, "device", "CUDA device ordinal", false, "0");
	options.add(0, "help", "print help", true);
	try {
		options.parse(argc, argv);
		if (options.count("help")) { cerr << options.help() << endl; return 0; }
		for (const char *req : {"list_kmers_files", "kmers_len", "all_kmers", "output"})
			if (!options.count(req)) { cerr << req << " is a required parameter" << endl; cerr << options.help() << endl; return 1; }
		const string fn_list = options.str("list_kmers_files"), fn_all = options.str("all_kmers"), output_base = options.str("output");
		const size_t kmer_len = options.as<size_t>("kmers_len");
		for (const string &f : {fn_list, fn_all})
			if (!is_file_exist(f)) { cerr << "Couldn't find file: " << f << endl; return 1; }
		if ((kmer_len > 31) || (kmer_len < 10)) { cerr << "kmer length has to be between 10-31" << endl; return 1; }
		const uint64_t range = max<uint64_t>(1, options.as<uint64_t>("range_kmers"));
		const int device = options.as<int>("device");

		const vector<AccessionPath> acc = read_accessions_path_list(fn_list);
		ofstream fout_names(output_base + ".names", ios::binary);
		vector<SortedKmerFile> files(acc.size());
		for (size_t i = 0; i < acc.size(); i++) {
			fout_names << acc[i].name << endl;
			if (!is_file_exist(acc[i].path)) { cerr << "Couldn't find file: " << acc[i].path << endl; return 1; }
			if (!files[i].open(acc[i].path)) throw logic_error("can't open file: " + acc[i].path);
			uint64_t k;
			if (!files[i].peek(k)) throw logic_error("sorted kmer file is empty: " + acc[i].path);
		}
		fout_names.close();
		cerr << "Create merger" << endl;
		SortedKmerFile all;
		if (!all.open(fn_all)) throw logic_error("can't open file: " + fn_all);
		cerr << "Opens file" << endl;
		ofstream fout(output_base + ".table", ios::binary);
		// header (kmers_merge_multiple_databaes.cpp:54-58): AA BB CC DD, u64 accessions, u32 k-mer length
		const uint64_t n_acc = acc.size();
		const uint32_t klen32 = (uint32_t)kmer_len;
		const unsigned char magic[4] = {0xAA, 0xBB, 0xCC, 0xDD};
		fout.write(reinterpret_cast<const char *>(magic), 4);
		fout.write(reinterpret_cast<const char *>(&n_acc), sizeof n_acc);
		fout.write(reinterpret_cast<const char *>(&klen32), sizeof klen32);
		const size_t w = (n_acc + 63) / 64;
		vector<uint64_t> all_range, packed, offsets(n_acc + 1), table;
		uint64_t rows_total = 0, pass = 0;
		for (;;) {
			all_range.clear();
			uint64_t k;
			while (all_range.size() < range && all.peek(k)) { all_range.push_back(k); all.pop(); }
			if (all_range.empty()) break;
			const uint64_t threshold = all_range.back();
			packed.clear();
			for (size_t a = 0; a < n_acc; a++) {
				offsets[a] = packed.size();
				files[a].take_upto(threshold, packed);
			}
			offsets[n_acc] = packed.size();
			table.resize(all_range.size() * (w + 1));
			if (kg_table_build(device, all_range.data(), all_range.size(), (uint32_t)n_acc, packed.data(), offsets.data(), table.data()) != KG_OK)
				throw runtime_error(string("kg_table_build: ") + kg_last_error(nullptr));
			fout.write(reinterpret_cast<const char *>(table.data()), (streamsize)(table.size() * 8));
			rows_total += all_range.size();
			cerr << ++pass << " : Wrote: kmers=" << all_range.size() << "\tpa words=" << all_range.size() * w << endl;
		}
		cerr << "close file" << endl;
		fout.close();
	} catch (const CliOptions::ParseError &e) {
		cerr << "error parsing options: " << e.what() << endl;
		cerr << options.help() << endl;
		return 1;
	} catch (const std::exception &e) {
		cerr << "build_kmers_table: " << e.what() << endl;
		return 2;
	}
	return 0;
}
