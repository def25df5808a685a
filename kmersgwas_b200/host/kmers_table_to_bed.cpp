// kmers_table_to_bed -- convert the k-mers table to PLINK .bed/.bim/.fam files (B200 build).
//
// Same flags and output files as the reference CLI (/root/reference/src/kmers_table_to_bed.cpp: flags :41-50, batch loop
// :104-125): every row that passes the MAC filter of load_kmers (kmers_multiple_databases.cpp:117-121) is written, in
// table order, into <output>.<batch>.{bed,bim,fam}; a batch file holds at most --batch_size KEPT rows
// (load_kmers stops reading at the batch_size-th kept row, :110, and a new batch is started while file rows remain,
// :103-108).  With -u only the first row of every distinct presence/absence pattern is written (:262-272).
// The MAC filter runs on the GPU (kg_mac_filter); new flags: --device D, --rows_per_load R (raw rows per device pass).
#include <cmath>
#include <exception>
#include <fstream>
#include <iostream>
#include <memory>

#include "cli_options.h"
#include "kmer_general.h"
#include "kmers_multiple_databases.h"

using namespace std;

int main(int argc, char *argv[]) {
	CliOptions options("kmers_table_to_bed", "Convert k-mers table to PLINK binary format");
	options.add('t', "kmers_table", "k-mers table path");
	options.add('k', "kmers_len", "length of k-mers");
	options.add('p', "phentype_file", "phenotype file, condense output only to individuals with a phenotype");
	options.add(0, "maf", "minor allele frequency");
	options.add(0, "mac", "minor allele count");
	options.add('b', "batch_size", "maximal number of variants in each PLINK bed file (seperate to many file  if needed)");
	options.add('o', "output", "prefix for output files");
	options.add('u', "unique_patterns", "output only unique presence/absence patterns", true);
	options.add(0, "device", "CUDA device ordinal", false, "0");
	options.add(0, "rows_per_load", "raw table rows per device pass", false, "4194304");
	options.add(0, "help", "print help", true);
	try {
		options.parse(argc, argv);
		if (options.count("help")) {
			cerr << options.help() << endl;
			return 0;
		}
		for (const char *req : {"kmers_table", "kmers_len", "phentype_file", "maf", "mac", "batch_size", "output"}) {
			if (options.count(req) == 0) {
				cerr << req << " is a required parameter" << endl;
				cerr << options.help() << endl;
				return 1;
			}
		}
		const string fn_kmers_table = options.str("kmers_table");
		const string fn_phenotypes = options.str("phentype_file");
		const double MAF = options.as<double>("maf");
		const size_t MAC = options.as<size_t>("mac");
		const size_t max_batch_size = options.as<size_t>("batch_size");
		const size_t kmer_len = options.as<size_t>("kmers_len");
		const string output_base = options.str("output");
		const bool unique_patterns = options.count("unique_patterns") != 0;
		for (const string &f : {fn_kmers_table + ".names", fn_kmers_table + ".table", fn_phenotypes}) {
			if (!is_file_exist(f)) {
				cerr << "Couldn't find file: " << f << endl;
				return 1;
			}
		}
		if ((kmer_len > 31) || (kmer_len < 10)) {
			cerr << "kmer length has to be between 10-31" << endl;
			return 1;
		}
		const uint64_t rows_per_load = max<uint64_t>(1, options.as<uint64_t>("rows_per_load"));

		// phenotypes: only the first column is used, and only for the accession list / the .fam values (:94-96)
		pair<vector<string>, vector<PhenotypeList> > phenotypes_file_info = load_phenotypes_file(fn_phenotypes);
		cerr << "using " << phenotypes_file_info.first[0] << endl;
		PhenotypeList pheno_info = intersect_phenotypes_to_present_DBs(phenotypes_file_info.second[0], fn_kmers_table, true);

		size_t min_count = (size_t)ceil(double(pheno_info.first.size()) * MAF);   // (:99-101)
		if (min_count < MAC) min_count = MAC;

		MultipleKmersDataBases::set_device(options.as<int>("device"));
		MultipleKmersDataBases multiDB(fn_kmers_table, pheno_info.first, (uint32_t)kmer_len);

		KmersSet pa_patterns;
		auto write_fam = [&](const string &base) {
			ofstream fout(base + ".fam");
			for (size_t i = 0; i < pheno_info.first.size(); i++)
				fout << pheno_info.first[i] << " " << pheno_info.first[i] << " 0 0 0 " << pheno_info.second[i] << endl;
		};

		size_t batch_index = 0;
		size_t kept_in_batch = 0;
		unique_ptr<BedBimFilesHandle> out;
		auto open_batch = [&]() {
			cerr << "Batch:\t" << batch_index + 1 << endl;
			out.reset(new BedBimFilesHandle(output_base + "." + to_string(batch_index)));
			kept_in_batch = 0;
		};
		auto close_batch = [&]() {
			out->close();
			out.reset();
			write_fam(output_base + "." + to_string(batch_index));
			batch_index++;
		};
		cerr << "loading.... " << endl;
		vector<uint8_t> keep;
		const uint64_t total_rows = multiDB.rows_in_file();
		uint64_t rows_done = 0;
		while (multiDB.load_kmers(rows_per_load, min_count)) {
			multiDB.mac_filter_loaded(min_count, keep);
			const uint64_t n = multiDB.rows_loaded();
			for (uint64_t r = 0; r < n; r++, rows_done++) {
				// the reference opens a batch whenever file rows remain at the start of load_kmers (:105-108)
				if (!out) open_batch();
				if (keep[r]) {
					kept_in_batch++;   // counts against the batch size whether or not its pattern is new (:110, :262-272)
					bool write = true;
					if (unique_patterns) {
						const uint64_t seed = multiDB.presence_absence_pattern_hash_loaded_row(r);
						write = pa_patterns.insert(seed).second;
					}
					if (write) multiDB.output_plink_loaded_row(*out, r);
					// load_kmers stops reading as soon as it holds batch_size kept rows
					if (kept_in_batch == max_batch_size) close_batch();
				}
			}
		}
		if (out) close_batch();
		(void)total_rows;
		(void)rows_done;
	} catch (const CliOptions::ParseError &e) {
		cerr << "error parsing options: " << e.what() << endl;
		cerr << options.help() << endl;
		return 1;
	} catch (const std::exception &e) {
		cerr << "kmers_table_to_bed: " << e.what() << endl;
		return 2;
	}
	return 0;
}
