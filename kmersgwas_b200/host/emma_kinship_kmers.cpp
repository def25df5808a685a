// emma_kinship_kmers -- kinship (IBS) matrix from the k-mers table, printed to stdout (B200 build).
//
// Same flags and output as the reference CLI (/root/reference/src/emma_kinship_kmers.cpp: flags :37-42,
// normalisation and printing :95-111).  The accumulation (:89-92 ->
// kmers_multiple_databases.cpp:418-438) runs on the GPU; new flags: --gpus G (row shards on G GPUs, their integer
// accumulators summed exactly by one NCCL all-reduce), --device D, --engine {0 auto, 1 popcount, 2 tensor cores}.
#include <cmath>
#include <exception>
#include <iostream>
#include <memory>
#include <thread>

#include "cli_options.h"
#include "kmer_general.h"
#include "kmers_multiple_databases.h"

using namespace std;

int main(int argc, char *argv[]) {
	CliOptions options("emma_kinship_kmers", "Calculate a kinship matrix from the k-mers table (output to stdout)");
	options.add('t', "kmers_table", "k-mers table path");
	options.add('k', "kmers_len", "length of k-mers");
	options.add(0, "maf", "minor allele frequency");
	options.add(0, "gpus", "Number of GPUs to shard the table over", false, "1");
	options.add(0, "device", "First CUDA device ordinal", false, "0");
	options.add(0, "engine", "Kinship engine: 0 auto, 1 popcount, 2 tensor cores", false, "0");
	options.add(0, "batch_size", "rows per load", false, "4194304");
	options.add(0, "help", "print help", true);
	try {
		options.parse(argc, argv);
		if (options.count("help")) {
			cerr << options.help() << endl;
			return 0;
		}
		for (const char *req : {"kmers_table", "kmers_len", "maf"}) {
			if (options.count(req) == 0) {
				cerr << req << " is a required parameter" << endl;
				cerr << options.help() << endl;
				return 1;
			}
		}
		const string fn_kmers_table = options.str("kmers_table");
		const double MAF = options.as<double>("maf");
		const size_t kmer_len = options.as<size_t>("kmers_len");
		for (const string &f : {fn_kmers_table + ".names", fn_kmers_table + ".table"}) {
			if (!is_file_exist(f)) {
				cerr << "Couldn't find file: " << f << endl;
				return 1;
			}
		}
		if ((kmer_len > 31) || (kmer_len < 10)) {
			cerr << "kmer length has to be between 10-31" << endl;
			return 1;
		}
		const size_t n_gpus = max<size_t>(1, options.as<size_t>("gpus"));
		const int device0 = options.as<int>("device");
		const int engine = options.as<int>("engine");
		const uint64_t batch = options.as<uint64_t>("batch_size");

		const vector<string> names = load_kmers_talbe_column_names(fn_kmers_table);
		const size_t n_acc = names.size();
		const size_t min_count = (size_t)ceil(static_cast<double>(n_acc) * MAF);
		cerr << "Min count = " << min_count << endl;
		uint64_t n_snps = 0;
		vector<vector<uint64_t> > K(n_acc, vector<uint64_t>(n_acc, 0));

		cerr << "loading..." << endl;
		vector<unique_ptr<MultipleKmersDataBases> > dbs(n_gpus);
		for (size_t g = 0; g < n_gpus; g++) {
			MultipleKmersDataBases::set_device(device0 + (int)g);
			dbs[g].reset(new MultipleKmersDataBases(fn_kmers_table, names, (uint32_t)kmer_len));
			dbs[g]->set_kinship_engine(engine);
		}
		const uint64_t total_rows = dbs[0]->rows_in_file();
		vector<exception_ptr> errors(n_gpus);
		auto work = [&](size_t g) {
			try {
				MultipleKmersDataBases &db = *dbs[g];
				if (n_gpus > 1) {
					const uint64_t first = total_rows * g / n_gpus, last = total_rows * (g + 1) / n_gpus;
					db.restrict_to_rows(first, last - first);
				}
				db.kinship_begin(min_count);
				while (db.load_kmers(batch, min_count)) {
					if (g == 0) { cerr << "."; cerr.flush(); }
					db.kinship_accumulate_loaded();
				}
			} catch (...) {
				errors[g] = current_exception();
			}
		};
		vector<thread> threads;
		for (size_t g = 1; g < n_gpus; g++) threads.emplace_back(work, g);
		work(0);
		for (auto &t : threads) t.join();
		for (size_t g = 0; g < n_gpus; g++)
			if (errors[g]) rethrow_exception(errors[g]);
		if (n_gpus > 1) {
			// the path's one exchange step: NCCL all-reduce (sum) of the u64 accumulators over NVLink; exact, shards add
			vector<kg_ctx *> ctxs(n_gpus);
			for (size_t g = 0; g < n_gpus; g++) ctxs[g] = dbs[g]->context();
			if (kg_comm_init_all(ctxs.data(), (int)n_gpus) != KG_OK)
				throw runtime_error(string("kg_comm_init_all: ") + kg_last_error(ctxs[0]));
			if (kg_kinship_allreduce_all(ctxs.data(), (int)n_gpus) != KG_OK)
				throw runtime_error(string("kg_kinship_allreduce_all: ") + kg_last_error(ctxs[0]));
		}
		dbs[0]->kinship_finish(K, n_snps);
		cerr << "#" << n_snps << endl;

		// normalise + print (reference :95-111): K/n_snps, symmetric, unit diagonal, default precision
		vector<vector<double> > K_norm(n_acc, vector<double>(n_acc, 0));
		for (size_t i = 0; i < n_acc; i++) {
			K_norm[i][i] = 1;
			for (size_t j = 0; j < i; j++) {
				K_norm[i][j] = static_cast<double>(K[i][j]) / static_cast<double>(n_snps);
				K_norm[j][i] = K_norm[i][j];
			}
		}
		for (size_t i = 0; i < n_acc; i++) {
			for (size_t j = 0; j < n_acc; j++) {
				if (j > 0) cout << "\t";
				cout << K_norm[i][j];
			}
			cout << "\n";
		}
	} catch (const CliOptions::ParseError &e) {
		cerr << "error parsing options: " << e.what() << endl;
		cerr << options.help() << endl;
		return 1;
	} catch (const std::exception &e) {
		cerr << "emma_kinship_kmers: " << e.what() << endl;
		return 2;
	}
	return 0;
}
