// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A small driver of OURS that links against the UNMODIFIED reference objects
// (built by oracle/Makefile from /root/reference/src) and calls only their
// PUBLIC methods, to extract full-precision data the reference CLIs do not print:
//
//   ref_harness scores  <table_base> <kmer_len> <pheno.tsv> <min_count> <batch_rows> <out_prefix>
//       every kept row's (kmer u64, score f64) per phenotype j -> <out_prefix>.<j>.scores
//       (ascending score = heap pop order), plus <out_prefix>.tested = #kept rows.
//       Uses MultipleKmersDataBases::load_kmers / add_kmers_to_heap
//       (src/kmers_multiple_databases.cpp:103-146,275-284) and
//       BestAssociationsHeap::output_to_file_with_scores (src/best_associations_heap.cpp:82-92)
//       with a heap large enough that nothing is ever evicted.
//
//   ref_harness kinship <table_base> <kmer_len> <min_count> <batch_rows> <out.bin>
//       raw u64 IBS counts: writes u64 n_acc, u64 n_snps, then n_acc*n_acc u64 (row-major,
//       only j<i filled) from update_emma_kinshhip_calculation (:418-438); the CLI
//       itself only prints 6 significant digits (src/emma_kinship_kmers.cpp:95-111).
#include "kmer_general.h"
#include "kmers_multiple_databases.h"
#include "best_associations_heap.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace std;

static int run_scores(int argc, char **argv) {
	if (argc != 8) { fprintf(stderr, "bad args\n"); return 2; }
	string table = argv[2];
	uint32_t klen = (uint32_t)atoi(argv[3]);
	string pheno = argv[4];
	size_t min_count = (size_t)atoll(argv[5]);
	uint64_t batch = (uint64_t)atoll(argv[6]);
	string out = argv[7];

	auto info = load_phenotypes_file(pheno);
	size_t P = info.first.size();
	for (size_t i = 0; i < P; i++)
		info.second[i] = intersect_phenotypes_to_present_DBs(info.second[i], table, true);
	MultipleKmersDataBases db(table, info.second[0].first, klen);
	vector<BestAssociationsHeap> heaps(P, BestAssociationsHeap((size_t)1 << 40));
	while (db.load_kmers(batch, min_count))
		for (size_t j = 0; j < P; j++)
			db.add_kmers_to_heap(heaps[j], info.second[j].second, min_count);
	for (size_t j = 0; j < P; j++)
		heaps[j].output_to_file_with_scores(out + "." + to_string(j) + ".scores");
	ofstream ft(out + ".tested");
	ft << heaps[0].number_of_insertion() << endl;
	return 0;
}

static int run_kinship(int argc, char **argv) {
	if (argc != 7) { fprintf(stderr, "bad args\n"); return 2; }
	string table = argv[2];
	uint32_t klen = (uint32_t)atoi(argv[3]);
	size_t min_count = (size_t)atoll(argv[4]);
	uint64_t batch = (uint64_t)atoll(argv[5]);
	string out = argv[6];

	vector<string> names = load_kmers_talbe_column_names(table);
	MultipleKmersDataBases db(table, names, klen);
	uint64_t n_acc = names.size(), n_snps = 0;
	vector<vector<uint64_t> > K(n_acc, vector<uint64_t>(n_acc, 0));
	while (db.load_kmers(batch, min_count))
		db.update_emma_kinshhip_calculation(K, n_snps);
	FILE *f = fopen(out.c_str(), "wb");
	if (!f) return 3;
	fwrite(&n_acc, 8, 1, f);
	fwrite(&n_snps, 8, 1, f);
	for (uint64_t i = 0; i < n_acc; i++) fwrite(K[i].data(), 8, n_acc, f);
	fclose(f);
	return 0;
}

int main(int argc, char **argv) {
	if (argc >= 2 && !strcmp(argv[1], "scores")) return run_scores(argc, argv);
	if (argc >= 2 && !strcmp(argv[1], "kinship")) return run_kinship(argc, argv);
	fprintf(stderr, "usage: ref_harness scores|kinship ...\n");
	return 2;
}
