// associate_snps -- best-N SNPs per phenotype from a PLINK bed/bim/fam triple (B200 build).
// Same positional arguments, outputs and stderr vocabulary as the reference CLI
// (/root/reference/src/associate_snps.cpp:29-73); all phenotypes are scored in one GPU pass.
#include <cmath>
#include <cstdlib>
#include <exception>
#include <iostream>

#include "kmer_general.h"
#include "snps_multiple_databases.h"

using namespace std;

int main(int argc, char *argv[]) {
	if (argc != 7) {
		cerr << "usage: " << argv[0] << " <phenotypes file> <base bedbim file> <base output files> <# snps to output> <maf> <mac>" << endl;
		return 1;
	}
	try {
		pair<vector<string>, vector<PhenotypeList> > phenotypes_info = load_phenotypes_file(argv[1]);
		cerr << "Loading snps information" << endl;
		if (phenotypes_info.second.empty()) throw logic_error(string("no phenotype columns in ") + argv[1]);
		MultipleSNPsDataBases snps_dataset(argv[2], phenotypes_info.second[0].first);
		const size_t n_samples = phenotypes_info.second[0].first.size();
		const size_t n_best = (size_t)atoi(argv[4]);
		const double maf = atof(argv[5]);
		cerr << "MAF = " << maf << " n_sample = " << n_samples << endl;
		double mac = atof(argv[6]);
		if (mac < ceil(maf * n_samples)) mac = ceil(maf * n_samples);
		cerr << "Minor allele count  = " << mac << endl;
		const size_t phenotype_n = phenotypes_info.first.size();
		cerr << "Associating phenotypes:";
		cerr.flush();
		const double t0 = get_time();
		vector<vector<float> > y(phenotype_n);
		for (size_t j = 0; j < phenotype_n; j++) {
			cerr << ".";
			y[j] = phenotypes_info.second[j].second;
		}
		const vector<vector<size_t> > best = snps_dataset.get_most_associated_snps(y, n_best, mac);
		cerr << "Average time per phenotype:\t" << (get_time() - t0) / phenotype_n << endl;
		cerr << "\noutputting best snps";
		vector<string> bases;
		for (size_t j = 0; j < phenotype_n; j++) bases.push_back(string(argv[3]) + "." + phenotypes_info.first[j]);
		snps_dataset.output_plink_bed_file(bases, best);
	} catch (const std::exception &e) {
		cerr << "associate_snps: " << e.what() << endl;
		return 2;
	}
	return 0;
}
