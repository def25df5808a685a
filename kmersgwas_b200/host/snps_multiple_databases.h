// snps_multiple_databases.h -- the SNP twin of the association scan (B200 build).
//
// Same class name and public methods as /root/reference/src/snps_multiple_databases.h:25-41.  The reference expands
// the PLINK .bed file into three bit planes in host memory and scores every SNP on one core; here the .bed payload
// goes to the GPU as it is and kg_snps_scores (include/kmersgwas_b200.h) builds the planes and scores every SNP
// against ALL phenotypes in one pass (bit-identical scores); the best-N selection stays the reference's heap replay.
#ifndef KGH_SNPS_MULTIDB_H
#define KGH_SNPS_MULTIDB_H
#include <stdint.h>

#include <string>
#include <vector>

class MultipleSNPsDataBases {
	public:
		MultipleSNPsDataBases(const std::string &bedbin_base_fn, const std::vector<std::string> &samples_to_use);
		MultipleSNPsDataBases() = delete;
		MultipleSNPsDataBases(const MultipleSNPsDataBases &) = delete;
		MultipleSNPsDataBases &operator=(const MultipleSNPsDataBases &) = delete;

		// indices of the best N SNPs for one phenotype, ascending (reference :225-236)
		std::vector<std::size_t> get_most_associated_snps(std::vector<float> phenotypes, const std::size_t &number_of_best_associations,
		                                                  const double &min_minor_allele_count) const;
		// the same for several phenotypes in one device pass
		std::vector<std::vector<std::size_t> > get_most_associated_snps(const std::vector<std::vector<float> > &phenotypes,
		                                                                const std::size_t &number_of_best_associations,
		                                                                const double &min_minor_allele_count) const;
		// copy the selected SNPs' .bed rows and .bim lines into one PLINK pair per phenotype (reference :246-286)
		void output_plink_bed_file(const std::vector<std::string> &files_base_names,
		                           std::vector<std::vector<std::size_t> > SNPs_indices) const;
		static void set_device(int device) { s_device = device; }
		std::size_t snps() const { return m_n_snps; }

	private:
		std::string m_base_name;
		std::vector<std::string> m_samples_names;
		std::size_t m_n_snps, m_n_bytes_per_snp;
		std::vector<uint8_t> m_bed;                    // .bed payload (after the 3-byte magic)
		std::vector<uint32_t> m_map_byte, m_map_shift; // sample i of the phenotype order -> byte / bit of a .bed row
		static int s_device;
};

#endif
