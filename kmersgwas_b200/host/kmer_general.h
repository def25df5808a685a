// kmer_general.h -- host-side utilities of the association hot path (B200 build).
//
// Mirrors the part of /root/reference/src/kmer_general.{h,cpp} that the hot path uses, with the same
// public names and file formats, so that code written against the reference compiles against this:
//   types        PhenotypeList, AssociationScoreHeap, AssociationOutputInfo, kmers_output_list,
//                cmp_second, AssociationsPriorityQueue, BedBimFilesHandle     (kmer_general.h:54,113-145)
//   functions    load_kmers_talbe_column_names (:45-53), load_phenotypes_file (:175-205),
//                intersect_phenotypes_to_present_DBs (:239-253), get_index_DB (:227-237),
//                write_fam_file (:207-225), bits2kmer31 (:77-87), get_time (:102-107), is_file_exist
// Table construction helpers (KMC adaptor, kmer2bits, ...) are out of scope.
#ifndef KGH_KMER_GENERAL_H
#define KGH_KMER_GENERAL_H

#include <cstddef>
#include <cstdint>
#include <fstream>
#include <queue>
#include <string>
#include <tuple>
#include <unordered_set>
#include <utility>
#include <vector>

#define MAX_KMER_LEN 31
#define MIN_KMER_LEN 15
#define WLEN 64
#define NULL_KEY 0xFFFFFFFFFFFFFFFFull

// 64-bit mixer used for presence/absence pattern hashing (kmer_general.h:32-41)
struct Hash64 {
	std::size_t operator()(uint64_t key) const {
		key = (key ^ (key >> 33)) * 0xff51afd7ed558ccdull;
		key = (key ^ (key >> 33)) * 0xc4ceb9fe1a85ec53ull;
		return key ^ (key >> 33);
	}
};

// The reference uses google::dense_hash_set; only membership and size are ever observed.
typedef std::unordered_set<uint64_t, Hash64> KmersSet;

typedef std::pair<std::vector<std::string>, std::vector<float> > PhenotypeList;

typedef std::tuple<uint64_t, double, std::size_t> AssociationScoreHeap;    // k-mer, score, row
typedef std::tuple<uint64_t, uint64_t, std::size_t> AssociationOutputInfo;  // k-mer, rank, row

struct kmers_output_list {
	std::vector<AssociationOutputInfo> list;
	std::size_t next_index;
};

// min-heap on the score: the queue's top() is the lowest kept score
struct cmp_second {
	inline bool operator()(const AssociationScoreHeap &l, const AssociationScoreHeap &r) const {
		return std::get<1>(l) > std::get<1>(r);
	}
};
typedef std::priority_queue<AssociationScoreHeap, std::vector<AssociationScoreHeap>, cmp_second> AssociationsPriorityQueue;

// PLINK .bed/.bim pair; the .bed magic (6C 1B 01) is written on open.
struct BedBimFilesHandle {
	explicit BedBimFilesHandle(const std::string &base_name);
	BedBimFilesHandle(BedBimFilesHandle &&o) = default;
	~BedBimFilesHandle() { close(); }
	void close();
	std::ofstream f_bed;
	std::ofstream f_bim;
};

std::vector<std::string> load_kmers_talbe_column_names(const std::string &kmers_table_base);
std::pair<std::vector<std::string>, std::vector<PhenotypeList> > load_phenotypes_file(const std::string &filename);
std::size_t get_index_DB(const std::string &name, const std::vector<std::string> &names);
PhenotypeList intersect_phenotypes_to_present_DBs(const PhenotypeList &pl, const std::string &kmers_table_base,
                                                  const bool &must_be_present);
void write_fam_file(const std::vector<PhenotypeList> &phenotypes, const std::string &fn);
void write_fam_file(const PhenotypeList &phenotype, const std::string &fn);
std::string bits2kmer31(uint64_t w, const std::size_t &k);
double get_time(void);
bool is_file_exist(const std::string &file_name);

#endif
