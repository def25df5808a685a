// kmers_multiple_databases.cpp -- see kmers_multiple_databases.h.
#include "kmers_multiple_databases.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <iostream>
#include <stdexcept>

using std::string;
using std::vector;

int MultipleKmersDataBases::s_device = 0;

namespace {
const uint32_t kTableMagic = 0xDDCCBBAAu;
const size_t kHeaderBytes = 4 + 8 + 4;
const uint64_t kSubTileRows = 1ull << 20;     // rows per device submit (H2D of tile i+1 overlaps kernels of tile i)

void read_fully(int fd, void *dst, size_t bytes, uint64_t offset, const string &path) {
	char *p = static_cast<char *>(dst);
	while (bytes > 0) {
		ssize_t got = pread(fd, p, bytes, (off_t)offset);
		if (got <= 0) throw std::logic_error("Couldn't read kmer table file: " + path);
		p += got;
		bytes -= (size_t)got;
		offset += (uint64_t)got;
	}
}
}  // namespace

void MultipleKmersDataBases::check(kg_status st, const char *what) const {
	if (st != KG_OK) throw std::runtime_error(string(what) + ": " + kg_last_error(m_ctx));
}

// Header parse + validation: reference ctor :39-94.  Layout: u32 magic, u64 N_file, u32 k, then rows.
MultipleKmersDataBases::MultipleKmersDataBases(const string &kmers_table_base, const vector<string> &db_to_use,
                                               const uint32_t &kmer_len)
    : m_db_names_db_file(load_kmers_talbe_column_names(kmers_table_base)),
      m_db_names_table(db_to_use),
      m_accessions_db_file(m_db_names_db_file.size()),
      m_accessions(db_to_use.size()),
      m_hash_words_db_file((m_accessions_db_file + WLEN - 1) / WLEN),
      m_hash_words(2 * ((m_accessions + (2 * WLEN) - 1) / (2 * WLEN))),
      m_kmer_len(kmer_len),
      m_table_path(kmers_table_base + ".table"),
      m_fd(-1), m_kmer_number(0), m_first_row(0), m_kmer_loaded(0), m_row_offset(0),
      m_batch(nullptr), m_batch_cap(0), m_rows_loaded(0), m_load_mac(0),
      m_ctx(nullptr), m_pheno_min_cnt(0), m_kinship_streaming(false) {
	m_fd = open(m_table_path.c_str(), O_RDONLY);
	if (m_fd < 0) throw std::logic_error("Couldn't open kmer table file: " + m_table_path);
	struct stat sb;
	if (fstat(m_fd, &sb) != 0) throw std::logic_error("Couldn't open kmer table file: " + m_table_path);
	const uint64_t file_size = (uint64_t)sb.st_size;
	if (file_size <= kHeaderBytes) throw std::logic_error("Kmer table size is too small");
	unsigned char hdr[kHeaderBytes];
	read_fully(m_fd, hdr, kHeaderBytes, 0, m_table_path);
	uint32_t magic, file_k;
	uint64_t file_n;
	memcpy(&magic, hdr, 4);
	memcpy(&file_n, hdr + 4, 8);
	memcpy(&file_k, hdr + 12, 4);
	if (magic != kTableMagic) throw std::logic_error("Incorrect prefix");
	if (file_n != m_accessions_db_file) throw std::logic_error("Number of accession in file not as defined in class");
	if (file_k != m_kmer_len) throw std::logic_error("Kmer length not as defined in class");
	const uint64_t row_bytes = 8ull * (1 + m_hash_words_db_file);
	if ((file_size - kHeaderBytes) % row_bytes != 0) throw std::logic_error("size of file not valid");
	m_kmer_number = (file_size - kHeaderBytes) / row_bytes;
	create_map_from_all_DBs();

	kg_shape shape;
	shape.n_file = m_accessions_db_file;
	shape.n_used = m_accessions;
	shape.map_word = m_map_word_index.data();
	shape.map_bit = m_map_bit_index.data();
	kg_status st = kg_ctx_create(s_device, &shape, nullptr, &m_ctx);
	if (st != KG_OK) throw std::runtime_error(string("kg_ctx_create: ") + kg_last_error(nullptr));
}

MultipleKmersDataBases::~MultipleKmersDataBases() {
	if (m_batch) kg_host_free(m_ctx, m_batch);
	if (m_ctx) kg_ctx_destroy(m_ctx);
	if (m_fd >= 0) close(m_fd);
}

// memory column i <- (file word, bit) of accession i of db_to_use; reference :297-311
void MultipleKmersDataBases::create_map_from_all_DBs() {
	m_map_word_index.clear();
	m_map_bit_index.clear();
	m_map_mask.assign(m_hash_words_db_file, 0);
	for (size_t i = 0; i < m_accessions; i++) {
		const auto it = std::find(m_db_names_db_file.begin(), m_db_names_db_file.end(), m_db_names_table[i]);
		if (it == m_db_names_db_file.end())
			throw std::logic_error("All accessions suppose to be in DB file: " + m_db_names_table[i]);
		const size_t col = (size_t)(it - m_db_names_db_file.begin());
		m_map_word_index.push_back((uint32_t)(col / WLEN));
		m_map_bit_index.push_back((uint32_t)(col % WLEN));
		m_map_mask[col / WLEN] |= 1ull << (col % WLEN);
	}
}

void MultipleKmersDataBases::restrict_to_rows(uint64_t first, uint64_t count) {
	if (first > m_kmer_number) first = m_kmer_number;
	if (count > m_kmer_number - first) count = m_kmer_number - first;
	m_first_row = first;
	m_kmer_number = first + count;  // exclusive end, see load_kmers
	m_kmer_loaded = first;
	m_row_offset = first;
	m_rows_loaded = 0;
}

bool MultipleKmersDataBases::load_kmers(const uint64_t &batch_size, const size_t &mac) {
	m_row_offset = m_kmer_loaded;
	m_rows_loaded = 0;
	m_load_mac = mac;
	const uint64_t left = m_kmer_number - m_kmer_loaded;
	if (left == 0) return false;
	const uint64_t n = std::min<uint64_t>(batch_size, left);
	const size_t stride = 1 + m_hash_words_db_file;
	if (m_batch_cap < n) {
		if (m_batch) kg_host_free(m_ctx, m_batch);
		m_batch = nullptr;
		void *p = nullptr;
		check(kg_host_alloc(m_ctx, (size_t)n * stride * 8, &p), "kg_host_alloc");
		m_batch = static_cast<uint64_t *>(p);
		m_batch_cap = n;
	}
	read_fully(m_fd, m_batch, (size_t)n * stride * 8, kHeaderBytes + m_kmer_loaded * stride * 8, m_table_path);
	m_kmer_loaded += n;
	m_rows_loaded = n;
	return true;
}

void MultipleKmersDataBases::set_scan_engine(int engine) const { check(kg_set_option(m_ctx, KG_OPT_SCAN_ENGINE, engine), "kg_set_option"); }
void MultipleKmersDataBases::set_kinship_engine(int engine) const { check(kg_set_option(m_ctx, KG_OPT_KINSHIP_ENGINE, engine), "kg_set_option"); }

// ---- scoring ---------------------------------------------------------------------------------------
void MultipleKmersDataBases::ensure_phenotypes(const vector<vector<float> > &scores, size_t min_cnt) const {
	const size_t P = scores.size();
	vector<float> flat;
	flat.reserve(P * m_accessions);
	for (size_t j = 0; j < P; j++) {
		if (scores[j].size() != m_accessions)
			throw std::logic_error("phenotype vector length differs from the number of accessions used");
		flat.insert(flat.end(), scores[j].begin(), scores[j].end());
	}
	const bool same = m_pheno_min_cnt == min_cnt && flat.size() == m_pheno_flat.size() &&
	                  (flat.empty() || memcmp(flat.data(), m_pheno_flat.data(), flat.size() * sizeof(float)) == 0);
	if (same) return;
	check(kg_scan_set_phenotypes(m_ctx, flat.data(), (uint32_t)P, min_cnt), "kg_scan_set_phenotypes");
	m_pheno_flat.swap(flat);
	m_pheno_min_cnt = min_cnt;
	m_driver.rows_scored = 0;
	m_driver.rows_submitted = 0;
	m_driver.kept_seen = 0;
}

void MultipleKmersDataBases::add_kmers_to_heap(BestAssociationsHeap &heap, vector<float> scores, const size_t &min_cnt) const {
	// The device scores all phenotypes it holds in one pass; a single heap is the P = 1 case.
	if (min_cnt != m_load_mac)
		throw std::logic_error("GPU path: load_kmers' minor allele count and add_kmers_to_heap's min_cnt must be equal");
	ensure_phenotypes(vector<vector<float> >(1, scores), min_cnt);
	BestAssociationsHeap *hp = &heap;
	kgh_associate_rows(m_ctx, &hp, 1, m_batch, m_rows_loaded, m_row_offset, 1 + m_hash_words_db_file, m_driver);
	kgh_associate_finish(m_ctx, &hp, 1, m_driver);
}

// All phenotypes of the loaded batch in one device pass (reference loop body associate_kmers.cpp:134-141
// + kmers_multiple_databases.cpp:275-284); the round / replay logic lives in association_driver.cpp.
void MultipleKmersDataBases::add_kmers_to_heaps(vector<BestAssociationsHeap> &heaps, const vector<vector<float> > &scores,
                                                const size_t &min_cnt) const {
	if (heaps.size() != scores.size()) throw std::logic_error("heaps and phenotypes differ in number");
	if (min_cnt != m_load_mac)
		throw std::logic_error("GPU path: load_kmers' minor allele count and add_kmers_to_heap's min_cnt must be equal");
	ensure_phenotypes(scores, min_cnt);
	vector<BestAssociationsHeap *> hp(heaps.size());
	for (size_t j = 0; j < heaps.size(); j++) hp[j] = &heaps[j];
	kgh_associate_rows(m_ctx, hp.data(), hp.size(), m_batch, m_rows_loaded, m_row_offset, 1 + m_hash_words_db_file, m_driver);
	kgh_associate_finish(m_ctx, hp.data(), hp.size(), m_driver);
}

// ---- kinship ---------------------------------------------------------------------------------------
void MultipleKmersDataBases::kinship_begin(const size_t &min_count) {
	check(kg_kinship_begin(m_ctx, min_count, nullptr), "kg_kinship_begin");
	m_kinship_streaming = true;
}

void MultipleKmersDataBases::kinship_accumulate_loaded() {
	const size_t stride = 1 + m_hash_words_db_file;
	for (uint64_t off = 0; off < m_rows_loaded; off += kSubTileRows) {
		const uint64_t n = std::min<uint64_t>(kSubTileRows, m_rows_loaded - off);
		check(kg_kinship_submit(m_ctx, m_batch + off * stride, n), "kg_kinship_submit");
	}
	// the pinned batch buffer is about to be reused by the next load_kmers
	check(kg_sync(m_ctx), "kg_sync");
}

void MultipleKmersDataBases::kinship_finish(vector<vector<uint64_t> > &K, uint64_t &count) {
	vector<uint64_t> flat(m_accessions * m_accessions);
	uint64_t kept = 0;
	check(kg_kinship_fetch(m_ctx, flat.data(), &kept), "kg_kinship_fetch");
	for (size_t i = 0; i < m_accessions; i++)
		for (size_t j = 0; j < i; j++) K[i][j] += flat[i * m_accessions + j];
	count += kept;
	m_kinship_streaming = false;
}

void MultipleKmersDataBases::update_emma_kinshhip_calculation(vector<vector<uint64_t> > &K, uint64_t &count) const {
	MultipleKmersDataBases *self = const_cast<MultipleKmersDataBases *>(this);
	self->kinship_begin(m_load_mac);
	self->kinship_accumulate_loaded();
	self->kinship_finish(K, count);
}

// ---- PLINK output ------------------------------------------------------------------------------------
// load_kmers' squeeze (:125-132) for one row, on the host (only the few selected rows are written).
void MultipleKmersDataBases::squeeze_row(const uint64_t *file_row, vector<uint64_t> &mem_row) const {
	mem_row.assign(m_hash_words, 0);
	for (size_t col = 0; col < m_accessions; col++) {
		const uint64_t bit = (file_row[1 + m_map_word_index[col]] >> m_map_bit_index[col]) & 1ull;
		mem_row[col >> 6] |= bit << (col & 63);
	}
}

// .bim line + .bed record (:218-239): 2 bits per accession, 11 = present, 00 = absent
void MultipleKmersDataBases::write_PA(const string &name, const vector<uint64_t> &mem_row, BedBimFilesHandle &f) const {
	f.f_bim << "0\t" << name << "\t0\t0\t0\t1\n";
	const size_t n_bytes = (m_accessions + 3) / 4;
	string rec(n_bytes, '\0');
	for (size_t s = 0; s < m_accessions; s++)
		if ((mem_row[s >> 6] >> (s & 63)) & 1ull) rec[s >> 2] = (char)(rec[s >> 2] | (3u << (2 * (s & 3))));
	f.f_bed.write(rec.data(), (std::streamsize)rec.size());
}

size_t MultipleKmersDataBases::output_plink_bed_file(BedBimFilesHandle &f, const vector<AssociationOutputInfo> &kmer_list,
                                                     size_t index) const {
	const size_t stride = 1 + m_hash_words_db_file;
	vector<uint64_t> mem_row;
	while (index < kmer_list.size()) {
		const uint64_t row = std::get<2>(kmer_list[index]);
		if (row < m_row_offset) { index++; continue; }       // not expected: list is sorted by row
		if (row >= m_row_offset + m_rows_loaded) break;
		squeeze_row(m_batch + (row - m_row_offset) * stride, mem_row);
		write_PA(bits2kmer31(std::get<0>(kmer_list[index]), m_kmer_len) + "_" + std::to_string(std::get<1>(kmer_list[index])),
		         mem_row, f);
		index++;
	}
	return index;
}

void MultipleKmersDataBases::output_plink_bed_file_selected(BedBimFilesHandle &f, const vector<AssociationOutputInfo> &kmer_list) {
	const size_t stride = 1 + m_hash_words_db_file;
	vector<uint64_t> file_row(stride), mem_row;
	for (size_t i = 0; i < kmer_list.size(); i++) {
		const uint64_t row = std::get<2>(kmer_list[i]);
		read_fully(m_fd, file_row.data(), stride * 8, kHeaderBytes + row * stride * 8, m_table_path);
		squeeze_row(file_row.data(), mem_row);
		write_PA(bits2kmer31(std::get<0>(kmer_list[i]), m_kmer_len) + "_" + std::to_string(std::get<1>(kmer_list[i])), mem_row, f);
	}
}

// ---- whole-table PLINK conversion (kmers_table_to_bed) ---------------------------------------------
void MultipleKmersDataBases::mac_filter_loaded(const size_t &min_count, vector<uint8_t> &keep) const {
	keep.assign(m_rows_loaded, 0);
	const size_t stride = 1 + m_hash_words_db_file;
	for (uint64_t off = 0; off < m_rows_loaded; off += kSubTileRows) {
		const uint64_t n = std::min<uint64_t>(kSubTileRows, m_rows_loaded - off);
		check(kg_mac_filter(m_ctx, m_batch + off * stride, n, min_count, keep.data() + off, nullptr), "kg_mac_filter");
	}
}

void MultipleKmersDataBases::output_plink_loaded_row(BedBimFilesHandle &f, size_t r) const {
	const uint64_t *row = m_batch + r * (1 + m_hash_words_db_file);
	vector<uint64_t> mem_row;
	squeeze_row(row, mem_row);
	write_PA(bits2kmer31(row[0], m_kmer_len), mem_row, f);
}

uint64_t MultipleKmersDataBases::presence_absence_pattern_hash_loaded_row(size_t r) const {
	static const Hash64 hasher;
	vector<uint64_t> mem_row;
	squeeze_row(m_batch + r * (1 + m_hash_words_db_file), mem_row);
	uint64_t seed = 0;
	for (size_t w = 0; w < m_hash_words; w++) seed ^= hasher(mem_row[w]) + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2);
	return seed;
}

// ---- presence/absence pattern counter (:367-380), host implementation ("next" row of SURVEY 8(f)) -----
void MultipleKmersDataBases::update_presence_absence_pattern_counter(KmersSet &pa_pattern_counter) const {
	static const Hash64 hasher;
	const size_t stride = 1 + m_hash_words_db_file;
	vector<uint64_t> mem_row;
	for (uint64_t r = 0; r < m_rows_loaded; r++) {
		const uint64_t *row = m_batch + r * stride;
		uint64_t cnt = 0;
		for (size_t w = 0; w < m_hash_words_db_file; w++) cnt += (uint64_t)__builtin_popcountll(row[1 + w] & m_map_mask[w]);
		if (!(cnt >= m_load_mac && cnt <= m_accessions - m_load_mac)) continue;
		squeeze_row(row, mem_row);
		uint64_t seed = 0;
		for (size_t w = 0; w < m_hash_words; w++)
			seed ^= hasher(mem_row[w]) + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2);
		pa_pattern_counter.insert(seed);
	}
}
