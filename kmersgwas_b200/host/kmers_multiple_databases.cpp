// kmers_multiple_databases.cpp -- see kmers_multiple_databases.h.
#include "kmers_multiple_databases.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <thread>

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

// A batch is read by several threads, each pread-ing its own slice: one thread copies ~5 GB/s out of the page cache,
// the PCIe link behind the pinned buffer takes ~55 GB/s.
void read_parallel(int fd, void *dst, size_t bytes, uint64_t offset, const string &path) {
	const size_t kSlice = 16u << 20;
	unsigned n_thr = (unsigned)std::min<size_t>((bytes + kSlice - 1) / kSlice, std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
	if (const char *e = getenv("KMERSGWAS_READ_THREADS")) n_thr = std::max(1, atoi(e));
	if (n_thr <= 1) { read_fully(fd, dst, bytes, offset, path); return; }
	std::vector<std::thread> thr;
	std::vector<std::exception_ptr> err(n_thr);
	const size_t per = ((bytes + n_thr - 1) / n_thr + 4095) & ~(size_t)4095;
	for (unsigned t = 0; t < n_thr; t++) {
		const size_t b = std::min(bytes, (size_t)t * per), e = std::min(bytes, b + per);
		if (b == e) continue;
		thr.emplace_back([=, &err, &path] {
			try { read_fully(fd, static_cast<char *>(dst) + b, e - b, offset + b, path); } catch (...) { err[t] = std::current_exception(); }
		});
	}
	for (auto &t : thr) t.join();
	for (auto &e : err) if (e) std::rethrow_exception(e);
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
      m_batch(nullptr), m_cur(0), m_prefetch_pending(false), m_prefetch_batch(0), m_prefetch_first(0),
      m_device_selection(false), m_rows_submitted_sel(0), m_rows_loaded(0), m_load_mac(0),
      m_ctx(nullptr), m_pheno_min_cnt(0), m_kinship_streaming(false) {
	for (int b = 0; b < kBuffers; b++) { m_buf[b] = nullptr; m_buf_cap[b] = 0; m_buf_ticket[b] = 0; m_buf_busy[b] = false; }
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
	m_device = s_device;
	kg_status st = kg_ctx_create(s_device, &shape, nullptr, &m_ctx);
	if (st != KG_OK) throw std::runtime_error(string("kg_ctx_create: ") + kg_last_error(nullptr));
}

MultipleKmersDataBases::~MultipleKmersDataBases() {
	try { cancel_prefetch(); } catch (...) {}
	if (m_ctx) kg_sync(m_ctx);
	for (int b = 0; b < kBuffers; b++) if (m_buf[b]) kg_host_free(m_ctx, m_buf[b]);
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
	cancel_prefetch();
	if (m_file_rows == 0) m_file_rows = m_kmer_number;
	m_kmer_number = m_file_rows;
	if (first > m_kmer_number) first = m_kmer_number;
	if (count > m_kmer_number - first) count = m_kmer_number - first;
	m_first_row = first;
	m_kmer_number = first + count;  // exclusive end, see load_kmers
	m_kmer_loaded = first;
	m_row_offset = first;
	m_rows_loaded = 0;
}

// read file rows [first_row, first_row + n_rows) into pinned buffer b (waiting until the device is done with it)
uint64_t MultipleKmersDataBases::read_batch_into(int b, uint64_t first_row, uint64_t n_rows) {
	const size_t stride = 1 + m_hash_words_db_file;
	if (m_buf_busy[b]) {
		check(kg_stream_wait(m_ctx, m_buf_ticket[b]), "kg_stream_wait");
		m_buf_busy[b] = false;
	}
	if (m_buf_cap[b] < n_rows) {
		if (m_buf[b]) kg_host_free(m_ctx, m_buf[b]);
		m_buf[b] = nullptr;
		m_buf_cap[b] = 0;
		void *p = nullptr;
		check(kg_host_alloc(m_ctx, (size_t)n_rows * stride * 8, &p), "kg_host_alloc");
		m_buf[b] = static_cast<uint64_t *>(p);
		m_buf_cap[b] = n_rows;
	}
	read_parallel(m_fd, m_buf[b], (size_t)n_rows * stride * 8, kHeaderBytes + first_row * stride * 8, m_table_path);
	return n_rows;
}

void MultipleKmersDataBases::start_prefetch(uint64_t batch_size) {
	const uint64_t left = m_kmer_number - m_kmer_loaded;
	if (left == 0 || getenv("KMERSGWAS_NO_PREFETCH")) return;
	const uint64_t n = std::min<uint64_t>(batch_size, left);
	const int b = (m_cur + 1) % kBuffers;
	m_prefetch_batch = batch_size;
	m_prefetch_first = m_kmer_loaded;
	m_prefetch = std::async(std::launch::async, [this, b, n] { return read_batch_into(b, m_prefetch_first, n); });
	m_prefetch_pending = true;
}

void MultipleKmersDataBases::cancel_prefetch() {
	if (m_prefetch_pending) {
		m_prefetch_pending = false;
		m_prefetch.get();
	}
}

// Same contract as the reference (:103-108, 145): true with a (possibly short) batch, false once the file is exhausted.
// The rows are RAW file rows in a pinned buffer; the next batch is already being read in the background.
bool MultipleKmersDataBases::load_kmers(const uint64_t &batch_size, const size_t &mac) {
	m_row_offset = m_kmer_loaded;
	m_rows_loaded = 0;
	m_load_mac = mac;
	const uint64_t left = m_kmer_number - m_kmer_loaded;
	if (left == 0) { cancel_prefetch(); return false; }
	const uint64_t n = std::min<uint64_t>(batch_size, left);
	const int b = (m_cur + 1) % kBuffers;
	uint64_t got;
	if (m_prefetch_pending && m_prefetch_batch == batch_size && m_prefetch_first == m_kmer_loaded) {
		m_prefetch_pending = false;
		got = m_prefetch.get();
	} else {
		cancel_prefetch();
		got = read_batch_into(b, m_kmer_loaded, n);
	}
	m_cur = b;
	m_batch = m_buf[b];
	m_kmer_loaded += got;
	m_rows_loaded = got;
	start_prefetch(batch_size);
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

// ---- device-resident heaps --------------------------------------------------------------------------
bool MultipleKmersDataBases::begin_device_selection(const vector<size_t> &capacities, const vector<vector<float> > &scores,
                                                    const size_t &min_cnt, bool log_admissions) {
	if (capacities.size() != scores.size()) throw std::logic_error("heaps and phenotypes differ in number");
	ensure_phenotypes(scores, min_cnt);
	vector<uint64_t> kb(capacities.begin(), capacities.end());
	const kg_status st = kg_select_begin(m_ctx, kb.data(), (uint32_t)kb.size(), log_admissions ? KG_SELECT_LOG : 0);
	if (st == KG_ERR_INVALID) return false;   // a heap does not fit the device: the caller uses the host replay path
	check(st, "kg_select_begin");
	m_device_selection = true;
	m_rows_submitted_sel = 0;
	return true;
}

void MultipleKmersDataBases::add_loaded_kmers_to_device_heaps() {
	if (!m_device_selection) throw std::logic_error("add_loaded_kmers_to_device_heaps without begin_device_selection");
	if (m_load_mac != m_pheno_min_cnt)
		throw std::logic_error("GPU path: load_kmers' minor allele count and the selection's min_cnt must be equal");
	if (m_rows_loaded == 0) return;
	check(kg_scan_submit(m_ctx, m_batch, m_rows_loaded, m_row_offset), "kg_scan_submit");
	// the reader may overwrite this buffer once everything submitted so far has left it
	check(kg_stream_mark(m_ctx, &m_buf_ticket[m_cur]), "kg_stream_mark");
	m_buf_busy[m_cur] = true;
	m_rows_submitted_sel += m_rows_loaded;
}

void MultipleKmersDataBases::device_selection_log_reset() {
	check(kg_select_log_reset(m_ctx), "kg_select_log_reset");
	m_rows_submitted_sel = 0;
}

static void throw_if_overflow(kg_ctx *ctx, kg_status st, const char *what) {
	if (st == KG_ERR_HITS_OVERFLOW) throw DeviceSelectionOverflow(string(what) + ": " + kg_last_error(ctx));
	if (st != KG_OK) throw std::runtime_error(string(what) + ": " + kg_last_error(ctx));
}

void MultipleKmersDataBases::device_selection_log(vector<uint64_t> &offsets, vector<uint64_t> &entries, uint64_t &rows, uint64_t &kept) {
	uint64_t applied = 0;
	throw_if_overflow(m_ctx, kg_select_sync(m_ctx, &applied, &kept), "kg_select_sync");
	rows = m_rows_submitted_sel;
	const size_t P = m_pheno_flat.size() / m_accessions;
	vector<uint64_t> counts(P);
	check(kg_select_log_counts(m_ctx, counts.data()), "kg_select_log_counts");
	offsets.assign(P + 1, 0);
	for (size_t p = 0; p < P; p++) offsets[p + 1] = offsets[p] + counts[p];
	entries.assign((size_t)offsets[P] * 3, 0);
	check(kg_select_log_export(m_ctx, entries.data(), offsets.data()), "kg_select_log_export");
}

void MultipleKmersDataBases::device_selection_replay(const vector<uint64_t> &offsets, const vector<uint64_t> &entries, uint64_t rows, uint64_t kept) {
	check(kg_select_replay(m_ctx, entries.data(), offsets.data(), rows, kept), "kg_select_replay");
}

void MultipleKmersDataBases::finish_device_selection(vector<BestAssociationsHeap> &heaps) {
	uint64_t applied = 0, kept = 0;
	throw_if_overflow(m_ctx, kg_select_sync(m_ctx, &applied, &kept), "kg_select_sync");
	const size_t P = heaps.size();
	vector<uint64_t> state(kg_select_state_len(m_ctx));
	check(kg_select_export(m_ctx, state.data()), "kg_select_export");
	const uint32_t kmax = kg_select_kmax(m_ctx);
	for (size_t p = 0; p < P; p++) {
		const uint64_t *hdr = state.data() + 4 * p;
		const uint64_t *ent = state.data() + 4 * P + (size_t)p * kmax * 3;
		heaps[p].load_layout(ent, (size_t)hdr[0], (size_t)kept, (size_t)hdr[1], (size_t)hdr[2]);
	}
	check(kg_select_end(m_ctx), "kg_select_end");
	m_device_selection = false;
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

// The selected rows of one phenotype -> its .bim / .bed pair, without re-streaming the table: each row is read with one
// pread, squeezed (memcpy when the column map is the identity) and expanded to PLINK's 2 bits per sample with a byte
// table (8 presence bits -> 2 output bytes); both files are assembled in memory and written once.  Thread-safe
// (const, pread), so the CLI writes the phenotypes' files concurrently.
void MultipleKmersDataBases::output_plink_bed_file_selected(BedBimFilesHandle &f, const vector<AssociationOutputInfo> &kmer_list) const {
	static const struct Lut {
		uint16_t v[256];
		Lut() {
			for (unsigned b = 0; b < 256; b++) {
				uint16_t o = 0;
				for (unsigned k = 0; k < 8; k++)
					if (b & (1u << k)) o = (uint16_t)(o | (3u << (2 * k)));   // 11 = present, 00 = absent (:218-239)
				v[b] = o;
			}
		}
	} lut;
	const size_t stride = 1 + m_hash_words_db_file;
	const size_t n_bytes = (m_accessions + 3) / 4;
	bool identity = m_accessions == m_accessions_db_file;
	for (size_t i = 0; identity && i < m_accessions; i++)
		identity = (size_t)m_map_word_index[i] * WLEN + m_map_bit_index[i] == i;
	vector<uint64_t> file_row(stride), mem_row;
	string bed(kmer_list.size() * n_bytes, '\0'), bim;
	bim.reserve(kmer_list.size() * (m_kmer_len + 24));
	vector<uint8_t> rec(2 * 8 * m_hash_words + 16);
	for (size_t i = 0; i < kmer_list.size(); i++) {
		const uint64_t row = std::get<2>(kmer_list[i]);
		read_fully(m_fd, file_row.data(), stride * 8, kHeaderBytes + row * stride * 8, m_table_path);
		const uint64_t *words;
		if (identity) {
			words = file_row.data() + 1;
		} else {
			squeeze_row(file_row.data(), mem_row);
			words = mem_row.data();
		}
		const size_t n_words = identity ? m_hash_words_db_file : m_hash_words;
		const uint8_t *src = reinterpret_cast<const uint8_t *>(words);
		for (size_t b = 0; b < n_words * 8; b++) {
			const uint16_t o = lut.v[src[b]];
			rec[2 * b] = (uint8_t)(o & 0xFF);
			rec[2 * b + 1] = (uint8_t)(o >> 8);
		}
		// samples beyond N are absent in the squeezed row; in the raw identity row the file has no such columns either
		char *dst = &bed[i * n_bytes];
		memcpy(dst, rec.data(), n_bytes);
		if (m_accessions % 4) dst[n_bytes - 1] = (char)(dst[n_bytes - 1] & ((1u << (2 * (m_accessions % 4))) - 1u));
		bim += "0\t";
		bim += bits2kmer31(std::get<0>(kmer_list[i]), m_kmer_len);
		bim += '_';
		bim += std::to_string(std::get<1>(kmer_list[i]));
		bim += "\t0\t0\t0\t1\n";
	}
	f.f_bim.write(bim.data(), (std::streamsize)bim.size());
	f.f_bed.write(bed.data(), (std::streamsize)bed.size());
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

// ---- presence/absence pattern counter on the device ------------------------------------------------------
void MultipleKmersDataBases::pattern_counter_begin(uint64_t max_rows, const size_t &min_count) {
	check(kg_patterns_begin(m_ctx, 0), "kg_patterns_begin");
	check(kg_patterns_attach(m_ctx, min_count, max_rows), "kg_patterns_attach");
}

uint64_t MultipleKmersDataBases::pattern_counter_size() {
	uint64_t n = 0;
	check(kg_patterns_count(m_ctx, &n, nullptr), "kg_patterns_count");
	return n;
}

void MultipleKmersDataBases::pattern_counter_export(vector<uint64_t> &keys) {
	uint64_t n = 0;
	check(kg_patterns_export(m_ctx, nullptr, 0, &n), "kg_patterns_export");
	keys.assign(n, 0);
	if (n) check(kg_patterns_export(m_ctx, keys.data(), n, &n), "kg_patterns_export");
}

void MultipleKmersDataBases::pattern_counter_insert(const vector<uint64_t> &keys) {
	check(kg_patterns_insert(m_ctx, keys.data(), keys.size()), "kg_patterns_insert");
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
