// associate_kmers -- associate k-mer presence/absence patterns with phenotypes (B200 build).
//
// Same flags, defaults and output files as the reference CLI (/root/reference/src/associate_kmers.cpp:
// flags :38-53, outputs :155-205).  Differences underneath:
//   * pass 1 (the hot loop :123-148) scores all phenotypes of a batch in one GPU pass
//     (MultipleKmersDataBases::add_kmers_to_heaps) instead of one CTPL task per phenotype;
//     --parallel is accepted and ignored;
//   * pass 2 does not re-stream the table: only the selected rows are read back for the PLINK files;
//   * the best-K heaps live on the GPU (kg_select_*: exact libstdc++ heap replay on the device) whenever their capacity
//     fits shared memory (the pipeline's -n 10001 does); larger capacities use the host replay path;
//   * new flags: --gpus G (row-shard the table over G GPUs of this box, exact merge), --device D,
//     --engine {0 auto, 1 exact, 2 tensor filter + exact refine}, --select {auto, device, host}.
#include <cmath>
#include <exception>
#include <iostream>
#include <memory>
#include <thread>

#include "association_driver.h"
#include "best_associations_heap.h"
#include "cli_options.h"
#include <cstdlib>

#include "kmer_general.h"
#include "kmers_multiple_databases.h"

using namespace std;

namespace {
struct Shard {
	unique_ptr<MultipleKmersDataBases> db;
	vector<BestAssociationsHeap> heaps;   // host replay path: local heaps (only supply thresholds when sharded)
	AssociationDriverState state;
	exception_ptr error;
	bool overflow = false;                // device path: a round overflowed (DeviceSelectionOverflow)
};

// Pass 1 of one shard through the DEVICE heaps (kg_select_*): batches are submitted asynchronously while the tile
// reader loads ahead; nothing comes back to the host until the end.  Shards other than the first warm-start from
// the first prefix_rows rows of the table and log what their heaps admit (merged on shard 0 afterwards).
void run_shard_device(Shard &sh, size_t min_count, size_t batch_size, bool verbose, bool count_patterns,
                      uint64_t first, uint64_t count, uint64_t prefix_rows, bool is_first_shard) {
	try {
		MultipleKmersDataBases &db = *sh.db;
		// this thread, its reader threads and the pinned tiles they fill live on the GPU's NUMA node (no-op on one node)
		kg_bind_host_to_device(db.device(), nullptr);
		// Nothing comes back per batch on this path, so --batch_size only bounds memory: the table is streamed in tiles of at
		// most 256 MB (three pinned buffers: one being read from the file, one in flight to the GPU, one being scanned)
		const size_t row_bytes = 8 * (1 + db.file_words());
		batch_size = min<size_t>(batch_size, max<size_t>(65536, (256u << 20) / row_bytes));
		if (!is_first_shard) {
			db.restrict_to_rows(0, prefix_rows);
			while (db.load_kmers(batch_size, min_count)) db.add_loaded_kmers_to_device_heaps();
			db.device_selection_log_reset();
		}
		db.restrict_to_rows(first, count);
		if (count_patterns) db.pattern_counter_begin(count, min_count);   // this shard's own rows only (not the shared prefix)
		double t0 = get_time(), t1;
		size_t batch_index = 0;
		while (db.load_kmers(batch_size, min_count)) {
			t1 = get_time();
			if (verbose) cerr << "Load [" << batch_index << "]\t" << (t1 - t0) / 60. << "min" << endl;
			t0 = get_time();
			db.add_loaded_kmers_to_device_heaps();
			t1 = get_time();
			if (verbose) cerr << "Associations [" << batch_index << "]\t" << (t1 - t0) / 60. << "min" << endl;
			t0 = get_time();
			batch_index++;
		}
	} catch (const DeviceSelectionOverflow &) {
		sh.overflow = true;
	} catch (...) {
		sh.error = current_exception();
	}
}

// Pass 1 of one shard through the host replay path (any heap capacity): device candidates -> BestAssociationsHeap.
void run_shard(Shard &sh, const vector<vector<float> > &y, size_t min_count, size_t batch_size, bool verbose,
               KmersSet *pattern_counter) {
	try {
		vector<BestAssociationsHeap *> hp(sh.heaps.size());
		for (size_t j = 0; j < hp.size(); j++) hp[j] = &sh.heaps[j];
		vector<float> flat;
		for (const auto &v : y) flat.insert(flat.end(), v.begin(), v.end());
		kg_ctx *ctx = sh.db->context();
		if (kg_scan_set_phenotypes(ctx, flat.data(), (uint32_t)y.size(), min_count) != KG_OK)
			throw runtime_error(string("kg_scan_set_phenotypes: ") + kg_last_error(ctx));
		double t0 = get_time(), t1;
		size_t batch_index = 0;
		while (sh.db->load_kmers(batch_size, min_count)) {
			t1 = get_time();
			if (verbose) cerr << "Load [" << batch_index << "]\t" << (t1 - t0) / 60. << "min" << endl;
			t0 = get_time();
			if (pattern_counter) sh.db->update_presence_absence_pattern_counter(*pattern_counter);
			kgh_associate_rows(ctx, hp.data(), hp.size(), sh.db->loaded_rows(), sh.db->rows_loaded(), sh.db->row_offset(),
			                   1 + sh.db->file_words(), sh.state);
			// the batch buffer is reused by the reader: drain the round still in flight
			kgh_associate_finish(ctx, hp.data(), hp.size(), sh.state);
			t1 = get_time();
			if (verbose) cerr << "Associations [" << batch_index << "]\t" << (t1 - t0) / 60. << "min" << endl;
			t0 = get_time();
			batch_index++;
		}
	} catch (...) {
		sh.error = current_exception();
	}
}
}  // namespace

int main(int argc, char *argv[]) {
	CliOptions options("associate_kmers", "Associate k-mers presence/absence pattern with a phenotype of interest");
	options.add('p', "phenotype_file", "phenotype file name");
	options.add('b', "base_name", "base name to use for all files");
	options.add('o', "output_dir", "where to save output files", false, ".");
	options.add(0, "kmers_table", "Presence/absemce k-mer file");
	options.add('n', "best", "Number of best k-mers to report", false, "1000000");
	options.add(0, "first_phenotype_best", "if provided will save a different number of k-mers for the first phenotype");
	options.add(0, "batch_size", "Loading only part of the presence absence info to memory", false, "10000000");
	options.add(0, "parallel", "Max number of threads to use (ignored: phenotypes are batched on the GPU)", false, "4");
	options.add(0, "kmer_len", "Length of the k-mers");
	options.add(0, "maf", "Minor allele frequency", false, "0.05");
	options.add(0, "mac", "Minor allele count", false, "5");
	options.add(0, "k_mers_scores", "output the best k_mers scores in binary format", true);
	options.add(0, "pattern_counter", "Count the number of unique presence/absence patterns", true);
	options.add(0, "gpus", "Number of GPUs to shard the table over", false, "1");
	options.add(0, "device", "First CUDA device ordinal", false, "0");
	options.add(0, "engine", "Scan engine: 0 auto, 1 exact, 2 tensor filter + exact refine", false, "0");
	options.add(0, "select", "Where the best-K heaps live: auto (device when every capacity fits), device, host", false, "auto");
	options.add(0, "help", "print help", true);
	try {
		options.parse(argc, argv);
		if (options.count("help")) {
			cerr << options.help() << endl;
			return 0;
		}
		for (const char *req : {"phenotype_file", "base_name", "kmers_table", "kmer_len"})
			if (!options.count(req)) throw CliOptions::ParseError(string("Option '") + req + "' is required");

		const string fn_base = options.str("output_dir") + "/" + options.str("base_name");
		const size_t heap_size = options.as<size_t>("best");
		const size_t batch_size = options.as<size_t>("batch_size");
		(void)options.as<size_t>("parallel");
		const uint32_t kmer_length = options.as<uint32_t>("kmer_len");
		if ((kmer_length > 31) || (kmer_length < 10)) {
			cerr << "kmer length has to be between 10-31" << endl;
			return 1;
		}
		const double maf = options.as<double>("maf");
		const size_t mac = options.as<size_t>("mac");
		const size_t n_gpus = max<size_t>(1, options.as<size_t>("gpus"));
		const int device0 = options.as<int>("device");
		const int engine = options.as<int>("engine");
		const string select_mode = options.str("select");
		if (select_mode != "auto" && select_mode != "device" && select_mode != "host")
			throw CliOptions::ParseError("--select must be auto, device or host");
		const string table = options.str("kmers_table");

		const double t_start = get_time();
		const bool phase_times = getenv("KMERSGWAS_PHASE_TIMES") != nullptr;
		auto phase = [&](const char *what) { if (phase_times) cerr << "[phase] " << what << "\t" << get_time() - t_start << " s" << endl; };
		// phenotypes (reference :81-88)
		pair<vector<string>, vector<PhenotypeList> > phenotypes_info = load_phenotypes_file(options.str("phenotype_file"));
		const size_t phenotypes_n = phenotypes_info.first.size();
		if (phenotypes_n == 0) throw logic_error("no phenotype columns in " + options.str("phenotype_file"));
		for (size_t i = 0; i < phenotypes_n; i++)
			phenotypes_info.second[i] = intersect_phenotypes_to_present_DBs(phenotypes_info.second[i], table, true);
		const vector<PhenotypeList> &p_list = phenotypes_info.second;
		vector<vector<float> > y(phenotypes_n);
		for (size_t j = 0; j < phenotypes_n; j++) y[j] = p_list[j].second;

		// heaps (reference :92-96)
		vector<BestAssociationsHeap> k_heap;
		if (options.count("first_phenotype_best")) k_heap.resize(1, BestAssociationsHeap(options.as<size_t>("first_phenotype_best")));
		k_heap.resize(phenotypes_n, BestAssociationsHeap(heap_size));

		// effective minor allele count (reference :99-103)
		const size_t n_accessions = p_list[0].first.size();
		size_t min_count = (size_t)ceil(static_cast<double>(n_accessions) * maf);
		if (min_count < mac) min_count = mac;
		cerr << "Effective minor allele count:\t" << min_count << endl;

		KmersSet pa_patterns_counter;
		const bool count_patterns = options.count("pattern_counter") > 0;

		phase("phenotypes loaded");
		// ---- pass 1: association scan ----------------------------------------------------------
		vector<unique_ptr<Shard> > shard_ptrs(n_gpus);
		for (auto &sp : shard_ptrs) sp.reset(new Shard());
		auto shards = [&](size_t g) -> Shard & { return *shard_ptrs[g]; };
		for (size_t g = 0; g < n_gpus; g++) {
			// KMERSGWAS_SHARDS_ON_ONE_DEVICE=1 (tests): every shard gets its own context on the same GPU
			const bool one_device = getenv("KMERSGWAS_SHARDS_ON_ONE_DEVICE") != nullptr;
			MultipleKmersDataBases::set_device(one_device ? device0 : device0 + (int)g);
			shards(g).db.reset(new MultipleKmersDataBases(table, p_list[0].first, kmer_length));
			shards(g).db->set_scan_engine(engine);
		}
		const uint64_t total_rows = shards(0).db->rows_in_file();
		vector<size_t> capacities(phenotypes_n);
		for (size_t j = 0; j < phenotypes_n; j++) capacities[j] = k_heap[j].capacity();

		// device heaps (kg_select_*) when every capacity fits them (the pipeline's -n 10001 does); else host replay
		phase("contexts created");
		bool device_heaps = select_mode != "host";
		if (device_heaps)
			for (size_t g = 0; g < n_gpus && device_heaps; g++)
				device_heaps = shards(g).db->begin_device_selection(capacities, y, min_count, g > 0);
		if (!device_heaps && select_mode == "device") throw logic_error("--select device: a heap capacity does not fit the device heaps");
		bool done = false, have_device_patterns = false;
		uint64_t device_patterns = 0;
		if (device_heaps) {
			const uint64_t prefix_rows = min<uint64_t>(total_rows / n_gpus, max<uint64_t>(1, (uint64_t)8 << 20));
			vector<thread> threads;
			for (size_t g = 0; g < n_gpus; g++) {
				const uint64_t first = total_rows * g / n_gpus, last = total_rows * (g + 1) / n_gpus;
				threads.emplace_back(run_shard_device, ref(shards(g)), min_count, batch_size, g == 0, count_patterns,
				                     first, last - first, prefix_rows, g == 0);
			}
			for (auto &t : threads) t.join();
			bool overflow = false;
			for (auto &sp : shard_ptrs) {
				if (sp->error) rethrow_exception(sp->error);
				overflow = overflow || sp->overflow;
			}
			try {
				// exact merge: shard 0's heaps are the sequential heaps of its block; the later shards' logs are replayed
				// through them in row order (kg_select_replay)
				for (size_t g = 1; g < n_gpus && !overflow; g++) {
					vector<uint64_t> off, ent;
					uint64_t rows = 0, kept = 0;
					shards(g).db->device_selection_log(off, ent, rows, kept);
					shards(0).db->device_selection_replay(off, ent, rows, kept);
				}
				if (!overflow) {
					if (count_patterns) {
						// distinct patterns of the whole table: union of the shards' device sets (by hash key)
						for (size_t g = 1; g < n_gpus; g++) {
							vector<uint64_t> keys;
							shards(g).db->pattern_counter_export(keys);
							shards(0).db->pattern_counter_insert(keys);
						}
						device_patterns = shards(0).db->pattern_counter_size();
						have_device_patterns = true;
					}
					shards(0).db->finish_device_selection(k_heap);
					done = true;
				}
			} catch (const DeviceSelectionOverflow &) {
				overflow = true;
			}
			if (overflow) {
				cerr << "device heaps: a candidate segment overflowed (scores rising along the table); re-running pass 1 through the host replay path" << endl;
				pa_patterns_counter.clear();
				for (size_t g = 0; g < n_gpus; g++) {
					MultipleKmersDataBases::set_device(shards(g).db->device());
					shard_ptrs[g].reset(new Shard());
					shards(g).db.reset(new MultipleKmersDataBases(table, p_list[0].first, kmer_length));
					shards(g).db->set_scan_engine(engine);
				}
			}
		}
		if (!done && n_gpus == 1) {
			Shard &sh = shards(0);
			sh.heaps.swap(k_heap);
			run_shard(sh, y, min_count, batch_size, true, count_patterns ? &pa_patterns_counter : nullptr);
			if (sh.error) rethrow_exception(sh.error);
			k_heap.swap(sh.heaps);
		} else if (!done) {
			if (count_patterns) throw logic_error("--pattern_counter with --gpus > 1 needs the device heaps (--select auto|device and -n <= ~14000)");
			vector<thread> threads;
			for (size_t g = 0; g < n_gpus; g++) {
				const uint64_t first = total_rows * g / n_gpus, last = total_rows * (g + 1) / n_gpus;
				shards(g).db->restrict_to_rows(first, last - first);
				shards(g).heaps = k_heap;  // same capacities, empty
				shards(g).state.log_hits = true;
				threads.emplace_back(run_shard, ref(shards(g)), cref(y), min_count, batch_size, g == 0, nullptr);
			}
			for (auto &t : threads) t.join();
			vector<AssociationDriverState *> states;
			for (auto &sp : shard_ptrs) {
				if (sp->error) rethrow_exception(sp->error);
				states.push_back(&sp->state);
			}
			vector<BestAssociationsHeap *> hp(phenotypes_n);
			for (size_t j = 0; j < phenotypes_n; j++) hp[j] = &k_heap[j];
			kgh_merge_shards(states, hp.data(), hp.size());
		}
		phase("pass 1 done (heaps final)");
		const uint64_t n_patterns = have_device_patterns ? device_patterns : (uint64_t)pa_patterns_counter.size();
		if (count_patterns) cerr << "Total patterns\t" << n_patterns << endl;

		// ---- outputs (reference :155-205) --------------------------------------------------------
		vector<kmers_output_list> best_kmers;
		for (size_t j = 0; j < phenotypes_n; j++) {
			if (options.count("k_mers_scores"))
				k_heap[j].output_to_file_with_scores(fn_base + "." + std::to_string(j) + ".best_kmers.scores");
			best_kmers.push_back(k_heap[j].get_kmers_for_output(kmer_length));
			k_heap[j].empty_heap();
		}
		phase("heaps read out");
		MultipleKmersDataBases &db0 = *shards(0).db;
		{
			// the phenotypes' PLINK triples are independent: written concurrently (the reference writes them from one pass 2)
			const unsigned n_thr = (unsigned)min<size_t>(phenotypes_n, max(1u, kgh_host_threads()));
			vector<thread> writers;
			vector<exception_ptr> werr(n_thr);
			for (unsigned t = 0; t < n_thr; t++)
				writers.emplace_back([&, t] {
					try {
						for (size_t j = t; j < phenotypes_n; j += n_thr) {
							const string base = fn_base + "." + std::to_string(j) + "." + phenotypes_info.first[j];
							BedBimFilesHandle handle(base);
							write_fam_file(p_list[j], base + ".fam");
							db0.output_plink_bed_file_selected(handle, best_kmers[j].list);
							handle.close();
						}
					} catch (...) {
						werr[t] = current_exception();
					}
				});
			for (auto &w : writers) w.join();
			for (auto &e : werr) if (e) rethrow_exception(e);
			for (size_t j = 0; j < phenotypes_n; j++) cerr << ".";
		}
		cerr << endl;
		if (count_patterns) {
			ofstream fout(fn_base + ".pattern_counter");
			fout << n_patterns << endl;
		}
		phase("PLINK files written");
		ofstream fout(fn_base + ".tested_kmers");
		fout << k_heap[0].number_of_insertion() << endl;
		fout.close();
	} catch (const CliOptions::ParseError &e) {
		cerr << "error parsing options: " << e.what() << endl;
		cerr << options.help() << endl;
		return 1;
	} catch (const std::exception &e) {
		cerr << "associate_kmers: " << e.what() << endl;
		return 2;
	}
	return 0;
}
