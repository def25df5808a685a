"""Pins the CPU oracle (oracle/oracle.c) against the UNMODIFIED reference built into oracle/_ref/
(oracle/Makefile).  The reference ships no tests or golden vectors, so this
differential test -- plus the fixtures it generated under tests/golden/ -- is what pins parity.

CPU-only; skipped when oracle/_ref is absent."""
import numpy as np
import pytest

import support as S

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref not built")]


def _case(tmp_path, n_file, n_rows, n_pheno, seed, subset=None, maf=0.05, mac=5):
    names = [f"s{i}" for i in range(n_file)]
    table = S.synth_table(seed, n_rows, n_file)
    base = tmp_path / "t"
    S.write_table(base, table, n_file, names)
    used = names if subset is None else [names[i] for i in subset]
    y = S.synth_phenotypes(seed + 1, len(used), n_pheno)
    S.write_pheno(tmp_path / "p.tsv", used, y)
    mw, mb = S.column_map(names, used)
    mc = S.min_count_of(len(used), maf, mac)
    return names, used, table, y, mw, mb, mc, base


@pytest.mark.parametrize("n_file", [64, 65, 127, 128, 129, 241])
def test_scores_bit_exact(tmp_path, n_file):
    names, used, table, y, mw, mb, mc, base = _case(tmp_path, n_file, 3000, 3, seed=n_file)
    keep, scores, kept = S.oracle_scan(table, n_file, mw, mb, y, mc)
    ref, tested = S.ref_scores(base, 31, tmp_path / "p.tsv", mc, 1000, tmp_path / "o", 3)
    assert tested == kept
    kmers_kept = table[keep, 0]
    for j in range(3):
        assert len(ref[j]) == kept
        order = np.argsort(ref[j]["kmer"])
        r_k, r_s = ref[j]["kmer"][order], ref[j]["score"][order]
        assert np.array_equal(r_k, kmers_kept)  # table rows ascend by k-mer
        assert np.array_equal(r_s.view(np.uint64), scores[j][keep].view(np.uint64))


def test_scores_subset_permuted_columns(tmp_path):
    rng = np.random.default_rng(7)
    subset = rng.permutation(300)[:211].tolist()
    names, used, table, y, mw, mb, mc, base = _case(tmp_path, 300, 2500, 2, seed=11, subset=subset)
    keep, scores, kept = S.oracle_scan(table, 300, mw, mb, y, mc)
    ref, tested = S.ref_scores(base, 31, tmp_path / "p.tsv", mc, 700, tmp_path / "o", 2)
    assert tested == kept and 0 < kept < 2500
    for j in range(2):
        order = np.argsort(ref[j]["kmer"])
        assert np.array_equal(ref[j]["kmer"][order], table[keep, 0])
        assert np.array_equal(ref[j]["score"][order].view(np.uint64), scores[j][keep].view(np.uint64))


def test_scores_n1135(tmp_path):
    names, used, table, y, mw, mb, mc, base = _case(tmp_path, 1135, 1500, 2, seed=5)
    keep, scores, kept = S.oracle_scan(table, 1135, mw, mb, y, mc)
    ref, tested = S.ref_scores(base, 31, tmp_path / "p.tsv", mc, 4096, tmp_path / "o", 2)
    assert tested == kept
    for j in range(2):
        order = np.argsort(ref[j]["kmer"])
        assert np.array_equal(ref[j]["score"][order].view(np.uint64), scores[j][keep].view(np.uint64))


@pytest.mark.parametrize("n_file", [64, 129, 241])
def test_kinship_integer_exact(tmp_path, n_file):
    names = [f"s{i}" for i in range(n_file)]
    table = S.synth_table(100 + n_file, 1500, n_file)
    base = tmp_path / "t"
    S.write_table(base, table, n_file, names)
    mw, mb = S.column_map(names, names)
    import math
    mc = int(math.ceil(n_file * 0.05))
    K, cnt = S.oracle_kinship(table, n_file, mw, mb, mc)
    Kr, cntr = S.ref_kinship(base, 31, mc, 400, tmp_path / "k.bin")
    assert cnt == cntr and cnt > 0
    assert np.array_equal(K, Kr)


def test_heap_ties_and_cli_scores_file(tmp_path):
    """Top-K with many exact ties (duplicate patterns): the oracle heap replay must reproduce the
    reference CLI's .scores file byte for byte."""
    n_file, n_rows, kbest = 96, 4000, 37
    names = [f"s{i}" for i in range(n_file)]
    table = S.synth_table(3, n_rows, n_file)
    # force heavy ties: every row pattern drawn from only 50 distinct source rows
    src = table[np.arange(n_rows) % 50, 1:].copy()
    table[:, 1:] = src
    base = tmp_path / "t"
    S.write_table(base, table, n_file, names)
    y = S.synth_phenotypes(9, n_file, 2)
    S.write_pheno(tmp_path / "p.tsv", names, y)
    out = tmp_path / "out"
    out.mkdir()
    S.run_ref_associate(["-p", str(tmp_path / "p.tsv"), "-b", "r", "-o", str(out), "--kmers_table", str(base),
                         "-n", str(kbest), "--kmer_len", "31", "--k_mers_scores", "--batch_size", "1500",
                         "--parallel", "2"])
    mw, mb = S.column_map(names, names)
    mc = S.min_count_of(n_file, 0.05, 5)
    keep, scores, kept = S.oracle_scan(table, n_file, mw, mb, y, mc)
    assert int(open(out / "r.tested_kmers").read().split()[0]) == kept
    for j in range(2):
        h = S.oracle_topk(table, keep, scores[j], kbest)
        k, s, r = h.dump()
        raw = np.fromfile(out / f"r.{j}.best_kmers.scores", dtype=np.dtype([("kmer", "<u8"), ("score", "<f8")]))
        assert np.array_equal(raw["kmer"], k)
        assert np.array_equal(raw["score"].view(np.uint64), s.view(np.uint64))


@pytest.mark.parametrize("n_samples,subset", [(64, False), (131, True), (260, True)])
def test_snp_twin_selection_equals_reference(tmp_path, n_samples, subset):
    """oracle.c's SNP score restatement + the oracle heap select the SNPs the reference's associate_snps writes
    (zero-score ties included: every SNP goes through the heap, src/snps_multiple_databases.cpp:230-234)"""
    import subprocess
    if not (S.REF_DIR / "associate_snps").exists():
        pytest.skip("oracle/_ref/associate_snps not built")
    n_snps, n_best, mac = 1500, 40, 6.0
    base, bed, names = S.synth_plink(tmp_path, n_samples, n_snps, 17 + n_samples)
    rng = np.random.default_rng(2)
    idx = rng.permutation(n_samples)[: n_samples - 9] if subset else np.arange(n_samples)
    used = [names[i] for i in idx]
    y = S.synth_phenotypes(3, len(used), 2)
    S.write_pheno(tmp_path / "p.tsv", used, y)
    subprocess.run([str(S.REF_DIR / "associate_snps"), str(tmp_path / "p.tsv"), base, str(tmp_path / "best"), str(n_best), "0.01", str(mac)],
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for j, pname in enumerate(["phenotype_value", "P1"]):
        got = [int(l.split("\t")[1][3:]) for l in open(tmp_path / f"best.{pname}.bim")]
        scores = S.oracle_snp_scores(bed, idx // 4, (idx % 4) * 2, y[j], mac)
        h = S.OracleHeap(n_best)
        h.add_many(np.zeros(n_snps, dtype=np.uint64), scores, np.arange(n_snps, dtype=np.uint64))
        want = np.sort(h.dump()[2])
        assert got == want.tolist()
