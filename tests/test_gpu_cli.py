"""GPU tests of the two CLIs (kmersgwas_b200/bin/associate_kmers, emma_kinship_kmers): same flags, same input files
and BYTE-IDENTICAL output files as the unmodified reference binaries (oracle/_ref, built from /root/reference by
oracle/Makefile and shipped to the GPU box)."""
import filecmp
import subprocess
from pathlib import Path

import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "kmersgwas_b200" / "bin"


@pytest.fixture(scope="module")
def bins(gpu_device):
    if not S.have_ref():
        pytest.skip("oracle/_ref not built")
    for exe in ("associate_kmers", "emma_kinship_kmers"):
        if not (BIN / exe).exists():
            pytest.skip("CLI binaries not built")
    return BIN


def _run(exe, args, cwd=None):
    return subprocess.run([str(exe)] + [str(a) for a in args], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def _same_dir(a: Path, b: Path):
    fa, fb = sorted(p.name for p in a.iterdir()), sorted(p.name for p in b.iterdir())
    assert fa == fb, (fa, fb)
    for n in fa:
        assert filecmp.cmp(a / n, b / n, shallow=False), f"{n} differs"
    return fa


@pytest.mark.parametrize("name,extra", [
    ("plumbing_n64", []),
    ("identity_n131", ["--k_mers_scores"]),
    ("subset_n300", ["--k_mers_scores", "--pattern_counter"]),
    ("ties_n96", ["--k_mers_scores"]),
    ("thaliana_n1135", ["--k_mers_scores", "--engine", "2"]),
])
def test_associate_kmers_outputs_byte_identical(bins, tmp_path, name, extra):
    g = S.Golden(name)
    table, pheno = g.write_inputs(tmp_path)
    outs = {}
    for tag, exe in (("ref", S.REF_DIR / "associate_kmers"), ("ours", bins / "associate_kmers")):
        out = tmp_path / tag
        out.mkdir()
        args = ["-p", pheno, "-b", "run", "-o", out, "--kmers_table", table, "-n", g.kbest, "--kmer_len", 31,
                "--maf", g.maf, "--mac", g.mac, "--batch_size", g.batch, "--parallel", 2]
        args += [a for a in extra if tag == "ours" or a not in ("--engine", "2")]
        r = _run(exe, args)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = (out, r)
    files = _same_dir(outs["ref"][0], outs["ours"][0])
    assert any(f.endswith(".bed") for f in files) and "run.tested_kmers" in files
    # the same progress vocabulary on stderr (timers differ)
    for key in ("Effective minor allele count", "Load [0]", "Associations [0]"):
        assert key in outs["ours"][1].stderr


@pytest.mark.parametrize("select", ["device", "host"])
def test_associate_kmers_first_phenotype_best_and_small_batches(bins, tmp_path, select):
    """both homes of the best-K heaps -- the device (kg_select_*) and the host replay path -- give the reference's files"""
    g = S.Golden("identity_n131")
    table, pheno = g.write_inputs(tmp_path)
    dirs = []
    for tag, exe in (("ref", S.REF_DIR / "associate_kmers"), ("ours", bins / "associate_kmers")):
        out = tmp_path / tag
        out.mkdir()
        r = _run(exe, ["-p", pheno, "-b", "x", "-o", out, "--kmers_table", table, "-n", 40, "--first_phenotype_best", 90,
                       "--kmer_len", 31, "--maf", 0.1, "--mac", 3, "--batch_size", 777, "--k_mers_scores"]
                 + (["--select", select] if tag == "ours" else []))
        assert r.returncode == 0, r.stderr[-2000:]
        dirs.append(out)
    _same_dir(*dirs)


def test_associate_kmers_default_capacity_uses_host_heaps(bins, tmp_path):
    """the binary's default -n 1000000 does not fit the device heaps: --select auto falls back to the host replay path,
    --select device refuses"""
    g = S.Golden("plumbing_n64")
    table, pheno = g.write_inputs(tmp_path)
    dirs = []
    for tag, exe in (("ref", S.REF_DIR / "associate_kmers"), ("ours", bins / "associate_kmers")):
        out = tmp_path / tag
        out.mkdir()
        r = _run(exe, ["-p", pheno, "-b", "x", "-o", out, "--kmers_table", table, "--kmer_len", 31, "--k_mers_scores"])
        assert r.returncode == 0, r.stderr[-2000:]
        dirs.append(out)
    _same_dir(*dirs)
    r = _run(bins / "associate_kmers", ["-p", pheno, "-b", "x", "-o", tmp_path, "--kmers_table", table, "--kmer_len", 31, "--select", "device"])
    assert r.returncode != 0


def test_associate_kmers_errors_like_reference(bins, tmp_path):
    g = S.Golden("plumbing_n64")
    table, pheno = g.write_inputs(tmp_path)
    # --help: exit 0; unknown flag / missing required flag / bad k-mer length: non-zero exit, nothing written
    assert _run(bins / "associate_kmers", ["--help"]).returncode == 0
    assert _run(bins / "associate_kmers", ["--nonsense"]).returncode != 0
    assert _run(bins / "associate_kmers", ["-p", pheno, "-b", "x", "--kmers_table", table]).returncode != 0
    r = _run(bins / "associate_kmers", ["-p", pheno, "-b", "x", "-o", tmp_path, "--kmers_table", table, "--kmer_len", 9])
    assert r.returncode == 1 and "kmer length has to be between 10-31" in r.stderr
    # wrong k in the table header, unknown accession in the phenotype file
    r = _run(bins / "associate_kmers", ["-p", pheno, "-b", "x", "-o", tmp_path, "--kmers_table", table, "--kmer_len", 25])
    assert r.returncode != 0
    bad = tmp_path / "bad.tsv"
    bad.write_text(Path(pheno).read_text().replace("s3\t", "nobody\t"))
    r = _run(bins / "associate_kmers", ["-p", bad, "-b", "x", "-o", tmp_path, "--kmers_table", table, "--kmer_len", 31])
    assert r.returncode != 0


@pytest.mark.parametrize("name,maf", [("plumbing_n64", 0.05), ("identity_n131", 0.1), ("thaliana_n1135", 0.05)])
def test_emma_kinship_kmers_stdout_identical(bins, tmp_path, name, maf):
    g = S.Golden(name)
    table, _ = g.write_inputs(tmp_path)
    args = ["-t", table, "-k", 31, "--maf", maf]
    ref = _run(S.REF_DIR / "emma_kinship_kmers", args)
    ours = _run(bins / "emma_kinship_kmers", args)
    assert ref.returncode == 0 and ours.returncode == 0, ours.stderr[-2000:]
    assert ours.stdout == ref.stdout
    assert len(ours.stdout.splitlines()) == g.n_file


def test_associate_kmers_two_shards_equal_one(bins, tmp_path):
    """--gpus 2 on one device (two contexts): sharded scan + exact merge == single scan."""
    import os
    import torch
    g = S.Golden("subset_n300")
    table, pheno = g.write_inputs(tmp_path)
    dirs = []
    if torch.cuda.device_count() < 2:
        os.environ["KMERSGWAS_SHARDS_ON_ONE_DEVICE"] = "1"
    for tag, extra in (("one", []), ("two", ["--gpus", 2]), ("three", ["--gpus", 3]), ("two_host", ["--gpus", 2, "--select", "host"])):
        out = tmp_path / tag
        out.mkdir()
        if tag in ("three", "two_host"):
            os.environ["KMERSGWAS_SHARDS_ON_ONE_DEVICE"] = "1"
        r = _run(bins / "associate_kmers", ["-p", pheno, "-b", "x", "-o", out, "--kmers_table", table, "-n", g.kbest,
                                           "--kmer_len", 31, "--maf", g.maf, "--mac", g.mac, "--batch_size", 500,
                                           "--k_mers_scores"] + extra + ([] if tag == "two_host" else ["--pattern_counter"]))
        assert r.returncode == 0, r.stderr[-2000:]
        dirs.append(out)
    os.environ.pop("KMERSGWAS_SHARDS_ON_ONE_DEVICE", None)
    for d in dirs[1:3]:
        _same_dir(dirs[0], d)              # incl. x.pattern_counter: the shards' device pattern sets are merged by key
    (dirs[0] / "x.pattern_counter").unlink()
    _same_dir(dirs[0], dirs[3])
    # and the single-GPU count is the reference's
    ref = tmp_path / "ref"
    ref.mkdir()
    r = _run(S.REF_DIR / "associate_kmers", ["-p", pheno, "-b", "x", "-o", ref, "--kmers_table", table, "-n", g.kbest, "--kmer_len", 31,
                                            "--maf", g.maf, "--mac", g.mac, "--batch_size", 500, "--k_mers_scores", "--pattern_counter"])
    assert r.returncode == 0
    assert (ref / "x.pattern_counter").read_text() == (dirs[1] / "x.pattern_counter").read_text()


@pytest.mark.parametrize("name,batch,unique,rows_per_load", [
    ("plumbing_n64", 1000000, False, 4194304),     # one batch
    ("subset_n300", 700, False, 333),              # several batch files, phenotype-order subset of the columns,
                                                   # device passes that do not line up with the batch boundaries
    ("ties_n96", 50, True, 64),                    # -u: duplicate presence/absence patterns dropped across batches
    ("identity_n131", 1, False, 1000),             # one kept row per file
])
def test_kmers_table_to_bed_outputs_byte_identical(bins, tmp_path, name, batch, unique, rows_per_load):
    """table -> PLINK conversion of every row that passes the MAC filter (device MAC filter)."""
    if not (bins / "kmers_table_to_bed").exists() or not (S.REF_DIR / "kmers_table_to_bed").exists():
        pytest.skip("kmers_table_to_bed not built")
    g = S.Golden(name)
    if batch == 1:
        g.table = g.table[:300]                    # one file triple per kept row: keep the directory small
    table, pheno = g.write_inputs(tmp_path)
    dirs = []
    for tag, exe in (("ref", S.REF_DIR / "kmers_table_to_bed"), ("ours", bins / "kmers_table_to_bed")):
        out = tmp_path / tag
        out.mkdir()
        args = ["-t", table, "-k", 31, "-p", pheno, "--maf", g.maf, "--mac", g.mac, "-b", batch, "-o", out / "conv"]
        if unique:
            args.append("-u")
        if tag == "ours":
            args += ["--rows_per_load", rows_per_load]
        r = _run(exe, args)
        assert r.returncode == 0, r.stderr[-2000:]
        dirs.append(out)
    files = _same_dir(*dirs)
    assert "conv.0.bed" in files and "conv.0.fam" in files



# ------------------------------------------------------------------------------ SNP twin, table construction
@pytest.mark.parametrize("n_samples,n_snps,n_pheno,subset", [(64, 500, 1, False), (131, 3000, 4, True), (300, 2000, 3, True)])
def test_associate_snps_outputs_byte_identical(bins, tmp_path, n_samples, n_snps, n_pheno, subset):
    """SNP twin of the scan: same best-N SNP selection (incl. the zero-score ties of SNPs that fail the MAC test) and the
    same PLINK files as the reference's associate_snps"""
    if not (bins / "associate_snps").exists() or not (S.REF_DIR / "associate_snps").exists():
        pytest.skip("associate_snps not built")
    base, bed, names = S.synth_plink(tmp_path, n_samples, n_snps, 40 + n_samples)
    rng = np.random.default_rng(3)
    used = [names[i] for i in (rng.permutation(n_samples)[: n_samples - 17] if subset else range(n_samples))]
    y = S.synth_phenotypes(9, len(used), n_pheno)
    S.write_pheno(tmp_path / "p.tsv", used, y)
    dirs = []
    for tag, exe in (("ref", S.REF_DIR / "associate_snps"), ("ours", bins / "associate_snps")):
        out = tmp_path / tag
        out.mkdir()
        r = _run(exe, [tmp_path / "p.tsv", base, out / "best", 57, 0.05, 5])
        assert r.returncode == 0, r.stderr[-2000:]
        dirs.append(out)
    files = _same_dir(*dirs)
    assert len(files) == 2 * n_pheno
    assert (dirs[1] / files[0]).stat().st_size > 0
    # bit-identical scores against the C restatement of the reference (oracle.c), all phenotypes in one device pass
    import kmersgwas_b200._abi as abi
    idx = np.array([names.index(u) for u in used])
    mb, ms = idx // 4, (idx % 4) * 2
    got = abi.snps_scores(bed, mb, ms, y, 7.0)
    for j in range(n_pheno):
        want = S.oracle_snp_scores(bed, mb, ms, y[j], 7.0)
        assert np.array_equal(got[j].view(np.uint64), want.view(np.uint64))
    assert (got == 0).any() and (got > 0).any()


@pytest.mark.parametrize("n_acc,n_all,rng_rows", [(3, 2000, 4194304), (70, 30000, 7001), (130, 5000, 1)])
def test_build_kmers_table_byte_identical(bins, tmp_path, n_acc, n_all, rng_rows):
    """table construction from sorted k-mer lists: .table / .names byte-identical to the reference's build_kmers_table"""
    if not (bins / "build_kmers_table").exists() or not (S.REF_DIR / "build_kmers_table").exists():
        pytest.skip("build_kmers_table not built")
    if rng_rows == 1:
        n_all = 300
    lst, allk, names, all_k, member = S.synth_kmer_lists(tmp_path, n_acc, n_all, 70 + n_acc)
    outs = []
    for tag, exe in (("ref", S.REF_DIR / "build_kmers_table"), ("ours", bins / "build_kmers_table")):
        out = tmp_path / tag
        out.mkdir()
        args = ["-l", lst, "-k", 31, "-a", allk, "-o", out / "t"]
        if tag == "ours":
            args += ["--range_kmers", rng_rows]
        r = _run(exe, args)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(out)
    _same_dir(*outs)
    # and the table says what the generator put in
    raw = np.fromfile(outs[1] / "t.table", dtype=np.uint8)
    rows = raw[16:].view(np.uint64).reshape(len(all_k), 1 + (n_acc + 63) // 64)
    assert np.array_equal(rows[:, 0], all_k)
    bits = np.unpackbits(np.ascontiguousarray(rows[:, 1:]).view(np.uint8), axis=1, bitorder="little")[:, :n_acc]
    assert np.array_equal(bits.astype(bool), member)


def test_associate_kmers_device_overflow_falls_back_to_host_replay(bins, tmp_path):
    """scores that rise along the table make every row a candidate: a round overflows its device candidate segment, the
    CLI notices at the end (DeviceSelectionOverflow) and re-runs pass 1 through the host replay path -- same files"""
    n_file, n_rows = 64, 300000
    names = [f"s{i}" for i in range(n_file)]
    table = S.synth_table(123, n_rows, n_file)
    y = S.synth_phenotypes(124, n_file, 1)
    idx = np.arange(n_file)
    mc = S.min_count_of(n_file, 0.05, 5)
    keep, scores, kept = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
    order = np.argsort(scores[0], kind="stable")          # ascending scores: every kept row beats the heap minimum
    table = np.ascontiguousarray(table[order])
    S.write_table(tmp_path / "t", table, n_file, names)
    S.write_pheno(tmp_path / "p.tsv", names, y)
    dirs = []
    for tag, exe in (("ref", S.REF_DIR / "associate_kmers"), ("ours", bins / "associate_kmers")):
        out = tmp_path / tag
        out.mkdir()
        r = _run(exe, ["-p", tmp_path / "p.tsv", "-b", "x", "-o", out, "--kmers_table", tmp_path / "t", "-n", 100, "--kmer_len", 31,
                       "--k_mers_scores", "--pattern_counter"])
        assert r.returncode == 0, r.stderr[-2000:]
        if tag == "ours":
            assert "re-running pass 1 through the host replay path" in r.stderr
        dirs.append(out)
    _same_dir(*dirs)


def test_emma_kinship_kmers_two_gpus_allreduce(bins, tmp_path):
    """--gpus 2: the shards' u64 accumulators are summed by one NCCL all-reduce inside the library
    (kg_comm_init_all + kg_kinship_allreduce_all); the matrix printed is the reference's.  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = S.Golden("thaliana_n1135")
    table, _ = g.write_inputs(tmp_path)
    args = ["-t", table, "-k", 31, "--maf", 0.05]
    ref = _run(S.REF_DIR / "emma_kinship_kmers", args)
    ours = _run(bins / "emma_kinship_kmers", args + ["--gpus", 2])
    assert ref.returncode == 0 and ours.returncode == 0, ours.stderr[-2000:]
    assert ours.stdout == ref.stdout
