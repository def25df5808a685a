"""Test support: ctypes binding of the CPU oracle (oracle/liboracle.so), synthetic table /
phenotype writers in the reference's on-disk formats, and runners for the
unmodified reference build in oracle/_ref/ (present in the build container and shipped to the GPU
box; tests that need it skip when it is absent).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"
GOLDEN_DIR = ROOT / "tests" / "golden"

_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


_oracle = None


def oracle():
    """Load (building if needed) oracle/liboracle.so."""
    global _oracle
    if _oracle is not None:
        return _oracle
    so = ORACLE_DIR / "liboracle.so"
    src = ORACLE_DIR / "oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(ORACLE_DIR), "oracle"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(str(so))
    lib.kgo_permute_scores.argtypes = [_f32p, C.c_size_t, _f32p]
    lib.kgo_update_scores_and_sum.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p]
    lib.kgo_update_scores_and_sum.restype = C.c_float
    lib.kgo_scan_scores.argtypes = [_u64p, C.c_uint64, C.c_size_t, _u32p, _u32p, C.c_size_t,
                                    _f32p, C.c_size_t, C.c_uint64, _u8p, _f64p]
    lib.kgo_scan_scores.restype = C.c_uint64
    lib.kgo_kinship.argtypes = [_u64p, C.c_uint64, C.c_size_t, _u32p, _u32p, C.c_size_t,
                                C.c_uint64, _u64p, _u64p]
    lib.kgo_kinship.restype = C.c_uint64
    lib.kgo_heap_new.argtypes = [C.c_uint64]
    lib.kgo_heap_new.restype = C.c_void_p
    lib.kgo_heap_free.argtypes = [C.c_void_p]
    lib.kgo_heap_add.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_uint64]
    lib.kgo_heap_size.argtypes = [C.c_void_p]
    lib.kgo_heap_size.restype = C.c_uint64
    lib.kgo_heap_insertions.argtypes = [C.c_void_p]
    lib.kgo_heap_insertions.restype = C.c_uint64
    lib.kgo_heap_dump.argtypes = [C.c_void_p, _u64p, _f64p, _u64p]
    lib.kgo_synth_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _u64p]
    lib.kgo_snp_scores.argtypes = [_u8p, C.c_uint64, C.c_size_t, _u32p, _u32p, C.c_size_t, _f32p, C.c_double, _f64p]
    _oracle = lib
    return lib


# ----------------------------------------------------------------------------- shapes
def w_file_of(n_file: int) -> int:
    return (n_file + 63) // 64


def w_mem_of(n: int) -> int:
    return 2 * ((n + 127) // 128)


def min_count_of(n: int, maf: float, mac: int) -> int:
    """associate_kmers.cpp:99-102"""
    import math
    return max(int(math.ceil(float(n) * maf)), mac)


# ----------------------------------------------------------------------------- synthetic data
def synth_table(seed: int, n_rows: int, n_file: int, first_row: int = 0) -> np.ndarray:
    """[n_rows, 1 + W_file] uint64: column 0 = k-mer id, rest = presence words."""
    out = np.zeros((n_rows, 1 + w_file_of(n_file)), dtype=np.uint64)
    if n_rows:
        oracle().kgo_synth_rows(seed, first_row, n_rows, n_file, _ptr(out, _u64p))
    return out


def synth_phenotypes(seed: int, n: int, n_pheno: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    y = rng.standard_normal((n_pheno, n)).astype(np.float32)
    # values must survive "%.9g" -> stof exactly (they do for float32)
    return np.ascontiguousarray(y)


def column_map(names_file: list[str], names_used: list[str]):
    idx = np.array([names_file.index(n) for n in names_used], dtype=np.int64)
    return (idx // 64).astype(np.uint32), (idx % 64).astype(np.uint32)


# ----------------------------------------------------------------------------- file formats
def write_table(base: str | os.PathLike, rows: np.ndarray, n_file: int, names: list[str], k: int = 31):
    """<base>.table / <base>.names (kmers_merge_multiple_databaes.cpp:54-73)."""
    base = str(base)
    assert rows.dtype == np.uint64 and rows.shape[1] == 1 + w_file_of(n_file)
    with open(base + ".table", "wb") as f:
        f.write(struct.pack("<IQI", 0xDDCCBBAA, n_file, k))
        f.write(np.ascontiguousarray(rows).tobytes())
    with open(base + ".names", "w") as f:
        for n in names:
            f.write(n + "\n")


def write_pheno(path: str | os.PathLike, names: list[str], y: np.ndarray, pheno_names=None):
    """Phenotype TSV (kmer_general.cpp:175-205). y: [P, N] float32."""
    p = y.shape[0]
    pheno_names = pheno_names or (["phenotype_value"] + [f"P{i}" for i in range(1, p)])
    with open(path, "w") as f:
        f.write("accession_id\t" + "\t".join(pheno_names) + "\n")
        for i, n in enumerate(names):
            f.write(n + "\t" + "\t".join("%.9g" % float(v) for v in y[:, i]) + "\n")


# ----------------------------------------------------------------------------- oracle wrappers
def oracle_scan(table: np.ndarray, n_file: int, map_word, map_bit, y: np.ndarray, min_count: int):
    """-> keep[n_rows] bool, scores[P, n_rows] float64 (0 where not kept), kept count."""
    n_rows = table.shape[0]
    n = len(map_word)
    p = y.shape[0]
    keep = np.zeros(n_rows, dtype=np.uint8)
    scores = np.zeros((p, n_rows), dtype=np.float64)
    table = np.ascontiguousarray(table)
    y = np.ascontiguousarray(y, dtype=np.float32)
    mw = np.ascontiguousarray(map_word, dtype=np.uint32)
    mb = np.ascontiguousarray(map_bit, dtype=np.uint32)
    kept = oracle().kgo_scan_scores(_ptr(table, _u64p), n_rows, w_file_of(n_file), _ptr(mw, _u32p),
                                    _ptr(mb, _u32p), n, _ptr(y, _f32p), p, min_count,
                                    _ptr(keep, _u8p), _ptr(scores, _f64p))
    return keep.astype(bool), scores, int(kept)


def oracle_kinship(table: np.ndarray, n_file: int, map_word, map_bit, min_count: int):
    n_rows = table.shape[0]
    n = len(map_word)
    K = np.zeros((n, n), dtype=np.uint64)
    cnt = C.c_uint64(0)
    table = np.ascontiguousarray(table)
    mw = np.ascontiguousarray(map_word, dtype=np.uint32)
    mb = np.ascontiguousarray(map_bit, dtype=np.uint32)
    oracle().kgo_kinship(_ptr(table, _u64p), n_rows, w_file_of(n_file), _ptr(mw, _u32p),
                         _ptr(mb, _u32p), n, min_count, _ptr(K, _u64p), C.byref(cnt))
    return K, int(cnt.value)


class OracleHeap:
    """BestAssociationsHeap restatement (oracle.c)."""

    def __init__(self, max_results: int):
        self._lib = oracle()
        self._h = self._lib.kgo_heap_new(max_results)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.kgo_heap_free(self._h)
            self._h = None

    def add(self, kmer: int, score: float, row: int):
        self._lib.kgo_heap_add(self._h, int(kmer), float(score), int(row))

    def add_many(self, kmers, scores, rows):
        add = self._lib.kgo_heap_add
        h = self._h
        for k, s, r in zip(kmers.tolist(), scores.tolist(), rows.tolist()):
            add(h, k, s, r)

    @property
    def insertions(self) -> int:
        return int(self._lib.kgo_heap_insertions(self._h))

    def dump(self):
        n = int(self._lib.kgo_heap_size(self._h))
        k = np.zeros(n, dtype=np.uint64)
        s = np.zeros(n, dtype=np.float64)
        r = np.zeros(n, dtype=np.uint64)
        if n:
            self._lib.kgo_heap_dump(self._h, _ptr(k, _u64p), _ptr(s, _f64p), _ptr(r, _u64p))
        return k, s, r


def oracle_topk(table, keep, scores_p, k_best):
    """Replay one phenotype's kept rows through the oracle heap, in row order.
    Row id = file row index (any monotone id gives the same outputs, SURVEY.md 7 hard part 5)."""
    h = OracleHeap(k_best)
    rows = np.nonzero(keep)[0]
    h.add_many(table[rows, 0], scores_p[rows], rows.astype(np.uint64))
    return h


# ----------------------------------------------------------------------------- reference runners
def have_ref() -> bool:
    return (REF_DIR / "associate_kmers").exists() and (REF_DIR / "ref_harness").exists()


def ref_scores(table_base, klen, pheno_path, min_count, batch_rows, out_prefix, n_pheno):
    """Run oracle/_ref/ref_harness scores -> list over phenotypes of dict kmer->score, tested."""
    subprocess.check_call([str(REF_DIR / "ref_harness"), "scores", str(table_base), str(klen),
                           str(pheno_path), str(min_count), str(batch_rows), str(out_prefix)],
                          stderr=subprocess.DEVNULL)
    res = []
    for j in range(n_pheno):
        raw = np.fromfile(f"{out_prefix}.{j}.scores", dtype=np.dtype([("kmer", "<u8"), ("score", "<f8")]))
        res.append(raw)
    tested = int(open(f"{out_prefix}.tested").read().split()[0])
    return res, tested


def ref_kinship(table_base, klen, min_count, batch_rows, out_path):
    subprocess.check_call([str(REF_DIR / "ref_harness"), "kinship", str(table_base), str(klen),
                           str(min_count), str(batch_rows), str(out_path)], stderr=subprocess.DEVNULL)
    raw = np.fromfile(out_path, dtype="<u8")
    n, cnt = int(raw[0]), int(raw[1])
    return raw[2:].reshape(n, n), cnt


def run_ref_associate(args: list[str], cwd=None):
    return subprocess.run([str(REF_DIR / "associate_kmers")] + args, cwd=cwd, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, check=True)


def run_ref_kinship_cli(args: list[str]):
    return subprocess.run([str(REF_DIR / "emma_kinship_kmers")] + args, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, check=True)


# ----------------------------------------------------------------------------- golden fixtures
GOLDEN_CASES = ["identity_n131", "subset_n300", "ties_n96", "plumbing_n64", "thaliana_n1135"]


class Golden:
    """A tests/golden/<name>.npz fixture (made by tests/golden/make_golden.py from oracle/_ref)."""

    def __init__(self, name: str):
        z = np.load(GOLDEN_DIR / f"{name}.npz")
        self.z = z
        self.name = name
        self.n_file = int(z["n_file"])
        self.n_rows = int(z["n_rows"])
        self.n_pheno = int(z["n_pheno"])
        self.seed = int(z["seed"])
        self.kbest = int(z["kbest"])
        self.maf = float(z["maf"])
        self.mac = int(z["mac"])
        self.min_count = int(z["min_count"])
        self.batch = int(z["batch"])
        self.y = np.ascontiguousarray(z["y"], dtype=np.float32)
        self.names = [f"s{i}" for i in range(self.n_file)]
        subset = z["subset"].tolist()
        self.used = self.names if not subset else [self.names[i] for i in subset]
        self.table = synth_table(self.seed, self.n_rows, self.n_file)
        tp = int(z["tie_patterns"])
        if tp:
            self.table[:, 1:] = self.table[np.arange(self.n_rows) % tp, 1:].copy()
        self.map_word, self.map_bit = column_map(self.names, self.used)

    def write_inputs(self, d):
        d = Path(d)
        write_table(d / "t", self.table, self.n_file, self.names)
        write_pheno(d / "p.tsv", self.used, self.y)
        return d / "t", d / "p.tsv"


# ----------------------------------------------------------------------------- SNP twin / table construction
def synth_plink(d, n_samples: int, n_snps: int, seed: int):
    """Random PLINK bed/bim/fam triple with all four genotype codes (00 a/a, 01 missing, 10 A/a, 11 A/A).
    -> (base path, bed payload uint8 [n_snps, bytes_per_snp], sample names)"""
    rng = np.random.default_rng(seed)
    names = [f"s{i}" for i in range(n_samples)]
    bps = (n_samples + 3) // 4
    g = rng.choice(4, size=(n_snps, bps * 4), p=[0.45, 0.05, 0.15, 0.35]).astype(np.uint8)
    g[:, n_samples:] = 0
    # rarer alleles for some SNPs so that the minor-allele-count test rejects them (zero scores: heap ties)
    rare = rng.random(n_snps) < 0.2
    g[rare] = np.where(rng.random((int(rare.sum()), bps * 4)) < 0.97, 0, g[rare])
    bed = (g[:, 0::4] | (g[:, 1::4] << 2) | (g[:, 2::4] << 4) | (g[:, 3::4] << 6)).astype(np.uint8)
    base = str(Path(d) / "snps")
    with open(base + ".bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        f.write(np.ascontiguousarray(bed).tobytes())
    with open(base + ".fam", "w") as f:
        for n in names:
            f.write(f"{n} {n} 0 0 0 -9\n")
    with open(base + ".bim", "w") as f:
        for i in range(n_snps):
            f.write(f"1\tsnp{i}\t0\t{100 + i}\tA\tC\n")
    return base, bed, names


def oracle_snp_scores(bed: np.ndarray, map_byte, map_shift, y1: np.ndarray, mac: float) -> np.ndarray:
    bed = np.ascontiguousarray(bed, dtype=np.uint8)
    mb = np.ascontiguousarray(map_byte, dtype=np.uint32)
    ms = np.ascontiguousarray(map_shift, dtype=np.uint32)
    y1 = np.ascontiguousarray(y1, dtype=np.float32)
    out = np.zeros(bed.shape[0], dtype=np.float64)
    oracle().kgo_snp_scores(_ptr(bed, _u8p), bed.shape[0], bed.shape[1], _ptr(mb, _u32p), _ptr(ms, _u32p), len(mb),
                            _ptr(y1, _f32p), float(mac), _ptr(out, _f64p))
    return out


def synth_kmer_lists(d, n_acc: int, n_all: int, seed: int):
    """Sorted k-mer list files for build_kmers_table: all_kmers (sorted, unique, < 2^62) and per accession a sorted
    subset + some k-mers that are not in all_kmers, with random strand flags in the two top bits.
    -> (list file, all_kmers file, names, all k-mers, membership [n_all, n_acc] bool)"""
    rng = np.random.default_rng(seed)
    pool = np.unique(rng.integers(0, 1 << 62, size=int(n_all * 1.3), dtype=np.uint64))
    rng.shuffle(pool)
    all_k = np.sort(pool[:n_all])
    extra = pool[n_all:]
    member = rng.random((len(all_k), n_acc)) < rng.random(n_acc)[None, :]
    d = Path(d)
    all_k.tofile(d / "all_kmers")
    names = [f"acc{i}" for i in range(n_acc)]
    with open(d / "list.txt", "w") as f:
        for a in range(n_acc):
            own = np.concatenate([all_k[member[:, a]], extra[rng.random(len(extra)) < 0.1]])
            if len(own) == 0:
                own = extra[:1]
            own = np.sort(own)
            flags = rng.integers(0, 4, size=len(own), dtype=np.uint64) << np.uint64(62)
            (own | flags).tofile(d / f"kmers_{a}")
            f.write(f"{d / ('kmers_%d' % a)}\t{names[a]}\n")
    return d / "list.txt", d / "all_kmers", names, all_k, member
