"""Generates tests/golden/*.npz from the UNMODIFIED reference built in oracle/_ref/ (oracle/Makefile).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
Each fixture stores the generator parameters (the synthetic table is a pure function of them,
tests/support.synth_table), the float32 phenotypes, and the REFERENCE outputs:
  ref_kmers/ref_scores[p]   every kept row's k-mer id and f64 score (ref_harness scores)
  tested                    number of rows passing the MAC filter (.tested_kmers)
  kin / kin_cnt             raw u64 IBS counts of update_emma_kinshhip_calculation (ref_harness kinship),
                            or kin_sha256 + 64x64 corner for the large case
  top_kmers/top_scores[p]   the reference CLI's .scores file (heap pop order) for -n K
  bim/bed/fam               reference CLI output files for phenotype 0 (bytes)
  kin_tsv                   emma_kinship_kmers stdout
"""
import hashlib
import math
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import support as S  # noqa: E402


def make_case(name, n_file, n_rows, n_pheno, seed, kbest, subset=None, tie_patterns=0, kin=True,
              maf=0.05, mac=5, batch=700):
    assert S.have_ref()
    names = [f"s{i}" for i in range(n_file)]
    table = S.synth_table(seed, n_rows, n_file)
    if tie_patterns:
        table[:, 1:] = table[np.arange(n_rows) % tie_patterns, 1:].copy()
    used = names if subset is None else [names[i] for i in subset]
    y = S.synth_phenotypes(seed + 1, len(used), n_pheno)
    mc = S.min_count_of(len(used), maf, mac)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        base = td / "t"
        S.write_table(base, table, n_file, names)
        S.write_pheno(td / "p.tsv", used, y)
        ref, tested = S.ref_scores(base, 31, td / "p.tsv", mc, batch, td / "o", n_pheno)
        rk, rs = [], []
        for j in range(n_pheno):
            order = np.argsort(ref[j]["kmer"])
            rk.append(ref[j]["kmer"][order])
            rs.append(ref[j]["score"][order])
        out["ref_kmers"] = np.stack(rk)
        out["ref_scores"] = np.stack(rs)
        out["tested"] = np.int64(tested)
        # CLI run with a bounded heap
        od = td / "out"
        od.mkdir()
        S.run_ref_associate(["-p", str(td / "p.tsv"), "-b", "r", "-o", str(od), "--kmers_table", str(base),
                             "-n", str(kbest), "--kmer_len", "31", "--k_mers_scores", "--batch_size",
                             str(batch), "--parallel", "2", "--maf", str(maf), "--mac", str(mac)])
        tk, ts = [], []
        for j in range(n_pheno):
            raw = np.fromfile(od / f"r.{j}.best_kmers.scores", dtype=np.dtype([("kmer", "<u8"), ("score", "<f8")]))
            tk.append(raw["kmer"])
            ts.append(raw["score"])
        out["top_kmers"] = np.stack(tk)
        out["top_scores"] = np.stack(ts)
        out["cli_tested"] = np.int64(int(open(od / "r.tested_kmers").read().split()[0]))
        pn = "phenotype_value"
        out["bim"] = np.frombuffer((od / f"r.0.{pn}.bim").read_bytes(), dtype=np.uint8)
        out["bed"] = np.frombuffer((od / f"r.0.{pn}.bed").read_bytes(), dtype=np.uint8)
        out["fam"] = np.frombuffer((od / f"r.0.{pn}.fam").read_bytes(), dtype=np.uint8)
        if kin is not None:
            mck = int(math.ceil(n_file * maf))
            K, cnt = S.ref_kinship(base, 31, mck, batch, td / "k.bin")
            out["kin_cnt"] = np.int64(cnt)
            out["kin_min_count"] = np.int64(mck)
            if kin:
                out["kin"] = K
                r = S.run_ref_kinship_cli(["-t", str(base), "-k", "31", "--maf", str(maf)])
                out["kin_tsv"] = np.frombuffer(r.stdout, dtype=np.uint8)
            else:
                out["kin_sha256"] = np.frombuffer(hashlib.sha256(K.tobytes()).digest(), dtype=np.uint8)
                out["kin_corner"] = K[:64, :64].copy()
    out.update(dict(n_file=np.int64(n_file), n_rows=np.int64(n_rows), n_pheno=np.int64(n_pheno),
                    seed=np.int64(seed), kbest=np.int64(kbest), maf=np.float64(maf), mac=np.int64(mac),
                    min_count=np.int64(mc), batch=np.int64(batch), tie_patterns=np.int64(tie_patterns),
                    subset=np.array(subset if subset is not None else [], dtype=np.int64), y=y))
    path = S.GOLDEN_DIR / f"{name}.npz"
    np.savez_compressed(path, **out)
    print(f"{name}: rows={n_rows} N_file={n_file} used={len(used)} kept={tested} -> {path.stat().st_size} B")


def main():
    make_case("identity_n131", 131, 1500, 3, seed=21, kbest=50)
    rng = np.random.default_rng(7)
    make_case("subset_n300", 300, 1200, 2, seed=22, kbest=64, subset=rng.permutation(300)[:211].tolist())
    make_case("ties_n96", 96, 4000, 2, seed=23, kbest=37, tie_patterns=50, batch=1500)
    make_case("plumbing_n64", 64, 5000, 1, seed=24, kbest=100)
    make_case("thaliana_n1135", 1135, 600, 2, seed=25, kbest=40, kin=False, batch=4096)


if __name__ == "__main__":
    main()
