#!/usr/bin/env python
"""Wall-clock of the associate_kmers CLI, ours vs the unmodified reference (oracle/_ref), on one synthetic table file
(BASELINE configs[1] shape, --rows rows; page-cached).  Run on the GPU box:
    python profiles/cli_wallclock.py --rows 10000000 > gpurun_out/cli_wallclock.json
Outputs are compared byte for byte (scores, PLINK files, tested_kmers)."""
import argparse
import filecmp
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import support as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--ref-rows", type=int, default=1_000_000, help="rows the reference CLI is run on (it is ~1000x slower; 0 = skip)")
    ap.add_argument("--batch-size", type=int, default=10_000_000)
    args = ap.parse_args()
    n, p = bench.N_SAMPLES, bench.N_PHENO
    out = {"shape": f"{args.rows} rows x {n} samples x {p} phenotypes, -n {bench.K_BEST}"}
    with tempfile.TemporaryDirectory(prefix="kgcli_", dir=os.environ.get("KG_TMPDIR")) as td:
        td = Path(td)
        t0 = time.perf_counter()
        names = [f"s{i}" for i in range(n)]
        chunk = 1_000_000
        base = td / "t"
        with open(str(base) + ".table", "wb") as f:
            import struct
            f.write(struct.pack("<IQI", 0xDDCCBBAA, n, 31))
            for r0 in range(0, args.rows, chunk):
                f.write(S.synth_table(bench.SEED_TABLE, min(chunk, args.rows - r0), n, first_row=r0).tobytes())
        with open(str(base) + ".names", "w") as f:
            f.write("\n".join(names) + "\n")
        S.write_pheno(td / "p.tsv", names, bench.phenotypes(n, p))
        out["generate_s"] = time.perf_counter() - t0
        out["table_bytes"] = os.path.getsize(str(base) + ".table")
        common = ["-p", str(td / "p.tsv"), "-b", "run", "--kmers_table", str(base), "-n", str(bench.K_BEST), "--kmer_len", "31",
                  "--maf", str(bench.MAF), "--mac", str(bench.MAC), "--k_mers_scores"]
        runs = {}
        for tag, exe, extra in (("ours", ROOT / "kmersgwas_b200" / "bin" / "associate_kmers", ["--batch_size", str(args.batch_size)]),
                                ("ours_again", ROOT / "kmersgwas_b200" / "bin" / "associate_kmers", ["--batch_size", str(args.batch_size)])):
            o = td / tag
            o.mkdir()
            t0 = time.perf_counter()
            r = subprocess.run([str(exe)] + common + ["-o", str(o)] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            wall = time.perf_counter() - t0
            assert r.returncode == 0, r.stderr[-1000:]
            runs[tag] = {"wall_s": wall, "rows_per_s": args.rows / wall, "gb_per_s": out["table_bytes"] / wall / 1e9}
        out["ours"] = runs["ours_again"]          # second run: file in the page cache, CUDA context creation still included
        out["ours_first_run"] = runs["ours"]
        if args.ref_rows:
            # the reference on a prefix of the same table (its cost is linear in rows), with our CLI on the same prefix for the byte comparison
            nref = min(args.ref_rows, args.rows)
            sub = td / "sub"
            with open(str(base) + ".table", "rb") as f, open(str(sub) + ".table", "wb") as g:
                g.write(f.read(16 + nref * 8 * (1 + (n + 63) // 64)))
            (td / "sub.names").write_text((td / "t.names").read_text())
            sub_common = [a if a != str(base) else str(sub) for a in common]
            dirs = {}
            for tag, exe, extra in (("ref", ROOT / "oracle" / "_ref" / "associate_kmers", ["--parallel", str(os.cpu_count() or 1)]),
                                    ("ours_sub", ROOT / "kmersgwas_b200" / "bin" / "associate_kmers", [])):
                o = td / tag
                o.mkdir()
                t0 = time.perf_counter()
                r = subprocess.run([str(exe)] + sub_common + ["-o", str(o)] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                wall = time.perf_counter() - t0
                assert r.returncode == 0, r.stderr[-1000:]
                dirs[tag] = o
                out[tag] = {"rows": nref, "wall_s": wall, "rows_per_s": nref / wall}
            fa = sorted(x.name for x in dirs["ref"].iterdir())
            out["outputs_byte_identical"] = fa == sorted(x.name for x in dirs["ours_sub"].iterdir()) and all(
                filecmp.cmp(dirs["ref"] / x, dirs["ours_sub"] / x, shallow=False) for x in fa)
            out["files_compared"] = len(fa)
            out["speedup_cli_wall_clock"] = out["ours"]["rows_per_s"] / out["ref"]["rows_per_s"]
        out["cores"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
