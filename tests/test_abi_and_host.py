"""CPU-only tests: the C-ABI library loads here (no GPU) and exports exactly the symbols include/kmersgwas_b200.h
declares; without a device it fails loudly (no CPU fallback); the host library's BestAssociationsHeap behaves like
the oracle's restatement of the reference heap (ties included); shard-log merge; host helpers."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import support as S

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def kg():
    import kmersgwas_b200 as kg
    kg.build.build_all()
    return kg


def _header_symbols():
    text = (ROOT / "include" / "kmersgwas_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(kg):
    lib = kg.load()
    declared = _header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(kg.ABI_SYMBOLS) == declared, "the Python binding's symbol list and the header disagree"
    # and the dynamic symbol table agrees (no C++-mangled stand-ins)
    out = subprocess.run(["nm", "-D", "--defined-only", str(kg.lib_path())], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared) <= exported
    assert lib.kg_abi_version() == 1


def test_no_cpu_fallback_without_a_device(kg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(kg.KgError) as ei:
        kg.Context.identity(64)
    assert ei.value.status == 2 and "no CPU fallback" in str(ei.value)


def test_sm100a_only_build_flags(kg):
    assert "arch=compute_100a,code=sm_100a" in " ".join(kg.build.NVCC_FLAGS)
    out = subprocess.run(["cuobjdump", "-lelf", str(kg.lib_path())], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_host_heap_matches_oracle_heap_with_ties(kg):
    rng = np.random.default_rng(5)
    n, kbest = 5000, 37
    scores = rng.integers(0, 60, size=n).astype(np.float64) / 7.0     # many exact ties
    kmers = rng.integers(0, 1 << 62, size=n, dtype=np.uint64)
    hs = kg.HeapSet(kbest, 1)
    oh = S.OracleHeap(kbest)
    for r in range(n):
        hs.add(0, int(kmers[r]), float(scores[r]), r)
        oh.add(int(kmers[r]), float(scores[r]), r)
    k, s, rr = hs.heap(0)
    ko, so, ro = oh.dump()
    assert np.array_equal(k, ko) and np.array_equal(s, so) and np.array_equal(rr, ro)
    assert hs.tested(0) == oh.insertions == n


def test_merge_of_shard_logs_equals_sequential_heap(kg):
    """4 shards with their own local heaps; the merged logs reproduce the sequential heap exactly."""
    rng = np.random.default_rng(6)
    n, kbest, P, shards = 8000, 25, 3, 4
    scores = np.round(rng.standard_normal((P, n)) ** 2, 1)             # ties
    kmers = np.arange(n, dtype=np.uint64) * 3 + 1
    seq = [S.OracleHeap(kbest) for _ in range(P)]
    for j in range(P):
        seq[j].add_many(kmers, scores[j], np.arange(n, dtype=np.uint64))
    logs = []
    for g in range(shards):
        lo, hi = n * g // shards, n * (g + 1) // shards
        local = kg.HeapSet(kbest, P)
        for j in range(P):
            thr = -1.0
            for r in range(lo, hi):
                if thr < 0 or scores[j, r] > thr:
                    local.add(j, int(kmers[r]), float(scores[j, r]), r)
                    logs.append((r, int(kmers[r]), float(scores[j, r]), j, 0))
                    k, s, _ = local.heap(j)
                    thr = s[0] if len(k) >= kbest else -1.0
    log = np.array(logs, dtype=kg.HIT_DTYPE)
    rng.shuffle(log)                                                    # any order
    merged = kg.HeapSet(kbest, P)
    merged.merge(log, n)
    for j in range(P):
        k, s, r = merged.heap(j)
        ko, so, ro = seq[j].dump()
        assert np.array_equal(k, ko) and np.array_equal(s, so) and np.array_equal(r, ro)
        assert merged.tested(j) == n


def test_cli_help_and_usage_errors_run_without_a_gpu(kg):
    exe = ROOT / "kmersgwas_b200" / "bin" / "associate_kmers"
    r = subprocess.run([str(exe), "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--kmers_table" in r.stderr and "--first_phenotype_best" in r.stderr
    r = subprocess.run([str(exe), "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1
    exe = ROOT / "kmersgwas_b200" / "bin" / "emma_kinship_kmers"
    r = subprocess.run([str(exe), "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--kmers_len" in r.stderr


def test_kmers_table_to_bed_usage_errors_run_without_a_gpu(kg, tmp_path):
    """Flag handling of the converter mirrors the reference (kmers_table_to_bed.cpp:52-90): --help exits 0, a missing
    required flag / a missing file / a bad k-mer length exit 1 before any device is touched."""
    exe = ROOT / "kmersgwas_b200" / "bin" / "kmers_table_to_bed"
    assert exe.exists()
    run = lambda *a: subprocess.run([str(exe)] + [str(x) for x in a], capture_output=True, text=True)
    assert run("--help").returncode == 0
    r = run("-t", tmp_path / "t", "-k", 31)
    assert r.returncode == 1 and "is a required parameter" in r.stderr
    r = run("-t", tmp_path / "missing", "-k", 31, "-p", tmp_path / "p.tsv", "--maf", 0.05, "--mac", 5, "-b", 10, "-o", tmp_path / "o")
    assert r.returncode == 1 and "Couldn't find file" in r.stderr
    assert run("--nonsense").returncode == 1


def test_bench_reference_arm_json_contract(tmp_path):
    """`bench.py --impl reference` (the unmodified reference binary on the host cores) prints ONE JSON line with the
    keys the driver reads; runs here without a GPU on a tiny sample."""
    if not S.have_ref():
        pytest.skip("oracle/_ref not built")
    import json
    r = subprocess.run(["python", str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup-ref", "0",
                        "--ref-rows", "3000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "k-mers scored/sec" and d["unit"] == "k-mers/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_device_heap_algorithms_match_oracle_heap_on_host(tmp_path):
    """kg_select.cuh's push_heap / pop_heap restatement (the code thread 0 of the replay kernel runs) is __host__
    __device__: compile it for the host and replay random candidate streams with many ties against oracle.c's heap."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = tmp_path / "heap_host_check"
    subprocess.check_call([nvcc, "-O1", "-o", str(exe), str(ROOT / "tests" / "heap_host_check.cu"), str(ROOT / "oracle" / "oracle.c"),
                           "-Xcompiler", "-w"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout
