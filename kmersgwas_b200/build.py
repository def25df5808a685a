"""In-tree build of the native pieces (no JIT cache: the .so files travel with the repo snapshot).

  lib/libkmersgwas_b200.so   CUDA kernels + the C ABI (include/kmersgwas_b200.h), sm_100a only
  lib/libkmersgwas_host.so   host-side C++ mirror of the reference classes (kmersgwas_b200/host/)
  bin/associate_kmers        CLI with the reference's flags (host C++ over the C ABI)
  bin/emma_kinship_kmers     CLI with the reference's flags
  bin/kmers_table_to_bed     CLI with the reference's flags (table -> PLINK conversion)
  bin/associate_snps         CLI with the reference's arguments (SNP twin of the scan)
  bin/build_kmers_table      CLI with the reference's flags (table construction from sorted k-mer lists)

`python -m kmersgwas_b200.build` builds everything that is stale.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
HOST = PKG / "host"
LIB = PKG / "lib"
BIN = PKG / "bin"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]
CXX_FLAGS = ["-std=c++17", "-O2", "-Wall", "-fPIC", "-pthread"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(str(c) for c in cmd), flush=True)
    r = subprocess.run([str(c) for c in cmd], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError(f"build step failed: {' '.join(str(c) for c in cmd[:3])} ...")
    if verbose and r.stdout.strip():
        print(r.stdout)


def cuda_lib_path() -> Path:
    return LIB / "libkmersgwas_b200.so"


def host_lib_path() -> Path:
    return LIB / "libkmersgwas_host.so"


def build_cuda(force=False, verbose=False) -> Path:
    LIB.mkdir(exist_ok=True)
    out = cuda_lib_path()
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "kmersgwas_b200.h"]
    if force or _stale(out, srcs):
        extra = os.environ.get("KMERSGWAS_NVCC_EXTRA", "").split()   # e.g. -DKG_PERF_SWITCHES for profiling experiments
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-shared", "-o", out] + sorted(CSRC.glob("*.cu"))
        _run(cmd, verbose)
    return out


def build_host(force=False, verbose=False):
    """Host C++ library + CLIs; they link against the CUDA library through the C ABI only."""
    LIB.mkdir(exist_ok=True)
    BIN.mkdir(exist_ok=True)
    if not HOST.exists():
        return None
    build_cuda(force=False, verbose=verbose)
    clis = ("associate_kmers", "emma_kinship_kmers", "kmers_table_to_bed", "associate_snps", "build_kmers_table")
    lib_srcs = [p for p in sorted(HOST.glob("*.cpp")) if p.stem not in clis]
    hdrs = sorted(HOST.glob("*.h")) + [ROOT / "include" / "kmersgwas_b200.h"]
    out = host_lib_path()
    link = ["-L", LIB, "-lkmersgwas_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../lib"]
    if lib_srcs and (force or _stale(out, lib_srcs + hdrs + [cuda_lib_path()])):
        _run(["g++"] + CXX_FLAGS + ["-shared", "-I", ROOT / "include", "-o", out] + lib_srcs + link, verbose)
    for cli in clis:
        src = HOST / f"{cli}.cpp"
        if not src.exists():
            continue
        exe = BIN / cli
        if force or _stale(exe, [src] + hdrs + [out]):
            _run(["g++"] + CXX_FLAGS + ["-I", ROOT / "include", "-o", exe, src, "-L", LIB, "-lkmersgwas_host"] + link, verbose)
    return out


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built:", *(str(p) for p in sorted(LIB.glob("*.so")) + sorted(BIN.glob("*"))))
