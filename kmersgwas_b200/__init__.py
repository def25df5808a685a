"""kmersgwas_b200 -- B200 (sm_100a) implementation of the kmersGWAS association hot path.

Layout:
  csrc/   CUDA kernels + the C ABI (include/kmersgwas_b200.h)
  host/   C++ host-side mirror of the reference's MultipleKmersDataBases / BestAssociationsHeap
          surface and the two CLIs (associate_kmers, emma_kinship_kmers)
  _abi.py ctypes binding of the C ABI (tests / bench plumbing)
"""
from . import build  # noqa: F401
from ._abi import Context, KgError, HIT_DTYPE, ABI_SYMBOLS, load, lib_path  # noqa: F401
from ._abi import OPT_SCAN_ENGINE, OPT_HIT_CAPACITY, OPT_KINSHIP_ENGINE, OPT_KERNEL_TIMING, OPT_FILTER_PAIR_LIMIT  # noqa: F401
from ._abi import KERNEL_CLASS_NAMES, kernel_times  # noqa: F401
from ._abi import OPT_SELECT_GROWTH_PERMILLE, OPT_SELECT_MAX_ROUND, OPT_SELECT_CAND_CAP, OPT_SELECT_LOG_CAP, SELECT_LOG, KG_ERR_HITS_OVERFLOW  # noqa: F401

from ._host import Session, HeapSet  # noqa: F401

__all__ = ["Context", "KgError", "HIT_DTYPE", "ABI_SYMBOLS", "load", "lib_path", "build", "Session", "HeapSet"]
