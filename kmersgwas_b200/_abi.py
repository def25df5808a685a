"""ctypes binding of include/kmersgwas_b200.h (lib/libkmersgwas_b200.so).

This is plumbing for tests and bench.py; the product host code is the C++ mirror of the reference
classes in kmersgwas_b200/host/.  There is no CPU fallback: if the CUDA library is missing or no
device is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

KG_OK = 0
KG_ERR_HITS_OVERFLOW = 4
OPT_SCAN_ENGINE, OPT_HIT_CAPACITY, OPT_KINSHIP_ENGINE, OPT_KERNEL_TIMING, OPT_FILTER_PAIR_LIMIT = 1, 2, 3, 4, 5
OPT_SELECT_GROWTH_PERMILLE, OPT_SELECT_MAX_ROUND, OPT_SELECT_CAND_CAP, OPT_SELECT_LOG_CAP = 6, 7, 8, 9
SELECT_LOG = 1
KERNEL_SCAN_EXACT, KERNEL_SCAN_FILTER, KERNEL_SCAN_REFINE, KERNEL_KINSHIP, KERNEL_AUX, KERNEL_SCAN_SELECT = 0, 1, 2, 3, 4, 5
KERNEL_CLASS_NAMES = ["scan_exact", "scan_filter", "scan_refine", "kinship", "aux", "scan_select"]

HIT_DTYPE = np.dtype([("row", "<u8"), ("kmer", "<u8"), ("score", "<f8"), ("pheno", "<u4"), ("pad", "<u4")])

# every symbol include/kmersgwas_b200.h declares
ABI_SYMBOLS = [
    "kg_abi_version", "kg_ctx_create", "kg_ctx_destroy", "kg_last_error", "kg_set_option", "kg_sync",
    "kg_scan_set_phenotypes", "kg_scan_set_thresholds", "kg_scan_submit", "kg_scan_mark", "kg_scan_fetch",
    "kg_scan_clear_hits", "kg_scan_discard", "kg_scan_scores_dense", "kg_kinship_begin", "kg_kinship_accum_len",
    "kg_kinship_submit", "kg_kinship_fetch", "kg_host_alloc", "kg_host_free", "kg_synth_rows_device",
    "kg_launch_count", "kg_kernel_time", "kg_kernel_time_reset", "kg_scan_filter_sums", "kg_mac_filter",
    "kg_select_begin", "kg_select_end", "kg_select_sync", "kg_select_state_len", "kg_select_export", "kg_select_import",
    "kg_select_digest", "kg_select_thresholds", "kg_select_log_reset", "kg_select_log_counts", "kg_select_log_export",
    "kg_select_replay", "kg_select_set_floor", "kg_select_export_scores", "kg_select_kmax", "kg_probe_int8_peak", "kg_select_stats",
    "kg_comm_unique_id", "kg_comm_init_rank", "kg_comm_init_all", "kg_kinship_allreduce", "kg_kinship_allreduce_all",
    "kg_stream_mark", "kg_stream_wait", "kg_snps_scores", "kg_table_build", "kg_scan_filter_shape",
    "kg_bind_host_to_device",
    "kg_patterns_begin", "kg_patterns_attach", "kg_patterns_submit", "kg_patterns_count", "kg_patterns_export", "kg_patterns_insert",
]


class KgShape(C.Structure):
    _fields_ = [("n_file", C.c_uint64), ("n_used", C.c_uint64),
                ("map_word", C.POINTER(C.c_uint32)), ("map_bit", C.POINTER(C.c_uint32))]


class KgError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"kg status {status}: {msg}")
        self.status = status


_lib = None


def lib_path() -> Path:
    return _build.cuda_lib_path()


def load():
    """Load the CUDA library (it must have been built in-tree: python -m kmersgwas_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise RuntimeError(f"{p} is missing: run `python -m kmersgwas_b200.build` (no CPU fallback exists)")
    lib = C.CDLL(str(p))
    vp, u64, u32p, u64p = C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    lib.kg_abi_version.restype = C.c_int
    lib.kg_ctx_create.argtypes = [C.c_int, C.POINTER(KgShape), vp, C.POINTER(vp)]
    lib.kg_ctx_destroy.argtypes = [vp]
    lib.kg_ctx_destroy.restype = None
    lib.kg_last_error.argtypes = [vp]
    lib.kg_last_error.restype = C.c_char_p
    lib.kg_set_option.argtypes = [vp, C.c_int, C.c_int64]
    lib.kg_sync.argtypes = [vp]
    lib.kg_scan_set_phenotypes.argtypes = [vp, C.POINTER(C.c_float), C.c_uint32, u64]
    lib.kg_scan_set_thresholds.argtypes = [vp, C.POINTER(C.c_double), C.c_uint32]
    lib.kg_scan_submit.argtypes = [vp, vp, u64, u64]
    lib.kg_scan_mark.argtypes = [vp]
    lib.kg_scan_fetch.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t), u64p, u64p]
    lib.kg_scan_clear_hits.argtypes = [vp]
    lib.kg_scan_discard.argtypes = [vp]
    lib.kg_scan_scores_dense.argtypes = [vp, vp, u64, C.POINTER(C.c_uint8), C.POINTER(C.c_double)]
    lib.kg_mac_filter.argtypes = [vp, vp, u64, u64, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
    lib.kg_kinship_begin.argtypes = [vp, u64, vp]
    lib.kg_kinship_accum_len.argtypes = [vp]
    lib.kg_kinship_accum_len.restype = C.c_size_t
    lib.kg_kinship_submit.argtypes = [vp, vp, u64]
    lib.kg_kinship_fetch.argtypes = [vp, u64p, u64p]
    lib.kg_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.kg_host_free.argtypes = [vp, vp]
    lib.kg_host_free.restype = None
    lib.kg_synth_rows_device.argtypes = [vp, u64, u64, u64, vp]
    lib.kg_launch_count.argtypes = [vp]
    lib.kg_launch_count.restype = u64
    lib.kg_scan_filter_sums.argtypes = [vp, vp, u64, vp, vp]
    lib.kg_kernel_time.argtypes = [vp, C.c_int, C.POINTER(C.c_double), u64p, u64p]
    lib.kg_kernel_time_reset.argtypes = [vp]
    lib.kg_select_begin.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    lib.kg_select_end.argtypes = [vp]
    lib.kg_select_sync.argtypes = [vp, u64p, u64p]
    lib.kg_select_state_len.argtypes = [vp]
    lib.kg_select_state_len.restype = C.c_size_t
    lib.kg_select_export.argtypes = [vp, vp]
    lib.kg_select_import.argtypes = [vp, vp, u64, u64]
    lib.kg_select_digest.argtypes = [vp, u64p]
    lib.kg_select_thresholds.argtypes = [vp, vp]
    lib.kg_select_log_reset.argtypes = [vp]
    lib.kg_select_log_counts.argtypes = [vp, vp]
    lib.kg_select_log_export.argtypes = [vp, vp, vp]
    lib.kg_select_replay.argtypes = [vp, vp, vp, u64, u64]
    lib.kg_select_set_floor.argtypes = [vp, vp, C.c_uint32]
    lib.kg_select_export_scores.argtypes = [vp, u64, vp]
    lib.kg_probe_int8_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.kg_bind_host_to_device.argtypes = [C.c_int, C.POINTER(C.c_int)]
    lib.kg_bind_host_to_device.restype = C.c_int
    lib.kg_select_stats.argtypes = [vp, u64p, u64p, u64p, u64p]
    lib.kg_comm_unique_id.argtypes = [vp]
    lib.kg_comm_init_rank.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.kg_comm_init_all.argtypes = [C.POINTER(vp), C.c_int]
    lib.kg_kinship_allreduce.argtypes = [vp]
    lib.kg_kinship_allreduce_all.argtypes = [C.POINTER(vp), C.c_int]
    lib.kg_patterns_begin.argtypes = [vp, u64]
    lib.kg_patterns_submit.argtypes = [vp, vp, u64, u64]
    lib.kg_patterns_attach.argtypes = [vp, u64, u64]
    lib.kg_patterns_count.argtypes = [vp, u64p, u64p]
    lib.kg_patterns_export.argtypes = [vp, vp, u64, u64p]
    lib.kg_patterns_insert.argtypes = [vp, vp, u64]
    lib.kg_snps_scores.argtypes = [C.c_int, vp, u64, C.c_uint32, vp, vp, C.c_uint32, vp, C.c_uint32, C.c_double, vp]
    lib.kg_table_build.argtypes = [C.c_int, vp, u64, C.c_uint32, vp, vp, vp]
    lib.kg_scan_filter_shape.argtypes = [vp, u32p, u32p, u32p, u32p]
    lib.kg_stream_mark.argtypes = [vp, u64p]
    lib.kg_stream_wait.argtypes = [vp, u64]
    lib.kg_select_kmax.argtypes = [vp]
    lib.kg_select_kmax.restype = C.c_uint32
    _lib = lib
    return lib


def bind_host_to_device(device: int):
    """kg_bind_host_to_device: (numa node or -1, cpus the calling thread may now use)."""
    n = C.c_int(0)
    node = load().kg_bind_host_to_device(int(device), C.byref(n))
    return int(node), int(n.value)


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (rank 0 creates it and ships it to the other ranks)."""
    buf = C.create_string_buffer(128)
    if load().kg_comm_unique_id(buf) != KG_OK:
        raise KgError(-1, load().kg_last_error(None).decode())
    return buf.raw


def snps_scores(bed_rows: np.ndarray, map_byte, map_shift, y: np.ndarray, mac: float, device: int = 0) -> np.ndarray:
    """bed_rows: uint8 [n_snps, bytes_per_snp]; y: float32 [P, n_samples] -> float64 [P, n_snps] (kg_snps_scores)"""
    lib = load()
    bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8)
    mb = np.ascontiguousarray(map_byte, dtype=np.uint32)
    ms = np.ascontiguousarray(map_shift, dtype=np.uint32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    out = np.zeros((y.shape[0], bed_rows.shape[0]), dtype=np.float64)
    st = lib.kg_snps_scores(device, bed_rows.ctypes.data, bed_rows.shape[0], bed_rows.shape[1], mb.ctypes.data, ms.ctypes.data,
                            len(mb), y.ctypes.data, y.shape[0], float(mac), out.ctypes.data)
    if st != KG_OK:
        raise KgError(st, lib.kg_last_error(None).decode())
    return out


def kernel_times(handle) -> dict:
    """{class name: (ms_total, launches, rows)} of a kg_ctx handle (needs OPT_KERNEL_TIMING = 1)."""
    lib = load()
    out = {}
    for i, name in enumerate(KERNEL_CLASS_NAMES):
        ms, n, r = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
        if lib.kg_kernel_time(handle, i, C.byref(ms), C.byref(n), C.byref(r)) != KG_OK:
            raise KgError(-1, lib.kg_last_error(handle).decode())
        out[name] = (ms.value, int(n.value), int(r.value))
    return out


def _rows_ptr(rows):
    """rows: numpy uint64 array (host) or an int device/host address."""
    if isinstance(rows, np.ndarray):
        assert rows.dtype == np.uint64 and rows.flags.c_contiguous
        return rows.ctypes.data
    return int(rows)


class Context:
    """One kg_ctx (one GPU).  Thin, 1:1 with the C ABI."""

    def __init__(self, n_file: int, map_word, map_bit, device: int = 0, stream: int | None = None):
        self._lib = load()
        self._mw = np.ascontiguousarray(map_word, dtype=np.uint32)
        self._mb = np.ascontiguousarray(map_bit, dtype=np.uint32)
        self.n_file = int(n_file)
        self.n_used = len(self._mw)
        self.w_file = (self.n_file + 63) // 64
        shape = KgShape(self.n_file, self.n_used, self._mw.ctypes.data_as(C.POINTER(C.c_uint32)),
                        self._mb.ctypes.data_as(C.POINTER(C.c_uint32)))
        h = C.c_void_p()
        st = self._lib.kg_ctx_create(device, C.byref(shape), C.c_void_p(stream or 0), C.byref(h))
        if st != KG_OK:
            raise KgError(st, self._lib.kg_last_error(None).decode())
        self._h = h
        self.n_pheno = 0

    @classmethod
    def identity(cls, n_file: int, **kw):
        idx = np.arange(n_file)
        return cls(n_file, idx // 64, idx % 64, **kw)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st):
        if st != KG_OK:
            raise KgError(st, self._lib.kg_last_error(self._h).decode())

    def set_option(self, opt: int, value: int):
        self._chk(self._lib.kg_set_option(self._h, opt, value))

    def sync(self):
        self._chk(self._lib.kg_sync(self._h))

    @property
    def launches(self) -> int:
        return int(self._lib.kg_launch_count(self._h))

    def kernel_times(self) -> dict:
        return kernel_times(self._h)

    def kernel_times_reset(self):
        self._chk(self._lib.kg_kernel_time_reset(self._h))

    # ---- scan
    def set_phenotypes(self, y: np.ndarray, min_count: int):
        y = np.ascontiguousarray(y, dtype=np.float32)
        assert y.ndim == 2 and y.shape[1] == self.n_used
        self.n_pheno = y.shape[0]
        self._chk(self._lib.kg_scan_set_phenotypes(self._h, y.ctypes.data_as(C.POINTER(C.c_float)),
                                                   self.n_pheno, int(min_count)))

    def set_thresholds(self, thr):
        thr = np.ascontiguousarray(thr, dtype=np.float64)
        assert thr.shape == (self.n_pheno,)
        self._chk(self._lib.kg_scan_set_thresholds(self._h, thr.ctypes.data_as(C.POINTER(C.c_double)), self.n_pheno))

    def scan_submit(self, rows, n_rows: int, first_row_id: int = 0):
        self._keepalive = rows
        self._chk(self._lib.kg_scan_submit(self._h, _rows_ptr(rows), int(n_rows), int(first_row_id)))

    def scan_discard(self):
        self._chk(self._lib.kg_scan_discard(self._h))

    def scan_mark(self):
        self._chk(self._lib.kg_scan_mark(self._h))

    def scan_fetch(self):
        """Fetch the oldest interval -> (hits structured array sorted by (pheno, row), rows_seen, rows_kept).
        (The C ABI returns hits in no particular order; sorting here is a convenience of this binding.)"""
        n = C.c_size_t(0)
        seen, kept = C.c_uint64(0), C.c_uint64(0)
        self._chk(self._lib.kg_scan_fetch(self._h, None, 0, C.byref(n), C.byref(seen), C.byref(kept)))
        hits = np.zeros(n.value, dtype=HIT_DTYPE)
        if n.value:
            self._chk(self._lib.kg_scan_fetch(self._h, hits.ctypes.data, n.value, C.byref(n), None, None))
            hits = hits[np.lexsort((hits["row"], hits["pheno"]))]
        return hits, int(seen.value), int(kept.value)

    def scores_dense(self, rows, n_rows: int):
        keep = np.zeros(n_rows, dtype=np.uint8)
        scores = np.zeros((self.n_pheno, n_rows), dtype=np.float64)
        self._chk(self._lib.kg_scan_scores_dense(self._h, _rows_ptr(rows), int(n_rows),
                                                 keep.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                 scores.ctypes.data_as(C.POINTER(C.c_double))))
        return keep.astype(bool), scores

    def mac_filter(self, rows, n_rows: int, min_count: int):
        """-> (keep[n_rows] bool, kept): load_kmers' MAC filter alone (no phenotypes needed)"""
        keep = np.zeros(n_rows, dtype=np.uint8)
        kept = C.c_uint64(0)
        self._chk(self._lib.kg_mac_filter(self._h, _rows_ptr(rows), int(n_rows), int(min_count),
                                          keep.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(kept)))
        return keep.astype(bool), int(kept.value)

    def filter_shape(self) -> dict:
        a, b, c_, d = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        self._chk(self._lib.kg_scan_filter_shape(self._h, C.byref(a), C.byref(b), C.byref(c_), C.byref(d)))
        return dict(n_pass=a.value, p_pad=b.value, k_pad=c_.value, raw_stages=d.value)

    def filter_sums(self, rows, n_rows: int):
        """-> (q[n_rows, P] int32 exact sums of the tensor-core filter, yq[P, 64*W_file] int8 quantised phenotypes)"""
        q = np.zeros((n_rows, self.n_pheno), dtype=np.int32)
        yq = np.zeros((self.n_pheno, 64 * self.w_file), dtype=np.int8)
        self._chk(self._lib.kg_scan_filter_sums(self._h, _rows_ptr(rows), int(n_rows), q.ctypes.data, yq.ctypes.data))
        return q, yq

    # ---- device-side selection (BestAssociationsHeap on the GPU)
    def select_begin(self, kbest, flags: int = 0):
        kb = np.ascontiguousarray(np.broadcast_to(np.asarray(kbest, dtype=np.uint64), (self.n_pheno,)))
        self._chk(self._lib.kg_select_begin(self._h, kb.ctypes.data, self.n_pheno, int(flags)))
        self._kmax = int(kb.max())

    def select_end(self):
        self._chk(self._lib.kg_select_end(self._h))

    def select_sync(self):
        """-> (rows_applied, rows_kept); raises KgError(status KG_ERR_HITS_OVERFLOW) when a round overflowed."""
        a, k = C.c_uint64(0), C.c_uint64(0)
        st = self._lib.kg_select_sync(self._h, C.byref(a), C.byref(k))
        if st != KG_OK:
            e = KgError(st, self._lib.kg_last_error(self._h).decode())
            e.rows_applied, e.rows_kept = int(a.value), int(k.value)
            raise e
        return int(a.value), int(k.value)

    def select_state_len(self) -> int:
        return int(self._lib.kg_select_state_len(self._h))

    def select_export(self, dev_ptr: int | None = None):
        """Heap state image: numpy uint64 array (host) or written to dev_ptr."""
        if dev_ptr is not None:
            self._chk(self._lib.kg_select_export(self._h, C.c_void_p(dev_ptr)))
            return None
        out = np.zeros(self.select_state_len(), dtype=np.uint64)
        self._chk(self._lib.kg_select_export(self._h, out.ctypes.data))
        return out

    def select_import(self, state, rows_applied: int = 0, rows_kept: int = 0):
        ptr = state.ctypes.data if isinstance(state, np.ndarray) else int(state)
        self._keepalive2 = state
        self._chk(self._lib.kg_select_import(self._h, C.c_void_p(ptr), int(rows_applied), int(rows_kept)))

    def select_heaps(self):
        """-> list over phenotypes of (kmers, scores, rows) in libstdc++ LAYOUT order (position 0 = top)."""
        st = self.select_export()
        P, kmax = self.n_pheno, self._kmax
        hdr = st[:4 * P].reshape(P, 4)
        ent = st[4 * P:].reshape(P, kmax, 3)
        out = []
        for p in range(P):
            n = int(hdr[p, 0])
            out.append((ent[p, :n, 0].copy(), ent[p, :n, 1].copy().view(np.float64), ent[p, :n, 2].copy()))
        return out

    def select_stats(self) -> dict:
        a, b, c_, d = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._chk(self._lib.kg_select_stats(self._h, C.byref(a), C.byref(b), C.byref(c_), C.byref(d)))
        return dict(rounds=int(a.value), candidates=int(b.value), admitted=int(c_.value), reorders=int(d.value))

    def select_digest(self) -> int:
        d = C.c_uint64(0)
        self._chk(self._lib.kg_select_digest(self._h, C.byref(d)))
        return int(d.value)

    def select_thresholds(self):
        thr = np.zeros(self.n_pheno, dtype=np.float64)
        self._chk(self._lib.kg_select_thresholds(self._h, thr.ctypes.data))
        return thr

    def select_log_reset(self):
        self._chk(self._lib.kg_select_log_reset(self._h))

    def select_log(self, dev_ptr: int | None = None):
        """-> (offsets[P + 1], entries[total, 3] uint64 {row, kmer, score bits}); entries go to dev_ptr if given."""
        counts = np.zeros(self.n_pheno, dtype=np.uint64)
        self._chk(self._lib.kg_select_log_counts(self._h, counts.ctypes.data))
        off = np.zeros(self.n_pheno + 1, dtype=np.uint64)
        off[1:] = np.cumsum(counts)
        if dev_ptr is not None:
            self._chk(self._lib.kg_select_log_export(self._h, C.c_void_p(dev_ptr), off.ctypes.data))
            return off, None
        ent = np.zeros((int(off[-1]), 3), dtype=np.uint64)
        self._chk(self._lib.kg_select_log_export(self._h, ent.ctypes.data if len(ent) else None, off.ctypes.data))
        return off, ent

    def select_log_counts(self):
        counts = np.zeros(self.n_pheno, dtype=np.uint64)
        self._chk(self._lib.kg_select_log_counts(self._h, counts.ctypes.data))
        return counts

    def select_replay(self, entries, offsets, rows: int = 0, kept: int = 0):
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        ptr = entries.ctypes.data if isinstance(entries, np.ndarray) else int(entries or 0)
        self._keepalive2 = entries
        self._chk(self._lib.kg_select_replay(self._h, C.c_void_p(ptr), off.ctypes.data, int(rows), int(kept)))

    def select_set_floor(self, scores_dev: int, n_heaps: int):
        """scores_dev: device [n_heaps][P][k_max] f64 as select_export_scores writes them (disjoint row sets!)"""
        self._chk(self._lib.kg_select_set_floor(self._h, C.c_void_p(scores_dev), int(n_heaps)))

    def select_export_scores(self, min_row: int, scores_dev: int):
        self._chk(self._lib.kg_select_export_scores(self._h, int(min_row), C.c_void_p(scores_dev)))

    # ---- kinship
    def kinship_accum_len(self) -> int:
        return int(self._lib.kg_kinship_accum_len(self._h))

    def kinship_begin(self, min_count: int, accum_dev: int | None = None):
        self._chk(self._lib.kg_kinship_begin(self._h, int(min_count), C.c_void_p(accum_dev or 0)))

    def kinship_submit(self, rows, n_rows: int):
        self._keepalive = rows
        self._chk(self._lib.kg_kinship_submit(self._h, _rows_ptr(rows), int(n_rows)))

    def kinship_fetch(self, want_matrix: bool = True):
        m = C.c_uint64(0)
        ibs = np.zeros((self.n_used, self.n_used), dtype=np.uint64) if want_matrix else None
        p = ibs.ctypes.data_as(C.POINTER(C.c_uint64)) if want_matrix else None
        self._chk(self._lib.kg_kinship_fetch(self._h, p, C.byref(m)))
        return ibs, int(m.value)

    # ---- distinct presence/absence patterns
    def patterns_begin(self, expected: int = 0):
        self._chk(self._lib.kg_patterns_begin(self._h, int(expected)))

    def patterns_submit(self, rows, n_rows: int, min_count: int):
        self._keepalive = rows
        self._chk(self._lib.kg_patterns_submit(self._h, _rows_ptr(rows), int(n_rows), int(min_count)))

    def patterns_attach(self, min_count: int, max_rows: int):
        self._chk(self._lib.kg_patterns_attach(self._h, int(min_count), int(max_rows)))

    def patterns_count(self):
        d, k = C.c_uint64(0), C.c_uint64(0)
        self._chk(self._lib.kg_patterns_count(self._h, C.byref(d), C.byref(k)))
        return int(d.value), int(k.value)

    def patterns_export(self):
        n = C.c_uint64(0)
        self._chk(self._lib.kg_patterns_export(self._h, None, 0, C.byref(n)))
        keys = np.zeros(int(n.value), dtype=np.uint64)
        if n.value:
            self._chk(self._lib.kg_patterns_export(self._h, keys.ctypes.data, int(n.value), C.byref(n)))
        return keys

    def patterns_insert(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        self._chk(self._lib.kg_patterns_insert(self._h, keys.ctypes.data, len(keys)))

    def kinship_allreduce(self):
        self._chk(self._lib.kg_kinship_allreduce(self._h))

    def comm_init_rank(self, id128: bytes, n_ranks: int, rank: int):
        buf = C.create_string_buffer(bytes(id128), 128)
        self._chk(self._lib.kg_comm_init_rank(self._h, buf, int(n_ranks), int(rank)))

    def probe_int8_peak(self) -> float:
        t = C.c_double(0)
        self._chk(self._lib.kg_probe_int8_peak(self._h, C.byref(t)))
        return float(t.value)

    # ---- synthetic
    def synth_rows_device(self, seed: int, first_row: int, n_rows: int, rows_dev: int):
        self._chk(self._lib.kg_synth_rows_device(self._h, int(seed), int(first_row), int(n_rows), C.c_void_p(rows_dev)))
