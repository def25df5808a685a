"""ctypes binding of the host library's C entry points (kmersgwas_b200/host/host_capi.cpp):
the product's association driver + BestAssociationsHeap over in-memory tiles (bench / tests)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build as _build
from ._abi import HIT_DTYPE, _rows_ptr

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    p = _build.host_lib_path()
    if not p.exists():
        raise RuntimeError(f"{p} is missing: run `python -m kmersgwas_b200.build`")
    lib = C.CDLL(str(p))
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    lib.kgh_last_error.restype = C.c_char_p
    lib.kgh_session_create.restype = vp
    lib.kgh_session_create.argtypes = [C.c_int, u64, u64, vp, vp, vp, u32, u64, vp, vp, C.c_int, C.c_int]
    lib.kgh_session_destroy.argtypes = [vp]
    lib.kgh_session_destroy.restype = None
    lib.kgh_session_associate.argtypes = [vp, vp, u64, u64]
    lib.kgh_session_finish.argtypes = [vp]
    lib.kgh_session_ctx.argtypes = [vp]
    lib.kgh_session_ctx.restype = vp
    lib.kgh_session_heap_size.argtypes = [vp, u32]
    lib.kgh_session_heap_size.restype = u64
    lib.kgh_session_tested.argtypes = [vp, u32]
    lib.kgh_session_tested.restype = u64
    lib.kgh_session_threshold.argtypes = [vp, u32]
    lib.kgh_session_threshold.restype = C.c_double
    lib.kgh_session_heap_dump.argtypes = [vp, u32, vp, vp, vp]
    lib.kgh_session_heap_dump.restype = None
    lib.kgh_session_stats.argtypes = [vp, vp, vp, vp, vp]
    lib.kgh_session_stats.restype = None
    lib.kgh_session_host_ns.argtypes = [vp, vp]
    lib.kgh_session_host_ns.restype = None
    lib.kgh_session_io_bytes.argtypes = [vp, vp, vp]
    lib.kgh_session_io_bytes.restype = None
    lib.kgh_session_log_size.argtypes = [vp]
    lib.kgh_session_log_size.restype = u64
    lib.kgh_session_log_copy.argtypes = [vp, vp]
    lib.kgh_session_log_copy.restype = None
    lib.kgh_heapset_create.argtypes = [vp, u32]
    lib.kgh_heapset_create.restype = vp
    lib.kgh_heapset_destroy.argtypes = [vp]
    lib.kgh_heapset_destroy.restype = None
    lib.kgh_heapset_merge.argtypes = [vp, vp, u64, u64]
    lib.kgh_heapset_merge.restype = None
    lib.kgh_heapset_size.argtypes = [vp, u32]
    lib.kgh_heapset_size.restype = u64
    lib.kgh_heapset_tested.argtypes = [vp, u32]
    lib.kgh_heapset_tested.restype = u64
    lib.kgh_heapset_dump.argtypes = [vp, u32, vp, vp, vp]
    lib.kgh_heapset_dump.restype = None
    lib.kgh_heapset_add.argtypes = [vp, u32, u64, C.c_double, u64]
    lib.kgh_heapset_add.restype = None
    _lib = lib
    return lib


def _dump(fn, h, p, n):
    k = np.zeros(n, dtype=np.uint64)
    s = np.zeros(n, dtype=np.float64)
    r = np.zeros(n, dtype=np.uint64)
    if n:
        fn(h, p, k.ctypes.data, s.ctypes.data, r.ctypes.data)
    return k, s, r


class Session:
    """The product's associate loop (association_driver.cpp) bound to one GPU context and P heaps."""

    def __init__(self, n_file, map_word, map_bit, y, min_count, kbest, device=0, stream=None,
                 scan_engine=0, log_hits=False):
        self._lib = load()
        mw = np.ascontiguousarray(map_word, dtype=np.uint32)
        mb = np.ascontiguousarray(map_bit, dtype=np.uint32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        self.n_pheno = y.shape[0]
        kb = np.ascontiguousarray(np.broadcast_to(np.asarray(kbest, dtype=np.uint64), (self.n_pheno,)))
        self._h = self._lib.kgh_session_create(device, int(n_file), len(mw), mw.ctypes.data, mb.ctypes.data,
                                               y.ctypes.data, self.n_pheno, int(min_count), kb.ctypes.data,
                                               C.c_void_p(stream or 0), int(scan_engine), int(log_hits))
        if not self._h:
            raise RuntimeError("kgh_session_create: " + self._lib.kgh_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kgh_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def associate(self, rows, n_rows, first_row_id=0):
        self._keepalive = rows
        if self._lib.kgh_session_associate(self._h, _rows_ptr(rows), int(n_rows), int(first_row_id)) != 0:
            raise RuntimeError("kgh_session_associate: " + self._lib.kgh_last_error().decode())

    def finish(self):
        """Replay the round still in flight on the device; afterwards the heaps are final for the rows given so far."""
        if self._lib.kgh_session_finish(self._h) != 0:
            raise RuntimeError("kgh_session_finish: " + self._lib.kgh_last_error().decode())

    @property
    def ctx_handle(self):
        return self._lib.kgh_session_ctx(self._h)

    def launches(self) -> int:
        from ._abi import load as load_abi
        return int(load_abi().kg_launch_count(self.ctx_handle))

    def heap(self, p):
        n = int(self._lib.kgh_session_heap_size(self._h, p))
        return _dump(self._lib.kgh_session_heap_dump, self._h, p, n)

    def tested(self, p=0) -> int:
        return int(self._lib.kgh_session_tested(self._h, p))

    def threshold(self, p) -> float:
        return float(self._lib.kgh_session_threshold(self._h, p))

    def stats(self):
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._lib.kgh_session_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return dict(rounds=a.value, hits_replayed=b.value, rows_scored=c.value, rows_kept=d.value)

    def host_ms(self):
        """Host wall time per driver phase so far (ms): wait for the device, copy hits, group, replay, thresholds + submit."""
        a = (C.c_uint64 * 5)()
        self._lib.kgh_session_host_ns(self._h, a)
        return dict(zip(("wait_device", "copy_hits", "group", "replay", "thresholds_submit"), (x / 1e6 for x in a)))

    def io_bytes(self):
        """(small host->device bytes: thresholds, device->host bytes: hits + counters) so far."""
        a, b = C.c_uint64(), C.c_uint64()
        self._lib.kgh_session_io_bytes(self._h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def set_option(self, opt: int, value: int):
        from ._abi import load as load_abi
        if load_abi().kg_set_option(self.ctx_handle, opt, value) != 0:
            raise RuntimeError("kg_set_option: " + load_abi().kg_last_error(self.ctx_handle).decode())

    def kernel_times(self) -> dict:
        from ._abi import kernel_times
        return kernel_times(self.ctx_handle)

    def kernel_times_reset(self):
        from ._abi import load as load_abi
        load_abi().kg_kernel_time_reset(self.ctx_handle)

    def hit_log(self):
        n = int(self._lib.kgh_session_log_size(self._h))
        out = np.zeros(n, dtype=HIT_DTYPE)
        if n:
            self._lib.kgh_session_log_copy(self._h, out.ctypes.data)
        return out


class HeapSet:
    """P BestAssociationsHeap objects of the host library (merge of shard logs; CPU-only heap tests)."""

    def __init__(self, kbest, n_pheno):
        self._lib = load()
        self.n_pheno = n_pheno
        kb = np.ascontiguousarray(np.broadcast_to(np.asarray(kbest, dtype=np.uint64), (n_pheno,)))
        self._h = self._lib.kgh_heapset_create(kb.ctypes.data, n_pheno)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.kgh_heapset_destroy(self._h)
            self._h = None

    def add(self, p, kmer, score, row):
        self._lib.kgh_heapset_add(self._h, p, int(kmer), float(score), int(row))

    def merge(self, hits: np.ndarray, rows_kept: int):
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        self._lib.kgh_heapset_merge(self._h, hits.ctypes.data, len(hits), int(rows_kept))

    def heap(self, p):
        n = int(self._lib.kgh_heapset_size(self._h, p))
        return _dump(self._lib.kgh_heapset_dump, self._h, p, n)

    def tested(self, p=0) -> int:
        return int(self._lib.kgh_heapset_tested(self._h, p))
