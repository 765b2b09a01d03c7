"""ctypes binding of libtss.so (include/tss.h) for the tests and bench.py.

The product is the C-ABI library; this module only forwards to it.  There is
no Python / NumPy / torch implementation of any data-path call here: if
libtss.so is missing or a call fails the binding raises.

The directory name contains '-', so import it with ``tss_loader.load()`` (repo
root) or ``importlib`` under the module name ``trie_semantic_search_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (TSS_LIB_PATH: diagnostics only -- A/B of two builds of the same library in one GPU session)
LIB_PATH = os.environ.get("TSS_LIB_PATH") or os.path.join(_HERE, "libtss.so")

TSS_OK, TSS_ERR_INVALID_ARG, TSS_ERR_CUDA, TSS_ERR_NCCL, TSS_ERR_OOM, TSS_ERR_STATE = range(6)
TSS_F32, TSS_BF16 = 0, 1
TSS_MASK_NONE, TSS_MASK_INCLUDE, TSS_MASK_EXCLUDE = 0, 1, 2
TSS_PREFIX_TOKEN, TSS_PREFIX_CHAR = 0, 1
TSS_MAX_K = 1024
TSS_MAX_FUSED_K = 128
TSS_ROW_NONE = 0xFFFFFFFF
TSS_MAX_PENDING = 4
TSS_PENDING_MAX_NQ = 4

# every symbol include/tss.h declares (tests/test_abi.py checks the .so exports all of them)
ABI_SYMBOLS = [
    "tss_abi_version", "tss_last_error", "tss_device_count",
    "tss_index_create", "tss_index_reserve", "tss_index_add", "tss_index_add_synthetic",
    "tss_index_finalize", "tss_index_size", "tss_index_dim", "tss_index_destroy",
    "tss_index_get_rows", "tss_index_search", "tss_index_search_device", "tss_unpack_keys",
    "tss_comm_unique_id", "tss_comm_create", "tss_comm_destroy", "tss_index_set_shard", "tss_index_set_batch_policy",
    "tss_mask_create", "tss_mask_clear", "tss_mask_set_rows", "tss_mask_upload",
    "tss_mask_download", "tss_mask_popcount", "tss_mask_nbits", "tss_mask_destroy",
    "tss_terms_create", "tss_terms_size", "tss_terms_destroy", "tss_prefix_mask",
    "tss_index_stream", "tss_index_sync", "tss_dev_alloc", "tss_dev_free", "tss_dev_h2d",
    "tss_dev_d2h", "tss_event_create", "tss_event_record", "tss_event_elapsed_ms",
    "tss_event_destroy", "tss_launch_count", "tss_index_debug_phases",
    "tss_index_save", "tss_index_load",
    "tss_mask_clear_rows", "tss_columns_create", "tss_columns_destroy", "tss_filter_mask",
    "tss_terms_build", "tss_terms_sizes", "tss_terms_export",
    "tss_prefix_mask_fresh", "tss_terms_bind_stream",
    "tss_terms_build_text", "tss_terms_save", "tss_terms_load",
    "tss_index_search_submit", "tss_index_search_collect", "tss_index_search_prefix",
    "tss_index_search_prefix_submit",
]


class TssError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libtss error {code}: {msg}")
        self.code = code


class PrefixStats(C.Structure):
    _fields_ = [("exact_lo", C.c_uint64), ("exact_hi", C.c_uint64), ("sub_lo", C.c_uint64),
                ("sub_hi", C.c_uint64), ("npostings", C.c_uint64)]


_lib = None


def lib() -> C.CDLL:
    """Load libtss.so; raises (never falls back) when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (there is no fallback path)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    pf, pu32, pu64 = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    sig = {
        "tss_abi_version": (i32, []),
        "tss_last_error": (C.c_char_p, []),
        "tss_device_count": (i32, []),
        "tss_index_create": (i32, [C.POINTER(vp), u32, i32, i32]),
        "tss_index_reserve": (i32, [vp, u64]),
        "tss_index_add": (i32, [vp, vp, u64]),
        "tss_index_add_synthetic": (i32, [vp, u64, u64, u64]),
        "tss_index_finalize": (i32, [vp]),
        "tss_index_size": (u64, [vp]),
        "tss_index_dim": (u32, [vp]),
        "tss_index_destroy": (None, [vp]),
        "tss_index_get_rows": (i32, [vp, u64, u64, vp]),
        "tss_index_search": (i32, [vp, vp, u32, u32, vp, i32, vp, vp, vp]),
        "tss_index_search_device": (i32, [vp, vp, u32, u32, vp, i32, vp]),
        "tss_index_search_submit": (i32, [vp, vp, u32, u32, vp, i32, pu64]),
        "tss_index_search_collect": (i32, [vp, u64, vp, vp, vp]),
        "tss_index_search_prefix": (i32, [vp, vp, C.c_char_p, u32, i32, vp, vp, u32, u32, vp, vp, vp]),
        "tss_index_search_prefix_submit": (i32, [vp, vp, C.c_char_p, u32, i32, vp, vp, u32, u32, pu64]),
        "tss_unpack_keys": (None, [vp, u64, vp, vp]),
        "tss_comm_unique_id": (i32, [vp]),
        "tss_comm_create": (i32, [C.POINTER(vp), vp, i32, i32, i32]),
        "tss_comm_destroy": (None, [vp]),
        "tss_index_set_shard": (i32, [vp, u64, vp]),
        "tss_index_set_batch_policy": (i32, [vp, u32, i32]),
        "tss_mask_create": (i32, [C.POINTER(vp), u64, i32]),
        "tss_mask_clear": (i32, [vp]),
        "tss_mask_set_rows": (i32, [vp, vp, u64, u64]),
        "tss_mask_upload": (i32, [vp, vp]),
        "tss_mask_download": (i32, [vp, vp]),
        "tss_mask_popcount": (i32, [vp, pu64]),
        "tss_mask_nbits": (u64, [vp]),
        "tss_mask_destroy": (None, [vp]),
        "tss_terms_create": (i32, [C.POINTER(vp), vp, vp, vp, vp, u64, i32]),
        "tss_terms_size": (u64, [vp]),
        "tss_terms_destroy": (None, [vp]),
        "tss_prefix_mask": (i32, [vp, C.c_char_p, u32, i32, vp, u64, C.POINTER(PrefixStats)]),
        "tss_prefix_mask_fresh": (i32, [vp, C.c_char_p, u32, i32, vp, u64, C.POINTER(PrefixStats)]),
        "tss_terms_bind_stream": (i32, [vp, vp]),
        "tss_terms_build_text": (i32, [C.POINTER(vp), vp, vp, vp, u64, i32, u32, i32]),
        "tss_terms_save": (i32, [vp, C.c_char_p]),
        "tss_terms_load": (i32, [C.POINTER(vp), C.c_char_p, i32]),
        "tss_index_stream": (vp, [vp]),
        "tss_index_sync": (i32, [vp]),
        "tss_dev_alloc": (i32, [i32, u64, C.POINTER(vp)]),
        "tss_dev_free": (i32, [i32, vp]),
        "tss_dev_h2d": (i32, [i32, vp, vp, u64]),
        "tss_dev_d2h": (i32, [i32, vp, vp, u64]),
        "tss_event_create": (i32, [i32, C.POINTER(vp)]),
        "tss_event_record": (i32, [vp, vp]),
        "tss_event_elapsed_ms": (i32, [vp, vp, pf]),
        "tss_event_destroy": (i32, [vp]),
        "tss_launch_count": (u64, []),
        "tss_index_debug_phases": (i32, [vp, vp]),
        "tss_mask_clear_rows": (i32, [vp, vp, u64, u64]),
        "tss_columns_create": (i32, [C.POINTER(vp), vp, vp, u64, i32]),
        "tss_columns_destroy": (None, [vp]),
        "tss_filter_mask": (i32, [vp, vp, u32, C.c_int32, C.c_int32, vp, i32]),
        "tss_terms_build": (i32, [C.POINTER(vp), vp, vp, u32, vp, u32, vp, u64, i32]),
        "tss_terms_sizes": (i32, [vp, pu64, pu64, pu64]),
        "tss_terms_export": (i32, [vp, vp, vp, vp, vp]),
        "tss_index_save": (i32, [vp, C.c_char_p]),
        "tss_index_load": (i32, [C.POINTER(vp), C.c_char_p, i32]),
    }
    del pu32
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != TSS_OK:
        raise TssError(rc, lib().tss_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return lib().tss_device_count()


def launch_count() -> int:
    return int(lib().tss_launch_count())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def unpack_keys(keys: np.ndarray):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    rows = np.empty(keys.shape, dtype=np.uint32)
    scores = np.empty(keys.shape, dtype=np.float32)
    lib().tss_unpack_keys(keys.ctypes.data, keys.size, rows.ctypes.data, scores.ctypes.data)
    return rows, scores


class DeviceBuffer:
    """Raw HBM allocation made through the ABI (tests and bench plumbing)."""

    def __init__(self, device: int, nbytes: int):
        self.device, self.nbytes = device, int(nbytes)
        p = C.c_void_p()
        _check(lib().tss_dev_alloc(device, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, arr: np.ndarray) -> "DeviceBuffer":
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        _check(lib().tss_dev_h2d(self.device, self.ptr, arr.ctypes.data, arr.nbytes))
        return self

    def download(self, dtype, count: int) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        assert out.nbytes <= self.nbytes
        _check(lib().tss_dev_d2h(self.device, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self) -> None:
        if self.ptr:
            lib().tss_dev_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Event:
    def __init__(self, device: int):
        p = C.c_void_p()
        _check(lib().tss_event_create(device, C.byref(p)))
        self.ptr = p.value

    def record(self, index: "FlatIndex") -> None:
        _check(lib().tss_event_record(index.handle, self.ptr))

    def elapsed_ms(self, later: "Event") -> float:
        ms = C.c_float()
        _check(lib().tss_event_elapsed_ms(self.ptr, later.ptr, C.byref(ms)))
        return float(ms.value)

    def __del__(self):
        try:
            if self.ptr:
                lib().tss_event_destroy(self.ptr)
        except Exception:
            pass


class Comm:
    """One rank of a row-sharded index (tss_comm_*)."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _check(lib().tss_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, unique_id: bytes, rank: int, nranks: int, device: int):
        assert len(unique_id) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        p = C.c_void_p()
        _check(lib().tss_comm_create(C.byref(p), buf, rank, nranks, device))
        self.handle, self.rank, self.nranks = p.value, rank, nranks

    def close(self) -> None:
        if self.handle:
            lib().tss_comm_destroy(self.handle)
            self.handle = None


class Mask:
    """Device bitmask over the rows of one shard (tss_mask_*)."""

    def __init__(self, nbits: int, device: int = 0):
        p = C.c_void_p()
        _check(lib().tss_mask_create(C.byref(p), int(nbits), device))
        self.handle, self.nbits, self.device = p.value, int(nbits), device

    def clear(self) -> None:
        _check(lib().tss_mask_clear(self.handle))

    def set_rows(self, rows, row_base: int = 0) -> None:
        r = np.ascontiguousarray(rows, dtype=np.uint32)
        _check(lib().tss_mask_set_rows(self.handle, r.ctypes.data, r.size, int(row_base)))

    def clear_rows(self, rows, row_base: int = 0) -> None:
        r = np.ascontiguousarray(rows, dtype=np.uint32)
        _check(lib().tss_mask_clear_rows(self.handle, r.ctypes.data, r.size, int(row_base)))

    def upload(self, words) -> None:
        w = np.ascontiguousarray(words, dtype=np.uint32)
        assert w.size == (self.nbits + 31) // 32
        _check(lib().tss_mask_upload(self.handle, w.ctypes.data))

    def download(self) -> np.ndarray:
        w = np.empty((self.nbits + 31) // 32, dtype=np.uint32)
        _check(lib().tss_mask_download(self.handle, w.ctypes.data))
        return w

    def popcount(self) -> int:
        out = C.c_uint64()
        _check(lib().tss_mask_popcount(self.handle, C.byref(out)))
        return int(out.value)

    def close(self) -> None:
        if self.handle:
            lib().tss_mask_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Columns:
    """Per-row metadata columns (u16 court id, i32 decision date) for the pre-filter (N3)."""

    def __init__(self, court_ids, dates, device: int = 0):
        c = np.ascontiguousarray(court_ids, dtype=np.uint16)
        d = np.ascontiguousarray(dates, dtype=np.int32)
        assert c.size == d.size
        p = C.c_void_p()
        _check(lib().tss_columns_create(C.byref(p), c.ctypes.data, d.ctypes.data, c.size, device))
        self.handle, self.nrows = p.value, c.size

    def filter_mask(self, mask: "Mask", allowed_courts=(), date_lo: int = -2**31,
                    date_hi: int = 2**31 - 1, combine_and: bool = False) -> None:
        a = np.ascontiguousarray(list(allowed_courts), dtype=np.uint16)
        _check(lib().tss_filter_mask(self.handle, a.ctypes.data if a.size else None, a.size,
                                     int(date_lo), int(date_hi), mask.handle, int(combine_and)))

    def close(self) -> None:
        if self.handle:
            lib().tss_columns_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Terms:
    """Flattened, byte-sorted term array with CSR postings (tss_terms_*)."""

    def __init__(self, terms: Sequence[bytes], postings: Sequence[Sequence[int]], device: int = 0):
        assert len(terms) == len(postings)
        pool = b"".join(terms)
        toff = np.zeros(len(terms) + 1, dtype=np.uint64)
        np.cumsum([len(t) for t in terms], out=toff[1:])
        poff = np.zeros(len(terms) + 1, dtype=np.uint64)
        np.cumsum([len(p) for p in postings], out=poff[1:])
        rows = np.fromiter((r for p in postings for r in p), dtype=np.uint32, count=int(poff[-1]))
        self._init_raw(pool, toff, poff, rows, device)

    @classmethod
    def from_arrays(cls, pool: bytes, term_off, post_off, post_rows, device: int = 0) -> "Terms":
        self = cls.__new__(cls)
        self._init_raw(pool, np.ascontiguousarray(term_off, dtype=np.uint64),
                       np.ascontiguousarray(post_off, dtype=np.uint64),
                       np.ascontiguousarray(post_rows, dtype=np.uint32), device)
        return self

    def _init_raw(self, pool, toff, poff, rows, device):
        p = C.c_void_p()
        pool_buf = np.frombuffer(pool, dtype=np.uint8) if len(pool) else np.zeros(1, dtype=np.uint8)
        rows_buf = rows if rows.size else np.zeros(1, dtype=np.uint32)
        _check(lib().tss_terms_create(C.byref(p), pool_buf.ctypes.data, toff.ctypes.data,
                                      poff.ctypes.data, rows_buf.ctypes.data, toff.size - 1, device))
        self.handle, self.device = p.value, device

    @classmethod
    def build(cls, vocab: Sequence[bytes], token_ids, rows, device: int = 0) -> "Terms":
        """N2: device-side construction from tokenised postings (token id = vocab index + 1)."""
        self = cls.__new__(cls)
        vpool = b"".join(vocab)
        voff = np.zeros(len(vocab) + 1, dtype=np.uint64)
        np.cumsum([len(v) for v in vocab], out=voff[1:])
        ids = np.ascontiguousarray(token_ids, dtype=np.uint32)
        if ids.ndim == 1:
            ids = ids.reshape(-1, 1)
        r = np.ascontiguousarray(rows, dtype=np.uint32)
        assert ids.shape[0] == r.size
        vbuf = np.frombuffer(vpool, dtype=np.uint8) if vpool else np.zeros(1, dtype=np.uint8)
        p = C.c_void_p()
        _check(lib().tss_terms_build(C.byref(p), vbuf.ctypes.data, voff.ctypes.data, len(vocab),
                                     ids.ctypes.data if ids.size else None, ids.shape[1],
                                     r.ctypes.data if r.size else None, r.size, device))
        self.handle, self.device = p.value, device
        return self

    @classmethod
    def build_text(cls, phrases: Sequence[bytes], rows, lowercase: bool = True,
                   max_tokens: int = 16, device: int = 0) -> "Terms":
        """N2 from raw strings: the device tokenises (ASCII whitespace), lower-cases (optional),
        dictionary-encodes and builds; phrase i posts rows[i]."""
        self = cls.__new__(cls)
        text = b"".join(phrases)
        off = np.zeros(len(phrases) + 1, dtype=np.uint64)
        np.cumsum([len(p) for p in phrases], out=off[1:])
        r = np.ascontiguousarray(rows, dtype=np.uint32)
        assert r.size == len(phrases)
        tbuf = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(1, dtype=np.uint8)
        p = C.c_void_p()
        _check(lib().tss_terms_build_text(C.byref(p), tbuf.ctypes.data, off.ctypes.data,
                                          r.ctypes.data if r.size else None, len(phrases),
                                          1 if lowercase else 0, int(max_tokens), device))
        self.handle, self.device = p.value, device
        return self

    def save(self, path: str) -> None:
        _check(lib().tss_terms_save(self.handle, path.encode()))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "Terms":
        p = C.c_void_p()
        _check(lib().tss_terms_load(C.byref(p), path.encode(), device))
        self = cls.__new__(cls)
        self.handle, self.device = p.value, device
        return self

    def export(self):
        """-> (terms: list[bytes], postings: list[list[int]]) copied back from the device"""
        nt, pb, npst = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().tss_terms_sizes(self.handle, C.byref(nt), C.byref(pb), C.byref(npst)))
        pool = np.zeros(max(pb.value, 1), dtype=np.uint8)
        toff = np.zeros(nt.value + 1, dtype=np.uint64)
        poff = np.zeros(nt.value + 1, dtype=np.uint64)
        rows = np.zeros(max(npst.value, 1), dtype=np.uint32)
        _check(lib().tss_terms_export(self.handle, pool.ctypes.data, toff.ctypes.data,
                                      poff.ctypes.data, rows.ctypes.data))
        raw = pool.tobytes()
        terms = [raw[int(toff[i]):int(toff[i + 1])] for i in range(nt.value)]
        posts = [rows[int(poff[i]):int(poff[i + 1])].tolist() for i in range(nt.value)]
        return terms, posts

    def size(self) -> int:
        return int(lib().tss_terms_size(self.handle))

    def prefix_mask(self, prefix: bytes, mask: Mask, kind: int = TSS_PREFIX_TOKEN,
                    row_base: int = 0, want_stats: bool = True,
                    fresh: bool = False) -> Optional[PrefixStats]:
        """fresh=True: clear + prefix mask as one enqueue (tss_prefix_mask_fresh)."""
        st = PrefixStats() if want_stats else None
        fn = lib().tss_prefix_mask_fresh if fresh else lib().tss_prefix_mask
        _check(fn(self.handle, prefix, len(prefix), kind, mask.handle,
                  int(row_base), C.byref(st) if st is not None else None))
        return st

    def bind_stream(self, index: Optional["FlatIndex"]) -> None:
        _check(lib().tss_terms_bind_stream(self.handle, index.handle if index else None))

    def close(self) -> None:
        if self.handle:
            lib().tss_terms_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FlatIndex:
    """The HnswIndex seam (reference src/vector.rs:184-208) on one B200.

    add / add_vector <- HnswIndex::add_vector, search <- HnswIndex::search
    (returns cosine similarity, best first), size <- HnswIndex::size.
    """

    def __init__(self, dim: int, storage: int = TSS_F32, device: int = 0):
        p = C.c_void_p()
        _check(lib().tss_index_create(C.byref(p), dim, storage, device))
        self.handle, self.dim, self.device, self.storage = p.value, dim, device, storage

    def save(self, path: str) -> None:
        _check(lib().tss_index_save(self.handle, path.encode()))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "FlatIndex":
        p = C.c_void_p()
        _check(lib().tss_index_load(C.byref(p), path.encode(), device))
        self = cls.__new__(cls)
        self.handle, self.device = p.value, device
        self.dim = int(lib().tss_index_dim(p.value))
        self.storage = None
        return self

    def reserve(self, nrows: int) -> None:
        _check(lib().tss_index_reserve(self.handle, int(nrows)))

    def add(self, rows) -> None:
        r = _f32(rows)
        if r.ndim == 1:
            r = r.reshape(1, -1)
        if r.shape[1] != self.dim:
            raise ValueError(f"rows have {r.shape[1]} columns, index dim is {self.dim}")
        _check(lib().tss_index_add(self.handle, r.ctypes.data, r.shape[0]))

    def add_vector(self, embedding) -> int:
        """one row, as HnswIndex::add_vector; returns the row id it got."""
        row = self.size()
        self.add(embedding)
        return row

    def add_synthetic(self, row_begin: int, nrows: int, seed: int) -> None:
        _check(lib().tss_index_add_synthetic(self.handle, int(row_begin), int(nrows), int(seed)))

    def finalize(self) -> None:
        _check(lib().tss_index_finalize(self.handle))

    def size(self) -> int:
        return int(lib().tss_index_size(self.handle))

    def get_rows(self, row_begin: int, nrows: int) -> np.ndarray:
        out = np.empty((nrows, self.dim), dtype=np.float32)
        _check(lib().tss_index_get_rows(self.handle, int(row_begin), int(nrows), out.ctypes.data))
        return out

    def set_batch_policy(self, min_queries: int = 0, build_shadow_now: bool = False) -> None:
        """Which batches take the tensor-core path (results are bit-identical either way); see
        tss_index_set_batch_policy in include/tss.h."""
        _check(lib().tss_index_set_batch_policy(self.handle, int(min_queries),
                                                1 if build_shadow_now else 0))

    def set_shard(self, row_base: int, comm: Optional[Comm]) -> None:
        _check(lib().tss_index_set_shard(self.handle, int(row_base), comm.handle if comm else None))

    def search(self, queries, k: int, mask: Optional[Mask] = None, mask_mode: int = TSS_MASK_NONE):
        """-> (rows [nq,k] uint32, scores [nq,k] float32, counts [nq] uint32)"""
        q = _f32(queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.shape[1] != self.dim:
            raise ValueError(f"queries have {q.shape[1]} columns, index dim is {self.dim}")
        nq = q.shape[0]
        rows = np.empty((nq, k), dtype=np.uint32)
        scores = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.uint32)
        _check(lib().tss_index_search(self.handle, q.ctypes.data, nq, k,
                                      mask.handle if mask else None, mask_mode, rows.ctypes.data,
                                      scores.ctypes.data, counts.ctypes.data))
        return rows, scores, counts

    def search_into(self, q_ptr: int, nq: int, k: int, rows_ptr: int, scores_ptr: int,
                    counts_ptr: int, mask: Optional[Mask] = None,
                    mask_mode: int = TSS_MASK_NONE) -> None:
        """tss_index_search on caller-owned host buffers given as raw addresses (no per-call
        numpy allocation or conversion): for latency-sensitive callers and bench.py's e2e leg."""
        rc = lib().tss_index_search(self.handle, q_ptr, nq, k, mask.handle if mask else None,
                                    mask_mode, rows_ptr, scores_ptr, counts_ptr)
        if rc != TSS_OK:
            _check(rc)

    def search_submit(self, q_ptr: int, nq: int, k: int, mask: Optional[Mask] = None,
                      mask_mode: int = TSS_MASK_NONE) -> int:
        """tss_index_search_submit: enqueue a search of 1..4 host queries (raw address), return its
        ticket without waiting; up to TSS_MAX_PENDING may be in flight."""
        t = C.c_uint64(0)
        rc = lib().tss_index_search_submit(self.handle, q_ptr, nq, k, mask.handle if mask else None,
                                           mask_mode, C.byref(t))
        if rc != TSS_OK:
            _check(rc)
        return t.value

    def search_collect(self, ticket: int, rows_ptr: int, scores_ptr: int, counts_ptr: int) -> None:
        """tss_index_search_collect: wait for that search alone, unpack into caller-owned buffers."""
        rc = lib().tss_index_search_collect(self.handle, ticket, rows_ptr, scores_ptr, counts_ptr)
        if rc != TSS_OK:
            _check(rc)

    def search_prefix_into(self, terms: "Terms", prefix: bytes, scratch: Mask, q_ptr: int, nq: int,
                           k: int, rows_ptr: int, scores_ptr: int, counts_ptr: int,
                           kind: int = TSS_PREFIX_TOKEN) -> None:
        """tss_index_search_prefix: prefix -> fresh mask -> masked top-k as one host call."""
        rc = lib().tss_index_search_prefix(self.handle, terms.handle, prefix, len(prefix), kind,
                                           scratch.handle, q_ptr, nq, k, rows_ptr, scores_ptr,
                                           counts_ptr)
        if rc != TSS_OK:
            _check(rc)

    def search_prefix_submit(self, terms: "Terms", prefix: bytes, scratch: Mask, q_ptr: int, nq: int,
                             k: int, kind: int = TSS_PREFIX_TOKEN) -> int:
        """tss_index_search_prefix_submit: the hybrid query enqueued, a ticket for search_collect."""
        t = C.c_uint64(0)
        rc = lib().tss_index_search_prefix_submit(self.handle, terms.handle, prefix, len(prefix), kind,
                                                  scratch.handle, q_ptr, nq, k, C.byref(t))
        if rc != TSS_OK:
            _check(rc)
        return t.value

    def search_device(self, d_queries: DeviceBuffer, nq: int, k: int, d_out_keys: DeviceBuffer,
                      mask: Optional[Mask] = None, mask_mode: int = TSS_MASK_NONE) -> None:
        _check(lib().tss_index_search_device(self.handle, d_queries.ptr, nq, k,
                                             mask.handle if mask else None, mask_mode,
                                             d_out_keys.ptr))

    def sync(self) -> None:
        _check(lib().tss_index_sync(self.handle))

    def stream(self) -> int:
        return int(lib().tss_index_stream(self.handle) or 0)

    def close(self) -> None:
        if self.handle:
            lib().tss_index_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
