"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Parity unpinned for scoring (see oracle/oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

ORDER_CANONICAL, ORDER_SEQUENTIAL = 0, 1
MASK_NONE, MASK_INCLUDE, MASK_EXCLUDE = 0, 1, 2
TRIE_CASE_NAME, TRIE_CONTENT, TRIE_CITATION = 0, 1, 2


class DocRef(C.Structure):
    _fields_ = [("case_id", C.c_uint8 * 16), ("paragraph_index", C.c_uint64),
                ("char_offset", C.c_int64)]

    def key(self):
        return (bytes(self.case_id), int(self.paragraph_index), int(self.char_offset))


class TrieResult(C.Structure):
    _fields_ = [("exact_matches", C.POINTER(DocRef)), ("n_exact", C.c_uint64),
                ("completions", C.POINTER(C.c_char_p)), ("n_completions", C.c_uint64),
                ("n_completions_unlimited", C.c_uint64), ("total_matches", C.c_uint64),
                ("frequency", C.c_uint32)]


class Hit(C.Structure):
    _fields_ = [("case_id", C.c_uint64), ("score", C.c_float), ("match_type", C.c_int)]


_lib = None


def build() -> None:
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
    L.orc_gen_rows.argtypes = [vp, u64, u64, u32, u64]
    L.orc_gen_rows.restype = None
    L.orc_scores.argtypes = [vp, u64, u32, vp, vp, i32, i32]
    L.orc_scores.restype = None
    L.orc_scores_bf16.argtypes = [vp, u64, u32, vp, vp, i32]
    L.orc_scores_bf16.restype = None
    L.orc_cosine_topk.argtypes = [vp, u64, u32, vp, u32, u32, vp, i32, u64, vp, vp, vp, i32, i32, i32]
    L.orc_cosine_topk.restype = i32
    L.orc_cosine_topk_synth.argtypes = [u64, u64, u32, u64, vp, u32, u32, vp, vp, vp, i32]
    L.orc_cosine_topk_synth.restype = i32
    L.orc_pack_key.argtypes = [f32, u32]
    L.orc_pack_key.restype = u64
    L.orc_trie_new.restype = vp
    L.orc_trie_free.argtypes = [vp]
    L.orc_trie_insert_case_name.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.orc_trie_insert_content.argtypes = [vp, C.POINTER(C.c_char_p), u32, C.POINTER(DocRef)]
    L.orc_trie_insert_citation.argtypes = [vp, C.c_char_p, C.POINTER(DocRef)]
    L.orc_trie_search_one.argtypes = [vp, i32, C.c_char_p]
    L.orc_trie_search_one.restype = C.POINTER(TrieResult)
    L.orc_trie_search.argtypes = [vp, C.c_char_p]
    L.orc_trie_search.restype = C.POINTER(TrieResult)
    L.orc_trie_result_free.argtypes = [C.POINTER(TrieResult)]
    L.orc_trie_prefix_postings.argtypes = [vp, i32, C.c_char_p, C.POINTER(C.POINTER(DocRef))]
    L.orc_trie_prefix_postings.restype = u64
    L.orc_free.argtypes = [vp]
    L.orc_hybrid_merge.argtypes = [vp, u32, vp, vp, u32, i32, i32, u32, C.c_int64, f32, f32,
                                   C.POINTER(Hit), u32]
    L.orc_hybrid_merge.restype = u32
    L.orc_num_threads.restype = i32
    _lib = L
    return L


def num_threads() -> int:
    return int(lib().orc_num_threads())


def gen_rows(row_begin: int, nrows: int, dim: int, seed: int) -> np.ndarray:
    out = np.empty((nrows, dim), dtype=np.float32)
    lib().orc_gen_rows(out.ctypes.data, row_begin, nrows, dim, seed)
    return out


def scores(rows: np.ndarray, query: np.ndarray, order: int = ORDER_CANONICAL, bf16: bool = False,
           threads: int = 0) -> np.ndarray:
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    query = np.ascontiguousarray(query, dtype=np.float32)
    out = np.empty(rows.shape[0], dtype=np.float32)
    if bf16:
        lib().orc_scores_bf16(rows.ctypes.data, rows.shape[0], rows.shape[1], query.ctypes.data,
                              out.ctypes.data, threads)
    else:
        lib().orc_scores(rows.ctypes.data, rows.shape[0], rows.shape[1], query.ctypes.data,
                         out.ctypes.data, order, threads)
    return out


def cosine_topk(rows, queries, k, mask_words=None, mask_mode=MASK_NONE, row_base=0,
                order=ORDER_CANONICAL, bf16=False, threads=0):
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    if queries.ndim == 1:
        queries = queries.reshape(1, -1)
    n, dim = rows.shape if rows.ndim == 2 else (0, queries.shape[1])
    nq = queries.shape[0]
    out_rows = np.empty((nq, k), dtype=np.uint32)
    out_scores = np.empty((nq, k), dtype=np.float32)
    out_counts = np.empty(nq, dtype=np.uint32)
    mw = None
    if mask_words is not None:
        mw = np.ascontiguousarray(mask_words, dtype=np.uint32)
    rc = lib().orc_cosine_topk(rows.ctypes.data if n else None, n, dim, queries.ctypes.data, nq, k,
                               mw.ctypes.data if mw is not None else None, mask_mode, row_base,
                               out_rows.ctypes.data, out_scores.ctypes.data,
                               out_counts.ctypes.data, order, 1 if bf16 else 0, threads)
    assert rc == 0
    return out_rows, out_scores, out_counts


def cosine_topk_synth(row_begin, nrows, dim, seed, queries, k, threads=0):
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    if queries.ndim == 1:
        queries = queries.reshape(1, -1)
    nq = queries.shape[0]
    out_rows = np.empty((nq, k), dtype=np.uint32)
    out_scores = np.empty((nq, k), dtype=np.float32)
    out_counts = np.empty(nq, dtype=np.uint32)
    rc = lib().orc_cosine_topk_synth(row_begin, nrows, dim, seed, queries.ctypes.data, nq, k,
                                     out_rows.ctypes.data, out_scores.ctypes.data,
                                     out_counts.ctypes.data, threads)
    assert rc == 0
    return out_rows, out_scores, out_counts


def pack_key(score: float, row: int) -> int:
    return int(lib().orc_pack_key(float(score), int(row)))


def docref(case_id: bytes, paragraph_index: int = 0, char_offset: int = -1) -> DocRef:
    d = DocRef()
    d.case_id = (C.c_uint8 * 16).from_buffer_copy(case_id)
    d.paragraph_index = paragraph_index
    d.char_offset = char_offset
    return d


class Trie:
    """TrieIndex of src/trie.rs (three token tries + cascade)."""

    def __init__(self):
        self.h = lib().orc_trie_new()

    def __del__(self):
        try:
            lib().orc_trie_free(self.h)
        except Exception:
            pass

    def insert_case_name(self, name: str, case_id: bytes):
        lib().orc_trie_insert_case_name(self.h, name.encode(), case_id)

    def insert_content(self, tokens, ref: DocRef):
        arr = (C.c_char_p * len(tokens))(*[t.encode() for t in tokens])
        lib().orc_trie_insert_content(self.h, arr, len(tokens), C.byref(ref))

    def insert_citation(self, citation: str, ref: DocRef):
        lib().orc_trie_insert_citation(self.h, citation.encode(), C.byref(ref))

    @staticmethod
    def _unpack(rp):
        r = rp.contents
        out = {
            "exact": [r.exact_matches[i].key() for i in range(r.n_exact)],
            "completions": [r.completions[i].decode() for i in range(r.n_completions)],
            "n_completions_unlimited": int(r.n_completions_unlimited),
            "total": int(r.total_matches),
            "frequency": int(r.frequency),
        }
        lib().orc_trie_result_free(rp)
        return out

    def search_one(self, which: int, query: str):
        return self._unpack(lib().orc_trie_search_one(self.h, which, query.encode()))

    def search(self, query: str):
        return self._unpack(lib().orc_trie_search(self.h, query.encode()))

    def prefix_postings(self, which: int, query: str):
        p = C.POINTER(DocRef)()
        n = lib().orc_trie_prefix_postings(self.h, which, query.encode(), C.byref(p))
        out = [p[i].key() for i in range(n)]
        if n:
            lib().orc_free(p)
        return out


def hybrid_merge(exact_cases, vec_cases, vec_scores, enable_prefix=True, enable_semantic=True,
                 cfg_max_results=10, query_max_results=None, min_similarity=0.5,
                 exact_match_weight=2.0):
    ec = np.ascontiguousarray(exact_cases, dtype=np.uint64)
    vc = np.ascontiguousarray(vec_cases, dtype=np.uint64)
    vs = np.ascontiguousarray(vec_scores, dtype=np.float32)
    cap = ec.size + vc.size + 1
    out = (Hit * cap)()
    n = lib().orc_hybrid_merge(ec.ctypes.data, ec.size, vc.ctypes.data, vs.ctypes.data, vc.size,
                               int(enable_prefix), int(enable_semantic), cfg_max_results,
                               -1 if query_max_results is None else query_max_results,
                               min_similarity, exact_match_weight, out, cap)
    return [(int(out[i].case_id), float(out[i].score), int(out[i].match_type)) for i in range(n)]
