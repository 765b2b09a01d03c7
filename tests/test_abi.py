"""The C-ABI library loads on a CPU box and exports exactly what include/tss.h declares.
No compute call is made here (there is no GPU); the product path must fail loudly."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tss.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tss_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(tss):
    declared = _declared()
    assert len(declared) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", tss.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in tss.h but not exported: {missing}"
    # nothing but the ABI leaks out of the library
    extra = [s for s in exported if not s.startswith("tss_")]
    assert not extra, f"non-ABI symbols exported: {extra[:5]}"
    assert sorted(tss.ABI_SYMBOLS) == declared


def test_rust_binding_source_declares_the_same_symbols():
    """ffi/tss.rs cannot be compiled here (no rustc), but its extern block must at least name
    every entry point of the header, no more and no fewer."""
    rs = open(os.path.join(ROOT, "trie-semantic-search_b200", "ffi", "tss.rs")).read()
    block = rs[rs.index('extern "C" {'):]
    block = block[:block.index("\n}")]
    assert sorted(set(re.findall(r"pub fn (tss_[a-z0-9_]+)\s*\(", block))) == _declared()


_C2RS = {"int": "c_int", "char": "c_char", "void": "c_void", "float": "f32", "uint8_t": "u8",
         "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "int64_t": "i64"}


def _c_signatures():
    """name -> (return, [params]) from tss.h, each type as (rust base name, pointer depth)"""
    src = open(os.path.join(ROOT, "include", "tss.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def norm(t):
        t = t.strip()
        depth = t.count("*") + (1 if re.search(r"\[\d*\]", t) else 0)
        t = re.sub(r"\[\d*\]", "", t.replace("*", " "))
        words = [w for w in t.split() if w not in ("const", "struct")]
        base = words[0] if len(words) == 1 or words[0] in _C2RS or words[0].startswith("tss_") else None
        assert base, t
        return (_C2RS.get(base, base), depth)

    out = {}
    for ret, name, params in re.findall(r"^\s*([A-Za-z_][\w\s\*]*?)\b(tss_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src,
                                        flags=re.M | re.S):
        ps = [] if params.strip() in ("", "void") else [norm(x) for x in params.split(",")]
        out[name] = (norm(ret), ps)
    return out


def _rs_signatures():
    rs = open(os.path.join(ROOT, "trie-semantic-search_b200", "ffi", "tss.rs")).read()
    block = rs[rs.index('extern "C" {'):]
    block = block[:block.index("\n}")]

    def norm(t):
        t = t.strip()
        depth = t.count("*")
        t = re.sub(r"\*\s*(const|mut)\s*", "", t).strip()
        return (t, depth)

    out = {}
    for name, params, ret in re.findall(r"pub fn (tss_[a-z0-9_]+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        ps = [norm(x.split(":", 1)[1]) for x in params.split(",") if ":" in x]
        out[name] = (norm(ret) if ret else ("c_void", 0), ps)
    return out


def test_rust_binding_signatures_match_the_header():
    """Beyond the names: every extern fn of ffi/tss.rs has the header's arity, and each parameter
    and return value the header's base type and pointer depth (u32 <-> uint32_t, *mut *mut tss_index
    <-> tss_index**, ...).  rustc is not available here, so this is the type check the binding gets."""
    c, rs = _c_signatures(), _rs_signatures()
    assert sorted(c) == sorted(rs) == _declared()
    for name in c:
        (cret, cps), (rret, rps) = c[name], rs[name]
        assert len(cps) == len(rps), (name, cps, rps)
        assert cps == rps, (name, cps, rps)
        assert cret == rret, (name, cret, rret)


def test_library_loads_and_reports_version(tss):
    L = tss.lib()
    assert L.tss_abi_version() == 1
    assert tss.launch_count() >= 0


def test_no_cpu_fallback(tss):
    if tss.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(tss.TssError) as ei:
        tss.FlatIndex(384)
    assert ei.value.code == tss.TSS_ERR_CUDA and "no CPU path" in str(ei.value)
    with pytest.raises(tss.TssError):
        tss.Mask(100)


def test_argument_validation_without_device(tss):
    import ctypes as C
    L = tss.lib()
    p = C.c_void_p()
    assert L.tss_index_create(C.byref(p), 0, tss.TSS_F32, 0) == tss.TSS_ERR_INVALID_ARG
    assert L.tss_index_create(C.byref(p), 2048, tss.TSS_F32, 0) == tss.TSS_ERR_INVALID_ARG
    assert L.tss_index_create(C.byref(p), 384, 7, 0) == tss.TSS_ERR_INVALID_ARG
    assert b"storage" in L.tss_last_error()
    assert L.tss_index_size(None) == 0
    # the in-flight entries reject a NULL index / ticket before touching anything
    t = C.c_uint64(7)
    assert L.tss_index_search_submit(None, None, 1, 10, None, 0, C.byref(t)) == tss.TSS_ERR_INVALID_ARG
    assert L.tss_index_search_collect(None, 1, None, None, None) == tss.TSS_ERR_INVALID_ARG
    assert L.tss_index_search_prefix(None, None, b"a", 1, 0, None, None, 1, 10, None, None, None) == \
        tss.TSS_ERR_INVALID_ARG
    assert L.tss_index_search_prefix_submit(None, None, b"a", 1, 0, None, None, 1, 10, C.byref(t)) == \
        tss.TSS_ERR_INVALID_ARG
    assert b"index is NULL" in L.tss_last_error()
    L.tss_index_destroy(None)  # no-op
    L.tss_mask_destroy(None)
    L.tss_terms_destroy(None)


def test_unpack_keys_is_pure_host(tss, orc):
    import numpy as np
    keys = np.array([orc.pack_key(0.75, 12), orc.pack_key(-0.5, 7), 0], dtype=np.uint64)
    rows, scores = tss.unpack_keys(keys)
    assert list(rows) == [12, 7, tss.TSS_ROW_NONE]
    assert list(scores) == [0.75, -0.5, 0.0]


def test_terms_validation_runs_before_any_device_call(tss):
    import numpy as np
    import ctypes as C
    L = tss.lib()
    p = C.c_void_p()
    pool = np.frombuffer(b"bbaa", dtype=np.uint8)
    toff = np.array([0, 2, 4], dtype=np.uint64)
    poff = np.array([0, 0, 0], dtype=np.uint64)
    rows = np.zeros(1, dtype=np.uint32)
    rc = L.tss_terms_create(C.byref(p), pool.ctypes.data, toff.ctypes.data, poff.ctypes.data,
                            rows.ctypes.data, 2, 0)
    assert rc == tss.TSS_ERR_INVALID_ARG and b"byte-sorted" in L.tss_last_error()


def test_index_load_rejects_bad_files_before_touching_a_device(tss, tmp_path):
    import ctypes as C
    import struct
    L = tss.lib()
    p = C.c_void_p()
    bad = tmp_path / "bad.tssidx"
    bad.write_bytes(b"NOTANIDX" + b"\0" * 56)
    assert L.tss_index_load(C.byref(p), str(bad).encode(), 0) == tss.TSS_ERR_INVALID_ARG
    assert b"TSSIDX01" in L.tss_last_error()
    assert L.tss_index_load(C.byref(p), str(tmp_path / "missing").encode(), 0) == tss.TSS_ERR_INVALID_ARG
    hdr = b"TSSIDX01" + struct.pack("<IIIIQQ", 384, 0, 999, 0, 10, 1536) + b"\0" * 24
    bad.write_bytes(hdr)
    assert L.tss_index_load(C.byref(p), str(bad).encode(), 0) == tss.TSS_ERR_INVALID_ARG
    assert b"inconsistent" in L.tss_last_error()
