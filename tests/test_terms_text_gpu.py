"""N2 from raw strings (tss_terms_build_text) and N1's term-array half (tss_terms_save/load).

The device tokenises (ASCII whitespace split), lower-cases (case names / content; citations keep
their case: src/trie.rs:147,158,171,177 vs :190,196), dictionary-encodes and builds the flattened
trie.  Checked against (a) a plain Python construction of the same structure and (b) the oracle's
literal restatement of the reference's TrieNode tree (oracle.cpp TokenTrieRef) -- terms, postings
in insertion order, and prefix masks, bit for bit."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WS = b" \t\n\r\x0b\x0c"


def _py_tokens(phrase: bytes, lowercase: bool):
    out, cur = [], bytearray()
    for b in phrase:
        if bytes([b]) in WS:
            if cur:
                out.append(bytes(cur)); cur = bytearray()
        else:
            cur.append(b + 32 if lowercase and 65 <= b <= 90 else b)
    if cur:
        out.append(bytes(cur))
    return out


def _host_build(phrases, rows, lowercase):
    d = {}
    for p, r in zip(phrases, rows):
        d.setdefault(b" ".join(_py_tokens(p, lowercase)), []).append(int(r))
    terms = sorted(d)
    return terms, [d[t] for t in terms]


def _mask_words(rows, n):
    w = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.asarray(sorted({int(r) for r in rows if r < n}), dtype=np.int64)
    if idx.size:
        np.bitwise_or.at(w, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
    return w


def _expected(terms, posts, prefix, n_rows):
    rows = []
    for t, ps in zip(terms, posts):
        if prefix == b"" or t == prefix or t.startswith(prefix + b" "):
            rows.extend(ps)
    return _mask_words(rows, n_rows)


@pytest.mark.parametrize("lowercase", [True, False])
def test_text_build_small_matches_host_and_trie(tss, orc, lowercase):
    phrases = [b"Brown v. Board of Education", b"  brown   V.\tboard ", b"Miranda v. Arizona", b"",
               b"   ", b"Roe v. Wade", b"ROE V. WADE", b"roe", b"\xc3\x89t\xc3\xa9 v. \xc3\xa9t\xc3\xa9",
               b"347 U.S. 483 (1954)", b"a", b"A", b"ab", b"a b", b"a\nb", b"x" * 128]
    rows = np.arange(len(phrases), dtype=np.uint32) * 3
    t = tss.Terms.build_text(phrases, rows, lowercase=lowercase, max_tokens=8)
    terms, posts = t.export()
    want_terms, want_posts = _host_build(phrases, rows, lowercase)
    assert terms == want_terms and posts == want_posts
    # the literal trie: case-name trie lower-cases, citation trie keeps case (same tokenisation)
    trie = orc.Trie()
    which = orc.TRIE_CASE_NAME if lowercase else orc.TRIE_CITATION
    for p, r in zip(phrases, rows):
        if not _py_tokens(p, lowercase):
            continue  # the wrappers never insert an empty token list; the root term is ours
        try:
            s = p.decode()
        except UnicodeDecodeError:
            continue
        if lowercase:
            cid = int(r).to_bytes(16, "little")
            trie.insert_case_name(s, cid)
        else:
            trie.insert_citation(s, orc.docref(b"\0" * 16, int(r), -1))
    n_rows = int(rows.max()) + 1
    m = tss.Mask(n_rows)
    for prefix in [b"brown", b"brown v.", b"roe", b"Roe", b"ROE V.", b"a", b"a b", b"x" * 128, b"347 U.S."]:
        t.prefix_mask(prefix, m, fresh=True)
        assert np.array_equal(m.download(), _expected(want_terms, want_posts, prefix, n_rows)), prefix
        try:
            got_trie = trie.prefix_postings(which, prefix.decode())
        except UnicodeDecodeError:
            continue
        if lowercase:
            trie_rows = {int.from_bytes(ref[0], "little") for ref in got_trie}
        else:
            trie_rows = {ref[1] for ref in got_trie}
        # the trie search folds the QUERY too (case-name) -- compare on the folded prefix
        folded = b" ".join(_py_tokens(prefix, lowercase))
        assert np.array_equal(_mask_words(trie_rows, n_rows),
                              _expected(want_terms, want_posts, folded, n_rows)), prefix


def test_text_build_one_million_postings(tss, orc):
    """>= 1M postings: device build == host construction; masks == numpy; a 20k-phrase slice is
    also cross-checked against the literal trie."""
    rng = np.random.default_rng(77)
    n, n_rows = 1_050_000, 800_000
    vocab = np.array([b"W%05d" % i if i % 3 else b"w%05d" % i for i in range(30_000)], dtype=object)
    ntok = rng.integers(1, 5, n)
    tok = (rng.zipf(1.15, size=(n, 4)) - 1) % len(vocab)
    seps = [b" ", b"  ", b"\t", b" \n "]
    phrases = []
    for i in range(n):
        parts = [vocab[tok[i, j]] for j in range(ntok[i])]
        s = seps[i & 3].join(parts)
        phrases.append(s if i % 5 else b" " + s + b"  ")
    rows = rng.integers(0, n_rows, n).astype(np.uint32)
    t = tss.Terms.build_text(phrases, rows, lowercase=True, max_tokens=4)
    want_terms, want_posts = _host_build(phrases, rows, True)
    got_terms, got_posts = t.export()
    assert got_terms == want_terms
    assert got_posts == want_posts
    # prefix masks vs numpy on the host structure
    import bisect
    flat = np.concatenate([np.asarray(p, dtype=np.int64) for p in want_posts])
    off = np.zeros(len(want_terms) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in want_posts], out=off[1:])
    m = tss.Mask(n_rows)
    for prefix in [b"w00000", b"w00001", b"w00000 w00001", b"w00002 w00000 w00001", b"w29999", b"zz", b""]:
        t.prefix_mask(prefix, m, fresh=True)
        if prefix == b"":
            r = flat
        else:
            lo = bisect.bisect_left(want_terms, prefix)
            hi = lo + (1 if lo < len(want_terms) and want_terms[lo] == prefix else 0)
            slo = bisect.bisect_left(want_terms, prefix + b" ")
            shi = bisect.bisect_left(want_terms, prefix + b"!")
            r = np.concatenate([flat[off[lo]:off[hi]], flat[off[slo]:off[shi]]])
        assert np.array_equal(m.download(), _mask_words(np.unique(r), n_rows)), prefix
    # the literal trie on a slice
    ns = 20_000
    small = tss.Terms.build_text(phrases[:ns], rows[:ns], lowercase=True, max_tokens=4)
    trie = orc.Trie()
    for p, r in zip(phrases[:ns], rows[:ns]):
        trie.insert_case_name(p.decode(), int(r).to_bytes(16, "little"))
    for prefix in [b"w00000", b"W00001", b"w00000 w00001", b"w00004"]:
        small.prefix_mask(prefix.lower(), m, fresh=True)
        trie_rows = {int.from_bytes(ref[0], "little")
                     for ref in trie.prefix_postings(orc.TRIE_CASE_NAME, prefix.decode())}
        assert np.array_equal(m.download(), _mask_words(trie_rows, n_rows)), prefix


def test_text_build_rejects_what_it_cannot_represent(tss):
    with pytest.raises(tss.TssError):  # control byte
        tss.Terms.build_text([b"a\x01b"], [0])
    with pytest.raises(tss.TssError):  # token > 128 bytes
        tss.Terms.build_text([b"y" * 129], [0])
    with pytest.raises(tss.TssError):  # more tokens than max_tokens
        tss.Terms.build_text([b"a b c"], [0], max_tokens=2)
    t = tss.Terms.build_text([], [])
    assert t.size() == 0
    t = tss.Terms.build_text([b"", b" "], [4, 5])
    assert t.export() == ([b""], [[4, 5]])


def test_terms_save_load_round_trip(tss):
    rng = np.random.default_rng(3)
    n_rows = 100_000
    phrases = [b" ".join(b"t%03d" % x for x in rng.integers(0, 300, rng.integers(1, 4))) for _ in range(50_000)]
    rows = rng.integers(0, n_rows, len(phrases)).astype(np.uint32)
    t = tss.Terms.build_text(phrases, rows, max_tokens=3)
    terms, posts = t.export()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "trie.terms")
        t.save(path)
        u = tss.Terms.load(path)
        assert u.export() == (terms, posts)
        m1, m2 = tss.Mask(n_rows), tss.Mask(n_rows)
        for prefix in [b"", b"t000", b"t001 t002", b"t299", b"nope"]:
            t.prefix_mask(prefix, m1, fresh=True)
            u.prefix_mask(prefix, m2, fresh=True)
            want = _expected(terms, posts, prefix, n_rows)
            assert np.array_equal(m1.download(), want) and np.array_equal(m2.download(), want), prefix
        # empty structure
        e = tss.Terms([], [])
        e.save(path)
        assert tss.Terms.load(path).export() == ([], [])
        # truncated / corrupt files are refused
        t.save(path)
        raw = open(path, "rb").read()
        open(path, "wb").write(raw[:-9])
        with pytest.raises(tss.TssError):
            tss.Terms.load(path)
        bad = bytearray(raw)
        # swap two terms' bytes so they are no longer sorted: find the pool (last section)
        bad[-20], bad[-40] = 0xFF, 0x00
        open(path, "wb").write(bytes(bad))
        try:
            v = tss.Terms.load(path)
            ok_terms, _ = v.export()
            assert ok_terms == sorted(set(ok_terms))  # accepted only if still strictly sorted
        except tss.TssError:
            pass
        open(path, "wb").write(b"NOTATRIE" + raw[8:])
        with pytest.raises(tss.TssError):
            tss.Terms.load(path)
