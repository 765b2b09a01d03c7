"""N2: device-side construction of the flattened trie == the host construction (unique terms in
byte order, postings per term in insertion order), and prefix masks from it match the trie."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _host_build(vocab, ids, rows):
    d = {}
    for tup, r in zip(ids, rows):
        term = b" ".join(vocab[i - 1] for i in tup if i)
        d.setdefault(term, []).append(int(r))
    terms = sorted(d)
    return terms, [d[t] for t in terms]


def _expected_mask(terms, posts, prefix, n_rows):
    """numpy/python restatement of the prefix posting set: the term equal to the token prefix and
    every term below it (trie.rs:223-278 without the limit), as mask words"""
    rows = []
    for t, ps in zip(terms, posts):
        if prefix == b"" or t == prefix or t.startswith(prefix + b" "):
            rows.extend(ps)
    w = np.zeros((n_rows + 31) // 32, dtype=np.uint32)
    idx = np.array(sorted({r for r in rows if r < n_rows}), dtype=np.int64)
    if idx.size:
        np.bitwise_or.at(w, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    return w


def test_device_build_matches_host_build(tss, orc):
    rng = np.random.default_rng(31)
    vocab = sorted({b"w%05d" % i for i in rng.integers(0, 100000, 3000)} |
                   {b"a", b"ab", b"ab!", b"abc", b"z~", b"\xc3\xa9t\xc3\xa9"})
    V, L, n, n_rows = len(vocab), 5, 60_000, 50_000
    ntok = rng.integers(0, L + 1, n)
    ntok[:3] = 0  # the root term (zero tokens)
    ids = (rng.zipf(1.2, size=(n, L)) % V + 1).astype(np.uint32)
    ids[np.arange(L)[None, :] >= ntok[:, None]] = 0
    rows = rng.integers(0, n_rows, n).astype(np.uint32)
    t = tss.Terms.build(vocab, ids, rows)
    got_terms, got_posts = t.export()
    want_terms, want_posts = _host_build(vocab, ids, rows)
    assert got_terms == want_terms
    assert got_posts == want_posts  # insertion order inside a term (stable sort)
    assert t.size() == len(want_terms) and got_terms[0] == b""
    # the built object behaves like an uploaded one
    ref = tss.Terms(want_terms, want_posts)
    m1, m2 = tss.Mask(n_rows), tss.Mask(n_rows)
    for p in [b"", want_terms[5], want_terms[100].split(b" ")[0], b"ab", b"a", b"nope",
              b" ".join(want_terms[-1].split(b" ")[:2])]:
        m1.clear(); m2.clear()
        s1 = t.prefix_mask(p, m1)
        s2 = ref.prefix_mask(p, m2)
        want = _expected_mask(want_terms, want_posts, p, n_rows)  # host-computed, not each other
        assert np.array_equal(m1.download(), want), p
        assert np.array_equal(m2.download(), want), p
        assert (s1.exact_lo, s1.exact_hi, s1.sub_lo, s1.sub_hi, s1.npostings) == \
               (s2.exact_lo, s2.exact_hi, s2.sub_lo, s2.sub_hi, s2.npostings)
    # and like the reference's trie: same prefix posting set as the literal restatement
    trie = orc.Trie()
    for tup, r in zip(ids[:5000], rows[:5000]):
        toks = [vocab[i - 1].decode() for i in tup if i]
        if toks:
            trie.insert_citation(" ".join(toks), orc.docref(b"\0" * 16, int(r), -1))
    small = tss.Terms.build(vocab, ids[:5000], rows[:5000])
    m = tss.Mask(n_rows)
    for p in [want_terms[7].split(b" ")[0], b"ab"]:
        m.clear()
        small.prefix_mask(p, m)
        want_rows = {ref_[1] for ref_ in trie.prefix_postings(orc.TRIE_CITATION, p.decode())}
        w = np.zeros((n_rows + 31) // 32, dtype=np.uint32)
        idx = np.array(sorted(want_rows), dtype=np.int64)
        if idx.size:
            np.bitwise_or.at(w, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
        assert np.array_equal(m.download(), w), p


def test_build_edge_cases_and_validation(tss):
    t = tss.Terms.build([b"x"], np.zeros((0, 3), np.uint32), np.zeros(0, np.uint32))
    assert t.size() == 0 and t.export() == ([], [])
    t = tss.Terms.build([b"b", b"c"], np.array([[2, 1], [1, 0], [2, 1], [1, 0]], np.uint32), [9, 8, 7, 6])
    assert t.export() == ([b"b", b"c b"], [[8, 6], [9, 7]])
    with pytest.raises(tss.TssError):  # vocabulary not sorted
        tss.Terms.build([b"b", b"a"], np.array([[1]], np.uint32), [0])
    with pytest.raises(tss.TssError):  # token with a space
        tss.Terms.build([b"a b"], np.array([[1]], np.uint32), [0])
    with pytest.raises(tss.TssError):  # id beyond the vocabulary
        tss.Terms.build([b"a"], np.array([[2]], np.uint32), [0])
    with pytest.raises(tss.TssError):  # hole in a tuple
        tss.Terms.build([b"a"], np.array([[0, 1]], np.uint32), [0])
