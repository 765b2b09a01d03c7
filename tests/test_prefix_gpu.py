"""K4 parity: prefix-match kernel + posting scatter vs the literal trie restatement.

Masks must match bit-exactly (north_star).  The flattened term array is what
TrieNode trees export to: unique terms (tokens joined by ' '), byte-sorted, CSR postings.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ID = [bytes([i]) * 16 for i in range(8)]


def _flatten(term_postings):
    """dict term(bytes) -> [rows]  ->  (sorted terms, postings)"""
    terms = sorted(term_postings)
    return terms, [term_postings[t] for t in terms]


def _bits(rows, n):
    w = np.zeros((n + 31) // 32, dtype=np.uint32)
    rows = np.asarray(sorted(set(int(r) for r in rows if 0 <= r < n)), dtype=np.int64)
    if rows.size:
        np.bitwise_or.at(w, rows >> 5, (np.uint32(1) << (rows & 31).astype(np.uint32)))
    return w


def test_k9_simple_demo_corpus(tss, orc):
    # case-name trie of the three simple_demo.rs cases; row i <-> case i
    names = ["brown v. board of education", "miranda v. arizona", "roe v. wade"]
    terms, posts = _flatten({n.encode(): [i] for i, n in enumerate(names)})
    t = tss.Terms(terms, posts)
    assert t.size() == 3
    m = tss.Mask(3)

    def mask_of(prefix, kind=tss.TSS_PREFIX_TOKEN):
        m.clear()
        st = t.prefix_mask(prefix, m, kind)
        return int(m.download()[0]), st

    assert mask_of(b"brown")[0] == 0b001              # K9
    assert mask_of(b"brown v.")[0] == 0b001           # K2's completion, as postings
    assert mask_of(b"bro")[0] == 0                    # K3: token-level, not char-level
    assert mask_of(b"bro", tss.TSS_PREFIX_CHAR)[0] == 0b001
    assert mask_of(b"brown board")[0] == 0            # K4
    bits, st = mask_of(b"roe v. wade")                # K1: exact node, no subtree
    assert bits == 0b100 and st.exact_hi - st.exact_lo == 1 and st.sub_hi == st.sub_lo
    bits, st = mask_of(b"")                           # K8/K9: root -> every posting
    assert bits == 0b111 and (st.sub_lo, st.sub_hi) == (0, 3) and st.npostings == 3
    assert mask_of(b"zzz")[0] == 0 and mask_of(b"a")[0] == 0
    assert mask_of(b"m", tss.TSS_PREFIX_CHAR)[0] == 0b010


def _random_terms(rng, nterms, vocab, n_rows):
    tp = {}
    zipf = rng.zipf(1.3, size=nterms * 4) % vocab
    zi = 0
    while len(tp) < nterms:
        ntok = int(rng.integers(1, 5))
        toks = [f"w{zipf[(zi + j) % zipf.size]:05d}" for j in range(ntok)]
        zi += ntok
        term = " ".join(toks).encode()
        if term in tp:
            continue
        tp[term] = [int(r) for r in rng.integers(0, n_rows, size=int(rng.geometric(0.25)))]
    return tp


def test_token_prefix_masks_match_trie(tss, orc):
    rng = np.random.default_rng(11)
    n_rows = 40_000
    tp = _random_terms(rng, 30_000, 600, n_rows)
    terms, posts = _flatten(tp)
    trie = orc.Trie()
    for term, rows in tp.items():  # citation trie: case-preserving, whitespace tokenised
        for r in rows:
            trie.insert_citation(term.decode(), orc.docref(ID[0], r, -1))
    t = tss.Terms(terms, posts)
    m = tss.Mask(n_rows)
    prefixes = [b"", b"w00001", b"w00001 w00002", b"w00000", b"w00003 w00001 w00000", b"w0000",
                b"nope", b"w00001 ", terms[0], terms[-1], terms[len(terms) // 2],
                terms[7].split(b" ")[0], b"w00002  w00001"]
    prefixes += [b" ".join(terms[i].split(b" ")[:2]) for i in range(0, len(terms), 3001)]
    for p in prefixes:
        m.clear()
        st = t.prefix_mask(p, m)
        want_refs = trie.prefix_postings(orc.TRIE_CITATION, p.decode())
        want = _bits([ref[1] for ref in want_refs], n_rows)
        # the ABI takes the prefix already joined by single spaces (the shim normalises);
        # un-normalised input is simply a different byte string
        if p != b" ".join(p.split()):
            continue
        assert np.array_equal(m.download(), want), p
        assert st.npostings == len(want_refs), p
        assert m.popcount() == int(np.unpackbits(want.view(np.uint8)).sum())


def test_char_prefix_masks_match_bruteforce(tss):
    rng = np.random.default_rng(12)
    n_rows = 10_000
    tp = _random_terms(rng, 5_000, 300, n_rows)
    tp[b"\xff\xff"] = [1]
    tp[b"\xff\xffz"] = [2]
    terms, posts = _flatten(tp)
    t = tss.Terms(terms, posts)
    m = tss.Mask(n_rows)
    for p in [b"w", b"w0", b"w001", b"w00012 w", b"x", b"", b"\xff", b"\xff\xff", terms[10][:7]]:
        m.clear()
        t.prefix_mask(p, m, tss.TSS_PREFIX_CHAR)
        want = _bits([r for term, rows in tp.items() if term.startswith(p) for r in rows], n_rows)
        assert np.array_equal(m.download(), want), p


def test_sharded_mask_and_accumulation(tss):
    tp = {b"a b": [0, 5, 99], b"a c": [100, 150], b"d": [199, 3]}
    terms, posts = _flatten(tp)
    t = tss.Terms(terms, posts)
    lo, hi = tss.Mask(100), tss.Mask(100)
    t.prefix_mask(b"a", lo, row_base=0)
    t.prefix_mask(b"a", hi, row_base=100)
    assert np.array_equal(lo.download(), _bits([0, 5, 99], 100))
    assert np.array_equal(hi.download(), _bits([0, 50], 100))
    # masks accumulate until cleared (OR semantics)
    t.prefix_mask(b"d", lo, row_base=0)
    assert np.array_equal(lo.download(), _bits([0, 5, 99, 3], 100))


def test_hybrid_prefix_filtered_topk(tss, orc):
    """config-4 shape, reduced: prefix -> mask -> masked top-10 == oracle on the same mask."""
    rng = np.random.default_rng(13)
    n_rows, dim = 120_000, 384
    tp = _random_terms(rng, 20_000, 400, n_rows)
    terms, posts = _flatten(tp)
    t = tss.Terms(terms, posts)
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n_rows, 0x5EED)
    ix.finalize()
    rows = orc.gen_rows(0, n_rows, dim, 0x5EED)
    q = orc.gen_rows(0, 2, dim, 0xBEEF)
    m = tss.Mask(n_rows)
    for p in [b"w00001", b"w00000 w00001", b""]:
        m.clear()
        t.prefix_mask(p, m)
        words = m.download()
        got = ix.search(q, 10, m, tss.TSS_MASK_INCLUDE)
        want = orc.cosine_topk(rows, q, 10, words, orc.MASK_INCLUDE)
        for g, w in zip(got, want):
            assert np.array_equal(g, w), p
        got = ix.search(q, 10, m, tss.TSS_MASK_EXCLUDE)
        want = orc.cosine_topk(rows, q, 10, words, orc.MASK_EXCLUDE)
        for g, w in zip(got, want):
            assert np.array_equal(g, w), p


def test_terms_edge_cases(tss):
    t = tss.Terms([], [])
    m = tss.Mask(10)
    st = t.prefix_mask(b"x", m)
    assert m.popcount() == 0 and st.npostings == 0
    one = tss.Terms([b"solo"], [[4, 4, 4]])  # duplicate postings are kept, the bit is set once
    st = one.prefix_mask(b"solo", m)
    assert m.popcount() == 1 and st.npostings == 3
    with pytest.raises(tss.TssError):
        tss.Terms([b"b", b"a"], [[], []])  # unsorted
    with pytest.raises(tss.TssError):
        tss.Terms([b"a", b"a"], [[], []])  # duplicate


def test_mask_ordering_stress_idle_then_clear_prefix(tss):
    """Round 1's GPUTEST failure: tss_mask_clear ran on the legacy stream and the scatter on the
    terms' non-blocking stream, so after the GPU had idled a queued memset could land on top of
    the scatter's bits.  Masks are now ordered by events across whatever streams touch them:
    idle the GPU, then clear(); clear(); prefix_mask(built); prefix_mask(uploaded), 200 times,
    every mask compared with numpy -- never with the other mask."""
    import time
    rng = np.random.default_rng(5)
    n_rows = 300_000
    vocab = sorted({b"w%04d" % i for i in range(400)})
    n = 120_000
    L = 3
    ntok = rng.integers(1, L + 1, n)
    ids = (rng.zipf(1.3, size=(n, L)) % len(vocab) + 1).astype(np.uint32)
    ids[np.arange(L)[None, :] >= ntok[:, None]] = 0
    rows = rng.integers(0, n_rows, n).astype(np.uint32)
    built = tss.Terms.build(vocab, ids, rows)
    terms, posts = built.export()
    uploaded = tss.Terms(terms, posts)
    # expectations: term ranges by bisect on the sorted terms, rows by numpy
    import bisect
    flat = np.concatenate([np.asarray(p, dtype=np.int64) for p in posts])
    off = np.zeros(len(terms) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in posts], out=off[1:])

    def expected(prefix):
        w = np.zeros((n_rows + 31) // 32, dtype=np.uint32)
        if prefix == b"":
            r = flat
        else:
            lo = bisect.bisect_left(terms, prefix)
            hi_exact = lo + (1 if lo < len(terms) and terms[lo] == prefix else 0)
            slo = bisect.bisect_left(terms, prefix + b" ")
            shi = bisect.bisect_left(terms, prefix + b"!")
            r = np.concatenate([flat[off[lo]:off[hi_exact]], flat[off[slo]:off[shi]]])
        if r.size:
            np.bitwise_or.at(w, r >> 5, (np.uint32(1) << (r & 31).astype(np.uint32)))
        return w

    prefixes = [b"", vocab[0], vocab[1], vocab[0] + b" " + vocab[0], vocab[7], b"nope", vocab[3] + b" " + vocab[1]]
    want = {p: expected(p) for p in prefixes}
    m1, m2 = tss.Mask(n_rows), tss.Mask(n_rows)
    full = np.full((n_rows + 31) // 32, 0xFFFFFFFF, dtype=np.uint32)
    for it in range(200):
        p = prefixes[it % len(prefixes)]
        if it % 50 == 0:
            time.sleep(1.5)  # the GPU idles: queued work from different streams starts together
        if it % 3 == 0:      # dirty masks, so a lost clear shows as well as a lost bit
            m1.upload(full)
            m2.upload(full)
        m1.clear(); m1.clear()
        m2.clear(); m2.clear()
        built.prefix_mask(p, m1, want_stats=False)
        uploaded.prefix_mask(p, m2, want_stats=False)
        a, b = m1.download(), m2.download()
        assert np.array_equal(a, want[p]), (it, p)
        assert np.array_equal(b, want[p]), (it, p)
    # the one-enqueue form (clear folded into the search kernel) on dirty masks, no host sync
    for it in range(100):
        p = prefixes[it % len(prefixes)]
        m1.upload(full)
        built.prefix_mask(p, m1, want_stats=False, fresh=True)
        assert np.array_equal(m1.download(), want[p]), (it, p)
        assert m1.popcount() == int(np.unpackbits(want[p].view(np.uint8)).sum())


def test_prefix_then_masked_search_is_ordered(tss, orc):
    """prefix mask on the terms' stream -> masked search on the index stream -> next query's
    clear: no host synchronisation anywhere in between, results == oracle with the numpy mask."""
    rng = np.random.default_rng(9)
    n, dim, k = 60_000, 128, 10
    rows = orc.gen_rows(0, n, dim, 0x5EED)
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, 0x5EED)
    ix.finalize()
    terms = [b"t%03d" % i for i in range(64)]
    posts = [sorted(set(int(r) for r in rng.integers(0, n, size=int(rng.integers(1, 4000))))) for _ in terms]
    t = tss.Terms(terms, posts)
    m = tss.Mask(n)
    q = orc.gen_rows(0, 40, dim, 0xBEEF)
    for bound in (False, True):
        if bound:
            t.bind_stream(ix)
        for i in range(40):
            p = terms[(i * 7) % len(terms)]
            t.prefix_mask(p, m, want_stats=False, fresh=True)
            got = ix.search(q[i], k, m, tss.TSS_MASK_INCLUDE)
            w = np.zeros((n + 31) // 32, dtype=np.uint32)
            idx = np.asarray(posts[(i * 7) % len(terms)], dtype=np.int64)
            np.bitwise_or.at(w, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
            want = orc.cosine_topk(rows, q[i], k, mask_words=w, mask_mode=orc.MASK_INCLUDE)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2]), (bound, i)
    t.bind_stream(None)


def test_list_driven_scan_and_its_invalidation(tss, orc):
    """tss_prefix_mask_fresh also leaves the list of its unique rows; the masked scan then fetches
    exactly those rows.  Whatever path runs -- list, mask walk after the list overflowed, mask walk
    after another call edited the mask -- the result equals the oracle under the numpy mask."""
    rng = np.random.default_rng(21)
    n, dim, k = 200_000, 384, 10
    rows = orc.gen_rows(0, n, dim, 0x5EED)
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, 0x5EED)
    ix.finalize()
    sizes = [1, 7, 8, 9, 100, 5_000, 16_385, 40_000, 131_072, 131_073, 150_000]   # postings; the list holds 131 072
    terms = [b"p%02d" % i for i in range(len(sizes))]
    posts = [[int(r) for r in rng.integers(0, n, size=s)] for s in sizes]  # duplicates included
    t = tss.Terms(terms, posts)
    t.bind_stream(ix)
    m = tss.Mask(n)
    q = orc.gen_rows(0, 4, dim, 0xBEEF)

    def words(rs):
        w = np.zeros((n + 31) // 32, dtype=np.uint32)
        idx = np.asarray(sorted(set(rs)), dtype=np.int64)
        if idx.size:
            np.bitwise_or.at(w, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
        return w

    def check(w, tag):
        got = ix.search(q, k, m, tss.TSS_MASK_INCLUDE)
        want = orc.cosine_topk(rows, q, k, mask_words=w, mask_mode=orc.MASK_INCLUDE)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2]), tag
        assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)), tag
        assert np.array_equal(m.download(), w), tag

    for term, ps in zip(terms, posts):
        t.prefix_mask(term, m, want_stats=False, fresh=True)
        check(words(ps), term)
        # an edit after the scatter: the list no longer describes the mask
        extra = [int(r) for r in rng.integers(0, n, size=5)]
        m.set_rows(np.asarray(extra, dtype=np.uint32))
        check(words(ps + extra), (term, "set_rows"))
        m.clear_rows(np.asarray(ps[:1], dtype=np.uint32))
        check(words([r for r in ps + extra if r != ps[0]]), (term, "clear_rows"))
        # OR-ing a second prefix into the mask
        t.prefix_mask(term, m, want_stats=False, fresh=True)
        t.prefix_mask(terms[2], m, want_stats=False)
        check(words(ps + posts[2]), (term, "or"))
    # the root prefix: every posting
    t.prefix_mask(b"", m, want_stats=False, fresh=True)
    check(words([r for ps in posts for r in ps]), "root")
    # EXCLUDE never uses the list
    t.prefix_mask(terms[4], m, want_stats=False, fresh=True)
    got = ix.search(q, k, m, tss.TSS_MASK_EXCLUDE)
    want = orc.cosine_topk(rows, q, k, mask_words=words(posts[4]), mask_mode=orc.MASK_EXCLUDE)
    assert np.array_equal(got[0], want[0])
    # bf16 storage and a shard offset
    ixb = tss.FlatIndex(dim, tss.TSS_BF16)
    ixb.add_synthetic(50_000, 100_000, 0x5EED)
    ixb.set_shard(50_000, None)
    ixb.finalize()
    t.bind_stream(ixb)
    mb = tss.Mask(100_000)
    for term, ps in zip(terms[3:7], posts[3:7]):
        t.prefix_mask(term, mb, want_stats=False, fresh=True, row_base=50_000)
        local = [r - 50_000 for r in ps if 50_000 <= r < 150_000]
        w = np.zeros((100_000 + 31) // 32, dtype=np.uint32)
        idx = np.asarray(sorted(set(local)), dtype=np.int64)
        if idx.size:
            np.bitwise_or.at(w, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
        got = ixb.search(q, k, mb, tss.TSS_MASK_INCLUDE)
        want = orc.cosine_topk(rows[50_000:150_000], q, k, mask_words=w, mask_mode=orc.MASK_INCLUDE,
                               row_base=50_000, bf16=True)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2]), term
        assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)), term
    t.bind_stream(None)
