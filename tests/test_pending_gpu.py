"""Searches in flight: tss_index_search_submit / _collect (one thread pipelining), blocking
searches of several threads overlapping on one handle, and the one-call hybrid query
tss_index_search_prefix -- every result against the CPU oracle, bit for bit.

Reference seam: HnswIndex::search (src/vector.rs:195-202) is an `async fn` behind a process-wide
write lock (src/search.rs:249-252); SURVEY section 8(b) asks for a re-entrant search with
per-call workspaces instead.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0x5EED


def _bufs(nq, k):
    return (np.full((nq, k), 7, np.uint32), np.full((nq, k), 7, np.float32), np.full(nq, 7, np.uint32))


def _same(got, want, sl):
    gr, gs, gc = got
    wr, ws, wc = want
    return (np.array_equal(gc, wc[sl]) and np.array_equal(gr, wr[sl]) and
            np.array_equal(gs.view(np.uint32), ws[sl].view(np.uint32)))


@pytest.mark.parametrize("dim,k", [(384, 10), (100, 33), (768, 128)])
def test_submit_collect_pipeline_matches_oracle(tss, orc, dim, k):
    n = 50_003
    rows = orc.gen_rows(0, n, dim, SEED)
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, SEED)
    ix.finalize()
    q = np.ascontiguousarray(orc.gen_rows(0, 40, dim, 0xBEEF))
    want = orc.cosine_topk(rows, q, k)
    # sizes 1..4, kept TSS_MAX_PENDING deep, collected in submission order
    plan, at = [], 0
    while at < 40:
        nq = min(1 + (len(plan) % 4), 40 - at)
        plan.append((at, nq))
        at += nq
    pending = []
    for (q0, nq) in plan:
        if len(pending) == tss.TSS_MAX_PENDING:
            t, p0, pn, b = pending.pop(0)
            ix.search_collect(t, b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
            assert _same(b, want, slice(p0, p0 + pn)), (p0, pn)
        b = _bufs(nq, k)
        pending.append((ix.search_submit(q[q0].ctypes.data, nq, k), q0, nq, b))
    # the rest out of order: a ticket waits for its own search only
    for t, p0, pn, b in reversed(pending):
        ix.search_collect(t, b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
        assert _same(b, want, slice(p0, p0 + pn)), (p0, pn)
    ix.close()


def test_submit_with_masks_and_rewritten_mask(tss, orc):
    """A mask handed to submit may be rewritten at once: the write is ordered behind the read."""
    n, dim, k = 30_000, 384, 10
    rows = orc.gen_rows(0, n, dim, SEED)
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, SEED)
    ix.finalize()
    q = np.ascontiguousarray(orc.gen_rows(0, 8, dim, 0xBEEF))
    m = tss.Mask(n)
    sets = [np.arange(i, n, 5 + i, dtype=np.uint32) for i in range(4)]

    def words(rs):
        w = np.zeros((n + 31) // 32, dtype=np.uint32)
        idx = rs.astype(np.int64)
        np.bitwise_or.at(w, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
        return w

    tickets = []
    for i, rs in enumerate(sets):
        m.clear()
        m.set_rows(rs)
        mode = tss.TSS_MASK_INCLUDE if i % 2 == 0 else tss.TSS_MASK_EXCLUDE
        tickets.append((ix.search_submit(q[i].ctypes.data, 1, k, m, mode), i, mode, _bufs(1, k)))
    for t, i, mode, b in tickets:
        ix.search_collect(t, b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
        omode = orc.MASK_INCLUDE if mode == tss.TSS_MASK_INCLUDE else orc.MASK_EXCLUDE
        want = orc.cosine_topk(rows, q[i:i + 1], k, mask_words=words(sets[i]), mask_mode=omode)
        assert _same(b, want, slice(0, 1)), i
    ix.close()


def test_submit_limits_and_ticket_errors(tss, orc):
    n, dim, k = 5_000, 384, 10
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, SEED)
    ix.finalize()
    q = np.ascontiguousarray(orc.gen_rows(0, 8, dim, 0xBEEF))
    b = _bufs(4, k)
    for bad in (lambda: ix.search_submit(q.ctypes.data, 5, k),
                lambda: ix.search_submit(q.ctypes.data, 0, k),
                lambda: ix.search_submit(q.ctypes.data, 1, 129)):
        with pytest.raises(tss.TssError) as ei:
            bad()
        assert ei.value.code == tss.TSS_ERR_INVALID_ARG
    qbad = q[:1].copy()
    qbad[0, 3] = np.nan
    with pytest.raises(tss.TssError) as ei:
        ix.search_submit(qbad.ctypes.data, 1, k)
    assert ei.value.code == tss.TSS_ERR_INVALID_ARG
    ts = [ix.search_submit(q[i].ctypes.data, 1, k) for i in range(tss.TSS_MAX_PENDING)]
    with pytest.raises(tss.TssError) as ei:   # a fifth in flight
        ix.search_submit(q[4].ctypes.data, 1, k)
    assert ei.value.code == tss.TSS_ERR_STATE and "pending" in str(ei.value)
    ix.search_collect(ts[0], b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
    with pytest.raises(tss.TssError) as ei:   # collected twice
        ix.search_collect(ts[0], b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
    assert ei.value.code == tss.TSS_ERR_STATE
    with pytest.raises(tss.TssError) as ei:   # not a ticket
        ix.search_collect(0, b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
    assert ei.value.code == tss.TSS_ERR_INVALID_ARG
    t5 = ix.search_submit(q[5].ctypes.data, 1, k)   # the freed slot is reused with a new ticket
    assert t5 != ts[0]
    with pytest.raises(tss.TssError):          # the old ticket of that slot stays dead
        ix.search_collect(ts[0], b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
    # a blocking search while this thread's own four tickets hold every slot must not wait for
    # a slot (nobody else will collect): it takes the serialised path and still answers
    rows = orc.gen_rows(0, n, dim, SEED)
    want = orc.cosine_topk(rows, q, k)
    got = ix.search(q[7], k)
    assert _same(got, want, slice(7, 8))
    got = ix.search(q[4:8], k)
    assert _same(got, want, slice(4, 8))
    ix.search_collect(ts[1], b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
    assert _same((b[0][:1], b[1][:1], b[2][:1]), want, slice(1, 2))
    got = ix.search(q[6], k)
    assert _same(got, want, slice(6, 7))
    for t, i in ((ts[2], 2), (ts[3], 3), (t5, 5)):
        ix.search_collect(t, b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
        assert _same((b[0][:1], b[1][:1], b[2][:1]), want, slice(i, i + 1)), i
    ix.close()


def test_blocking_searches_wait_for_a_free_slot(tss, orc):
    """More host threads than result slots: the extra callers wait inside the library until a
    slot is handed back, nobody fails, every answer is the oracle's."""
    import threading
    n, dim, k = 80_000, 384, 10
    rows = orc.gen_rows(0, n, dim, SEED)
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, SEED)
    ix.finalize()
    q = np.ascontiguousarray(orc.gen_rows(0, 32, dim, 0xBEEF))
    want = orc.cosine_topk(rows, q, k)
    errors = []

    def worker(t):
        try:
            for i in range(60):
                j = (t * 5 + i) % 32
                if not _same(ix.search(q[j], k), want, slice(j, j + 1)):
                    errors.append((t, i))
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(10)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:5]
    ix.close()


def test_search_prefix_is_prefix_mask_then_search(tss, orc):
    n, dim, k = 40_000, 384, 10
    rows = orc.gen_rows(0, n, dim, SEED)
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, SEED)
    ix.finalize()
    rng = np.random.default_rng(3)
    terms = sorted({b"w%03d w%03d" % (a, b) for a, b in rng.integers(0, 40, size=(3000, 2))} |
                   {b"w%03d" % a for a in range(0, 40, 3)})
    postings = [sorted(set(rng.integers(0, n, size=int(rng.integers(1, 9))).tolist())) for _ in terms]
    t = tss.Terms(terms, postings)
    m = tss.Mask(n)
    q = np.ascontiguousarray(orc.gen_rows(0, 6, dim, 0xBEEF))
    for bound in (False, True):
        if bound:
            t.bind_stream(ix)
        for pi, prefix in enumerate((b"w003", b"w004", b"w003 w01", b"nothing", b"")):
            live = sorted({r for term, ps in zip(terms, postings)
                           if term == prefix or term.startswith(prefix + b" ") or prefix == b""
                           for r in ps})
            w = np.zeros((n + 31) // 32, dtype=np.uint32)
            idx = np.asarray(live, dtype=np.int64)
            np.bitwise_or.at(w, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
            nq = 1 + pi % 3
            b = _bufs(nq, k)
            ix.search_prefix_into(t, prefix, m, q[pi].ctypes.data, nq, k, b[0].ctypes.data,
                                  b[1].ctypes.data, b[2].ctypes.data)
            want = orc.cosine_topk(rows, q[pi:pi + nq], k, mask_words=w, mask_mode=orc.MASK_INCLUDE)
            assert _same(b, want, slice(0, nq)), (bound, prefix)
            assert np.array_equal(m.download(), w)      # scratch holds the prefix's row set
    t.bind_stream(None)
    with pytest.raises(tss.TssError):
        ix.search_prefix_into(t, b"w003", tss.Mask(100), q.ctypes.data, 1, k, b[0].ctypes.data,
                              b[1].ctypes.data, b[2].ctypes.data)
    t.close()
    ix.close()


@pytest.mark.parametrize("bound", [True, False])
def test_hybrid_queries_in_flight(tss, orc, bound):
    """tss_index_search_prefix_submit: hybrid queries pipelined, one scratch mask per query in
    flight, every result the oracle's top-k over exactly the prefix's rows.  bound=False leaves the
    prefix searches on the terms' own stream: query i+1's K4 may then run while query i's scan
    does, ordered only by the masks' events."""
    n, dim, k = 60_000, 384, 10
    rows = orc.gen_rows(0, n, dim, SEED)
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, SEED)
    ix.finalize()
    rng = np.random.default_rng(11)
    terms = sorted({b"w%03d w%03d" % (a, b) for a, b in rng.integers(0, 30, size=(2500, 2))})
    postings = [sorted(set(rng.integers(0, n, size=int(rng.integers(1, 12))).tolist())) for _ in terms]
    t = tss.Terms(terms, postings)
    if bound:
        t.bind_stream(ix)
    q = np.ascontiguousarray(orc.gen_rows(0, 24, dim, 0xBEEF))
    masks = [tss.Mask(n) for _ in range(3)]
    prefixes = [b"w%03d" % (i % 30) for i in range(24)]

    def want_for(i):
        live = sorted({r for term, ps in zip(terms, postings) if term.startswith(prefixes[i] + b" ") for r in ps})
        w = np.zeros((n + 31) // 32, dtype=np.uint32)
        idx = np.asarray(live, dtype=np.int64)
        np.bitwise_or.at(w, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
        return orc.cosine_topk(rows, q[i:i + 1], k, mask_words=w, mask_mode=orc.MASK_INCLUDE)

    pending = []
    for i in range(24):
        if len(pending) == 3:
            j, tk, b = pending.pop(0)
            ix.search_collect(tk, b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
            assert _same(b, want_for(j), slice(0, 1)), j
        b = _bufs(1, k)
        tk = ix.search_prefix_submit(t, prefixes[i], masks[i % 3], q[i].ctypes.data, 1, k)
        pending.append((i, tk, b))
    for j, tk, b in pending:
        ix.search_collect(tk, b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data)
        assert _same(b, want_for(j), slice(0, 1)), j
    with pytest.raises(tss.TssError) as ei:    # rejected before the scratch mask is touched
        ix.search_prefix_submit(t, b"w001", masks[0], q.ctypes.data, 5, k)
    assert ei.value.code == tss.TSS_ERR_INVALID_ARG
    t.bind_stream(None)
    t.close()
    ix.close()
