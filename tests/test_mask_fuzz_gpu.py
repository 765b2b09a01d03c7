"""Stateful fuzz of everything that writes or reads a mask, against a numpy model.

Masks are touched from four kinds of streams (the mask's own, a terms handle's, a columns
handle's, an index's) and ordered only by the events they carry (include/tss.h, masks).  A random
sequence of clear / set_rows / clear_rows / upload / prefix_mask (fresh or OR-ing, from a terms
handle on its own stream or bound to the index) / filter_mask (overwrite or AND) / masked search
(INCLUDE or EXCLUDE, batch 1-4) / download / popcount must behave exactly like the same sequence
applied to a boolean array -- including which searches may use the prefix row list."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _words(bits):
    w = np.zeros((bits.size + 31) // 32, dtype=np.uint32)
    idx = np.nonzero(bits)[0]
    if idx.size:
        np.bitwise_or.at(w, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
    return w


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_mask_op_sequences(tss, orc, seed):
    rng = np.random.default_rng(seed)
    n, dim, k = 50_000, 128, 10
    rows = orc.gen_rows(0, n, dim, 0x5EED)
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, 0x5EED)
    ix.finalize()
    q = orc.gen_rows(0, 16, dim, 0xBEEF)
    sizes = [1, 5, 300, 5_000, 16_384, 140_000, 1]
    terms = [b"t%02d" % i for i in range(len(sizes))] + [b"t00 x", b"t01 y z"]
    posts = [[int(r) for r in rng.integers(0, n, size=s)] for s in sizes] + [[7, 8, 9], [n - 1]]
    order = np.argsort(np.array(terms, dtype=object))
    terms = [terms[i] for i in order]
    posts = [posts[i] for i in order]
    t_own = tss.Terms(terms, posts)      # scatters on its own stream
    t_bound = tss.Terms(terms, posts)    # scatters on the index's stream
    t_bound.bind_stream(ix)
    court = rng.integers(0, 20, n).astype(np.uint16)
    date = rng.integers(0, 1000, n).astype(np.int32)
    cols = tss.Columns(court, date)
    m = tss.Mask(n)
    model = np.zeros(n, dtype=bool)

    def prefix_rows(p):
        out = []
        for t, ps in zip(terms, posts):
            if p == b"" or t == p or t.startswith(p + b" "):
                out.extend(ps)
        return np.asarray(sorted(set(out)), dtype=np.int64)

    for step in range(160):
        op = rng.integers(0, 10)
        if op == 0:
            m.clear()
            model[:] = False
        elif op == 1:
            r = rng.integers(0, n, size=int(rng.integers(1, 2000))).astype(np.uint32)
            m.set_rows(r)
            model[r] = True
        elif op == 2:
            r = rng.integers(0, n, size=int(rng.integers(1, 2000))).astype(np.uint32)
            m.clear_rows(r)
            model[r] = False
        elif op == 3:
            bits = rng.random(n) < rng.choice([0.0, 0.001, 0.3, 1.0])
            m.upload(_words(bits))
            model[:] = bits
        elif op in (4, 5, 6):
            p = [b"", b"t00", b"t01", b"t02", b"t03", b"t04", b"t05", b"t06", b"nope", b"t01 y"][int(rng.integers(0, 10))]
            fresh = bool(rng.integers(0, 2))
            (t_bound if op == 4 else t_own).prefix_mask(p, m, want_stats=bool(rng.integers(0, 2)), fresh=fresh)
            if fresh:
                model[:] = False
            model[prefix_rows(p)] = True
        elif op == 7:
            allowed = [int(c) for c in rng.integers(0, 20, size=int(rng.integers(0, 4)))]
            lo, hi = sorted(int(x) for x in rng.integers(0, 1000, size=2))
            combine = bool(rng.integers(0, 2))
            cols.filter_mask(m, allowed, lo, hi, combine_and=combine)
            ok = (date >= lo) & (date <= hi)
            if allowed:
                ok &= np.isin(court, allowed)
            model[:] = (model & ok) if combine else ok
        elif op == 8:
            if rng.integers(0, 2):
                assert np.array_equal(m.download(), _words(model)), step
            else:
                assert m.popcount() == int(model.sum()), step
        # every step ends with a search under the current mask (the list-driven path is live
        # right after a fresh selective prefix and must not be after anything else)
        nq = int(rng.integers(1, 5))
        mode = tss.TSS_MASK_INCLUDE if rng.integers(0, 3) else tss.TSS_MASK_EXCLUDE
        j = int(rng.integers(0, 12))
        got = ix.search(q[j:j + nq], k, m, mode)
        want = orc.cosine_topk(rows, q[j:j + nq], k, mask_words=_words(model),
                               mask_mode=orc.MASK_INCLUDE if mode == tss.TSS_MASK_INCLUDE else orc.MASK_EXCLUDE)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2]), (step, op)
        assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)), (step, op)
    assert np.array_equal(m.download(), _words(model))
    t_bound.bind_stream(None)
