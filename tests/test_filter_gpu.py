"""N3: device-side metadata pre-filter (court / date columns -> include mask) vs numpy, and the
filtered search vs the oracle on the same mask.  Semantics = apply_filters' conjunction
(reference src/search.rs:255-274), applied BEFORE top-k instead of after it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _words(bits):
    n = bits.size
    w = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.nonzero(bits)[0]
    np.bitwise_or.at(w, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    return w


def test_filter_mask_matches_numpy_and_search(tss, orc):
    rng = np.random.default_rng(21)
    n, dim = 70_007, 384
    court = rng.integers(0, 300, n).astype(np.uint16)
    court[::1000] = 65535  # extreme id
    date = rng.integers(-5000, 20000, n).astype(np.int32)
    cols = tss.Columns(court, date)
    m = tss.Mask(n)
    rows = orc.gen_rows(0, n, dim, 0x5EED)
    ix = tss.FlatIndex(dim)
    ix.add(rows)
    ix.finalize()
    q = orc.gen_rows(0, 2, dim, 0xBEEF)
    cases = [
        ([], -2**31, 2**31 - 1),
        ([3, 17, 65535], -2**31, 2**31 - 1),
        ([], 0, 9999),
        ([5], 100, 15000),
        ([299], 19999, 19999),
        ([1], 30000, 40000),  # empty
    ]
    for allowed, lo, hi in cases:
        cols.filter_mask(m, allowed, lo, hi)
        want_bits = (date >= lo) & (date <= hi)
        if allowed:
            want_bits &= np.isin(court, np.array(allowed, dtype=np.uint16))
        assert np.array_equal(m.download(), _words(want_bits)), (allowed, lo, hi)
        got = ix.search(q, 10, m, tss.TSS_MASK_INCLUDE)
        want = orc.cosine_topk(rows, q, 10, _words(want_bits), orc.MASK_INCLUDE)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
    # combine with another mask (AND), then knock out "seen" rows
    pre = rng.random(n) < 0.4
    m.upload(_words(pre))
    cols.filter_mask(m, [3, 17], 0, 15000, combine_and=True)
    want_bits = pre & (date >= 0) & (date <= 15000) & np.isin(court, np.array([3, 17], dtype=np.uint16))
    assert np.array_equal(m.download(), _words(want_bits))
    seen = np.nonzero(want_bits)[0][:50].astype(np.uint32)
    m.clear_rows(seen)
    want_bits[seen] = False
    assert np.array_equal(m.download(), _words(want_bits))
    with pytest.raises(tss.TssError):
        cols.filter_mask(tss.Mask(10), [], 0, 1)
