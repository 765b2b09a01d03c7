"""K1 parity: the CUDA scan (through the C ABI) vs the CPU oracle, bit for bit.

fp32 path bar (north_star): top-k row ids bit-exact, scores within 1e-5 relative --
here scores are required to be bit-identical too, because kernel and oracle
share one canonical reduction order.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0x5EED


def _mk_index(tss, rows, storage=None):
    ix = tss.FlatIndex(rows.shape[1], tss.TSS_F32 if storage is None else storage)
    ix.add(rows)
    ix.finalize()
    return ix


def _assert_same(got, want):
    gr, gs, gc = got
    wr, ws, wc = want
    assert np.array_equal(gc, wc), (gc, wc)
    assert np.array_equal(gr, wr), np.argwhere(gr != wr)[:5]
    assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32))


def test_device_generator_matches_oracle(tss, orc):
    for dim in (384, 100, 5, 768):
        ix = tss.FlatIndex(dim)
        ix.add_synthetic(1000, 777, SEED)
        ix.finalize()
        assert ix.size() == 777
        assert np.array_equal(ix.get_rows(0, 777), orc.gen_rows(1000, 777, dim, SEED))


@pytest.mark.parametrize("n", [0, 1, 3, 7, 8, 9, 1000, 2368 * 8 + 5, 100_003])
def test_parity_sizes_k10(tss, orc, n):
    dim = 384
    rows = orc.gen_rows(0, n, dim, SEED) if n else np.zeros((0, dim), np.float32)
    q = orc.gen_rows(5, 3, dim, 0xBEEF)
    ix = tss.FlatIndex(dim)
    if n:
        ix.add(rows)
    ix.finalize()
    for qi in range(3):
        _assert_same(ix.search(q[qi], 10), orc.cosine_topk(rows, q[qi], 10))


@pytest.mark.parametrize("k", [1, 2, 10, 16, 17, 32, 33, 50, 64, 100, 128])
def test_parity_k(tss, orc, k):
    rows = orc.gen_rows(0, 60_000, 384, SEED)
    q = orc.gen_rows(0, 2, 384, 0xBEEF)
    ix = _mk_index(tss, rows)
    _assert_same(ix.search(q, k), orc.cosine_topk(rows, q, k))


@pytest.mark.parametrize("dim", [1, 4, 100, 128, 129, 256, 384, 512, 600, 768, 1000, 1024])
def test_parity_dims(tss, orc, dim):
    rng = np.random.default_rng(dim)
    rows = rng.standard_normal((20_011, dim)).astype(np.float32)
    q = rng.standard_normal((2, dim)).astype(np.float32)
    ix = _mk_index(tss, rows)
    _assert_same(ix.search(q, 10), orc.cosine_topk(rows, q, 10))
    # tolerance-only second opinion: sequential-sum oracle, 1e-5 relative on scores
    _, gs, _ = ix.search(q[0], 10)
    ss = orc.scores(rows, q[0], orc.ORDER_SEQUENTIAL)
    gr, _, _ = ix.search(q[0], 10)
    np.testing.assert_allclose(gs[0], ss[gr[0]], rtol=1e-5, atol=2e-7)


@pytest.mark.parametrize("nq", [1, 2, 3, 4, 5, 9, 33])
def test_parity_batches(tss, orc, nq):
    rows = orc.gen_rows(0, 30_000, 384, SEED)
    q = orc.gen_rows(0, nq, 384, 0xBEEF)
    q[0] = rows[1234] + 0.125 * q[0]  # planted query: a clear winner
    ix = _mk_index(tss, rows)
    got = ix.search(q, 10)
    _assert_same(got, orc.cosine_topk(rows, q, 10))
    assert got[0][0][0] == 1234


def test_ties_duplicates_and_zero_rows(tss, orc):
    rng = np.random.default_rng(3)
    base = rng.standard_normal((500, 384)).astype(np.float32)
    rows = np.concatenate([base, np.zeros((40, 384), np.float32), base, base[:100]])
    q = base[17].copy()
    ix = _mk_index(tss, rows)
    got = ix.search(q, 20)
    _assert_same(got, orc.cosine_topk(rows, q, 20))
    assert list(got[0][0][:3]) == [17, 557, 1057]  # exact ties resolve to ascending row id
    # zero-norm query (what the reference's stub embedding is, src/vector.rs:173): all scores 0
    z = np.zeros(384, np.float32)
    got = ix.search(z, 10)
    _assert_same(got, orc.cosine_topk(rows, z, 10))
    assert list(got[0][0]) == list(range(10)) and np.all(got[1] == 0)


def test_negative_scores_and_small_corpus_padding(tss, orc):
    rng = np.random.default_rng(5)
    rows = rng.standard_normal((6, 384)).astype(np.float32)
    q = -rows[2]
    ix = _mk_index(tss, rows)
    got = ix.search(q, 10)
    _assert_same(got, orc.cosine_topk(rows, q, 10))
    assert got[2][0] == 6 and np.all(got[0][0][6:] == tss.TSS_ROW_NONE)
    assert got[0][0][5] == 2  # the anti-parallel row is last


@pytest.mark.parametrize("mode", ["include", "exclude"])
@pytest.mark.parametrize("density", [0.0, 0.001, 0.3, 0.97, 1.0])
def test_masked_parity(tss, orc, mode, density):
    n = 50_003
    rows = orc.gen_rows(0, n, 384, SEED)
    q = orc.gen_rows(0, 2, 384, 0xBEEF)
    rng = np.random.default_rng(int(density * 1000))
    bits = rng.random(n) < density
    words = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.nonzero(bits)[0]
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    ix = _mk_index(tss, rows)
    m = tss.Mask(n)
    m.upload(words)
    assert m.popcount() == int(bits.sum())
    tm = tss.TSS_MASK_INCLUDE if mode == "include" else tss.TSS_MASK_EXCLUDE
    om = orc.MASK_INCLUDE if mode == "include" else orc.MASK_EXCLUDE
    _assert_same(ix.search(q, 10, m, tm), orc.cosine_topk(rows, q, 10, words, om))
    # set_rows builds the same mask from a row list
    m2 = tss.Mask(n)
    m2.set_rows(idx.astype(np.uint32))
    assert np.array_equal(m2.download(), words)


def test_incremental_add_and_refinalize(tss, orc):
    rows = orc.gen_rows(0, 9000, 384, SEED)
    q = orc.gen_rows(0, 1, 384, 0xBEEF)[0]
    ix = tss.FlatIndex(384)
    ix.add(rows[:10])
    for r in rows[10:20]:
        assert ix.add_vector(r) == ix.size() - 1
    with pytest.raises(tss.TssError) as ei:  # search before finalize
        ix.search(q, 10)
    assert ei.value.code == tss.TSS_ERR_STATE
    ix.finalize()
    _assert_same(ix.search(q, 10), orc.cosine_topk(rows[:20], q, 10))
    ix.add(rows[20:])
    ix.finalize()
    assert ix.size() == 9000
    _assert_same(ix.search(q, 10), orc.cosine_topk(rows, q, 10))


def test_rejects_nan_inf_and_bad_args(tss):
    ix = tss.FlatIndex(384)
    good = np.ones((4, 384), np.float32)
    ix.add(good)
    bad = good.copy()
    bad[2, 100] = np.nan
    with pytest.raises(tss.TssError) as ei:
        ix.add(bad)
    assert ei.value.code == tss.TSS_ERR_INVALID_ARG and ix.size() == 4
    bad[2, 100] = np.inf
    with pytest.raises(tss.TssError):
        ix.add(bad)
    ix.finalize()
    q = np.ones(384, np.float32)
    with pytest.raises(tss.TssError):
        ix.search(q, 0)
    with pytest.raises(tss.TssError):
        ix.search(q, tss.TSS_MAX_K + 1)
    q[3] = np.nan
    with pytest.raises(tss.TssError):
        ix.search(q, 10)
    with pytest.raises(tss.TssError):  # mask shorter than the shard
        ix.search(np.ones(384, np.float32), 10, tss.Mask(2), tss.TSS_MASK_INCLUDE)


def test_bf16_storage_scan_matches_bf16_oracle(tss, orc):
    rows = orc.gen_rows(0, 40_001, 384, SEED)
    q = orc.gen_rows(0, 3, 384, 0xBEEF)
    ix = _mk_index(tss, rows, tss.TSS_BF16)
    _assert_same(ix.search(q, 10), orc.cosine_topk(rows, q, 10, bf16=True))
    u = rows[:100].view(np.uint32)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)
    assert np.array_equal(ix.get_rows(0, 100), rounded)


def test_sharded_row_base_equals_unsharded(tss, orc):
    """Two shards searched independently and merged on the host == one index (no comm)."""
    n, dim = 70_001, 384
    rows = orc.gen_rows(0, n, dim, SEED)
    q = orc.gen_rows(0, 2, dim, 0xBEEF)
    want = orc.cosine_topk(rows, q, 10)
    cut = 31_337
    a = _mk_index(tss, rows[:cut])
    b = tss.FlatIndex(dim)
    b.add(rows[cut:])
    b.set_shard(cut, None)
    b.finalize()
    ra, sa, _ = a.search(q, 10)
    rb, sb, _ = b.search(q, 10)
    for qi in range(2):
        keys = sorted([(orc.pack_key(s, r)) for r, s in zip(ra[qi], sa[qi])] +
                      [(orc.pack_key(s, r)) for r, s in zip(rb[qi], sb[qi])], reverse=True)[:10]
        rows_m, scores_m = tss.unpack_keys(np.array(keys, dtype=np.uint64))
        assert np.array_equal(rows_m, want[0][qi])
        assert np.array_equal(scores_m.view(np.uint32), want[1][qi].view(np.uint32))


def test_search_device_matches_host_api(tss, orc):
    rows = orc.gen_rows(0, 25_000, 384, SEED)
    q = orc.gen_rows(0, 4, 384, 0xBEEF)
    ix = _mk_index(tss, rows)
    dq = tss.DeviceBuffer(0, q.nbytes).upload(q)
    dk = tss.DeviceBuffer(0, 4 * 10 * 8)
    before = tss.launch_count()
    ix.search_device(dq, 4, 10, dk)
    ix.sync()
    assert tss.launch_count() == before + 1
    r, s = tss.unpack_keys(dk.download(np.uint64, 40).reshape(4, 10))
    want = ix.search(q, 10)
    assert np.array_equal(r, want[0]) and np.array_equal(s, want[1])


def test_full_size_properties_1m(tss, orc):
    """BASELINE config 2 size: 1M x 384 generated on the device; size-independent checks."""
    n, dim, k = 1_000_000, 384, 10
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, SEED)
    ix.finalize()
    planted = 777_777
    q = orc.gen_rows(planted, 1, dim, SEED)[0] + 0.125 * orc.gen_rows(0, 1, dim, 0xBEEF)[0]
    r, s, c = ix.search(q, k)
    assert c[0] == k and r[0][0] == planted
    keys = [orc.pack_key(float(sc), int(ro)) for ro, sc in zip(r[0], s[0])]
    assert keys == sorted(keys, reverse=True)           # sortedness under the total order
    # recompute the winners' scores on the CPU from the generator: bit-exact
    for ro, sc in zip(r[0], s[0]):
        e = orc.gen_rows(int(ro), 1, dim, SEED)
        assert orc.scores(e, q)[0].view(np.uint32) == sc.view(np.uint32)
    # no sampled row beats the k-th key unless it is in the result
    rng = np.random.default_rng(0)
    sample = rng.integers(0, n, 20_000)
    sc = orc.scores(np.concatenate([orc.gen_rows(int(i), 1, dim, SEED) for i in sample[:2000]]), q)
    kth = keys[-1]
    for i, v in zip(sample[:2000], sc):
        assert orc.pack_key(float(v), int(i)) <= kth or int(i) in set(int(x) for x in r[0])
    # idempotence
    r2, s2, _ = ix.search(q, k)
    assert np.array_equal(r, r2) and np.array_equal(s, s2)
    # the oracle streaming the same generator agrees on the whole top-k
    _assert_same((r, s, c), orc.cosine_topk_synth(0, n, dim, SEED, q, k))


@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_save_load_round_trip(tss, orc, tmp_path, storage):
    """N1: on-disk index -> load -> identical search results; truncated files are rejected."""
    import os
    n, dim = 33_333, 200  # dim pads to 256 in storage
    rows = orc.gen_rows(0, n, dim, SEED)
    q = orc.gen_rows(0, 3, dim, 0xBEEF)
    st = tss.TSS_F32 if storage == "f32" else tss.TSS_BF16
    ix = tss.FlatIndex(dim, st)
    ix.add(rows)
    ix.finalize()
    want = ix.search(q, 10)
    path = str(tmp_path / "idx.tssidx")
    ix.save(path)
    assert os.path.getsize(path) == 64 + n * 256 * (4 if storage == "f32" else 2)
    ix2 = tss.FlatIndex.load(path)
    assert ix2.size() == n and ix2.dim == dim
    got = ix2.search(q, 10)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert np.array_equal(ix2.get_rows(5, 7), ix.get_rows(5, 7))
    with open(path, "r+b") as f:
        f.truncate(64 + 1000)
    with pytest.raises(tss.TssError):
        tss.FlatIndex.load(path)


@pytest.mark.parametrize("k", [129, 300, 1024])
def test_large_k_rounds_on_the_scan_path(tss, orc, k):
    """k > 128 on an fp32 index: exact, by rounds of 128 that exclude what was already found."""
    n = 5000
    rows = orc.gen_rows(0, n, 384, SEED)
    rows[4000:4100] = rows[17]  # ties across a round boundary
    q = orc.gen_rows(0, 2, 384, 0xBEEF)
    q[1] = rows[17]
    ix = _mk_index(tss, rows)
    _assert_same(ix.search(q, k), orc.cosine_topk(rows, q, k))
    bits = np.random.default_rng(k).random(n) < 0.5
    words = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.nonzero(bits)[0]
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    m = tss.Mask(n)
    m.upload(words)
    _assert_same(ix.search(q, k, m, tss.TSS_MASK_INCLUDE),
                 orc.cosine_topk(rows, q, k, words, orc.MASK_INCLUDE))
    _assert_same(ix.search(q, k, m, tss.TSS_MASK_EXCLUDE),
                 orc.cosine_topk(rows, q, k, words, orc.MASK_EXCLUDE))
    assert np.array_equal(m.download(), words)  # the caller's mask is untouched


def test_back_to_back_launches_overlap_safely(tss, orc):
    """Hundreds of queued scans (they overlap under programmatic dependent launch and rotate
    through the workspace slots) each produce their own exact result."""
    class _Off:
        def __init__(self, ptr):
            self.ptr = ptr
    for n in (40, 3000, 200_000):
        rows = orc.gen_rows(0, n, 384, SEED)
        ix = _mk_index(tss, rows)
        q = orc.gen_rows(0, 6, 384, 0xBEEF)
        want = orc.cosine_topk(rows, q, 10)
        reps = 300
        dq = tss.DeviceBuffer(0, q.nbytes).upload(q)
        dk = tss.DeviceBuffer(0, reps * 10 * 8)
        for i in range(reps):
            ix.search_device(_Off(dq.ptr + (i % 6) * 384 * 4), 1, 10, _Off(dk.ptr + i * 80))
        ix.sync()
        r, s = tss.unpack_keys(dk.download(np.uint64, reps * 10).reshape(reps, 10))
        for i in range(reps):
            assert np.array_equal(r[i], want[0][i % 6]), (n, i)
            assert np.array_equal(s[i].view(np.uint32), want[1][i % 6].view(np.uint32))


def test_concurrent_searches_on_one_handle(tss, orc):
    """Several host threads search ONE index handle at the same time (ctypes drops the GIL for
    the call): the handle's lock serialises them, every thread gets the oracle's answer."""
    import threading
    n, dim, k = 120_000, 384, 10
    rows = orc.gen_rows(0, n, dim, 0x5EED)
    ix = tss.FlatIndex(dim)
    ix.add_synthetic(0, n, 0x5EED)
    ix.finalize()
    q = orc.gen_rows(0, 64, dim, 0xBEEF)
    want = orc.cosine_topk(rows, q, k)
    m = tss.Mask(n)
    m.set_rows(np.arange(0, n, 3, dtype=np.uint32))
    w = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.arange(0, n, 3, dtype=np.int64)
    np.bitwise_or.at(w, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
    want_m = orc.cosine_topk(rows, q, k, mask_words=w, mask_mode=orc.MASK_EXCLUDE)
    errors = []

    def worker(t):
        try:
            for i in range(40):
                j = (t * 7 + i) % 60
                if (t + i) % 3 == 0:   # a masked search reads the shared mask concurrently
                    got = ix.search(q[j], k, m, tss.TSS_MASK_EXCLUDE)
                    ref = want_m
                    sl = slice(j, j + 1)
                elif (t + i) % 3 == 1:  # a small batch
                    got = ix.search(q[j:j + 4], k)
                    ref, sl = want, slice(j, j + 4)
                else:
                    got = ix.search(q[j], k)
                    ref, sl = want, slice(j, j + 1)
                if not (np.array_equal(got[0], ref[0][sl]) and
                        np.array_equal(got[1].view(np.uint32), ref[1][sl].view(np.uint32))):
                    errors.append((t, i))
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:5]


@pytest.mark.parametrize("storage_bf16", [False, True])
def test_gpu_scores_equal_exact_rational_arithmetic(tss, storage_bf16):
    """The kernel against the SPECIFICATION directly (no oracle in between): every score the scan
    returns equals DESIGN.md section 3 evaluated in exact rational arithmetic with explicit
    round-to-nearest-even (tests/test_exact_arithmetic.py), bit for bit -- and so does the order."""
    import test_exact_arithmetic as exact
    rng = np.random.default_rng(77)
    for dim in (3, 384, 520):
        rows = rng.standard_normal((9, dim)).astype(np.float32)
        rows[1] *= 1e-3
        rows[2] *= 200.0
        rows[4] = 0.0
        rows[7] = rows[0]                       # an exact duplicate: a tie, broken by row id
        q = rng.standard_normal(dim).astype(np.float32)
        ix = tss.FlatIndex(dim, tss.TSS_BF16 if storage_bf16 else tss.TSS_F32)
        ix.add(rows)
        ix.finalize()
        gr, gs, gc = ix.search(q, 9)
        assert gc[0] == 9
        want = {r: exact._score_bits(rows[r], q, bf16=storage_bf16) for r in range(9)}
        for r, s in zip(gr[0], gs[0]):
            assert int(s.view(np.uint32)) == want[int(r)], (dim, int(r))
        # (score desc, row asc) under the exact scores
        order = sorted(range(9), key=lambda r: (-np.uint32(want[r]).view(np.float32), r))
        assert [int(r) for r in gr[0]] == order
