"""K2 (tcgen05 tensor-core path, bf16 index, large query batches).

north_star: the bf16 tensor-core path is reported as recall@k against fp32 exact, not
bit-compared (tensor-core accumulation order is the hardware's).  The path goes further: the
survivors within a rigorous error margin of the k-th best tensor-core score are re-scored with
the scan's arithmetic, so rows AND scores are bit-identical to the oracle on the same bf16 index
(`_exact`).  With TSS_GEMM_RESCORE=0 the result is the top-k of the raw bf16 x bf16 tensor-core
scores, checked against float64 on the CPU (`_own_arithmetic`).  Recall against the fp32 exact
oracle is printed and bounded either way, as is the exact fallback for overflowing lists.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0x5EED


def _bf16(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)


def _recall(got_rows, want_rows):
    hits = sum(len(set(g.tolist()) & set(w.tolist())) for g, w in zip(got_rows, want_rows))
    return hits / want_rows.size


def _f64_topk(rows_bf, q_bf, k):
    r = rows_bf.astype(np.float64)
    s = (q_bf.astype(np.float64) @ r.T) / (np.linalg.norm(q_bf.astype(np.float64), axis=1)[:, None]
                                           * np.linalg.norm(r, axis=1)[None, :])
    idx = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return idx, np.take_along_axis(s, idx, axis=1)


def _exact(orc, rows, q, k, got):
    """rows and score bits equal the oracle's top-k on the bf16-rounded rows"""
    want = orc.cosine_topk(rows, q, k, bf16=True)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))


@pytest.fixture
def no_rescore():
    os.environ["TSS_GEMM_RESCORE"] = "0"
    yield
    del os.environ["TSS_GEMM_RESCORE"]


@pytest.mark.parametrize("n,nq,k", [(150_000, 256, 100), (120_001, 200, 10), (300_000, 64, 100)])
def test_raw_tensor_core_scores(tss, orc, no_rescore, n, nq, k):
    """TSS_GEMM_RESCORE=0: the exact top-k of the bf16 x bf16 scores."""
    dim = 384
    rows = orc.gen_rows(0, n, dim, SEED)
    q = orc.gen_rows(0, nq, dim, 0xBEEF)
    ix = tss.FlatIndex(dim, tss.TSS_BF16)
    ix.add(rows)
    ix.finalize()
    gr, gs, gc = ix.search(q, k)
    assert np.all(gc == k) and np.all(np.diff(gs, axis=1) <= 0)
    wi, ws = _f64_topk(_bf16(rows), _bf16(q), k)
    assert _recall(gr, wi) >= 0.995
    np.testing.assert_allclose(gs[:, 0], ws[:, 0], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("n,nq,k", [(150_000, 256, 100), (120_001, 200, 10), (300_000, 64, 100),
                                    (400_000, 1000, 50)])
def test_recall_vs_fp32_exact(tss, orc, n, nq, k):
    dim = 384
    rows = orc.gen_rows(0, n, dim, SEED)
    q = orc.gen_rows(0, nq, dim, 0xBEEF)
    q[0] = rows[4242] + 0.125 * q[0]  # planted
    ix = tss.FlatIndex(dim, tss.TSS_BF16)
    ix.add(rows)
    ix.finalize()
    before = tss.launch_count()
    gr, gs, gc = ix.search(q, k)
    assert tss.launch_count() - before >= 5  # prep + 2 GEMM passes + threshold + select
    assert np.all(gc == k) and gr[0][0] == 4242
    assert np.all(np.diff(gs, axis=1) <= 0)
    _exact(orc, rows, q, k, (gr, gs))
    # reported quality: recall@k against the fp32 exact oracle
    er, _, _ = orc.cosine_topk(rows, q, k)
    rec = _recall(gr, er)
    print(f"recall@{k} vs fp32 exact: {rec:.4f} (n={n}, nq={nq})")
    assert rec >= 0.90
    if k >= 10:
        assert _recall(gr[:, :10], er[:, :10]) >= 0.85


def test_large_k(tss, orc):
    n, nq, k, dim = 600_000, 128, 500, 384
    ix = tss.FlatIndex(dim, tss.TSS_BF16)
    ix.add_synthetic(0, n, SEED)
    ix.finalize()
    q = orc.gen_rows(0, nq, dim, 0xBEEF)
    gr, gs, gc = ix.search(q, k)
    assert np.all(gc == k)
    rows = orc.gen_rows(0, n, dim, SEED)
    _exact(orc, rows, q, k, (gr, gs))
    # a small batch with the same large k takes the scan path by rounds: exact in bf16 storage
    sr, ss, sc = ix.search(q[:2], k)
    want = orc.cosine_topk(rows, q[:2], k, bf16=True)
    assert np.array_equal(sr, want[0]) and np.array_equal(ss.view(np.uint32), want[1].view(np.uint32))


def test_overflow_falls_back_to_exact_scan(tss, orc):
    """50k copies of one row score identically: every tile maximum ties, the survivor list
    overflows, and the query is redone by the exact K1 scan (ties -> ascending row id)."""
    n, dim, k = 200_000, 384, 10
    rows = orc.gen_rows(0, n, dim, SEED)
    rows[100_000:150_000] = rows[7]
    q = orc.gen_rows(0, 64, dim, 0xBEEF)
    q[3] = rows[7]
    ix = tss.FlatIndex(dim, tss.TSS_BF16)
    ix.add(rows)
    ix.finalize()
    gr, gs, gc = ix.search(q, k)
    want = orc.cosine_topk(rows, q[3], k, bf16=True)
    assert list(gr[3]) == list(want[0][0])  # 7, 100000, 100001, ...
    assert np.array_equal(gs[3].view(np.uint32), want[1][0].view(np.uint32))


def test_small_batches_keep_using_the_scan(tss, orc):
    rows = orc.gen_rows(0, 150_000, 384, SEED)
    q = orc.gen_rows(0, 8, 384, 0xBEEF)
    ix = tss.FlatIndex(384, tss.TSS_BF16)
    ix.add(rows)
    ix.finalize()
    got = ix.search(q, 10)
    want = orc.cosine_topk(rows, q, 10, bf16=True)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))


@pytest.mark.parametrize("dim,nq", [(64, 130), (100, 130), (256, 130), (300, 130), (512, 130),
                                    (768, 130), (1000, 130), (256, 64), (768, 64)])
def test_other_dimensions(tss, orc, dim, nq):
    """D pads to 128 ... 1024 columns: 2 ... 16 k-blocks of the same kernel (beyond 384 the query
    tile streams through the ring instead of staying resident).  130 queries = two query blocks =
    one cta_group::2 pair per slice; 64 queries = one block = independent CTAs."""
    n, k = 80_000, 10
    rng = np.random.default_rng(dim)
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    q[5] = rows[777] + 0.1 * q[5]
    ix = tss.FlatIndex(dim, tss.TSS_BF16)
    ix.add(rows)
    ix.finalize()
    gr, gs, gc = ix.search(q, k)
    assert np.all(gc == k) and gr[5][0] == 777
    _exact(orc, rows, q, k, (gr, gs))


@pytest.mark.parametrize("mode,nq", [("include", 96), ("exclude", 96), ("include", 200)])
def test_masked_large_batch(tss, orc, mode, nq):
    """Masks ride along on the tensor-core path: a masked row's 1/|row| is NaN in the epilogue
    (96 queries: independent CTAs; 200: CTA pairs)."""
    n, k, dim = 160_000, 20, 384
    rows = orc.gen_rows(0, n, dim, SEED)
    q = orc.gen_rows(0, nq, dim, 0xBEEF)
    rng = np.random.default_rng(4)
    bits = rng.random(n) < (0.3 if mode == "include" else 0.6)
    bits[:2048] = mode == "exclude"  # whole tiles without a live row
    words = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.nonzero(bits)[0]
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    ix = tss.FlatIndex(dim, tss.TSS_BF16)
    ix.add(rows)
    ix.finalize()
    m = tss.Mask(n)
    m.upload(words)
    before = tss.launch_count()
    gr, gs, gc = ix.search(q, k, m, tss.TSS_MASK_INCLUDE if mode == "include" else tss.TSS_MASK_EXCLUDE)
    assert tss.launch_count() - before <= 8  # the K2 pipeline + list compaction (+ the one-off row norms / shadow), not nq/4 scans
    live = bits if mode == "include" else ~bits
    assert np.all(gc == k) and np.all(live[gr])  # only live rows come back
    want = orc.cosine_topk(rows, q, k, mask_words=words,
                           mask_mode=orc.MASK_INCLUDE if mode == "include" else orc.MASK_EXCLUDE,
                           bf16=True)
    assert np.array_equal(gr, want[0])
    assert np.array_equal(gs.view(np.uint32), want[1].view(np.uint32))


@pytest.mark.parametrize("n,nq,k,dim", [(150_000, 256, 100, 384), (120_001, 40, 10, 384),
                                        (90_000, 130, 10, 200), (60_000, 64, 5, 768)])
def test_fp32_index_large_batches_are_exact(tss, orc, n, nq, k, dim):
    """An fp32 index sends large batches through the tensor cores too (a bf16 shadow of the rows
    picks the candidates), and the survivors are re-scored from the fp32 rows: rows and score bits
    equal the fp32 oracle -- north_star's bit-exact bar for the fp32 path."""
    rng = np.random.default_rng(n + nq)
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    q[1] = rows[31337] + 0.125 * q[1]
    ix = tss.FlatIndex(dim, tss.TSS_F32)
    ix.add(rows)
    ix.finalize()
    before = tss.launch_count()
    gr, gs, gc = ix.search(q, k)
    assert tss.launch_count() - before <= 8  # shadow + norms + the K2 pipeline, not nq/4 scans
    assert np.all(gc == k) and gr[1][0] == 31337
    want = orc.cosine_topk(rows, q, k)
    assert np.array_equal(gr, want[0])
    assert np.array_equal(gs.view(np.uint32), want[1].view(np.uint32))
    # a second batch reuses the shadow; a masked one too
    bits = rng.random(n) < 0.5
    words = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.nonzero(bits)[0]
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    m = tss.Mask(n)
    m.upload(words)
    before = tss.launch_count()
    gr, gs, gc = ix.search(q, k, m, tss.TSS_MASK_INCLUDE)
    assert tss.launch_count() - before == 6  # prep, sample pass, threshold, collect pass, select, fix-up list
    want = orc.cosine_topk(rows, q, k, mask_words=words, mask_mode=orc.MASK_INCLUDE)
    assert np.array_equal(gr, want[0])
    assert np.array_equal(gs.view(np.uint32), want[1].view(np.uint32))


@pytest.mark.parametrize("storage", ["bf16", "f32"])
def test_k_1000_stays_on_the_tensor_cores(tss, orc, storage):
    """k = 1000: the sorter takes 4k keys, the tile sample and the batch split keep the expected
    survivors inside the pool -- no fallback scans (which would be 8 rounds per query)."""
    n, nq, k, dim = 1_100_000, 48, 1000, 384
    ix = tss.FlatIndex(dim, tss.TSS_BF16 if storage == "bf16" else tss.TSS_F32)
    ix.add_synthetic(0, n, SEED)
    ix.finalize()
    q = orc.gen_rows(0, nq, dim, 0xBEEF)
    ix.search(q[:16], k)  # builds the one-off workspaces (norms, shadow)
    before = tss.launch_count()
    gr, gs, gc = ix.search(q, k)
    assert tss.launch_count() - before == 6
    rows = orc.gen_rows(0, n, dim, SEED)
    want = orc.cosine_topk(rows, q, k, bf16=storage == "bf16")
    assert np.array_equal(gr, want[0])
    assert np.array_equal(gs.view(np.uint32), want[1].view(np.uint32))


def test_batch_policy_single_queries_through_the_shadow(tss, orc):
    """tss_index_set_batch_policy(1, build_shadow_now): an fp32 index answers one or two queries
    per call from its bf16 shadow (K1 streams the shadow for the top-64, refine_kernel proves the
    fp32 top-k is among them and re-scores from the fp32 rows) -- same bits as the fp32 scan, half
    the bytes.  Three or more queries per call take K2."""
    n, dim, k = 200_000, 384, 10
    rows = orc.gen_rows(0, n, dim, SEED)
    rows[150_000:150_300] = rows[9]          # 301 identical rows: no top-64 can be proven complete
    q = orc.gen_rows(0, 6, dim, 0xBEEF)
    q[2] = rows[77] + 0.125 * q[2]
    q[4] = rows[9]
    ix = tss.FlatIndex(dim, tss.TSS_F32)
    ix.add(rows)
    ix.finalize()
    scan = ix.search(q, k)                      # default policy: K1 on the fp32 rows
    want = orc.cosine_topk(rows, q, k)
    assert np.array_equal(scan[0], want[0]) and scan[0][2][0] == 77
    assert list(scan[0][4][:3]) == [9, 150_000, 150_001]
    ix.set_batch_policy(1, build_shadow_now=True)
    for i in range(6):
        before = tss.launch_count()
        got = ix.search(q[i:i + 1], k)
        used = tss.launch_count() - before
        assert used == (4 if i == 4 else 3), (i, used)   # shadow scan + refine + redo list (+ fp32 redo)
        assert np.array_equal(got[0][0], scan[0][i])
        assert np.array_equal(got[1][0].view(np.uint32), scan[1][i].view(np.uint32))
    got = ix.search(q[:2], k)                   # two per call: two shadow scans, one refine
    assert np.array_equal(got[0], scan[0][:2])
    before = tss.launch_count()
    got = ix.search(q[:3], k)                   # three: the K2 pipeline
    assert tss.launch_count() - before == 6
    assert np.array_equal(got[0], scan[0][:3])
    assert np.array_equal(got[1].view(np.uint32), scan[1][:3].view(np.uint32))
    # masked single query
    bits = np.random.default_rng(3).random(n) < 0.4
    words = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.nonzero(bits)[0]
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    m = tss.Mask(n)
    m.upload(words)
    got = ix.search(q[:1], k, m, tss.TSS_MASK_INCLUDE)
    wm = orc.cosine_topk(rows, q[:1], k, mask_words=words, mask_mode=orc.MASK_INCLUDE)
    assert np.array_equal(got[0], wm[0]) and np.array_equal(got[1].view(np.uint32), wm[1].view(np.uint32))
    ix.set_batch_policy(0)                       # default again: single queries scan the fp32 rows
    before = tss.launch_count()
    ix.search(q[:1], k)
    assert tss.launch_count() - before == 1


@pytest.mark.parametrize("storage_f32", [False, True])
def test_sparse_include_mask_never_returns_rows_past_the_shard(tss, orc, storage_f32):
    """n % 256 != 0, nq >= 16 and an INCLUDE mask that leaves fewer than k live rows.
    The threshold degenerates to -inf; the columns of the last partial tile beyond n_rows used to
    be -inf as well and `-inf >= -inf` pushed row ids that do not exist.  They are NaN now."""
    n, dim, k, nq = 50_003, 384, 10, 32   # 50_003 % 256 = 83; K2 needs n >= 4*256*k
    rows = orc.gen_rows(0, n, dim, SEED)
    q = orc.gen_rows(0, nq, dim, 0xBEEF)
    ix = tss.FlatIndex(dim, tss.TSS_F32 if storage_f32 else tss.TSS_BF16)
    ix.add(rows)
    ix.finalize()
    for live in ([5, 49_990, 50_002], [50_001], [], [0, 1, 2, 3, 255, 256, 49_919, 49_920, 50_002]):
        m = tss.Mask(n)
        m.set_rows(np.asarray(live, dtype=np.uint32))
        w = np.zeros((n + 31) // 32, dtype=np.uint32)
        for r in live:
            w[r >> 5] |= np.uint32(1 << (r & 31))
        before = tss.launch_count()
        gr, gs, gc = ix.search(q, k, m, tss.TSS_MASK_INCLUDE)
        assert tss.launch_count() - before >= 5  # took K2
        want = orc.cosine_topk(rows, q, k, mask_words=w, mask_mode=orc.MASK_INCLUDE,
                               bf16=not storage_f32)
        assert np.all(gc == len(live))
        assert np.array_equal(gr, want[0]), live
        assert np.array_equal(gs.view(np.uint32), want[1].view(np.uint32)), live
        assert np.all((gr < n) | (gr == tss.TSS_ROW_NONE))


def test_device_resident_batches_fix_up_on_the_device(tss, orc):
    """tss_index_search_device never synchronises the host (include/tss.h): queries whose survivor
    list overflows are redone by guarded scan launches enqueued behind the batch.  One flagged
    query: exact without any host involvement.  More flagged queries than fix-up launches: the
    condition is reported by tss_index_sync and the next call has twice the launches."""
    n, dim, k = 200_000, 384, 10
    rows = orc.gen_rows(0, n, dim, SEED)
    rows[100_000:150_000] = rows[7]   # 50k copies: every list of a query near row 7 overflows
    ix = tss.FlatIndex(dim, tss.TSS_BF16)
    ix.add(rows)
    ix.finalize()
    nq = 64
    q = orc.gen_rows(0, nq, dim, 0xBEEF)
    q[3] = rows[7]
    want = orc.cosine_topk(rows, q, k, bf16=True)
    dq = tss.DeviceBuffer(0, q.nbytes).upload(q)
    dk = tss.DeviceBuffer(0, nq * k * 8)
    before = tss.launch_count()
    ix.search_device(dq, nq, k, dk)
    assert tss.launch_count() - before >= 5 + 1 + 2  # K2 pipeline + compaction + guarded fix-ups
    ix.sync()
    gr, gs = tss.unpack_keys(dk.download(np.uint64, nq * k).reshape(nq, k))
    assert np.array_equal(gr, want[0]) and np.array_equal(gs.view(np.uint32), want[1].view(np.uint32))
    # ten hostile queries > two fix-up launches
    for j in range(10):
        q[20 + j] = rows[7] * (1.0 + 0.01 * j)
    want = orc.cosine_topk(rows, q, k, bf16=True)
    dq.upload(q)
    for attempt in range(6):
        ix.search_device(dq, nq, k, dk)
        try:
            ix.sync()
            break
        except tss.TssError as e:
            assert e.code == tss.TSS_ERR_STATE
    else:
        raise AssertionError("the fix-up launches never caught up")
    assert attempt >= 1  # the first try could not cover eleven flagged queries with two launches
    gr, gs = tss.unpack_keys(dk.download(np.uint64, nq * k).reshape(nq, k))
    assert np.array_equal(gr, want[0]) and np.array_equal(gs.view(np.uint32), want[1].view(np.uint32))
    # the host entry covers any number of them in one call
    got = ix.search(q, k)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))


@pytest.mark.parametrize("seed", range(8))
def test_random_shapes_and_routes(tss, orc, seed):
    """Random corpus size / dimension / batch / k / storage / mask: whichever route the library
    picks (scan, pairs, quads, a left-over cluster serving several query groups, weights or the
    weight-free epilogue), rows, scores and counts equal the oracle's."""
    rng = np.random.default_rng(1000 + seed)
    dim = int(rng.choice([64, 128, 200, 384, 768]))
    n = int(rng.integers(30_000, 260_000))
    nq = int(rng.choice([16, 100, 129, 256, 300, 512, 600, 1024]))
    k = int(rng.choice([1, 10, 37, 100]))
    f32 = bool(rng.integers(0, 2))
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    rows[::977] *= 50.0
    rows[5] = 0.0                                   # a zero row scores 0
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    q[1] = rows[n // 2] * 0.5
    ix = tss.FlatIndex(dim, tss.TSS_F32 if f32 else tss.TSS_BF16)
    ix.add(rows)
    ix.finalize()
    mode = int(rng.integers(0, 3))
    m = w = None
    if mode:
        bits = rng.random(n) < rng.choice([0.02, 0.5, 0.97])
        w = np.zeros((n + 31) // 32, dtype=np.uint32)
        idx = np.nonzero(bits)[0]
        np.bitwise_or.at(w, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
        m = tss.Mask(n)
        m.upload(w)
    got = ix.search(q, k, m, mode)
    want = orc.cosine_topk(rows, q, k, mask_words=w, mask_mode=mode, bf16=not f32)
    assert np.array_equal(got[2], want[2]), (dim, n, nq, k, f32, mode)
    assert np.array_equal(got[0], want[0]), (dim, n, nq, k, f32, mode)
    assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)), (dim, n, nq, k, f32, mode)
