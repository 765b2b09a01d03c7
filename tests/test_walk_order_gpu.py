"""The masked scan's walk order (scan.cuh: runs of 2^w consecutive tiles per warp) must not
change results: every run length and both run-to-warp assignments, at row geometries of 16 / 8 /
4 / 2 / 1 rows per tile, ragged sizes, random / contiguous / blocky masks of both polarities and
both storages -- rows and score bits against the CPU oracle.

Reference: the mask is the seen-cases exclude set of src/search.rs:187,214 (EXCLUDE) or the prefix
filter of BASELINE.json config 4 (INCLUDE); SURVEY section 8(a) "semantic gap".
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 0x5EED


def _words(rows, n):
    w = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.asarray(rows, dtype=np.int64)
    np.bitwise_or.at(w, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
    return w


def _masks(n, rng):
    yield "random 11 %", np.flatnonzero(rng.random(n) < 0.11)
    yield "random 60 %", np.flatnonzero(rng.random(n) < 0.6)
    yield "sparse", rng.choice(n, size=min(n, 37), replace=False)
    yield "contiguous", np.arange(n // 3, n // 3 + max(1, n // 7))
    yield "tail", np.arange(max(0, n - 45), n)
    yield "blocks", np.flatnonzero((np.arange(n) // 300) % 5 == 2)
    yield "all", np.arange(n)
    yield "none", np.zeros(0, np.int64)


@pytest.mark.parametrize("walk", ["0", "2", "5", "13"])
@pytest.mark.parametrize("dim,storage", [(384, "f32"), (64, "f32"), (200, "bf16"), (768, "f32"), (1024, "bf16")])
def test_every_walk_order_gives_the_oracles_answer(tss, orc, walk, dim, storage):
    n, k = 70_019, 10
    rng = np.random.default_rng(int(walk) * 4099 + dim)
    rows = orc.gen_rows(0, n, dim, SEED)
    q = orc.gen_rows(0, 2, dim, 0xBEEF)
    old = os.environ.get("TSS_WALK_RUN")
    os.environ["TSS_WALK_RUN"] = walk     # read when the index is created
    try:
        ix = tss.FlatIndex(dim, tss.TSS_BF16 if storage == "bf16" else tss.TSS_F32)
    finally:
        if old is None:
            del os.environ["TSS_WALK_RUN"]
        else:
            os.environ["TSS_WALK_RUN"] = old
    ix.add_synthetic(0, n, SEED)
    ix.finalize()
    m = tss.Mask(n)
    for name, live in _masks(n, rng):
        w = _words(live, n)
        m.upload(w)
        for mode, omode in ((tss.TSS_MASK_INCLUDE, orc.MASK_INCLUDE), (tss.TSS_MASK_EXCLUDE, orc.MASK_EXCLUDE)):
            got = ix.search(q, k, m, mode)
            want = orc.cosine_topk(rows, q, k, mask_words=w, mask_mode=omode, bf16=storage == "bf16")
            assert np.array_equal(got[2], want[2]), (name, mode)
            assert np.array_equal(got[0], want[0]), (name, mode)
            assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)), (name, mode)
    ix.close()
