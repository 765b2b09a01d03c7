"""Committed fixtures under tests/golden/ (see make_golden.py for what they are and are not).

not gpu: the oracle still reproduces them (and the hand-derived trie/merge KATs);
gpu:     libtss reproduces scan_golden.npz bit for bit and the KAT prefix masks.
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "scan_golden.npz"))
KATS = json.load(open(os.path.join(HERE, "golden", "trie_kats.json")))
TRIE = {"case_name": 0, "content": 1, "citation": 2}


def _cid(i):
    return bytes([i]) * 16


@pytest.mark.parametrize("case", make_golden.CASES, ids=[c[0] for c in make_golden.CASES])
def test_oracle_reproduces_scan_golden(orc, case):
    name, n, dim, nq, k, storage, density, mode = case
    rows, q, words = make_golden.case_inputs(*case)
    r, s, c = orc.cosine_topk(rows, q, k, words, mode, bf16=(storage == "bf16"))
    assert np.array_equal(r, GOLD[name + "_rows"])
    assert np.array_equal(s.view(np.uint32), GOLD[name + "_score_bits"])
    assert np.array_equal(c, GOLD[name + "_counts"])


def test_oracle_reproduces_trie_and_merge_kats(orc):
    t = orc.Trie()
    for i, (name, cit) in enumerate(zip(KATS["case_names"], KATS["citations"]), start=1):
        t.insert_case_name(name, _cid(i))
        t.insert_citation(cit, orc.docref(_cid(i), 0, -1))
    for kat in KATS["search_one"]:
        r = t.search_one(TRIE[kat["trie"]], kat["query"])
        assert [x[0] for x in r["exact"]] == [_cid(i) for i in kat["exact_ids"]], kat["kat"]
        assert sorted(r["completions"]) == kat["completions"], kat["kat"]
        assert r["total"] == kat["total"], kat["kat"]
    for kat in KATS["cascade"]:
        r = t.search(kat["query"])
        assert [x[0] for x in r["exact"]] == [_cid(i) for i in kat["exact_ids"]], kat["kat"]
        assert r["total"] == kat["total"], kat["kat"]
    for kat in KATS["prefix_rows"]:
        got = sorted(x[0] for x in t.prefix_postings(TRIE[kat["trie"]], kat["query"]))
        assert got == sorted(_cid(i) for i in kat["ids"]), kat["kat"]
    for kat in KATS["merge"]:
        out = orc.hybrid_merge(kat["exact"], [v[0] for v in kat["vec"]], [v[1] for v in kat["vec"]])
        if "want" in kat:
            assert [(h[0], round(h[1], 6), "Exact" if h[2] == 0 else "Semantic") for h in out] == \
                   [(w[0], w[1], w[2]) for w in kat["want"]], kat["kat"]
        else:
            assert [h[0] for h in out] == kat["want_ids"], kat["kat"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", make_golden.CASES, ids=[c[0] for c in make_golden.CASES])
def test_libtss_reproduces_scan_golden(tss, case):
    name, n, dim, nq, k, storage, density, mode = case
    rows, q, words = make_golden.case_inputs(*case)
    ix = tss.FlatIndex(dim, tss.TSS_F32 if storage == "f32" else tss.TSS_BF16)
    ix.add(rows)
    ix.finalize()
    m = None
    if words is not None:
        m = tss.Mask(n)
        m.upload(words)
    r, s, c = ix.search(q, k, m, mode)
    assert np.array_equal(r, GOLD[name + "_rows"])
    assert np.array_equal(s.view(np.uint32), GOLD[name + "_score_bits"])
    assert np.array_equal(c, GOLD[name + "_counts"])


@pytest.mark.gpu
def test_libtss_reproduces_prefix_kats(tss):
    """K9 rows via the device prefix kernel; row i-1 <-> case id i."""
    for kat in KATS["prefix_rows"]:
        src = KATS["case_names"] if kat["trie"] == "case_name" else KATS["citations"]
        lower = kat["trie"] == "case_name"
        terms = {}
        for i, text in enumerate(src):
            key = " ".join(text.lower().split() if lower else text.split()).encode()
            terms.setdefault(key, []).append(i)
        keys = sorted(terms)
        t = tss.Terms(keys, [terms[x] for x in keys])
        m = tss.Mask(3)
        p = " ".join(kat["query"].lower().split() if lower else kat["query"].split()).encode()
        t.prefix_mask(p, m)
        bits = int(m.download()[0])
        assert sorted(i + 1 for i in range(3) if bits >> i & 1) == sorted(kat["ids"]), kat["kat"]
