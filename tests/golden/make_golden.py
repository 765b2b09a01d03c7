"""Regenerates tests/golden/*.  Run from the repo root: python tests/golden/make_golden.py

The reference holds no golden vectors for this path (its only tests are src/utils.rs:205-227)
and cannot run, so these fixtures are NOT outputs of the reference:
  * trie_kats.json  -- hand-derived from reading reference src/trie.rs / src/search.rs
                       (SURVEY.md section 8c, K1-K9 and M1-M3); typed in, not generated;
  * scan_golden.npz -- outputs of the CPU oracle (oracle/oracle.cpp, canonical arithmetic of
                       DESIGN.md section 3) on seeded inputs, frozen so that neither the oracle nor the
                       kernels can drift silently.  Inputs are regenerated from the seeds.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402

CASES = [  # name, n, dim, nq, k, storage, mask density (None = no mask), mask mode
    ("f32_384", 5000, 384, 3, 10, "f32", None, 0),
    ("f32_384_k50", 5000, 384, 2, 50, "f32", None, 0),
    ("f32_100", 3000, 100, 2, 10, "f32", None, 0),
    ("f32_768", 2000, 768, 2, 10, "f32", None, 0),
    ("bf16_384", 5000, 384, 3, 10, "bf16", None, 0),
    ("f32_384_include", 5000, 384, 2, 10, "f32", 0.2, 1),
    ("f32_384_exclude", 5000, 384, 2, 10, "f32", 0.2, 2),
]


def case_inputs(name, n, dim, nq, k, storage, density, mode):
    rows = orc.gen_rows(0, n, dim, 0x5EED)
    q = orc.gen_rows(0, nq, dim, 0xBEEF)
    q[0] = rows[n // 3] + 0.125 * q[0]
    words = None
    if density is not None:
        bits = np.random.default_rng(len(name)).random(n) < density
        words = np.zeros((n + 31) // 32, dtype=np.uint32)
        idx = np.nonzero(bits)[0]
        np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    return rows, q, words


def main():
    out = {}
    for c in CASES:
        name, n, dim, nq, k, storage, density, mode = c
        rows, q, words = case_inputs(*c)
        r, s, cnt = orc.cosine_topk(rows, q, k, words, mode, bf16=(storage == "bf16"))
        out[name + "_rows"] = r
        out[name + "_score_bits"] = s.view(np.uint32)
        out[name + "_counts"] = cnt
    np.savez_compressed(os.path.join(HERE, "scan_golden.npz"), **out)
    print("wrote scan_golden.npz with", len(out), "arrays")
    assert os.path.exists(os.path.join(HERE, "trie_kats.json")), "trie_kats.json is hand-written"
    json.load(open(os.path.join(HERE, "trie_kats.json")))


if __name__ == "__main__":
    main()
