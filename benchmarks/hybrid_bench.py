"""BASELINE config 4: T-term synthetic flattened trie -> prefix mask -> masked top-10 over N x 384.

Terms are 1-4 tokens "wNNNNNN" drawn Zipf(1.1) from a 200k vocabulary, unique, byte-sorted;
each term posts Geometric(mean 4) rows uniform in [0,N).  For prefixes of three selectivities it
times tss_mask_clear + tss_prefix_mask (K4) + the masked scan (K1) and checks the mask bit for
bit against numpy (the expected row set of the matching term range)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tss_loader
tss = tss_loader.load()
from _common import make_queries

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--terms", type=int, default=5_000_000)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
N, T, dim, k = a.rows, a.terms, 384, 10
rng = np.random.default_rng(5)

t0 = time.time()
ntok = rng.integers(1, 5, size=int(T * 1.15))
tok = (rng.zipf(1.1, size=(ntok.size, 4)) - 1) % 200_000
tok[np.arange(4)[None, :] >= ntok[:, None]] = -1
tok = np.unique(tok, axis=0)[:T]          # lexicographic on ids == byte order of the joined strings
T = tok.shape[0]
ntok = (tok >= 0).sum(axis=1)
# byte matrix [T, 31]: 'w' + 6 digits, ' ' between tokens
mat = np.full((T, 32), ord(" "), dtype=np.uint8)
for s in range(4):
    v = np.where(tok[:, s] >= 0, tok[:, s], 0)
    mat[:, 8 * s] = ord("w")
    for d in range(6):
        mat[:, 8 * s + 1 + d] = ord("0") + (v // 10 ** (5 - d)) % 10
lens = 8 * ntok - 1
keep = np.arange(32)[None, :] < lens[:, None]
pool = mat[keep].tobytes()
term_off = np.zeros(T + 1, dtype=np.uint64); np.cumsum(lens, out=term_off[1:])
npost = rng.geometric(0.25, size=T)
post_off = np.zeros(T + 1, dtype=np.uint64); np.cumsum(npost, out=post_off[1:])
post_rows = rng.integers(0, N, size=int(post_off[-1]), dtype=np.uint32)
build_s = time.time() - t0

terms = tss.Terms.from_arrays(pool, term_off, post_off, post_rows)
# N2: the same structure built on the device from the tokenised postings (one tuple per posting)
vocab = [b"w%06d" % i for i in range(200_000)]
tuple_of_posting = np.repeat(np.arange(T), npost)
perm = rng.permutation(tuple_of_posting.size)          # postings arrive in arbitrary order
ids_dev = (tok[tuple_of_posting[perm]] + 1).astype(np.uint32)
t0 = time.time()
built = tss.Terms.build(vocab, ids_dev, post_rows[perm])
device_build_s = time.time() - t0
assert built.size() == T
ix = tss.FlatIndex(dim)
ix.reserve(N); ix.add_synthetic(0, N, 0x5EED); ix.finalize()
mask = tss.Mask(N)
q = make_queries(4, dim, 0xBEEF)

# postings under each first token -> pick prefixes by selectivity
first = tok[:, 0]
per_first = np.bincount(first, weights=npost, minlength=200_000)
out = {"rows": N, "terms": int(T), "postings": int(post_off[-1]), "host_build_seconds": build_s,
       "device_build_seconds_incl_h2d_and_validation": device_build_s, "cases": []}
mask_b = tss.Mask(N)
dense = []
for _ in range(3):
    ix.search(q[0], k)
e0, e1 = tss.Event(0), tss.Event(0)
t = time.perf_counter()
for i in range(a.iters):
    ix.search(q[i % 4], k)
dense_us = (time.perf_counter() - t) / a.iters * 1e6
out["dense_search_us"] = dense_us
for target in (N * 1e-5, N * 1e-3, N * 1e-1):
    cand = int(np.argmin(np.abs(per_first - target)))
    prefix = b"w%06d" % cand
    lo, hi = np.searchsorted(first, cand), np.searchsorted(first, cand, side="right")
    want_rows = np.unique(post_rows[int(post_off[lo]):int(post_off[hi])])
    want = np.zeros((N + 31) // 32, dtype=np.uint32)
    np.bitwise_or.at(want, want_rows >> 5, np.uint32(1) << (want_rows & 31).astype(np.uint32))
    mask.clear(); st = terms.prefix_mask(prefix, mask)
    ok = bool(np.array_equal(mask.download(), want))
    mask_b.clear(); built.prefix_mask(prefix, mask_b)
    ok = ok and bool(np.array_equal(mask_b.download(), want))  # device-built trie: same mask
    for _ in range(2):
        mask.clear(); terms.prefix_mask(prefix, mask, want_stats=False); ix.search(q[0], k, mask, tss.TSS_MASK_INCLUDE)
    tm = ts = 0.0
    for i in range(a.iters):
        t = time.perf_counter(); mask.clear(); terms.prefix_mask(prefix, mask, want_stats=False); tm += time.perf_counter() - t
        t = time.perf_counter(); r = ix.search(q[i % 4], k, mask, tss.TSS_MASK_INCLUDE); ts += time.perf_counter() - t
    pc = int(want_rows.size)
    algo = pc * dim * 4 + 2 * (N // 8) + 4 * int(st.npostings)
    out["cases"].append({"prefix": prefix.decode(), "selectivity": pc / N, "mask_popcount": pc,
                         "postings_in_range": int(st.npostings), "terms_in_range": int(st.sub_hi - st.sub_lo + st.exact_hi - st.exact_lo),
                         "mask_bit_exact_vs_numpy": ok, "clear_plus_prefix_mask_us": tm / a.iters * 1e6,
                         "masked_search_us": ts / a.iters * 1e6, "algorithmic_bytes": algo,
                         "masked_scan_gbs": pc * dim * 4 / (ts / a.iters) / 1e9,
                         "speedup_vs_dense_search": dense_us / ((tm + ts) / a.iters * 1e6)})
print(json.dumps(out))
