"""Masked K1 + K4 at chosen selectivities over 10M x 384 fp32 (for ncu captures and quick timing).

A tiny flattened trie: one term per selectivity whose postings are that many random rows.  Per
case: tss_prefix_mask_fresh (K4, one launch) then the masked scan, device-timed with events on the
index stream; masks and top-10 checked against numpy / the oracle generator."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tss_loader
tss = tss_loader.load()
from _common import make_queries

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--sel", type=float, nargs="+", default=[1e-5, 1e-3, 1e-2, 0.11])
ap.add_argument("--contig", type=float, nargs="*", default=[],
                help="extra cases whose live rows are ONE contiguous range of this fraction of the rows "
                     "(a date filter over rows added in date order)")
a = ap.parse_args()
N, dim, k = a.rows, 384, 10
rng = np.random.default_rng(9)
posts = [rng.integers(0, N, size=max(1, int(N * s)), dtype=np.uint32) for s in a.sel]
posts += [np.arange(N // 3, N // 3 + max(1, int(N * s)), dtype=np.uint32) for s in a.contig]
a.sel = list(a.sel) + [-s for s in a.contig]   # (negative: contiguous)
terms = [b"s%02d" % i for i in range(len(a.sel))]
poff = np.zeros(len(terms) + 1, dtype=np.uint64); np.cumsum([p.size for p in posts], out=poff[1:])
toff = np.zeros(len(terms) + 1, dtype=np.uint64); np.cumsum([len(t) for t in terms], out=toff[1:])
t = tss.Terms.from_arrays(b"".join(terms), toff, poff, np.concatenate(posts))
ix = tss.FlatIndex(dim); ix.reserve(N); ix.add_synthetic(0, N, 0x5EED); ix.finalize()
t.bind_stream(ix)
m = tss.Mask(N)
q = make_queries(8, dim, 0xBEEF)
dq = tss.DeviceBuffer(0, q.nbytes).upload(q)
dk = tss.DeviceBuffer(0, 8 * k * 8)
class S:
    def __init__(s, p): s.ptr = p
e0, e1 = tss.Event(0), tss.Event(0)
out = []
for term, ps, sel in zip(terms, posts, a.sel):
    uniq = np.unique(ps)
    want = np.zeros((N + 31) // 32, dtype=np.uint32)
    np.bitwise_or.at(want, (uniq >> 5).astype(np.int64), np.uint32(1) << (uniq & 31).astype(np.uint32))
    t.prefix_mask(term, m, want_stats=False, fresh=True)
    ok = bool(np.array_equal(m.download(), want))
    for _ in range(3):
        t.prefix_mask(term, m, want_stats=False, fresh=True)
        ix.search_device(dq, 1, k, dk, m, tss.TSS_MASK_INCLUDE)
    ix.sync()
    e0.record(ix)
    for _ in range(a.iters):
        t.prefix_mask(term, m, want_stats=False, fresh=True)
    e1.record(ix); ix.sync()
    k4 = e0.elapsed_ms(e1) / a.iters * 1e3
    t.prefix_mask(term, m, want_stats=False, fresh=True)
    ix.sync()
    e0.record(ix)
    for i in range(a.iters):
        ix.search_device(S(dq.ptr + (i % 8) * dim * 4), 1, k, S(dk.ptr + (i % 8) * k * 8), m, tss.TSS_MASK_INCLUDE)
    e1.record(ix); ix.sync()
    scan = e0.elapsed_ms(e1) / a.iters * 1e3
    out.append({"selectivity": sel, "live_rows": int(uniq.size), "postings": int(ps.size),
                "mask_ok": ok, "k4_us": k4, "masked_scan_us": scan,
                "live_gbs": uniq.size * dim * 4 / (scan * 1e-6) / 1e9,
                "list_driven": bool(ps.size <= 131072)})
print(json.dumps({"rows": N, "cases": out}))
