"""The reference-faithful device path: a dense scan with an EXCLUDE mask of a few seen-case rows
(SearchEngine MaskPolicy::ExcludeOnDevice; src/search.rs:187,214) against the unmasked scan.

  python benchmarks/exclude_probe.py [--rows N] [--iters I]
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tss_loader
tss = tss_loader.load()
from _common import make_queries

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
N, dim, k = a.rows, 384, 10
ix = tss.FlatIndex(dim); ix.reserve(N); ix.add_synthetic(0, N, 0x5EED); ix.finalize()
q = make_queries(8, dim, 0xBEEF)
dq = tss.DeviceBuffer(0, q.nbytes).upload(q)
dk = tss.DeviceBuffer(0, 8 * k * 8)
class S:
    def __init__(s, p): s.ptr = p
e0, e1 = tss.Event(0), tss.Event(0)
rng = np.random.default_rng(4)
out = []
def timed(mask, mode):
    for i in range(5):
        ix.search_device(S(dq.ptr + (i % 8) * dim * 4), 1, k, S(dk.ptr + (i % 8) * k * 8), mask, mode)
    ix.sync(); e0.record(ix)
    for i in range(a.iters):
        ix.search_device(S(dq.ptr + (i % 8) * dim * 4), 1, k, S(dk.ptr + (i % 8) * k * 8), mask, mode)
    e1.record(ix); ix.sync()
    return e0.elapsed_ms(e1) / a.iters * 1e3
base = timed(None, tss.TSS_MASK_NONE)
keys0 = dk.download(np.uint64, 8 * k).copy()
out.append({"case": "unmasked", "us": base, "gbs": N * dim * 4 / base / 1e3})
for nset in (0, 10, 1000, 100_000):
    m = tss.Mask(N)
    rows = rng.choice(N, size=nset, replace=False).astype(np.uint32) if nset else np.zeros(0, np.uint32)
    if nset:
        m.set_rows(rows)
    us = timed(m, tss.TSS_MASK_EXCLUDE)
    same = bool(np.array_equal(dk.download(np.uint64, 8 * k), keys0)) if nset == 0 else None
    out.append({"case": f"EXCLUDE mask with {nset} rows set", "us": us,
                "gbs": (N - nset) * dim * 4 / us / 1e3, "keys_equal_unmasked": same})
    m.close()
# sparse INCLUDE masks set row by row (no row list: the scan walks the mask)
for nset in (0, 1000, 10_000, 100_000):
    m = tss.Mask(N)
    if nset:
        m.set_rows(rng.choice(N, size=nset, replace=False).astype(np.uint32))
    us = timed(m, tss.TSS_MASK_INCLUDE)
    out.append({"case": f"INCLUDE mask with {nset} rows set (mask walk)", "us": us,
                "gbs": nset * dim * 4 / us / 1e3})
    m.close()
print(json.dumps({"rows": N, "cases": out}))
