"""Small batches (1..12 queries): queries/s, kernel launches per batch and which path took them
(K1 scans of 1/2/4 queries per launch, or -- 3+ queries on a big corpus whose bf16 matrix exists --
one K2 pass).  --shadow builds the bf16 shadow of an fp32 index first (one 16-query batch)."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tss_loader
from _common import make_queries
tss = tss_loader.load()
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--storage", default="f32")
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--shadow", action="store_true")
a = ap.parse_args()
dim = 384
st = tss.TSS_F32 if a.storage == "f32" else tss.TSS_BF16
ix = tss.FlatIndex(dim, st); ix.reserve(a.rows); ix.add_synthetic(0, a.rows, 0x5EED); ix.finalize()
elem = 4 if a.storage == "f32" else 2
if a.shadow:
    q16 = make_queries(16, dim, 0xBEEF)
    ix.search(q16, 10)
for nq in (1, 2, 3, 4, 8, 12):
    q = make_queries(nq, dim, 0xBEEF)
    dq = tss.DeviceBuffer(0, q.nbytes).upload(q); dk = tss.DeviceBuffer(0, nq * 10 * 8)
    for _ in range(3): ix.search_device(dq, nq, 10, dk)
    ix.sync()
    e0, e1 = tss.Event(0), tss.Event(0)
    l0 = tss.launch_count()
    e0.record(ix)
    for _ in range(a.iters): ix.search_device(dq, nq, 10, dk)
    e1.record(ix); ix.sync()
    ms = e0.elapsed_ms(e1) / a.iters
    launches = (tss.launch_count() - l0) / a.iters
    path = "K2" if launches >= 5 else "K1"  # (prep, sample, threshold, main, select, fix-up list + guarded scans)
    print(json.dumps({"storage": a.storage, "shadow": a.shadow, "nq": nq, "ms_per_batch": ms,
                      "queries_per_s": nq / ms * 1e3, "launches": launches, "path": path}))
