"""K1 across dimensions / storage: batch-1 queries/s and GB/s (rows sized to ~7.7 GB per index)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tss_loader
from _common import make_queries
tss = tss_loader.load()
for dim, storage in ((128, "f32"), (256, "f32"), (384, "f32"), (512, "f32"), (768, "f32"), (1024, "f32"),
                     (100, "f32"), (384, "bf16"), (768, "bf16")):
    elem = 4 if storage == "f32" else 2
    ns = next(n for n in (1, 2, 3, 4, 6, 8) if n * 128 >= dim)
    rows = int(7.68e9 // (ns * 128 * elem))
    ix = tss.FlatIndex(dim, tss.TSS_F32 if storage == "f32" else tss.TSS_BF16)
    ix.reserve(rows); ix.add_synthetic(0, rows, 1); ix.finalize()
    q = make_queries(1, dim, 2)
    dq = tss.DeviceBuffer(0, q.nbytes).upload(q); dk = tss.DeviceBuffer(0, 80)
    for _ in range(5): ix.search_device(dq, 1, 10, dk)
    ix.sync()
    e0, e1 = tss.Event(0), tss.Event(0)
    e0.record(ix)
    for _ in range(100): ix.search_device(dq, 1, 10, dk)
    e1.record(ix); ix.sync()
    ms = e0.elapsed_ms(e1) / 100
    print(json.dumps({"dim": dim, "storage": storage, "rows": rows, "ms": round(ms, 4), "q_per_s": round(1e3 / ms, 1),
                      "stored_gbs": round(rows * ns * 128 * elem / ms / 1e6)}))
    ix.close()
