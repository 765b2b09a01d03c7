"""Diagnostics: when each CTA of a dense scan finishes its scan loop (tss_index_debug_phases),
unmasked vs an EXCLUDE mask with a few rows set.   python benchmarks/phase_probe_masked.py [rows]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tss_loader
tss = tss_loader.load()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ix = tss.FlatIndex(384)
ix.reserve(rows)
ix.add_synthetic(0, rows, 1)
ix.finalize()
q = np.random.default_rng(0).standard_normal((1, 384)).astype(np.float32)
dq = tss.DeviceBuffer(0, q.nbytes).upload(q)
dk = tss.DeviceBuffer(0, 80)
dbg = tss.DeviceBuffer(0, 256 * 8 * 8)
m = tss.Mask(rows)
m.set_rows(np.arange(10, dtype=np.uint32) * 1000)
for name, mask, mode in (("unmasked", None, tss.TSS_MASK_NONE), ("EXCLUDE 10 rows", m, tss.TSS_MASK_EXCLUDE)):
    for _ in range(3):
        ix.search_device(dq, 1, 10, dk, mask, mode)
    ix.sync()
    tss.lib().tss_index_debug_phases(ix.handle, dbg.ptr)
    for it in range(3):
        dbg.upload(np.zeros(256 * 8, np.uint64))
        ix.search_device(dq, 1, 10, dk, mask, mode)
        ix.sync()
        t = dbg.download(np.uint64, 256 * 8).reshape(256, 8)[:148].astype(np.int64)
        t0 = t[:, 0].min()
        end_scan = (t[:, 2] - t0) / 1e3
        print(name, "iter", it, "scan-loop end per CTA: min %.1f median %.1f max %.1f us | kernel total %.1f us"
              % (end_scan.min(), np.median(end_scan), end_scan.max(), (t[:, 5].max() - t0) / 1e3),
              "| CTAs ending >50 us before the last:", int((end_scan < end_scan.max() - 50).sum()))
    tss.lib().tss_index_debug_phases(ix.handle, None)
