"""Diagnostics: per-phase %globaltimer stamps of the scan kernel (tss_index_debug_phases)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tss_loader
tss = tss_loader.load()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ix = tss.FlatIndex(384)
ix.add_synthetic(0, rows, 1)
ix.finalize()
q = np.random.default_rng(0).standard_normal((1, 384)).astype(np.float32)
dq = tss.DeviceBuffer(0, q.nbytes).upload(q)
dk = tss.DeviceBuffer(0, 80)
dbg = tss.DeviceBuffer(0, 256 * 8 * 8)
for _ in range(5):
    ix.search_device(dq, 1, 10, dk)
ix.sync()
tss.lib().tss_index_debug_phases(ix.handle, dbg.ptr)
res = []
for it in range(5):
    dbg.upload(np.zeros(256 * 8, np.uint64))
    ix.search_device(dq, 1, 10, dk)
    ix.sync()
    t = dbg.download(np.uint64, 256 * 8).reshape(256, 8)[:148].astype(np.int64)
    t0 = t[:, 0].min()
    last = int(np.argmax(t[:, 5]))
    print("iter", it, "start spread %.1f us" % ((t[:, 0].max() - t0) / 1e3),
          "| median: qload %.1f scan %.1f prune+sync %.1f merge+ticket %.1f" % tuple(
              np.median(t[:, i + 1] - t[:, i]) / 1e3 for i in range(4)),
          "| last cta %d: final %.1f, kernel total %.1f us" % (
              last, (t[last, 5] - t[last, 4]) / 1e3, (t[last, 5] - t0) / 1e3),
          "| ticket time max-min %.1f" % ((t[:, 4].max() - t[:, 4].min()) / 1e3))
