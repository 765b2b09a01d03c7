"""tss_index_add throughput from pageable host memory, by staging-copy thread count
(TSS_UPLOAD_THREADS is read once per process, so each count runs in a child process).

  python benchmarks/upload_probe.py            # sweep 1, 2, 4, 8 threads
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import tss_loader
    tss = tss_loader.load()
    n, dim = 1_048_576, 384
    rows = np.random.default_rng(1).standard_normal((n, dim), dtype=np.float32)
    ix = tss.FlatIndex(dim)
    ix.reserve(3 * n)
    ix.add(rows[:4096])        # staging buffers allocated
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        ix.add(rows)
        ix.finalize()
        best = min(best, time.perf_counter() - t0)
    back = ix.get_rows(4096 + n + 777, 1)
    print(json.dumps({"threads": os.environ.get("TSS_UPLOAD_THREADS", "default"),
                      "bytes": rows.nbytes, "seconds": best, "gbs": rows.nbytes / best / 1e9,
                      "check": "ok" if np.array_equal(back[0], rows[777]) else "FAILED"}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for t in ("1", "2", "4", "8", None):
            env = dict(os.environ)
            if t is None:
                env.pop("TSS_UPLOAD_THREADS", None)
            else:
                env["TSS_UPLOAD_THREADS"] = t
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False)
