"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tss_loader, orc
tss = tss_loader.load()
n, dim = 12_345, 384
rows = orc.gen_rows(0, n, dim, 1)
q = orc.gen_rows(0, 5, dim, 2)
for storage, bf in ((tss.TSS_F32, False), (tss.TSS_BF16, True)):
    ix = tss.FlatIndex(dim, storage)
    ix.add(rows[:5000]); ix.add_synthetic(5000, n - 5000, 1); ix.finalize()
    for k in (10, 100, 200):
        got = ix.search(q, k)
        want = orc.cosine_topk(rows, q, k, bf16=bf)
        assert np.array_equal(got[0], want[0]), (storage, k)
    m = tss.Mask(n); m.set_rows(np.arange(0, n, 7, dtype=np.uint32))
    assert m.popcount() == len(range(0, n, 7))
    for mode, om in ((tss.TSS_MASK_INCLUDE, orc.MASK_INCLUDE), (tss.TSS_MASK_EXCLUDE, orc.MASK_EXCLUDE)):
        got = ix.search(q[:2], 10, m, mode)
        want = orc.cosine_topk(rows, q[:2], 10, m.download(), om, bf16=bf)
        assert np.array_equal(got[0], want[0])
terms = tss.Terms([b"a b", b"a c", b"d"], [[0, 5, 99], [100, 150], [199, 3]])
m = tss.Mask(n); terms.prefix_mask(b"a", m); assert m.popcount() == 5
# tensor-core path, small
ixb = tss.FlatIndex(dim, tss.TSS_BF16); ixb.add(rows); ixb.finalize()
qq = orc.gen_rows(0, 64, dim, 3)
r, s, c = ixb.search(qq, 10)
assert np.all(c == 10)
print("SANITIZE_SMOKE_OK launches", tss.launch_count())
