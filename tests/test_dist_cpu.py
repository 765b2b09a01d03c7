"""Host-side sharding logic, world_size 2 over gloo on the CPU (no GPU involved).

What the ranks exchange on the real path is an all-gather of nq*k packed keys followed
by a k-way merge; here the per-rank keys come from the oracle (test infrastructure) and
the exchange runs over gloo, checking the row partition, the global row ids, and that
merge(all_gather(local top-k)) == global top-k bit for bit.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _partition(n, world, rank):
    per = (n + world - 1) // world
    b = min(rank * per, n)
    return b, min(per, n - b)


def _worker(rank, world, port, n, k, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dim = 384
    q = orc.gen_rows(0, 3, dim, 0xBEEF)
    b, cnt = _partition(n, world, rank)
    rows = orc.gen_rows(b, cnt, dim, 0x5EED)
    r, s, c = orc.cosine_topk(rows, q, k, row_base=b, threads=1)
    keys = np.zeros((3, k), dtype=np.uint64)
    for qi in range(3):
        for j in range(c[qi]):
            keys[qi, j] = orc.pack_key(float(s[qi, j]), int(r[qi, j]))
    local = torch.from_numpy(keys.view(np.int64))
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    allk = np.stack([g.numpy().view(np.uint64) for g in gathered])  # [P][nq][k]
    merged = np.sort(allk.transpose(1, 0, 2).reshape(3, -1), axis=1)[:, ::-1][:, :k]
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), merged)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(20_001, 10), (7, 10)])
def test_sharded_merge_equals_global_topk(tmp_path, orc, tss, n, k):
    import torch.multiprocessing as mp
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, k, str(tmp_path)), nprocs=world, join=True)
    rows = orc.gen_rows(0, n, 384, 0x5EED)
    q = orc.gen_rows(0, 3, 384, 0xBEEF)
    wr, ws, wc = orc.cosine_topk(rows, q, k)
    for rank in range(world):
        merged = np.load(os.path.join(str(tmp_path), f"merged_{rank}.npy"))
        gr, gs = tss.unpack_keys(merged)
        assert np.array_equal(gr, wr)
        assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32))
        assert np.array_equal((merged != 0).sum(axis=1), wc)


def test_partition_covers_rows_exactly_once():
    for n in (0, 1, 7, 10_000_000, 100_000_003):
        for world in (1, 2, 4, 8):
            spans = [_partition(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for b, c in spans:
                if c:
                    assert b == pos
                pos += c
