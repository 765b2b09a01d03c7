"""Run under torchrun on >= 2 GPUs: sharded search (fused peer-memory exchange, or NCCL
all-gather + merge kernel) == oracle on the unsharded corpus, bit for bit.

run_checks() is also what `bench.py --gpus N` (N > 1) runs as its `selftest`, so the driver's
multi-GPU lease exercises the equivalence even when the 1-GPU test lease skips test_dist_gpu.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def run_checks(tss, orc, comm, rank, world, local, quick=False):
    """-> True when every sharded result on this rank equals the oracle's on the whole corpus"""
    dim, seed = 384, 0x5EED
    ok = True
    # every rank runs the oracle at the same time: share the cores (torchrun sets OMP_NUM_THREADS=1)
    thr = max(1, len(os.sched_getaffinity(0)) // world)
    cases = [(200_003, 10, 5), (1000, 50, 2), (5, 10, 1), (300_000, 128, 3)]
    if quick:
        cases = [(200_003, 10, 5), (5, 10, 1), (100_000, 128, 3)]
    for n, k, nq in cases:
        per = (n + world - 1) // world
        b = min(rank * per, n)
        cnt = min(per, n - b)
        ix = tss.FlatIndex(dim, tss.TSS_F32, local)
        ix.add_synthetic(b, cnt, seed)
        ix.set_shard(b, comm)
        ix.finalize()
        q = orc.gen_rows(0, nq, dim, 0xBEEF)
        got = ix.search(q, k)
        # device-resident variant too
        dq = tss.DeviceBuffer(local, q.nbytes).upload(q)
        dk = tss.DeviceBuffer(local, nq * k * 8)
        ix.search_device(dq, nq, k, dk)
        ix.sync()
        r2, s2 = tss.unpack_keys(dk.download(np.uint64, nq * k).reshape(nq, k))
        rows = orc.gen_rows(0, n, dim, seed)
        want = orc.cosine_topk(rows, q, k, threads=thr)
        same = (np.array_equal(got[0], want[0]) and
                np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)) and
                np.array_equal(got[2], want[2]) and np.array_equal(r2, want[0]) and
                np.array_equal(s2.view(np.uint32), want[1].view(np.uint32)))
        if not same:
            print(f"rank {rank}: MISMATCH n={n} k={k}", flush=True)
        ok &= same
        ix.close()
    # stress: hundreds of back-to-back tiny sharded scans.  Launches overlap under programmatic
    # dependent launch, so their peer-memory exchanges must still happen in launch order.
    n, k, reps = 300, 10, 200 if quick else 400
    per = (n + world - 1) // world
    b = min(rank * per, n)
    cnt = min(per, n - b)
    ix = tss.FlatIndex(dim, tss.TSS_F32, local)
    ix.add_synthetic(b, cnt, seed)
    ix.set_shard(b, comm)
    ix.finalize()
    q = orc.gen_rows(0, 4, dim, 0xBEEF)
    rows = orc.gen_rows(0, n, dim, seed)
    want = orc.cosine_topk(rows, q, k, threads=thr)
    dq = tss.DeviceBuffer(local, q.nbytes).upload(q)
    dk = tss.DeviceBuffer(local, reps * k * 8)

    class _Off:
        def __init__(self, ptr):
            self.ptr = ptr
    for i in range(reps):
        ix.search_device(_Off(dq.ptr + (i % 4) * dim * 4), 1, k, _Off(dk.ptr + i * k * 8))
    ix.sync()
    rr, ss = tss.unpack_keys(dk.download(np.uint64, reps * k).reshape(reps, k))
    same = all(np.array_equal(rr[i], want[0][i % 4]) and
               np.array_equal(ss[i].view(np.uint32), want[1][i % 4].view(np.uint32)) for i in range(reps))
    if not same:
        print(f"rank {rank}: back-to-back sharded scans MISMATCH", flush=True)
    ok &= same
    # the same from host memory with searches in flight (tss_index_search_submit / _collect, one
    # thread per rank, every rank in the same order): merged result on every rank
    qc = np.ascontiguousarray(q)
    outs = [(np.empty((1, k), np.uint32), np.empty((1, k), np.float32), np.empty(1, np.uint32))
            for _ in range(60)]
    tickets = []
    for i in range(60):
        if len(tickets) == 3:
            j, t = tickets.pop(0)
            ix.search_collect(t, *(a.ctypes.data for a in outs[j]))
        tickets.append((i, ix.search_submit(qc[i % 4].ctypes.data, 1, k)))
    for j, t in tickets:
        ix.search_collect(t, *(a.ctypes.data for a in outs[j]))
    same = all(np.array_equal(outs[i][0][0], want[0][i % 4]) and
               np.array_equal(outs[i][1][0].view(np.uint32), want[1][i % 4].view(np.uint32)) and
               int(outs[i][2][0]) == k for i in range(60))
    if not same:
        print(f"rank {rank}: pipelined sharded searches MISMATCH", flush=True)
    ok &= same
    ix.close()
    # K2 (tensor-core path) on a sharded bf16 index: local GEMM top-k, NCCL all-gather, merge
    # (survivors are re-scored with the scan's arithmetic, so the merged result is bit-identical
    # to the oracle on the unsharded bf16 corpus; 64 queries: independent CTAs, 200: CTA pairs)
    n, k = 400_000, 10
    per = (n + world - 1) // world
    b = min(rank * per, n)
    cnt = min(per, n - b)
    ix = tss.FlatIndex(dim, tss.TSS_BF16, local)
    ix.add_synthetic(b, cnt, seed)
    ix.set_shard(b, comm)
    ix.finalize()
    rows = orc.gen_rows(0, n, dim, seed)
    for nq in ((200,) if quick else (64, 200, 1030)):  # 1030: a 1024 chunk + a 6-query tail
        q = orc.gen_rows(0, nq, dim, 0xBEEF)
        q[0] = rows[123_456] + 0.125 * q[0]
        gr, gs, gc = ix.search(q, k)
        want = orc.cosine_topk(rows, q, k, bf16=True, threads=thr)
        same = (bool(np.all(gc == k)) and gr[0][0] == 123_456 and np.array_equal(gr, want[0])
                and np.array_equal(gs.view(np.uint32), want[1].view(np.uint32)))
        if not same:
            hits = sum(len(set(g.tolist()) & set(w.tolist())) for g, w in zip(gr, want[0]))
            print(f"rank {rank}: K2 sharded MISMATCH nq {nq} recall {hits / (nq * k):.3f}", flush=True)
        ok &= same
        # the device-resident entry routes the WHOLE call one way (a short tail chunk
        # used to take the fused scan and its results never reached the output)
        dq = tss.DeviceBuffer(local, q.nbytes).upload(q)
        dk = tss.DeviceBuffer(local, nq * k * 8)
        ix.search_device(dq, nq, k, dk)
        ix.sync()
        r2, s2 = tss.unpack_keys(dk.download(np.uint64, nq * k).reshape(nq, k))
        same = np.array_equal(r2, want[0]) and np.array_equal(s2.view(np.uint32), want[1].view(np.uint32))
        if not same:
            print(f"rank {rank}: K2 sharded search_device MISMATCH nq {nq}", flush=True)
        ok &= same
    ix.close()
    return ok


def main():
    import torch
    import torch.distributed as dist
    import tss_loader
    import orc
    tss = tss_loader.load()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(tss.Comm.unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    comm = tss.Comm(idt.cpu().numpy().tobytes(), rank, world, local)
    ok = run_checks(tss, orc, comm, rank, world, local)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAILED", flush=True)
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
