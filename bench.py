#!/usr/bin/env python
"""bench.py -- queries/sec, exact cosine top-10 over a 10M x 384 fp32 corpus (BASELINE.json metric).

A "step" is one batch-1 query over the whole corpus.  `value` times K steps on the
device (queries and result buffers resident in HBM, CUDA events on the index stream);
`e2e` times the same K steps through the host-pointer C-ABI call tss_index_search
(H2D of the query and D2H of the result inside the timed region).  With N>1 ranks the
corpus is row-sharded (fixed total size: strong scaling) and each step ends in the
all-gather + merge of the per-rank top-k.

At N=1 the line also carries `batched`: the same metric for one 1024-query batch on the same
index (K2: tcgen05 GEMM over a bf16 shadow + exact re-scoring), whose keys must equal the
batch-1 legs' bit for bit -- a full-size cross-check of the two kernels -- with its own tensor
roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...       (one rank per GPU)

The oracle (oracle/, test infrastructure) appears here in three roles only: it synthesises the
queries (its seeded generator is the one the device uses to fill the corpus, so "planted" queries
can be built without reading rows back), it CHECKS the timed results (`check`), and it is the
thing timed in `cpu_baseline` / `--impl reference`.  No timed GPU leg calls it.

--impl reference times the CPU restatement of the reference contract (oracle/, all host
threads) on a bounded sample of the same workload; the reference itself has no scoring
implementation to run (reference src/vector.rs:195-202 is a stub).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec, exact top-10 over 10M x 384"
UNIT = "queries/s"
SEED_ROWS, SEED_Q = 0x5EED, 0xBEEF


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--storage", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true",
                    help="skip the extra 1024-query leg (same index, same metric, tensor-core path)")
    return ap.parse_args()


def load_orc():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    orc.lib()
    return orc


def make_queries(orc, n, dim, total_rows):
    """n seeded queries; every 4th is 'planted' (a corpus row plus noise -> a clear winner)."""
    q = orc.gen_rows(0, n, dim, SEED_Q)
    planted = {}
    for i in range(0, n, 4):
        row = (i * 2654435761 + 12345) % total_rows
        q[i] = orc.gen_rows(row, 1, dim, SEED_ROWS)[0] + 0.125 * q[i]
        planted[i] = row
    return q, planted


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v == "Active":
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm),
                       power_w_max=max(power), reasons=sorted(reasons))
        return out


def cpu_baseline(orc, args, queries):
    """The oracle (a port: the reference has no scoring code) on all host threads, on a bounded
    sample: cpu_sample_rows rows of the same synthetic corpus held in RAM; scaled to the full
    corpus by rows (the scan is linear in rows)."""
    n = min(args.cpu_sample_rows, args.rows)
    rows = orc.gen_rows(0, n, args.dim, SEED_ROWS)
    threads = orc.num_threads()
    orc.cosine_topk(rows, queries[0], args.k)  # warm-up (page in, spin up the OpenMP team)
    t0 = time.perf_counter()
    nqs = 0
    while True:
        orc.cosine_topk(rows, queries[nqs % len(queries)], args.k)
        nqs += 1
        dt = time.perf_counter() - t0
        if dt > 10.0 or nqs >= 1000:
            break
    qps_sample = nqs / dt
    scale = n / args.rows
    return {
        "value": qps_sample * scale, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": (f"{nqs} batch-1 queries over a {n}-row in-RAM slice of the same synthetic corpus "
                   f"({dt:.1f} s, {qps_sample:.2f} q/s on the slice), scaled by {n}/{args.rows} rows"),
    }, rows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    orc = load_orc()
    nq = args.steps + args.warmup
    queries, _ = make_queries(orc, max(nq, 1), args.dim, args.rows)
    n = min(args.cpu_sample_rows, args.rows)
    rows = orc.gen_rows(0, n, args.dim, SEED_ROWS)
    threads = orc.num_threads()
    # keep the whole run within a few minutes whatever K is
    budget_s, done = 120.0, 0
    for i in range(args.warmup):
        orc.cosine_topk(rows, queries[i], args.k)
    t0 = time.perf_counter()
    for i in range(args.steps):
        orc.cosine_topk(rows, queries[args.warmup + i], args.k)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    scale = n / args.rows
    value = done / dt * scale
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": dt / done * 1e3 / scale,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"each step = one batch-1 query over a {n}-row in-RAM slice, all {threads} host "
                       f"threads (OpenMP); value scaled by {n}/{args.rows} rows to the full corpus"),
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("the reference (Rust) has no scoring implementation and no toolchain here; this arm "
                 "is the CPU restatement of its contract in oracle/ (kind=port)"),
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n_gpus):
    return {
        "workload": (f"synthetic {args.rows}x{args.dim} {args.storage} corpus, batch-1 query, exact "
                     f"cosine top-{args.k} (BASELINE.json metric config)"),
        "rows": args.rows, "dim": args.dim, "k": args.k, "batch": 1, "storage": args.storage,
        "sharding": f"rows split over {n_gpus} GPU(s), all-gather of nq*k keys + merge" if n_gpus > 1
        else "single GPU",
        "l2": "corpus shard is larger than L2 (126 MB); no flush needed",
    }


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import tss_loader
    tss = tss_loader.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    if tss.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libtss has no CPU path")

    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    device = local_rank
    storage = tss.TSS_F32 if args.storage == "f32" else tss.TSS_BF16
    elem = 4 if args.storage == "f32" else 2

    # ---- shard + index -------------------------------------------------------------------
    per = (args.rows + world - 1) // world
    row_begin = min(rank * per, args.rows)
    n_local = min(per, args.rows - row_begin)
    comm = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(tss.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = tss.Comm(bytes(idt.cpu().numpy().tobytes()), rank, world, device)
    ix = tss.FlatIndex(args.dim, storage, device)
    ix.reserve(n_local)
    ix.add_synthetic(row_begin, n_local, SEED_ROWS)
    ix.set_shard(row_begin, comm)
    ix.finalize()

    orc = load_orc() if rank == 0 else None
    nq_total = args.steps + args.warmup
    if rank == 0:
        queries, planted = make_queries(orc, nq_total, args.dim, args.rows)
    else:
        queries, planted = np.empty((nq_total, args.dim), np.float32), {}
    if world > 1:
        qt = torch.from_numpy(queries).cuda()
        dist.broadcast(qt, 0)
        queries = qt.cpu().numpy()

    d_q = tss.DeviceBuffer(device, queries.nbytes).upload(queries)
    d_out = tss.DeviceBuffer(device, nq_total * args.k * 8)
    qstride, ostride = args.dim * 4, args.k * 8

    class _Slice:  # a view into a DeviceBuffer (the ABI takes raw pointers)
        def __init__(self, ptr):
            self.ptr = ptr

    def barrier():
        ix.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_leg(first, count):
        for i in range(first, first + count):
            ix.search_device(_Slice(d_q.ptr + i * qstride), 1, args.k, _Slice(d_out.ptr + i * ostride))

    # ---- value: device-resident, CUDA events on the index stream ------------------------------
    ev0, ev1 = tss.Event(device), tss.Event(device)
    device_leg(0, args.warmup)
    barrier()
    sampler = ClockSampler(device) if rank == 0 else None
    time.sleep(0.15)
    launches0 = tss.launch_count()
    ev0.record(ix)
    device_leg(args.warmup, args.steps)
    ev1.record(ix)
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_ms(ev1))
    launches = tss.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step
    keys_value_leg = (d_out.download(np.uint64, nq_total * args.k).reshape(nq_total, args.k)
                      if rank == 0 else None)

    # ---- roofline leg: the scan kernel alone (no gather/merge) --------------------------------
    if comm is not None:
        ix.set_shard(row_begin, None)
        device_leg(0, args.warmup)
        barrier()
        ev0.record(ix)
        device_leg(args.warmup, args.steps)
        ev1.record(ix)
        barrier()
        scan_ms = ev0.elapsed_ms(ev1) / args.steps
        ix.set_shard(row_begin, comm)
    else:
        scan_ms = ev0.elapsed_ms(ev1) / args.steps
    scan_ms = max_over_ranks(scan_ms)
    peaks, peak_kind = measured_peaks()
    algo_bytes = n_local * args.dim * elem
    achieved = algo_bytes / (scan_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("rows") == n_local and tj.get("dim") == args.dim and tj.get("storage") == args.storage:
            traffic = tj.get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": "scan_topk_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
        "peak_kind": f"of {peak_kind} (MEASURED_PEAKS.json copy bandwidth)", "unit": "GB/s",
        "frac": achieved / peaks["hbm_gbs"], "frac_of_nominal_8TBs": achieved / 8000.0,
        "traffic": traffic, "algorithmic_bytes_per_launch": algo_bytes,
        "launch_ms": scan_ms,
    }

    # ---- e2e: host pointers through tss_index_search ------------------------------------------
    # host buffers owned by the caller, addresses computed once: each step is exactly one
    # tss_index_search(host query pointer, host result pointers) call
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    out_rows = np.empty((nq_total, args.k), np.uint32)
    out_scores = np.empty((nq_total, args.k), np.float32)
    out_counts = np.empty(nq_total, np.uint32)
    qp, rp, sp, cp = (a.ctypes.data for a in (queries, out_rows, out_scores, out_counts))
    qs, rs = args.dim * 4, args.k * 4
    for i in range(args.warmup):
        ix.search_into(qp + i * qs, 1, args.k, rp + i * rs, sp + i * rs, cp + i * 4)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        ix.search_into(qp + i * qs, 1, args.k, rp + i * rs, sp + i * rs, cp + i * 4)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    last = (out_rows[nq_total - 1:nq_total], out_scores[nq_total - 1:nq_total])
    e2e = {"value": args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": args.dim * 4,
           "d2h_bytes_per_step": args.k * 8, "ms_per_step": e2e_s / args.steps * 1e3}

    # ---- sanity: the timed work is the real work ----------------------------------------------
    check = "skipped"
    if rank == 0:
        rows_out, scores_out = tss.unpack_keys(keys_value_leg)
        ok = True
        for i, row in planted.items():
            ok &= int(rows_out[i][0]) == row
        # recompute the last query's winners on the CPU from the generator: bit-exact scores
        i = nq_total - 1
        ok &= np.array_equal(rows_out[i], last[0][0])
        if args.storage == "f32":
            for r, s in zip(rows_out[i], scores_out[i]):
                e = orc.gen_rows(int(r), 1, args.dim, SEED_ROWS)
                ok &= orc.scores(e, queries[i])[0].view(np.uint32) == s.view(np.uint32)
        check = "ok" if ok else "FAILED"

    # ---- extra (N=1): the same metric for a 1024-query batch on the same index -----------------
    # K2: tcgen05 GEMM picks candidates (from a bf16 shadow of an fp32 index), survivors are
    # re-scored with the scan's arithmetic -- the keys must equal the batch-1 legs' bit for bit.
    batched = None
    if world == 1 and not args.no_batched and args.rows >= 4 * 256 * args.k:
        nb = 1024
        qb = orc.gen_rows(0, nb, args.dim, SEED_Q ^ 0x1234)
        qb[:min(nb, nq_total)] = queries[:min(nb, nq_total)]
        d_qb = tss.DeviceBuffer(device, qb.nbytes).upload(qb)
        d_kb = tss.DeviceBuffer(device, nb * args.k * 8)
        for _ in range(2):
            ix.search_device(d_qb, nb, args.k, d_kb)
        ix.sync()
        iters = 5
        l0 = tss.launch_count()
        ev0.record(ix)
        for _ in range(iters):
            ix.search_device(d_qb, nb, args.k, d_kb)
        ev1.record(ix)
        ix.sync()
        bms = ev0.elapsed_ms(ev1) / iters
        kb = d_kb.download(np.uint64, nb * args.k).reshape(nb, args.k)
        same = bool(np.array_equal(kb[:min(nb, nq_total)], keys_value_leg[:min(nb, nq_total)]))
        tf = 2.0 * nb * args.rows * args.dim / (bms * 1e-3) / 1e12
        batched = {
            "workload": f"{nb}-query batch, same index, exact cosine top-{args.k}", "batch": nb,
            "value": nb / bms * 1e3, "unit": UNIT, "ms_per_batch": bms,
            "gpu_launches_per_batch": (tss.launch_count() - l0) / iters,
            "keys_equal_batch1_leg": same,
            "roofline": {"bound": "tensor", "kernel": "gemm_topk_kernel", "achieved": tf,
                         "peak": peaks.get("bf16_tflops"), "unit": "TFLOP/s",
                         "frac": tf / peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None,
                         "algorithmic_flops_per_batch": 2.0 * nb * args.rows * args.dim},
        }
        if not same:
            check = "FAILED"

    # ---- extra (N=1): batch-1 queries answered from the bf16 shadow + exact re-scoring ---------
    # tss_index_set_batch_policy(1): the scan kernel streams the bf16 shadow of the fp32 index
    # (2 bytes per element instead of 4) for the top-64, a refine kernel proves the fp32 top-k is
    # among them and re-scores them from the fp32 rows.  Same keys as the fp32 scan; NOT the
    # configuration the HBM roofline above is quoted for.
    prefiltered = None
    if batched is not None:
        ix.set_batch_policy(1, True)
        npf = min(args.steps, 200)
        device_leg(0, args.warmup)
        ix.sync()
        ev0.record(ix)
        device_leg(args.warmup, npf)
        ev1.record(ix)
        ix.sync()
        pms = ev0.elapsed_ms(ev1) / npf
        kp = d_out.download(np.uint64, nq_total * args.k).reshape(nq_total, args.k)
        same = bool(np.array_equal(kp[:args.warmup + npf], keys_value_leg[:args.warmup + npf]))
        ix.set_batch_policy(0, False)
        prefiltered = {
            "workload": ("batch-1 queries, tss_index_set_batch_policy(1): scan of the bf16 shadow "
                         "+ proof + exact re-scoring from the fp32 rows"),
            "value": 1e3 / pms, "unit": UNIT, "ms_per_step": pms, "steps": npf,
            "keys_equal_batch1_leg": same,
            "bytes_streamed_per_query": n_local * args.dim * 2,
        }
        if not same:
            check = "FAILED"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_baseline(orc, args, queries)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.storage == "f32" else "bf16->f32",
            "data": "synthetic", "config": workload_config(args, world),
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "clocks": clocks, "check": check,
        }
        if batched is not None:
            line["batched"] = batched
        if prefiltered is not None:
            line["prefiltered"] = prefiltered
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        if comm is not None:
            ix.close()
            comm.close()
        dist.destroy_process_group()
    return 1 if check == "FAILED" else 0


if __name__ == "__main__":
    sys.exit(main())
