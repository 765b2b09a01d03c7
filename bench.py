#!/usr/bin/env python
"""bench.py -- queries/sec, exact cosine top-10 over a 10M x 384 fp32 corpus (BASELINE.json metric).

A "step" is one batch-1 query over the whole corpus.  `value` times K steps on the
device (queries and result buffers resident in HBM, CUDA events on the index stream,
a barrier immediately before the first event so no rank's events time another rank's
start-up); `e2e` times the same K steps through the host-pointer C-ABI call
tss_index_search (H2D of the query and D2H of the result inside the timed region).
With N>1 ranks the corpus is row-sharded (fixed total size: strong scaling) and each
step ends in the exchange + merge of the per-rank top-k.

The same JSON line carries the other BASELINE.json configs as objects of their own:
  batched      (N=1) a 1024-query batch on the SAME fp32 index through K2 (tcgen05 GEMM over a
               bf16 shadow + exact re-scoring); its keys must equal the batch-1 legs' bit for bit
  prefiltered  (N=1) batch-1 queries answered from the bf16 shadow + proof + exact re-scoring
  config3      (N=1) 10M x 384 bf16 index, 1024 queries, top-100: burst and >= 2 s sustained
               tensor roofline, recall@100 / @10 against the fp32 exact result
  config4      (N=1) hybrid: ~5M-term flattened trie -> prefix mask (K4) -> masked top-10 (K1) at
               three selectivities; masks checked bit for bit against numpy
  config5      100M x 384 fp32 row-sharded over the N GPUs (N=1: all 153.6 GB in one B200's HBM):
               q/s, per-GPU GB/s, result checked against the 10M-row index (the first 10M rows
               are the same rows)
  top50        batch-1 exact top-50, the call the reference's merge makes (src/search.rs:251)
  selftest     (N>1) sharded == unsharded oracle, bit for bit (tests/dist_worker.py:run_checks)

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
                    help="skip the 1024-query and prefiltered legs on the metric index")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the config3 / config4 / config5 / selftest legs")
    ap.add_argument("--terms", type=int, default=5_000_000, help="config4: terms of the flattened trie")
    ap.add_argument("--config5-rows", type=int, default=100_000_000)
    return ap.parse_args()


def host_threads():
    """threads the CPU legs use: every core this process may run on (torchrun exports
    OMP_NUM_THREADS=1, which is not what "all host cores" means)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_orc():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    orc.lib()
    return orc


def make_queries(orc, n, dim, total_rows, seed=SEED_Q):
    """n seeded queries; every 4th is 'planted' (a corpus row plus noise -> a clear winner)."""
    q = orc.gen_rows(0, n, dim, seed)
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
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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
    threads = host_threads()
    orc.cosine_topk(rows, queries[0], args.k, threads=threads)  # warm-up (page in, spin up the team)
    t0 = time.perf_counter()
    nqs = 0
    while True:
        orc.cosine_topk(rows, queries[nqs % len(queries)], args.k, threads=threads)
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
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    orc = load_orc()
    nq = args.steps + args.warmup
    queries, _ = make_queries(orc, max(nq, 1), args.dim, args.rows)
    n = min(args.cpu_sample_rows, args.rows)
    rows = orc.gen_rows(0, n, args.dim, SEED_ROWS)
    threads = host_threads()  # stated explicitly: never inherited from OMP_NUM_THREADS
    # keep the whole run within a few minutes whatever K is
    budget_s, done = 120.0, 0
    for i in range(args.warmup):
        orc.cosine_topk(rows, queries[i], args.k, threads=threads)
    t0 = time.perf_counter()
    for i in range(args.steps):
        orc.cosine_topk(rows, queries[args.warmup + i], args.k, threads=threads)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    scale = n / args.rows
    value = done / dt * scale
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup,
        # the time one step really took (a step = one query over the SAMPLE, below)
        "ms_per_step": dt / done * 1e3,
        "ms_per_query_full_corpus_equivalent": dt / done * 1e3 / scale,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"each step = one batch-1 query over a {n}-row in-RAM slice, {threads} host "
                       f"threads (OpenMP, set explicitly); value = steps/s scaled by {n}/{args.rows} "
                       "rows to the full corpus (the scan is linear in rows)"),
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
        "sharding": f"rows split over {n_gpus} GPU(s), exchange of nq*k keys + merge" if n_gpus > 1
        else "single GPU",
        "l2": "corpus shard is larger than L2 (126 MB); no flush needed",
    }


def _json_default(o):
    if isinstance(o, np.generic):
        return o.item()
    raise TypeError(f"Object of type {o.__class__.__name__} is not JSON serializable")


class _Slice:  # a view into a DeviceBuffer (the ABI takes raw pointers)
    def __init__(self, ptr):
        self.ptr = ptr


def _mask_words(rows, n):
    w = np.zeros((n + 31) // 32, dtype=np.uint32)
    rows = np.asarray(rows, dtype=np.int64)
    if rows.size:
        np.bitwise_or.at(w, rows >> 5, np.uint32(1) << (rows & 31).astype(np.uint32))
    return w


# ---- config 3: bf16 index, 1024-query batch, top-100, recall vs fp32 exact --------------------
def config3_leg(tss, orc, args, ix_f32, device, peaks):
    rows, dim, nb, k = args.rows, args.dim, 1024, 100
    ix = tss.FlatIndex(dim, tss.TSS_BF16, device)
    ix.reserve(rows)
    ix.add_synthetic(0, rows, SEED_ROWS)
    ix.finalize()
    q, planted = make_queries(orc, nb, dim, rows, SEED_Q ^ 0x3333)
    d_q = tss.DeviceBuffer(device, q.nbytes).upload(q)
    d_k = tss.DeviceBuffer(device, nb * k * 8)
    ev0, ev1 = tss.Event(device), tss.Event(device)
    for _ in range(3):
        ix.search_device(d_q, nb, k, d_k)
    ix.sync()
    # burst: 5 batches from idle clocks
    iters = 5
    l0 = tss.launch_count()
    ev0.record(ix)
    for _ in range(iters):
        ix.search_device(d_q, nb, k, d_k)
    ev1.record(ix)
    ix.sync()
    burst_ms = ev0.elapsed_ms(ev1) / iters
    launches = (tss.launch_count() - l0) / iters
    # sustained: back to back for >= 2 s (the chip reaches its power cap and steady clocks)
    n_sus = max(20, int(2200.0 / burst_ms))
    ev0.record(ix)
    for _ in range(n_sus):
        ix.search_device(d_q, nb, k, d_k)
    ev1.record(ix)
    ix.sync()
    sus_total = ev0.elapsed_ms(ev1)
    sus_ms = sus_total / n_sus
    keys = d_k.download(np.uint64, nb * k).reshape(nb, k)
    gr, gs = tss.unpack_keys(keys)
    flops = 2.0 * nb * rows * dim
    # fp32 exact top-100 of the same queries: the fp32 index through the same library (K2 over the
    # shadow + exact re-scoring, bit-identical to the fp32 scan and to the oracle: see `check`,
    # `batched.keys_equal_batch1_leg` and tests/test_gemm_gpu.py) ...
    nrec = 256
    d_q2 = tss.DeviceBuffer(device, q[:nrec].nbytes).upload(q[:nrec])
    d_k2 = tss.DeviceBuffer(device, nrec * k * 8)
    ix_f32.search_device(d_q2, nrec, k, d_k2)
    ix_f32.sync()
    er, es = tss.unpack_keys(d_k2.download(np.uint64, nrec * k).reshape(nrec, k))
    # ... pinned to the CPU oracle on the full 10M-row corpus for a few of them
    nspot = 4
    t0 = time.perf_counter()
    orr, ors, _ = orc.cosine_topk_synth(0, rows, dim, SEED_ROWS, q[:nspot], k, threads=host_threads())
    oracle_s = time.perf_counter() - t0
    spot_ok = bool(np.array_equal(er[:nspot], orr) and
                   np.array_equal(es[:nspot].view(np.uint32), ors.view(np.uint32)))
    rec100 = float(np.mean([len(set(gr[i].tolist()) & set(er[i].tolist())) / k for i in range(nrec)]))
    rec10 = float(np.mean([len(set(gr[i][:10].tolist()) & set(er[i][:10].tolist())) / 10
                           for i in range(nrec)]))
    planted_ok = all(int(gr[i][0]) == r for i, r in planted.items())
    ok = spot_ok and planted_ok and rec100 >= 0.98 and rec10 >= 0.97
    ix.close()
    tf_b = flops / (burst_ms * 1e-3) / 1e12
    tf_s = flops / (sus_ms * 1e-3) / 1e12
    return {
        "workload": f"synthetic {rows}x{dim} bf16 corpus, {nb}-query batch, tcgen05 GEMM + top-{k}",
        "value": nb / burst_ms * 1e3, "unit": UNIT, "ms_per_batch": burst_ms,
        "sustained": {"value": nb / sus_ms * 1e3, "unit": UNIT, "ms_per_batch": sus_ms,
                      "batches": n_sus, "seconds": sus_total / 1e3},
        "gpu_launches_per_batch": launches,
        "roofline": {"bound": "tensor", "kernel": "gemm_topk_kernel", "achieved": tf_b,
                     "peak": peaks.get("bf16_tflops"), "unit": "TFLOP/s",
                     "frac": tf_b / peaks["bf16_tflops"],
                     "achieved_sustained": tf_s, "peak_sustained": peaks.get("bf16_tflops_sustained"),
                     "frac_sustained": tf_s / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]),
                     "algorithmic_flops_per_batch": flops},
        "recall_at_100_vs_fp32_exact": rec100, "recall_at_10_vs_fp32_exact": rec10,
        "recall_queries": nrec,
        "fp32_exact_reference": ("the fp32 index through the same library; its top-100 of "
                                 f"{nspot} of these queries equals the CPU oracle's over all {rows} "
                                 "rows bit for bit: " + str(spot_ok)),
        "oracle_seconds": oracle_s, "check": "ok" if ok else "FAILED",
    }


# ---- config 4: flattened trie prefix filter -> masked top-10 ---------------------------------
def config4_leg(tss, orc, args, ix, device):
    N, T, dim, k = args.rows, args.terms, args.dim, args.k
    rng = np.random.default_rng(5)
    t0 = time.perf_counter()
    ntok = rng.integers(1, 5, size=int(T * 1.15))
    tok = (rng.zipf(1.1, size=(ntok.size, 4)) - 1) % 200_000
    tok[np.arange(4)[None, :] >= ntok[:, None]] = -1
    tok = np.unique(tok, axis=0)[:T]  # lexicographic on ids == byte order of the joined strings
    T = tok.shape[0]
    ntok = (tok >= 0).sum(axis=1)
    npost = rng.geometric(0.25, size=T)
    post_off = np.zeros(T + 1, dtype=np.uint64)
    np.cumsum(npost, out=post_off[1:])
    post_rows = rng.integers(0, N, size=int(post_off[-1]), dtype=np.uint32)
    gen_s = time.perf_counter() - t0
    # the flattened trie is built on the device (N2) from tokenised postings in arbitrary order
    vocab = [b"w%06d" % i for i in range(200_000)]
    tuple_of_posting = np.repeat(np.arange(T), npost)
    perm = rng.permutation(tuple_of_posting.size)
    ids_dev = (tok[tuple_of_posting[perm]] + 1).astype(np.uint32)
    t0 = time.perf_counter()
    terms = tss.Terms.build(vocab, ids_dev, post_rows[perm], device)
    build_s = time.perf_counter() - t0
    assert terms.size() == T
    terms.bind_stream(ix)  # prefix -> mask -> masked search back to back on the index stream
    mask = tss.Mask(N, device)
    q, _ = make_queries(orc, 8, dim, N, SEED_Q ^ 0x4444)
    d_q = tss.DeviceBuffer(device, q.nbytes).upload(q)
    d_k = tss.DeviceBuffer(device, 8 * k * 8)
    out_rows = np.empty((1, k), np.uint32)
    out_scores = np.empty((1, k), np.float32)
    out_counts = np.empty(1, np.uint32)
    first = tok[:, 0]
    per_first = np.bincount(first, weights=npost, minlength=200_000)
    ev0, ev1 = tss.Event(device), tss.Event(device)
    iters = 30
    cases, ok_all = [], True
    for target in (N * 1e-5, N * 1e-3, N * 1e-1):
        cand = int(np.argmin(np.abs(per_first - target)))
        prefix = b"w%06d" % cand
        lo, hi = np.searchsorted(first, cand), np.searchsorted(first, cand, side="right")
        # posting order inside a term is the arrival order (perm), the SET is what the mask holds
        want_rows = np.unique(post_rows[int(post_off[lo]):int(post_off[hi])])
        want = _mask_words(want_rows, N)
        st = terms.prefix_mask(prefix, mask, fresh=True)
        got = mask.download()
        ok = bool(np.array_equal(got, want)) and mask.popcount() == int(want_rows.size)
        ok = ok and int(st.npostings) == int(post_off[hi] - post_off[lo])
        # device-timed: K4 alone, then K4 + masked K1 per query, no host synchronisation inside
        for _ in range(3):
            terms.prefix_mask(prefix, mask, want_stats=False, fresh=True)
            ix.search_device(d_q, 1, k, d_k, mask, tss.TSS_MASK_INCLUDE)
        ix.sync()
        ev0.record(ix)
        for _ in range(iters):
            terms.prefix_mask(prefix, mask, want_stats=False, fresh=True)
        ev1.record(ix)
        ix.sync()
        k4_us = ev0.elapsed_ms(ev1) / iters * 1e3
        ev0.record(ix)
        for i in range(iters):
            terms.prefix_mask(prefix, mask, want_stats=False, fresh=True)
            ix.search_device(_Slice(d_q.ptr + (i % 8) * dim * 4), 1, k, _Slice(d_k.ptr + (i % 8) * k * 8),
                             mask, tss.TSS_MASK_INCLUDE)
        ev1.record(ix)
        ix.sync()
        dev_us = ev0.elapsed_ms(ev1) / iters * 1e3
        # end to end through host pointers: prefix (host bytes) -> mask -> masked search -> rows
        t0 = time.perf_counter()
        for i in range(iters):
            terms.prefix_mask(prefix, mask, want_stats=False, fresh=True)
            ix.search_into(q[i % 8].ctypes.data, 1, k, out_rows.ctypes.data, out_scores.ctypes.data,
                           out_counts.ctypes.data, mask, tss.TSS_MASK_INCLUDE)
        e2e_us = (time.perf_counter() - t0) / iters * 1e6
        # ... and as ONE host call (tss_index_search_prefix)
        two_call = (out_rows.copy(), out_scores.copy(), out_counts.copy())
        for i in range(3):
            ix.search_prefix_into(terms, prefix, mask, q[i % 8].ctypes.data, 1, k, out_rows.ctypes.data,
                                  out_scores.ctypes.data, out_counts.ctypes.data)
        t0 = time.perf_counter()
        for i in range(iters):
            ix.search_prefix_into(terms, prefix, mask, q[i % 8].ctypes.data, 1, k, out_rows.ctypes.data,
                                  out_scores.ctypes.data, out_counts.ctypes.data)
        e2e1_us = (time.perf_counter() - t0) / iters * 1e6
        ok = ok and all(bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
                        for a, b in zip(two_call, (out_rows, out_scores, out_counts)))
        # ... and with two hybrid queries in flight (tss_index_search_prefix_submit, one scratch
        # mask each): the next query's prefix search runs while this one's rows stream
        mask2 = tss.Mask(N, device)
        pm = (mask, mask2)
        tk = []
        t0 = time.perf_counter()
        for i in range(iters):
            if len(tk) == 2:
                ix.search_collect(tk.pop(0), out_rows.ctypes.data, out_scores.ctypes.data, out_counts.ctypes.data)
            tk.append(ix.search_prefix_submit(terms, prefix, pm[i % 2], q[i % 8].ctypes.data, 1, k))
        for t_ in tk:
            ix.search_collect(t_, out_rows.ctypes.data, out_scores.ctypes.data, out_counts.ctypes.data)
        e2e2_us = (time.perf_counter() - t0) / iters * 1e6
        ok = ok and all(bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
                        for a, b in zip(two_call, (out_rows, out_scores, out_counts)))
        # ... and with the prefix searches on the terms' OWN stream (three in flight): query i+1's
        # prefix search then overlaps query i's scan on the device; the masks' events order them
        terms.bind_stream(None)
        mask3 = tss.Mask(N, device)
        pm3 = (mask, mask2, mask3)
        tk = []
        for rep in range(2):   # (first pass warms the cross-stream path)
            t0 = time.perf_counter()
            for i in range(iters):
                if len(tk) == 3:
                    ix.search_collect(tk.pop(0), out_rows.ctypes.data, out_scores.ctypes.data, out_counts.ctypes.data)
                tk.append(ix.search_prefix_submit(terms, prefix, pm3[i % 3], q[i % 8].ctypes.data, 1, k))
            for t_ in tk:
                ix.search_collect(t_, out_rows.ctypes.data, out_scores.ctypes.data, out_counts.ctypes.data)
            tk = []
            e2e3_us = (time.perf_counter() - t0) / iters * 1e6
        ok = ok and all(bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
                        for a, b in zip(two_call, (out_rows, out_scores, out_counts)))
        ix.sync()
        terms.bind_stream(ix)
        mask3.close()
        mask2.close()
        # the last query's result against the oracle: every live row scored on the CPU when the
        # mask is small, else the returned rows' score bits + membership
        qi = (iters - 1) % 8
        pc = int(want_rows.size)
        if pc <= 20_000:
            live = np.concatenate([orc.gen_rows(int(r), 1, dim, SEED_ROWS) for r in want_rows]) \
                if pc else np.zeros((0, dim), np.float32)
            sc = orc.scores(live, q[qi]) if pc else np.zeros(0, np.float32)
            order = sorted(range(pc), key=lambda j: (-float(sc[j]), int(want_rows[j])))[:k]
            exp_rows = [int(want_rows[j]) for j in order]
            exp_bits = [int(sc[j].view(np.uint32)) for j in order]
            n_out = int(out_counts[0])
            ok = ok and n_out == min(k, pc) and out_rows[0][:n_out].tolist() == exp_rows
            ok = ok and [int(x) for x in out_scores[0][:n_out].view(np.uint32)] == exp_bits
        else:
            ok = ok and int(out_counts[0]) == k and bool(np.all(np.isin(out_rows[0], want_rows)))
            for r, s in zip(out_rows[0], out_scores[0]):
                e = orc.gen_rows(int(r), 1, dim, SEED_ROWS)
                ok = ok and orc.scores(e, q[qi])[0].view(np.uint32) == s.view(np.uint32)
        ok_all = ok_all and ok
        scan_us = max(dev_us - k4_us, 1e-3)
        algo = pc * dim * 4 + 2 * (N // 8) + 4 * int(st.npostings)
        cases.append({
            "prefix": prefix.decode(), "selectivity": pc / N, "mask_popcount": pc,
            "postings_in_range": int(st.npostings),
            "terms_in_range": int(st.sub_hi - st.sub_lo + st.exact_hi - st.exact_lo),
            "mask_bit_exact_vs_numpy": ok, "prefix_to_mask_us_device": k4_us,
            "query_us_device": dev_us, "query_us_e2e_host_pointers": e2e_us,
            "query_us_e2e_one_call": e2e1_us, "query_us_e2e_two_in_flight": e2e2_us,
            "query_us_e2e_three_in_flight_two_streams": e2e3_us,
            "algorithmic_bytes": algo, "dense_scan_bytes": N * dim * 4,
            "did": "row-skipping scan of the live rows (not a dense scan)",
            "masked_scan_gbs_on_live_rows": pc * dim * 4 / (scan_us * 1e-6) / 1e9,
            "algorithmic_gbs": algo / (dev_us * 1e-6) / 1e9,
        })
    terms.bind_stream(None)
    terms.close()
    mask.close()
    return {
        "workload": (f"hybrid: {T}-term flattened trie (built on the device from "
                     f"{int(post_off[-1])} tokenised postings), token-prefix mask over {N} rows, "
                     f"masked exact top-{k}"),
        "terms": int(T), "postings": int(post_off[-1]), "host_generation_seconds": gen_s,
        "device_build_seconds_incl_h2d_and_validation": build_s,
        "bytes_formula": "popcount(mask)*D*4 + N/8 (mask read) + 4*postings_in_range + N/8 (mask clear)",
        "cases": cases, "check": "ok" if ok_all else "FAILED",
    }


# ---- config 5: 100M x 384 fp32 row-sharded over the ranks ---------------------------------------
def config5_leg(tss, orc, args, ix10, comm, rank, world, device, barrier, max_over_ranks, dist, torch):
    total, dim, k = args.config5_rows, args.dim, args.k
    per = (total + world - 1) // world
    b = min(rank * per, total)
    n_local = min(per, total - b)
    need = n_local * dim * 4 * 1.02 + (2 << 30)
    ix = tss.FlatIndex(dim, tss.TSS_F32, device)
    if world > 1:
        free_b = torch.cuda.mem_get_info()[0]
        fits = torch.tensor([1 if free_b > need else 0], device="cuda")
        dist.all_reduce(fits, op=dist.ReduceOp.MIN)
        if int(fits.item()) == 0:
            ix.close()
            return {"skipped": f"a {n_local}-row shard ({need / 1e9:.1f} GB) does not fit next to the "
                               f"10M-row index on every GPU ({free_b / 1e9:.1f} GB free on rank {rank})"}
        ix.reserve(n_local)
    else:
        # one GPU: the whole 100M x 384 fp32 matrix (153.6 GB) in one B200's HBM, next to the
        # 10M-row index it is checked against; the library reports OOM, it does not die of it
        try:
            ix.reserve(n_local)
        except tss.TssError as e:
            ix.close()
            return {"skipped": f"{n_local} rows ({need / 1e9:.1f} GB) do not fit on this GPU: {e}"}
    ix.add_synthetic(b, n_local, SEED_ROWS)
    ix.set_shard(b, comm)
    ix.finalize()
    steps, warm = (60, 5) if world > 1 else (30, 3)
    nq = steps + warm
    if rank == 0:
        q, planted = make_queries(orc, nq, dim, total, SEED_Q ^ 0x5555)
    else:
        q, planted = np.empty((nq, dim), np.float32), {}
    if world > 1:
        qt = torch.from_numpy(q).cuda()
        dist.broadcast(qt, 0)
        q = qt.cpu().numpy()
    d_q = tss.DeviceBuffer(device, q.nbytes).upload(q)
    d_k = tss.DeviceBuffer(device, nq * k * 8)
    ev0, ev1 = tss.Event(device), tss.Event(device)

    def leg(index, first, count, out):
        for i in range(first, first + count):
            index.search_device(_Slice(d_q.ptr + i * dim * 4), 1, k, _Slice(out.ptr + i * k * 8))

    leg(ix, 0, warm, d_k)
    ix.sync()
    barrier()
    ev0.record(ix)
    leg(ix, warm, steps, d_k)
    ev1.record(ix)
    ix.sync()
    ms = max_over_ranks(ev0.elapsed_ms(ev1)) / steps
    barrier()
    # the scan alone (no exchange): per-GPU bandwidth on a 19.2 GB (N=8) shard
    scan_ms = ms
    if world > 1:
        ix.set_shard(b, None)
        d_tmp = tss.DeviceBuffer(device, nq * k * 8)
        leg(ix, 0, warm, d_tmp)
        ix.sync()
        barrier()
        ev0.record(ix)
        leg(ix, warm, steps, d_tmp)
        ev1.record(ix)
        ix.sync()
        scan_ms = max_over_ranks(ev0.elapsed_ms(ev1)) / steps
        ix.set_shard(b, comm)
    # SURVEY 8(d) config 5 also names B=16: one 16-query call per step.  A sharded fp32 index
    # answers it with one tensor-core pass over its bf16 shadow (K2, results bit-identical); where
    # the shadow does not fit (N=1: 153.6 + 76.8 GB) it is four 4-query scans.
    nb16, it16 = 16, (5 if world > 1 else 3)
    d_k16 = tss.DeviceBuffer(device, nb16 * k * 8)
    b16 = None
    try:
        ix.search_device(d_q, nb16, k, d_k16)
        ix.sync()
        barrier()
        l0 = tss.launch_count()
        ev0.record(ix)
        for _ in range(it16):
            ix.search_device(d_q, nb16, k, d_k16)
        ev1.record(ix)
        ix.sync()
        ms16 = max_over_ranks(ev0.elapsed_ms(ev1)) / it16
        b16 = {"batch": nb16, "value": nb16 / ms16 * 1e3, "unit": UNIT, "ms_per_batch": ms16,
               "gpu_launches_per_batch": (tss.launch_count() - l0) / it16}
    except tss.TssError as e:   # (a rank without room for its shadow fails the call, loudly)
        b16 = {"batch": nb16, "skipped": str(e)}
    barrier()
    # the same queries over the 10M-row index (also sharded): rows [0, 10M) are the same rows
    d_k10 = tss.DeviceBuffer(device, nq * k * 8)
    leg(ix10, 0, nq, d_k10)
    ix10.sync()
    barrier()
    check = "skipped"
    if rank == 0:
        k100 = d_k.download(np.uint64, nq * k).reshape(nq, k)
        k10 = d_k10.download(np.uint64, nq * k).reshape(nq, k)
        r100, s100 = tss.unpack_keys(k100)
        ok = all(int(r100[i][0]) == row for i, row in planted.items())
        # slice consistency: the keys of the 100M result that lie in [0, 10M) are exactly the keys
        # of the 10M result above the 100M result's last key (keys order (score desc, row asc))
        for i in range(nq):
            last = k100[i][k - 1]
            in_slice = [int(x) for x, r in zip(k100[i], r100[i]) if r < args.rows]
            above = [int(x) for x in k10[i] if x >= last]
            ok = ok and in_slice == above
        # score bits of two queries' winners recomputed on the CPU from the generator
        for i in (warm, nq - 1):
            for r, s in zip(r100[i], s100[i]):
                e = orc.gen_rows(int(r), 1, dim, SEED_ROWS)
                ok = ok and orc.scores(e, q[i])[0].view(np.uint32) == s.view(np.uint32)
        if b16 is not None and "value" in b16:
            k16 = d_k16.download(np.uint64, nb16 * k).reshape(nb16, k)
            b16["keys_equal_batch1_leg"] = bool(np.array_equal(k16, k100[:nb16]))
            ok = ok and b16["keys_equal_batch1_leg"]
        check = "ok" if ok else "FAILED"
    ix.close()
    algo = n_local * dim * 4
    return {
        "workload": (f"synthetic {total}x{dim} f32 corpus row-sharded over {world} GPUs "
                     f"({n_local} rows = {algo / 1e9:.2f} GB per GPU), batch-1 exact top-{k}, fused "
                     "peer-memory exchange + merge") if world > 1 else
                    (f"synthetic {total}x{dim} f32 corpus ({algo / 1e9:.1f} GB) resident in ONE B200's "
                     f"HBM, batch-1 exact top-{k}: the N=1 point of config 5's scaling curve"),
        "value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "steps": steps, "scaling": "weak-ish: "
        "total rows fixed at 100M, so the shard shrinks with N (SURVEY 8d config 5)",
        "scan_only_ms": scan_ms, "per_gpu_gbs": algo / (scan_ms * 1e-3) / 1e9,
        "per_gpu_gbs_incl_exchange": algo / (ms * 1e-3) / 1e9,
        "ideal_qps_at_8TBs": 1.0 / (algo / 8e12), "batch16": b16, "check": check,
        "checks": "planted winners; 100M result vs the 10M-row index on rows < 10M (key for key); "
                  "score bits of 20 winners recomputed by the oracle",
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
    sampler = ClockSampler(device) if rank == 0 else None  # runs through warm-up and timed region

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

    orc = load_orc() if (rank == 0 or world > 1) else None
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

    def barrier():
        ix.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

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
    device_leg(0, args.warmup)  # (the clock sampler has been running since before this warm-up)
    launches0 = tss.launch_count()
    barrier()  # immediately before the first event: every rank starts its timed steps together
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

    # per-step medians: an event between consecutive steps (which costs the overlap that
    # programmatic dependent launch buys, so this is a separate leg, not `value`)
    nmed = min(args.steps, 40)
    evs = [tss.Event(device) for _ in range(nmed + 1)]
    barrier()
    evs[0].record(ix)
    for i in range(nmed):
        device_leg(args.warmup + i, 1)
        evs[i + 1].record(ix)
    barrier()
    per_step = [evs[i].elapsed_ms(evs[i + 1]) for i in range(nmed)]
    ms_median = max_over_ranks(statistics.median(per_step))

    # ---- roofline leg: the scan kernel alone (no exchange / merge across ranks) ---------------
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
    e2e_local = time.perf_counter() - t0
    barrier()
    e2e_s = max_over_ranks(e2e_local)
    last = (out_rows[nq_total - 1:nq_total], out_scores[nq_total - 1:nq_total])
    e2e = {"value": args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": args.dim * 4,
           "d2h_bytes_per_step": args.k * 8, "ms_per_step": e2e_s / args.steps * 1e3}

    # the same steps with two searches in flight (tss_index_search_submit / _collect, one host
    # thread): every step still copies its query from host memory and unpacks its result into host
    # memory, but the next scan is enqueued while this one runs, so launch and wake-up latency and
    # the merges' tail hide behind the stream of the corpus.  Reported beside the blocking call's
    # number, not instead of it.
    depth = 2
    p_rows = np.empty((nq_total, args.k), np.uint32)
    p_scores = np.empty((nq_total, args.k), np.float32)
    p_counts = np.empty(nq_total, np.uint32)
    prp, psp, pcp = (a.ctypes.data for a in (p_rows, p_scores, p_counts))

    def pipelined_leg(first, count):
        tickets = []
        for i in range(first, first + count):
            if len(tickets) == depth:
                j, t = tickets.pop(0)
                ix.search_collect(t, prp + j * rs, psp + j * rs, pcp + j * 4)
            tickets.append((i, ix.search_submit(qp + i * qs, 1, args.k)))
        for j, t in tickets:
            ix.search_collect(t, prp + j * rs, psp + j * rs, pcp + j * 4)

    pipelined_leg(0, args.warmup)
    barrier()
    t0 = time.perf_counter()
    pipelined_leg(args.warmup, args.steps)
    pipe_local = time.perf_counter() - t0
    barrier()
    pipe_s = max_over_ranks(pipe_local)
    e2e["pipelined"] = {
        "value": args.steps / pipe_s, "unit": UNIT, "ms_per_step": pipe_s / args.steps * 1e3,
        "in_flight": depth, "api": "tss_index_search_submit / tss_index_search_collect, one host thread",
        "h2d_bytes_per_step": args.dim * 4, "d2h_bytes_per_step": args.k * 8,
    }

    # ---- sanity: the timed work is the real work ----------------------------------------------
    check = "skipped"
    notes = []
    if rank == 0:
        rows_out, scores_out = tss.unpack_keys(keys_value_leg)
        ok = True
        for i, row in planted.items():
            ok &= int(rows_out[i][0]) == row
        # the e2e leg returned the same rows and score bits as the device leg, for every step
        ok &= bool(np.array_equal(rows_out[args.warmup:], out_rows[args.warmup:]))
        ok &= bool(np.array_equal(scores_out[args.warmup:].view(np.uint32),
                                  out_scores[args.warmup:].view(np.uint32)))
        # recompute the last query's winners on the CPU from the generator: bit-exact scores
        i = nq_total - 1
        ok &= np.array_equal(rows_out[i], last[0][0])
        if args.storage == "f32":
            for r, s in zip(rows_out[i], scores_out[i]):
                e = orc.gen_rows(int(r), 1, args.dim, SEED_ROWS)
                ok &= orc.scores(e, queries[i])[0].view(np.uint32) == s.view(np.uint32)
        # ... and so did the pipelined leg
        same_p = bool(np.array_equal(out_rows, p_rows) and
                      np.array_equal(out_scores.view(np.uint32), p_scores.view(np.uint32)) and
                      np.array_equal(out_counts, p_counts))
        e2e["pipelined"]["results_equal_blocking_leg"] = same_p
        ok &= same_p
        # an end-to-end number above the device-only number is impossible for real work
        for name, v in (("e2e", e2e["value"]), ("e2e.pipelined", e2e["pipelined"]["value"])):
            if v > value * 1.03:
                ok = False
                notes.append(f"{name} {v:.1f} q/s exceeds the device-timed value {value:.1f} q/s")
        check = "ok" if ok else "FAILED"

    # ---- extra: the call the reference's merge actually makes -- top-50 (src/search.rs:251) -------
    k50 = None
    if args.k != 50 and n_local >= 50:
        n50, w50 = min(args.steps, 40), min(args.warmup, 3)
        d_o50 = tss.DeviceBuffer(device, (n50 + w50) * 50 * 8)
        for i in range(n50 + w50):
            if i == w50:
                barrier()
                ev0.record(ix)
            ix.search_device(_Slice(d_q.ptr + i * qstride), 1, 50, _Slice(d_o50.ptr + i * 50 * 8))
        ev1.record(ix)
        barrier()
        ms50 = max_over_ranks(ev0.elapsed_ms(ev1)) / n50
        same50 = True
        if rank == 0:   # its first k keys are the top-k leg's keys
            k50k = d_o50.download(np.uint64, (n50 + w50) * 50).reshape(n50 + w50, 50)
            same50 = bool(np.array_equal(k50k[:, :args.k], keys_value_leg[:n50 + w50])) if args.k < 50 else True
            if not same50:
                check = "FAILED"
                notes.append("top-50 leg disagrees with the top-k leg")
        k50 = {"workload": "batch-1, exact top-50: what SearchEngine::search_vector asks of the index "
                           "(src/search.rs:251)", "value": 1e3 / ms50, "unit": UNIT, "ms_per_step": ms50,
               "steps": n50, "first_k_keys_equal_topk_leg": same50}

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
        launches_b = (tss.launch_count() - l0) / iters
        kb = d_kb.download(np.uint64, nb * args.k).reshape(nb, args.k)
        same = bool(np.array_equal(kb[:min(nb, nq_total)], keys_value_leg[:min(nb, nq_total)]))
        # the same batch end to end: host queries in (1.5 MB H2D), host rows + scores out
        qb = np.ascontiguousarray(qb)
        b_rows, b_scores = np.empty((nb, args.k), np.uint32), np.empty((nb, args.k), np.float32)
        b_counts = np.empty(nb, np.uint32)
        ix.search_into(qb.ctypes.data, nb, args.k, b_rows.ctypes.data, b_scores.ctypes.data, b_counts.ctypes.data)
        t0 = time.perf_counter()
        for _ in range(iters):
            ix.search_into(qb.ctypes.data, nb, args.k, b_rows.ctypes.data, b_scores.ctypes.data,
                           b_counts.ctypes.data)
        be2e_ms = (time.perf_counter() - t0) / iters * 1e3
        rr, ss = tss.unpack_keys(kb)
        same = same and bool(np.array_equal(rr, b_rows) and
                             np.array_equal(ss.view(np.uint32), b_scores.view(np.uint32)))
        tf = 2.0 * nb * args.rows * args.dim / (bms * 1e-3) / 1e12
        batched = {
            "workload": f"{nb}-query batch, same index, exact cosine top-{args.k}", "batch": nb,
            "value": nb / bms * 1e3, "unit": UNIT, "ms_per_batch": bms,
            "e2e": {"value": nb / be2e_ms * 1e3, "unit": UNIT, "ms_per_batch": be2e_ms,
                    "h2d_bytes_per_batch": nb * args.dim * 4, "d2h_bytes_per_batch": nb * args.k * 8},
            "gpu_launches_per_batch": launches_b,
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

    # ---- the other BASELINE.json configs ---------------------------------------------------------
    extras = {}
    extras_ok = True
    if not args.no_extras and args.storage == "f32":
        if world == 1 and args.rows >= 4 * 256 * 100:
            extras["config3"] = config3_leg(tss, orc, args, ix, device, peaks)
            extras["config4"] = config4_leg(tss, orc, args, ix, device)
        if world == 1:
            # how fast rows arrive through tss_index_add (host fp32 -> pinned staging -> HBM, checked
            # for NaN/Inf and packed on the device; the host copy of chunk i+1 overlaps chunk i)
            nup = 262_144
            up_rows = orc.gen_rows(0, nup, args.dim, SEED_ROWS)
            ixu = tss.FlatIndex(args.dim, storage, device)
            ixu.reserve(2 * nup)
            ixu.add(up_rows)  # warm-up: allocates the staging buffers
            t0 = time.perf_counter()
            ixu.add(up_rows)
            ixu.finalize()
            up_s = time.perf_counter() - t0
            back = ixu.get_rows(nup + 12345, 2)
            extras["upload"] = {
                "what": "tss_index_add of host fp32 rows (pageable numpy memory) + finalize",
                "rows": nup, "bytes": int(up_rows.nbytes), "seconds": up_s,
                "gbs": up_rows.nbytes / up_s / 1e9,
                "check": "ok" if np.array_equal(back, up_rows[12345:12347]) else "FAILED"}
            ixu.close()
        if world > 1:
            # a 1024-query batch on the SHARDED index: local K2 (bf16 shadow + exact re-scoring),
            # NCCL all-gather of the keys, merge kernel
            nb = 1024
            if rank == 0:
                qb = orc.gen_rows(0, nb, args.dim, SEED_Q ^ 0x1234)
                qb[:min(nb, nq_total)] = queries[:min(nb, nq_total)]
            else:
                qb = np.empty((nb, args.dim), np.float32)
            qbt = torch.from_numpy(qb).cuda()
            dist.broadcast(qbt, 0)
            qb = qbt.cpu().numpy()
            d_qb = tss.DeviceBuffer(device, qb.nbytes).upload(qb)
            d_kb = tss.DeviceBuffer(device, nb * args.k * 8)
            for _ in range(2):
                ix.search_device(d_qb, nb, args.k, d_kb)
            barrier()
            iters = 5
            ev0.record(ix)
            for _ in range(iters):
                ix.search_device(d_qb, nb, args.k, d_kb)
            ev1.record(ix)
            barrier()
            bms = max_over_ranks(ev0.elapsed_ms(ev1)) / iters
            same = True
            if rank == 0:
                kb = d_kb.download(np.uint64, nb * args.k).reshape(nb, args.k)
                same = bool(np.array_equal(kb[:min(nb, nq_total)], keys_value_leg[:min(nb, nq_total)]))
            extras["batched_sharded"] = {
                "workload": (f"{nb}-query batch on the index sharded over {world} GPUs: local tcgen05 "
                             "GEMM top-k, ncclAllGather of the keys, merge kernel"),
                "value": nb / bms * 1e3, "unit": UNIT, "ms_per_batch": bms,
                "keys_equal_batch1_leg": same, "check": "ok" if same else "FAILED"}
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import dist_worker
            t0 = time.perf_counter()
            ok_local = dist_worker.run_checks(tss, orc, comm, rank, world, device, quick=True)
            flag = torch.tensor([1 if ok_local else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            extras["selftest"] = {
                "what": ("sharded fp32 / bf16 searches (fused exchange, NCCL gather + merge, K2, "
                         "back-to-back scans under PDL) == the oracle on the unsharded corpus, bit "
                         "for bit, on every rank (tests/dist_worker.py:run_checks)"),
                "ranks": world, "seconds": time.perf_counter() - t0,
                "check": "ok" if int(flag.item()) == 1 else "FAILED"}
            extras["config5"] = config5_leg(tss, orc, args, ix, comm, rank, world, device, barrier,
                                            max_over_ranks, dist, torch)
        for name, obj in extras.items():
            if obj.get("check") == "FAILED":
                extras_ok = False
                notes.append(f"{name} check failed")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(orc, args, queries)

    # config 5 at N=1: the 100M-row corpus on a single GPU (last: it takes most of the HBM)
    if (world == 1 and not args.no_extras and args.storage == "f32" and "config4" in extras
            and args.config5_rows * args.dim * 4 < 170e9):
        extras["config5"] = config5_leg(tss, orc, args, ix, None, rank, world, device, barrier,
                                        max_over_ranks, None, None)
        if extras["config5"].get("check") == "FAILED":
            extras_ok = False
            notes.append("config5 check failed")

    if not extras_ok and check == "ok":
        check = "FAILED"
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "ms_per_step_median": ms_median,
            "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.storage == "f32" else "bf16->f32",
            "data": "synthetic", "config": workload_config(args, world),
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "clocks": clocks, "check": check,
        }
        if notes:
            line["check_notes"] = notes
        if k50 is not None:
            line["top50"] = k50
        if batched is not None:
            line["batched"] = batched
        if prefiltered is not None:
            line["prefiltered"] = prefiltered
        line.update(extras)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line, default=_json_default), flush=True)
    if dist is not None:
        dist.barrier()
        if comm is not None:
            ix.close()
            comm.close()
        dist.destroy_process_group()
    return 1 if check == "FAILED" else 0


if __name__ == "__main__":
    sys.exit(main())
