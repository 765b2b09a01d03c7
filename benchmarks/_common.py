"""Shared helpers of the benchmark scripts.  Benchmarks measure libtss; they do not import the
oracle (test infrastructure) except where one explicitly CHECKS a result (gemm_bench.py
--recall-queries)."""
import numpy as np


def make_queries(nq: int, dim: int, seed: int) -> np.ndarray:
    """nq seeded N(0,1) fp32 queries (not normalised: the kernels fuse the norms)."""
    return np.random.default_rng(seed).standard_normal((nq, dim)).astype(np.float32)
