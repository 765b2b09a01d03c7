"""The error bound behind K2's exact re-scoring and the shadow prefilter.

Both paths pick candidates by a cheaper score A and prove the exact top-k (by the scan's score B)
is among them using |A - B| <= eps.  This checks the bound itself on the CPU, in float64:
  bf16 index:        A = q_bf16 . e / |e|        B = q . e / |e|       eps = |q| 2^-8
  fp32 + bf16 shadow: A = q_bf16 . e_t / |e_t|    B = q . e / |e|       eps = |q| 2^-7
  shadow prefilter:   A = q . e_t / (|q||e_t|)    B = q . e / (|q||e|)  eps = 2^-8 (1 + 2^-8)  (cosine)
(e_t = bf16(e).  bf16 keeps 8 significant bits, so round-to-nearest moves a vector by at most 2^-8
of its norm -- NOT 2^-9, which the first version of the kernels assumed and this test caught; the
(D + 4) * 2^-22 slack the kernels add for fp32 accumulation is not needed in float64) and
the selection argument: keeping every row with A >= t - 2 eps, t the k-th largest A, keeps the
whole top-k by B.
"""
import numpy as np
import pytest


def _bf16(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)


def _cases():
    rng = np.random.default_rng(11)
    for dim in (3, 64, 384, 1000):
        for scale in (1e-3, 1.0, 250.0):
            rows = (rng.standard_normal((4000, dim)) * scale).astype(np.float32)
            rows[::7] *= rng.uniform(0.01, 100.0, size=(len(rows[::7]), 1)).astype(np.float32)
            rows[5] = rows[4]                      # duplicates
            rows[6] = -rows[4]
            q = (rng.standard_normal((16, dim)) * rng.uniform(1e-2, 1e2)).astype(np.float32)
            q[0] = rows[17] * 3.0                  # a query parallel to a row
            q[1] = np.abs(q[1])                    # all components one sign: worst case for rounding
            yield dim, rows, q


@pytest.mark.parametrize("dim,rows,q", list(_cases()), ids=lambda v: str(v) if isinstance(v, int) else None)
def test_score_differences_stay_inside_eps(dim, rows, q):
    r64, q64 = rows.astype(np.float64), q.astype(np.float64)
    rt64, qb64 = _bf16(rows).astype(np.float64), _bf16(q).astype(np.float64)
    rn = np.linalg.norm(r64, axis=1)
    rtn = np.linalg.norm(rt64, axis=1)
    qn = np.linalg.norm(q64, axis=1)[:, None]
    b = (q64 @ r64.T) / rn                                   # unscaled by |q|, like the K2 epilogue
    a_bf16_index = (qb64 @ r64.T) / rn                       # the index rows ARE bf16 there: e_t = e
    a_shadow = (qb64 @ rt64.T) / rtn
    # the rounding itself: at most 2^-8 of the norm, and it does get close to it
    dq = np.linalg.norm(qb64 - q64, axis=1) / qn[:, 0]
    de = np.linalg.norm(rt64 - r64, axis=1) / rn
    assert dq.max() <= 2.0 ** -8 and de.max() <= 2.0 ** -8
    assert np.all(np.abs(a_bf16_index - b) <= qn * 2.0 ** -8 * (1 + 1e-12))
    assert np.all(np.abs(a_shadow - b) <= qn * 2.0 ** -7 * (1 + 1e-12))
    cos_b = b / qn
    cos_a = (q64 @ rt64.T) / rtn / qn
    assert np.all(np.abs(cos_a - cos_b) <= 2.0 ** -8 * (1 + 2.0 ** -8))


@pytest.mark.parametrize("k", [1, 10, 100])
def test_margin_keeps_the_whole_topk(k):
    rng = np.random.default_rng(k)
    dim, n = 384, 20000
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    rows[100:140] = rows[7] + 1e-3 * rng.standard_normal((40, dim)).astype(np.float32)  # a cluster
    q = rng.standard_normal((8, dim)).astype(np.float32)
    q[0] = rows[7]
    r64, q64 = rows.astype(np.float64), q.astype(np.float64)
    rt64, qb64 = _bf16(rows).astype(np.float64), _bf16(q).astype(np.float64)
    b = (q64 @ r64.T) / np.linalg.norm(r64, axis=1)
    a = (qb64 @ rt64.T) / np.linalg.norm(rt64, axis=1)
    eps = np.linalg.norm(q64, axis=1) * 2.0 ** -7
    for i in range(len(q)):
        t = np.sort(a[i])[-k]
        kept = set(np.nonzero(a[i] >= t - 2 * eps[i])[0].tolist())
        top_b = set(np.argsort(-b[i], kind="stable")[:k].tolist())
        assert top_b <= kept
        assert len(kept) < 60 * k + 400                      # and the margin stays selective


@pytest.mark.parametrize("dim", [3, 100, 384, 1024])
def test_scan_arithmetic_stays_inside_the_accumulation_slack(orc, dim):
    """The margins reserve (D + 4) * 2^-22 (relative to |q|, i.e. in cosine units) for the fp32
    accumulation of BOTH paths; the scan's share -- its canonical fp32 order against exact
    arithmetic -- has to fit in a quarter of that with room to spare."""
    rng = np.random.default_rng(dim)
    rows = rng.standard_normal((3000, dim)).astype(np.float32)
    rows[:500] = np.abs(rows[:500])            # no cancellation: rounding errors all push one way
    for q in (rng.standard_normal(dim).astype(np.float32), np.abs(rng.standard_normal(dim)).astype(np.float32)):
        for bf16 in (False, True):
            got = orc.scores(rows, q, bf16=bf16).astype(np.float64)
            r64 = (_bf16(rows) if bf16 else rows).astype(np.float64)
            exact = (r64 @ q.astype(np.float64)) / (np.linalg.norm(r64, axis=1) * np.linalg.norm(q.astype(np.float64)))
            assert np.max(np.abs(got - exact)) <= (dim + 4) * 2.0 ** -24


def test_prefilter_proof_on_the_oracle_scores(orc):
    """refine_kernel's logic restated with the oracle's own fp32 arithmetic at 400k rows: A = the
    scan's score over bf16-rounded rows, B = over the fp32 rows.  Whenever the kc-th best A lies
    more than 2 eps under the k-th, the top-k by (B desc, row asc) is inside the top-kc by A; and
    for ordinary data the proof goes through almost always (it must, or the prefilter would keep
    falling back to the fp32 scan)."""
    n, dim, k, kc = 400_000, 384, 10, 64
    rows = orc.gen_rows(0, n, dim, 0x5EED)
    qs = orc.gen_rows(0, 24, dim, 0xBEEF)
    qs[3] = rows[1234] + 0.125 * qs[3]
    two_eps = 2 * (2.0 ** -8 * (1 + 2.0 ** -8) + (dim + 4) * 2.0 ** -22)
    proven = 0
    for q in qs:
        a = orc.scores(rows, q, bf16=True)
        b = orc.scores(rows, q)
        order_a = np.lexsort((np.arange(n), -a.astype(np.float64)))[:kc]
        top_b = np.lexsort((np.arange(n), -b.astype(np.float64)))[:k]
        if a[order_a[kc - 1]] < a[order_a[k - 1]] - np.float32(two_eps):
            proven += 1
            assert set(top_b.tolist()) <= set(order_a.tolist())
    assert proven >= len(qs) - 2


@pytest.mark.parametrize("dim,rows,q", list(_cases()), ids=lambda v: str(v) if isinstance(v, int) else None)
@pytest.mark.parametrize("stored_bf16", [False, True])
def test_unit_row_shadow_stays_inside_eps(dim, rows, q, stored_bf16):
    """The shadow the tensor cores read holds UNIT rows: e_n = bf16(e * rsqrt(sum e^2)), computed
    in fp32 as launch_normalize_rows does.  Unmasked batches use the raw dot product
    A' = q_bf16 . e_n as the score (no per-row weight), masked ones A = q_bf16 . e_n / |e_n|;
    both stay within eps = |q| 2^-7 of B = q . e / |e| (the margin prep_queries_kernel uses for
    a shadow, minus its accumulation slack), for fp32 and for bf16 stored rows; and the cosine the
    shadow prefilter scans, cos(q, e_n), within 2^-8 (1 + 2^-8) of cos(q, e)."""
    stored = _bf16(rows) if stored_bf16 else rows
    ss = np.einsum("ij,ij->i", stored, stored, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(ss > 0, np.float32(1.0) / np.sqrt(ss, dtype=np.float32), np.float32(0)).astype(np.float32)
    unit = _bf16((stored * inv[:, None]).astype(np.float32))
    r64, q64 = stored.astype(np.float64), q.astype(np.float64)
    u64, qb64 = unit.astype(np.float64), _bf16(q).astype(np.float64)
    rn = np.linalg.norm(r64, axis=1)
    un = np.linalg.norm(u64, axis=1)
    qn = np.linalg.norm(q64, axis=1)[:, None]
    b = (q64 @ r64.T) / rn
    slack = qn * (dim + 4) * 2.0 ** -22            # fp32 rsqrt / multiply of the normalisation
    assert np.all(np.abs(un - 1.0) <= 2.0 ** -8 + 1e-6)
    a_raw = qb64 @ u64.T
    a_weighted = a_raw / un
    assert np.all(np.abs(a_raw - b) <= qn * 2.0 ** -7 + slack)
    assert np.all(np.abs(a_weighted - b) <= qn * 2.0 ** -7 + slack)
    cos_a = (q64 @ u64.T) / un / qn
    assert np.all(np.abs(cos_a - b / qn) <= 2.0 ** -8 * (1 + 2.0 ** -8) + (dim + 4) * 2.0 ** -22)


@pytest.mark.parametrize("dim,rows,q", list(_cases()), ids=lambda v: str(v) if isinstance(v, int) else None)
@pytest.mark.parametrize("stored_bf16", [False, True])
def test_measured_margins_bound_the_difference(dim, rows, q, stored_bf16):
    """Round 2 builds the margins from MEASURED rounding distances instead of their worst case:
    dq = |q_bf16 - q| of the query at hand (prep_queries_kernel) and de = max over rows of
    |e_n - e/|e|| recorded when the unit-row shadow was built (normalize_rows_kernel):
        |A - B| <= dq (1 + de) + |q| de (1 + de)      (raw dot and weighted, both)
        |cos(q, e_n) - cos(q, e)| <= de (1 + de)      (shadow prefilter)
    and without a shadow (a bf16 index's own rows) |A - B| <= dq.  On ordinary data both
    distances are ~0.4-0.5 x 2^-8, so the margin is less than half its worst case."""
    stored = _bf16(rows) if stored_bf16 else rows
    ss = np.einsum("ij,ij->i", stored, stored, dtype=np.float32)
    inv = (np.float32(1.0) / np.sqrt(ss, dtype=np.float32)).astype(np.float32)
    x = (stored * inv[:, None]).astype(np.float32)
    unit = _bf16(x)
    de = float(np.max(np.linalg.norm(unit.astype(np.float64) - x.astype(np.float64), axis=1))) * 1.002
    assert de <= 2.0 ** -8 * 1.002
    r64, q64 = stored.astype(np.float64), q.astype(np.float64)
    u64, qb64 = unit.astype(np.float64), _bf16(q).astype(np.float64)
    rn, un = np.linalg.norm(r64, axis=1), np.linalg.norm(u64, axis=1)
    qn = np.linalg.norm(q64, axis=1)[:, None]
    dq = np.linalg.norm(qb64 - q64, axis=1)[:, None]
    assert np.all(dq <= qn * 2.0 ** -8)
    b = (q64 @ r64.T) / rn
    slack = qn * (dim + 4) * 2.0 ** -22
    eps = dq * (1 + de) + qn * de * (1 + de) + slack
    a_raw = qb64 @ u64.T
    assert np.all(np.abs(a_raw - b) <= eps)
    assert np.all(np.abs(a_raw / un - b) <= eps)
    cos_a = (q64 @ u64.T) / un / qn
    assert np.all(np.abs(cos_a - b / qn) <= de * (1 + de) + (dim + 4) * 2.0 ** -22)
    # a bf16 index read directly (stored rows ARE what the tensor cores see): only the query moves
    if stored_bf16:
        assert np.all(np.abs((qb64 @ r64.T) / rn - b) <= dq * (1 + 1e-12))
    if dim >= 64:  # the point of measuring: well under the worst case on ordinary data
        assert de < 0.7 * 2.0 ** -8 and np.median(dq / qn) < 0.7 * 2.0 ** -8
