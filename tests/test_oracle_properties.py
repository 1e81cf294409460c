"""Property tests of the CPU oracle (hypothesis): the sort-free counting form the kernels implement equals the
reference's sort form (rank.py:117-169 under a stable argsort) on arbitrary small inputs -- heavy ties, junk,
invalid queries, +-inf, NaN, -0.0 -- and, when oracle/_ref is built, equals the compiled reference on tie-free inputs."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import ref, restatement as R

LEVELS = np.array([-np.inf, -1.5, -0.0, 0.0, 0.25, 0.5, 1.0, 2.0, 7.0, np.inf, np.nan], dtype=np.float32)


@st.composite
def retrieval_case(draw, tie_free=False):
    Q = draw(st.integers(1, 6))
    G = draw(st.integers(1, 24))
    n_pid = draw(st.integers(1, 5))
    n_cam = draw(st.integers(1, 3))
    rng = np.random.RandomState(draw(st.integers(0, 2 ** 31 - 1)))
    if tie_free:
        d = rng.permutation(Q * G).reshape(Q, G).astype(np.float32) / 3.0
    else:
        d = LEVELS[rng.randint(0, len(LEVELS), size=(Q, G))]
    return (d, rng.randint(0, n_pid, Q), rng.randint(0, n_pid, G), rng.randint(0, n_cam, Q), rng.randint(0, n_cam, G),
            draw(st.integers(1, 8)))


@settings(max_examples=150, deadline=None)
@given(retrieval_case())
def test_counting_form_equals_sort_form_everywhere(case):
    d, qp, gp, qc, gc, max_rank = case
    pos = R.kept_positions(d, qp, gp, qc, gc)
    valid = [p for p in pos if p.size]
    kept_min = min((int(((gp != qp[q]) | (gc != qc[q])).sum()) for q in range(len(qp)) if pos[q].size), default=0)
    if not valid:
        with pytest.raises(AssertionError):
            R.eval_market1501(d, qp, gp, qc, gc, max_rank)
        return
    if kept_min < min(max_rank, d.shape[1]):
        return                                    # the reference's ragged case: contract excludes it
    cmc_s, map_s = R.eval_market1501(d, qp, gp, qc, gc, max_rank)
    cmc_c, map_c, n_valid = R.metrics_from_positions(pos, max_rank, d.shape[1])
    assert np.array_equal(cmc_s, cmc_c) and abs(map_s - map_c) < 1e-12 and n_valid == len(valid)
    minp = float(np.mean([p.size / (p[-1] + 1.0) for p in valid]))
    assert abs(R.mean_inverse_negative_penalty(d, qp, gp, qc, gc) - minp) < 1e-12
    k = min(5, d.shape[1])
    idx, val = R.topk_kept(d, qp, gp, qc, gc, k)
    for q in range(len(qp)):                      # the ranked list is what the positions count against
        keep = ~((gp == qp[q]) & (gc == qc[q]))
        order = np.argsort(d[q], kind="stable")
        want = order[keep[order]][:k]
        assert np.array_equal(idx[q, :want.size], want) and (idx[q, want.size:] == -1).all()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@settings(max_examples=60, deadline=None)
@given(retrieval_case(tie_free=True))
def test_restatement_equals_compiled_reference_when_tie_free(case):
    d, qp, gp, qc, gc, max_rank = case
    pos = R.kept_positions(d, qp, gp, qc, gc)
    if not any(p.size for p in pos):
        return
    kept_min = min(int(((gp != qp[q]) | (gc != qc[q])).sum()) for q in range(len(qp)) if pos[q].size)
    if kept_min < min(max_rank, d.shape[1]):
        return
    cmc_r, map_r = ref.eval_market1501(d, qp, gp, qc, gc, max_rank)          # rank.py:103, unmodified, compiled
    cmc_o, map_o = R.eval_market1501(d, qp, gp, qc, gc, max_rank)
    assert np.array_equal(np.asarray(cmc_r, dtype=np.float32), cmc_o) and abs(map_r - map_o) < 1e-12
    cmc_y, map_y = ref.evaluate_cy(d, qp, gp, qc, gc, max_rank)              # rank_cy.pyx:26 (float32 AP)
    assert np.array_equal(np.asarray(cmc_y, dtype=np.float32), cmc_o) and abs(map_y - map_o) < 1e-5
