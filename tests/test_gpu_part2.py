"""GPU parity of the Part2 score kernels against the oracle (bit-exact scores and picks)."""
import numpy as np
import pytest

import hica_oracle as O
import c_oracle as CO
from hic_genome_assembler_b200.synth import make_genome

pytestmark = pytest.mark.gpu


def _chrom(n=160, seed=5, nseg=8):
    rng = np.random.RandomState(seed)
    g = make_genome(n, n_chrom=1, seed=seed, scaffolds_per_chrom=nseg)
    A = g.A.numpy()
    segs = []
    for s in range(len(g.scaffold_names)):
        segs.append(np.nonzero(g.scaffold_of_bin == s)[0])
    return A, segs, rng


def _explicit(segs, cand):
    out = []
    for e in cand:
        b = segs[e >> 1]
        out.extend(b[::-1] if (e & 1) else b)
    return np.asarray(out, dtype=np.int32)


@pytest.mark.parametrize("n", [2, 3, 9, 40, 129, 160, 301])
def test_scores_of_explicit_orders(hica, n):
    rng = np.random.RandomState(n)
    M = rng.rand(n, n) * 5
    M = M + M.T
    hica.p2_upload(M)
    hica.p2_select(np.arange(n))
    orders = np.stack([rng.permutation(n) for _ in range(37)]).astype(np.int32)
    total = O.total_weight(M)
    best, idx, scores = hica.p2_score_orders(orders, float(total), 0.0, want_scores=True)
    ref = np.array([O.ordering_cost_of(M, o, total) for o in orders])
    assert np.array_equal(scores, ref)
    assert idx == int(np.argmax(ref)) and best == ref.max()


def test_asymmetric_matrix_uses_upper_positions(hica):
    rng = np.random.RandomState(1)
    n = 50
    M = rng.rand(n, n)
    hica.p2_upload(M)
    hica.p2_select(np.arange(n))
    orders = np.stack([rng.permutation(n) for _ in range(9)]).astype(np.int32)
    _, _, scores = hica.p2_score_orders(orders, 3.0, 0.0, want_scores=True)
    ref = np.array([O.ordering_cost_of(M, o, 3.0) for o in orders])
    assert np.array_equal(scores, ref)


def test_total_weight_bitwise(hica):
    for n, seed in [(7, 1), (9, 2), (40, 3), (129, 4), (160, 5), (300, 6), (801, 7)]:
        A, segs, rng = _chrom(n, seed, nseg=min(8, n // 2))
        hica.p2_upload(A)
        hica.p2_select(np.arange(n))
        ptr = np.concatenate([[0], np.cumsum([len(s) for s in segs])])
        hica.p2_set_segments(ptr, np.concatenate(segs))
        cand = np.array([(s << 1) | int(rng.rand() < .5) for s in rng.permutation(len(segs))], dtype=np.int32)
        o = _explicit(segs, cand)
        assert hica.p2_total(cand) == O.total_weight(A[np.ix_(o, o)])


def test_candidates_and_carried_threshold(hica):
    A, segs, rng = _chrom(160, 9, nseg=8)
    hica.p2_upload(A)
    hica.p2_select(np.arange(160))
    ptr = np.concatenate([[0], np.cumsum([len(s) for s in segs])])
    hica.p2_set_segments(ptr, np.concatenate(segs))
    cands = np.stack([[(s << 1) | int(rng.rand() < .5) for s in rng.permutation(8)] for _ in range(64)]).astype(np.int32)
    total = hica.p2_total(cands[0])
    best, idx, scores = hica.p2_score_candidates(cands, total, 0.0, want_scores=True)
    ref = np.array([O.ordering_cost_of(A, _explicit(segs, c), total) for c in cands])
    assert np.array_equal(scores, ref)
    assert idx == int(np.argmax(ref))
    # carried best above / equal to the maximum: nothing is accepted (strict '>')
    b2, i2, _ = hica.p2_score_candidates(cands, total, float(ref.max()))
    assert i2 == -1 and b2 == ref.max()
    # duplicate maximum: first index wins
    cands2 = np.concatenate([cands, cands[idx:idx + 1]])
    _, i3, _ = hica.p2_score_candidates(cands2, total, 0.0)
    assert i3 == idx


def test_window_enumeration_matches_reference_order(hica):
    A, segs, rng = _chrom(200, 13, nseg=9)
    nseg = len(segs)
    hica.p2_upload(A)
    hica.p2_select(np.arange(200))
    ptr = np.concatenate([[0], np.cumsum([len(s) for s in segs])])
    hica.p2_set_segments(ptr, np.concatenate(segs))
    base = np.array([(s << 1) | int(rng.rand() < .5) for s in rng.permutation(nseg)], dtype=np.int32)
    total = hica.p2_total(base)
    win_n = 4
    orders = np.array(O.order_candidates(win_n), dtype=np.uint8)
    orients = np.array([[0 if c == '+' else 1 for c in o] for o in O.orient_candidates(win_n)], dtype=np.uint8)
    for ws in (0, 2, nseg - win_n):
        best, idx, scores = hica.p2_score_window(base, ws, win_n, orders, orients, total, 0.0, want_scores=True)
        ref = []
        for od in orders:
            for orr in orients:
                cand = base.copy()
                for j in range(win_n):
                    cand[ws + j] = ((base[ws + od[j]] >> 1) << 1) | int(orr[j])
                ref.append(O.ordering_cost_of(A, _explicit(segs, cand), total))
        ref = np.array(ref)
        assert np.array_equal(scores, ref)
        assert idx == int(np.argmax(ref)) and best == ref.max()
        # a sub-range returns absolute indices
        b2, i2, _ = hica.p2_score_window(base, ws, win_n, orders, orients, total, 0.0, cand_lo=50, cand_hi=150)
        assert i2 == 50 + int(np.argmax(ref[50:150]))
