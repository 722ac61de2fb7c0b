"""The tie-order sensitivity bounds of the oracle (the specification of the CUDA cut kernels' x_lo / x_hi) against
brute force: every order of the tied runs that numpy.argsort could legally produce."""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hica_oracle as O  # noqa: E402


def test_count_bounds_equal_brute_force_over_all_tie_orders():
    rng = np.random.RandomState(4)
    for trial in range(200):
        n = rng.randint(3, 9)
        vals = rng.randint(0, 4, size=n).astype(float)     # many exact ties
        order = np.argsort(vals, kind="stable")[::-1]
        lo = rng.randint(0, n)
        hi = rng.randint(lo, n)
        c = rng.randint(0, n + 1)
        x, xlo, xhi = O.count_bounds(vals[order], order, c, lo, hi)
        # all permutations that keep the values non-increasing
        seen = set()
        for perm in itertools.permutations(range(n)):
            p = np.asarray(perm)
            if np.all(np.diff(vals[p]) <= 0):
                seen.add(int(np.count_nonzero((p[:c] >= lo) & (p[:c] <= hi))))
        assert x in seen
        if c == 0 or c == n:
            assert xlo == xhi == x
        assert (xlo, xhi) == (min(seen), max(seen)), (vals, c, lo, hi)


def test_tie_free_rows_report_nothing():
    rng = np.random.RandomState(1)
    S = rng.rand(40, 40)
    R = O.rank_matrix_stable(S)
    assert O.tie_sensitive_breakpoints(S, R, 5, 0.0) == 0
    assert O.tie_sensitive_filter(S, R, O.pre_process_breakpoints(R, 5, 0.0)) == 0
