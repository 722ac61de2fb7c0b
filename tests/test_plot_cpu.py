"""CPU: the host half of the plot-input path (hic_genome_assembler_b200/plot.py) - numpy.percentile's index and
interpolation arithmetic restated on top of order statistics (plotContactMaps.py:30-31), with numpy.sort standing in
for the device's radix select."""
import numpy as np
import pytest

from hic_genome_assembler_b200.plot import percentile_from_order_stats


def _fetch(x):
    s = np.sort(x[~np.isnan(x)])
    nan = int(np.isnan(x).sum())
    # as the device: a rank outside the non-NaN values answers NaN
    return lambda ranks: (np.asarray([s[r] if 0 <= r < len(s) else np.nan for r in ranks]), nan)


@pytest.mark.parametrize("seed", range(6))
def test_interpolation_equals_numpy_bit_for_bit(seed):
    rng = np.random.default_rng(seed)
    for t in range(400):
        n = int(rng.integers(1, 500))
        x = rng.random(n) * rng.choice([1.0, 1e-6, 1e6]) - rng.choice([0.0, 0.5])
        if t % 3 == 0:
            x = np.round(x, 1)
        if t % 40 == 0 and n > 2:
            x[rng.integers(0, n)] = np.nan
        q = [1, 98, 0, 100, 50, float(rng.random() * 100), int(rng.integers(0, 101))][t % 7]
        if t % 5 == 0:
            q = [1, 98, float(rng.random() * 100)]
        with np.errstate(all="ignore"):
            got, want = percentile_from_order_stats(n, q, _fetch(x)), np.percentile(x, q)
        assert type(got) is type(want)
        a, b = np.atleast_1d(got), np.atleast_1d(want)
        assert np.array_equal(a, b, equal_nan=True)             # -0.0 / +0.0 tie order is numpy's own business
        nz = (a != 0) & ~np.isnan(a)
        assert a[nz].tobytes() == b[nz].tobytes()


def test_rejects_what_numpy_rejects():
    x = np.arange(10.0)
    for q in (-1, 100.5, [5, 101]):
        with pytest.raises(ValueError, match="range"):
            percentile_from_order_stats(10, q, _fetch(x))
