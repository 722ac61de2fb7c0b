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


# ---- a numpy model of the device radix select (csrc/hica_plot.cu: k_os_count + k_os_decide) ----------------------

def _keys(x):
    u = np.where(x == 0, 0.0, x).view(np.uint64)            # -0.0 is keyed as +0.0
    neg = (u >> np.uint64(63)).astype(bool)
    return np.where(neg, ~u, u | np.uint64(1 << 63))


def _unkey(k):
    k = np.uint64(k)
    u = (k & np.uint64((1 << 63) - 1)) if (k >> np.uint64(63)) else ~k
    return np.array([u], dtype=np.uint64).view(np.float64)[0]


def _radix_select_model(x, ranks):
    """Six passes (5 x 11 bits + 9): histogram of the next digit over the elements that match a rank's prefix,
    pick the bucket that holds the rank, descend; ranks with equal prefixes share a histogram."""
    keep = ~np.isnan(x)
    keys = _keys(x[keep])
    prefix = [0] * len(ranks)
    k = list(ranks)
    bad = [not (0 <= r < keys.size) for r in ranks]
    k = [0 if b else r for r, b in zip(k, bad)]
    for p in range(6):
        shift, bits = (53 - 11 * p, 11) if p < 5 else (0, 9)
        hists = {}
        for r in range(len(ranks)):
            if prefix[r] not in hists:
                m = keys if p == 0 else keys[(keys >> np.uint64(shift + bits)) == np.uint64(prefix[r])]
                digit = ((m >> np.uint64(shift)) & np.uint64((1 << bits) - 1)).astype(np.int64)
                hists[prefix[r]] = np.bincount(digit, minlength=1 << bits)
            h = hists[prefix[r]]
            c = np.cumsum(h)
            b = int(np.searchsorted(c, k[r], side="right")) if h.sum() else (1 << bits) - 1
            b = min(b, (1 << bits) - 1)
            k[r] -= int(c[b] - h[b])
            prefix[r] = (prefix[r] << bits) | b
    return [np.nan if bad[r] else _unkey(prefix[r]) for r in range(len(ranks))], int((~keep).sum())


@pytest.mark.parametrize("kind", ["random", "zeros", "signed", "inf", "constant", "nan"])
def test_radix_select_model_returns_the_exact_order_statistics(kind):
    rng = np.random.RandomState(len(kind))
    n = 5000
    x = rng.rand(n) * 10
    if kind == "zeros":
        x[rng.rand(n) < 0.7] = 0.0
    elif kind == "signed":
        x = (x - 5) * 10.0 ** rng.randint(-300, 300, n)
        x[::7] = -0.0
    elif kind == "inf":
        x[::5] = np.inf
        x[1::9] = -np.inf
    elif kind == "constant":
        x[:] = 2.5
    elif kind == "nan":
        x[n // 2] = np.nan
    s = np.sort(x[~np.isnan(x)])
    ranks = [0, 1, 49, 50, 2500, s.size - 2, s.size - 1, s.size, -1]
    got, nan = _radix_select_model(x, ranks)
    assert nan == int(np.isnan(x).sum())
    for r, v in zip(ranks, got):
        if 0 <= r < s.size:
            assert v == s[r]
        else:
            assert np.isnan(v)


def test_plot_numpy_shim_routes_only_the_reference_call():
    """launcher._PlotNumpyShim: numpy.percentile(adjMat, q) on a big square matrix goes to the seam (the host copy
    of a device-backed matrix is recognised), everything else is numpy's."""
    from hic_genome_assembler_b200.launcher import _PlotNumpyShim
    from hic_genome_assembler_b200.lazy import LazyMatrix

    calls = []

    class FakeSeams:
        def percentile(self, a, q):
            calls.append((a, q))
            return 42.0

    shim = _PlotNumpyShim(np, FakeSeams())
    small = np.arange(16.0).reshape(4, 4)
    assert shim.percentile(small, 50) == np.percentile(small, 50) and not calls           # small: numpy
    big = np.zeros((256, 256))
    assert shim.percentile(big, 50, axis=0).shape == (256,) and not calls                  # other signature: numpy
    assert shim.percentile(big, 98) == 42.0 and calls[-1][0] is big
    lazy = LazyMatrix((256, 256), lambda: np.ones((256, 256)), tag=("Dperm", 1), kind="Dperm")
    host = shim.asarray(lazy)                                                              # plotContactMap :16-18
    assert isinstance(host, np.ndarray) and host.shape == (256, 256)
    assert shim.percentile(host, 1) == 42.0 and calls[-1][0] is lazy                       # the resident matrix
    assert shim.zeros is np.zeros and shim.asarray([1, 2]).tolist() == [1, 2]
