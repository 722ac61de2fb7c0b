"""GPU: the colour-scale clip of plotContactMap (plotContactMaps.py:30-31) - device order statistics
(hica_p1_order_stats / hica_order_stats_host) + numpy's interpolation arithmetic against numpy.percentile itself,
bit for bit."""
import numpy as np
import pytest

from hic_genome_assembler_b200 import plot as hplot
from hic_genome_assembler_b200.part1 import Seams
from hic_genome_assembler_b200.synth import make_genome

pytestmark = pytest.mark.gpu

QS = [1, 98, 0, 100, 50, 37.5, [1, 98], [0.0, 99.999, 12.5, 100.0]]


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    assert np.array_equal(a, b, equal_nan=True)           # -0.0 == +0.0 (the device keys them alike)
    nz = (a != 0) & ~np.isnan(a)
    assert a[nz].tobytes() == b[nz].tobytes()


def _values(kind, count, seed):
    rng = np.random.RandomState(seed)
    x = rng.rand(count) * 10
    if kind == "zeros":                                   # real maps are mostly exact zeros
        x[rng.rand(count) < 0.7] = 0.0
    elif kind == "rounded":                               # %.6f-style duplicates
        x = np.round(x, 1)
    elif kind == "signed":
        x = (x - 5) * 10.0 ** rng.randint(-300, 300, count)
        x[::7] = -0.0
        x[::11] = 0.0
    elif kind == "inf":
        x[::5] = np.inf
        x[1::9] = -np.inf
    elif kind == "constant":
        x[:] = 2.5
    elif kind == "nan":
        x[count // 2] = np.nan
    return x


@pytest.mark.parametrize("kind", ["random", "zeros", "rounded", "signed", "inf", "constant", "nan"])
@pytest.mark.parametrize("count", [1, 2, 33, 5000, 65536, 70001, (1 << 20) + 3])
def test_host_array_percentiles_equal_numpy(hica, kind, count):
    x = _values(kind, count, count % 1000 + len(kind))
    for q in QS:
        with np.errstate(all="ignore"):
            want = np.percentile(x, q)
            got = hplot.percentile_device(hica, x, q)
        _same(got, want)


def test_order_statistics_any_number_of_ranks(hica):
    x = _values("zeros", 300000, 5)
    s = np.sort(x)
    ranks = [0, 1, 17, 209999, 210000, 210001, 250000, 299998, 299999, 123456, 5]
    vals, nan = hica.order_stats_host(x, ranks)
    assert nan == 0 and np.array_equal(vals, s[ranks])
    vals, nan = hica.order_stats_host(x, [300000, -1, 7])          # out of range: NaN, the valid one is right
    assert np.isnan(vals[0]) and np.isnan(vals[1]) and vals[2] == s[7]
    x[3] = np.nan
    vals, nan = hica.order_stats_host(x, [])
    assert nan == 1 and vals.size == 0


@pytest.mark.parametrize("n", [37, 300, 1201])
def test_resident_matrix_percentiles(hica, n):
    g = make_genome(n, n_chrom=4, seed=n, scaffolds_per_chrom=3)
    A = g.A.numpy().copy()
    A[np.random.RandomState(n).rand(n, n) < 0.4] = 0.0
    A = np.triu(A) + np.triu(A, 1).T + np.eye(n)
    hica.p1_upload(A)
    hica.p1_set_row_sums(None)
    hica.p1_distance()
    D = hica.p1_get_distance()
    order = np.random.RandomState(n + 1).permutation(n).astype(np.int32)
    hica.p1_set_order(order)
    S = hica.p1_get_reordered(1)
    for q in QS:
        _same(hplot.percentile_device(hica, None, q, what=0), np.percentile(D, q))
        _same(hplot.percentile_device(hica, None, q, what=1), np.percentile(S, q))
        _same(hplot.percentile_device(hica, None, q, what=2), np.percentile(A, q))


def test_percentile_seam_reads_the_resident_matrix(hica):
    """Seams.percentile on the device-backed objects runPipeline hands to plotContactMap (:1124): no read-back."""
    n = 500
    g = make_genome(n, n_chrom=5, seed=3, scaffolds_per_chrom=4)
    A = g.A.numpy()

    class B:
        def __init__(self, rs):
            self.rowSum = rs
    bins = [B(float(v)) for v in A.sum(axis=1)]
    s = Seams(hica)
    D = s.convertMatrix(A.copy(), bins, distance=True)
    dend = s.averageClusterNodes(D, [str(i) for i in range(n)])
    Dp, bins2 = s.reorderMatrix(D, bins, dend["leaves"])
    got = [s.percentile(Dp, q) for q in (1, 98)]
    assert not Dp.materialised                                    # answered from the resident matrix
    want = [np.percentile(np.asarray(Dp), q) for q in (1, 98)]
    _same(got, want)
    host = np.asarray(Dp)
    _same(s.percentile(host, 98), want[1])                        # a plain host array: uploaded, same answer
    Sm = s.convertMatrix(Dp, bins2, distance=False, similarity=True)
    gs = s.percentile(Sm, [1, 98])
    assert not Sm.materialised
    _same(gs, np.percentile(np.asarray(Sm), [1, 98]))
